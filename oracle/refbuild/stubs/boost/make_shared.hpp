// Stand-in for boost/make_shared.hpp (Boost is absent): nothing on the compiled path uses it.
#pragma once
#include <memory>

// Stand-in for boost/shared_ptr.hpp (Boost is absent): nothing on the compiled path uses it.
#pragma once
#include <memory>

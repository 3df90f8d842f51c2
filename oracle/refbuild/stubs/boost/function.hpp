// Stand-in for <boost/function.hpp> (Boost is absent from this image): used by the reference's
// src/math/coordinates.hpp:28 only as a type-erased callable.  Test infrastructure, not product code.
#pragma once
#include <functional>
namespace boost { template <class S> using function = std::function<S>; }

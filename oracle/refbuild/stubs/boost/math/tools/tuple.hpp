// Stand-in for <boost/math/tools/tuple.hpp> (only make_tuple is named, PBSM3D.cpp:1066, on an optional branch).
#pragma once
#include <tuple>
namespace boost { namespace math { using std::make_tuple; } }

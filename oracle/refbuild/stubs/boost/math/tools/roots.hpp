// Stand-in for <boost/math/tools/roots.hpp>.  PBSM3D.cpp calls bracket_and_solve_root / newton_raphson_iterate only on
// the optional z0_ustar_coupling, use_subgrid_topo_V2 and iterative_subl branches (PBSM3D.cpp:552,708,1090), which the
// B200 path refuses (PBSM3D_ERR_UNSUPPORTED); the harness aborts loudly rather than guess Boost's iteration.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <utility>
namespace boost {
typedef std::uintmax_t uintmax_t;
namespace math { namespace tools {
[[noreturn]] inline void refharness_unsupported(const char* what)
{
    std::fprintf(stderr, "oracle/refbuild: %s is not available in the reference harness (optional PBSM3D branch)\n", what);
    std::abort();
}
template <class F, class T, class Tol>
std::pair<T, T> bracket_and_solve_root(F, const T&, T, bool, Tol, boost::uintmax_t&)
{
    refharness_unsupported("boost::math::tools::bracket_and_solve_root");
}
template <class F, class T> T newton_raphson_iterate(F, T, T, T, int) { refharness_unsupported("boost::math::tools::newton_raphson_iterate"); }
}}}

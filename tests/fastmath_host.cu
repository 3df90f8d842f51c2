// Host build of chm_b200/csrc/pbsm3d_math.cuh (the same source the device compiles): prints the largest relative error of
// flog / fexp / fpow against libm over the argument ranges of the PBSM3D assembly.  Driven by tests/test_fastmath.py.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../chm_b200/csrc/pbsm3d_math.cuh"
using namespace pbsm3d;
int main() {
    std::mt19937_64 rng(12345);
    auto uni = [&](double a, double b) { return a + (b - a) * (double)(rng() >> 11) * (1.0 / 9007199254740992.0); };
    double elog = 0, eexp = 0, epow = 0, ecbrt = 0;
    for (int i = 0; i < 2000000; ++i) {
        // heights 0.01..60 m, radii 1e-6..1e-3, wind 0.01..60, temperatures 200..320, densities/ratios 1e-12..1e6
        const double x = std::exp(uni(std::log(1e-12), std::log(1e6)));
        const double l = flog(x), lr = std::log(x);
        elog = std::fmax(elog, std::fabs(l - lr) / std::fmax(std::fabs(lr), 1e-3));  // absolute near ln 1 = 0
        const double a = uni(-40.0, 40.0);
        const double e = fexp(a), er = std::exp(a);
        eexp = std::fmax(eexp, std::fabs(e - er) / er);
        const double y = uni(-2.0, 2.0), xb = std::exp(uni(std::log(1e-6), std::log(1e3)));
        const double p = fpow(xb, y), pr = std::pow(xb, y);
        epow = std::fmax(epow, std::fabs(p - pr) / pr);
        const double c = uni(1.0, 1.5);
        ecbrt = std::fmax(ecbrt, std::fabs(fcbrt_1_15(c) - std::cbrt(c)) / std::cbrt(c));
    }
    std::printf("{\"flog\": %.3e, \"fexp\": %.3e, \"fpow\": %.3e, \"fcbrt\": %.3e}\n", elog, eexp, epow, ecbrt);
    return 0;
}

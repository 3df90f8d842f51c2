// Stand-in for interpolation/interpolation.hpp.  PBSM3D.hpp:27 includes it and never interpolates; scale_wind_vert uses
// interp_alg::tpspline through interpolation::init / operator() (scale_wind_vert.cpp:151,206).
// The thin plate spline below is a RESTATEMENT of interpolation/TPSpline.cpp:40-173 (same matrix, same basis
// Rd = -(log x + 0.577215 + E1(x)), x = (d*0.01/2)^2, same evaluation), because that file needs Eigen (FullPivLU), GSL
// (gsl_sf_expint_E1) and FunC, none of which exist here: E1 is the convergent series / continued fraction of its definition,
// the solve is Gaussian elimination with full pivoting.  It is pinned separately on the reference's own known-answer tests
// (src/tests/test_interpolation.cpp:47-170) in tests/test_wind_oracle.py.  Test infrastructure, never linked into the product.
#pragma once
#include <boost/tuple/tuple.hpp>
#include <cmath>
#include <map>
#include <string>
#include <utility>
#include <vector>

enum class interp_alg { tpspline, idw, nearest_sta };

namespace chmref_tps {
inline double expint_E1(double x)
{
    const double euler = 0.57721566490153286061;
    if (x <= 1.0) {  // E1(x) = -gamma - ln x + sum_{k>=1} (-1)^{k+1} x^k / (k k!)
        double sum = 0.0, term = 1.0;
        for (int k = 1; k < 60; ++k) {
            term *= -x / k;
            const double add = -term / k;
            sum += add;
            if (std::fabs(add) < 1e-18 * std::fabs(sum)) break;
        }
        return -euler - std::log(x) + sum;
    }
    // modified Lentz continued fraction for x > 1
    double b = x + 1.0, c = 1e300, d = 1.0 / b, h = d;
    for (int i = 1; i < 200; ++i) {
        const double an = -1.0 * i * i;
        b += 2.0;
        d = 1.0 / (an * d + b);
        c = b + an / c;
        const double del = c * d;
        h *= del;
        if (std::fabs(del - 1.0) < 1e-16) break;
    }
    return h * std::exp(-x);
}
inline double basis(double dist)
{
    const double weight = 0.01, c = 0.577215;
    double dij = (dist * weight / 2.0) * (dist * weight / 2.0);
    return -(std::log(dij) + c + expint_E1(dij));
}
}

class interpolation {
public:
    interpolation() {}
    interpolation(interp_alg ia, std::size_t size = 0, std::map<std::string, std::string> config = {}) { init(ia, size, config); }
    void init(interp_alg ia, std::size_t = 0, std::map<std::string, std::string> = {}) { _ia = ia; }

    double operator()(std::vector<boost::tuple<double, double, double>>& s, boost::tuple<double, double, double>& q)
    {
        const int n = (int)s.size(), size = n + 1;
        std::vector<double> A((std::size_t)size * size, 0.0), b(size, 0.0), x(size, 0.0);
        auto a = [&](int r, int c) -> double& { return A[(std::size_t)r * size + c]; };
        for (int i = 0; i < n; ++i)
            for (int j = i; j < n; ++j) {
                const double xd = s[i].get<0>() - s[j].get<0>(), yd = s[i].get<1>() - s[j].get<1>();
                if (xd == 0. && yd == 0.) continue;
                const double Rd = chmref_tps::basis(std::sqrt(xd * xd + yd * yd));
                a(i, j + 1) = Rd;
                a(j, i + 1) = Rd;
            }
        for (int i = 0; i < size; ++i) { a(i, 0) = 1; a(size - 1, i) = 1; }
        a(size - 1, 0) = 0;
        for (int i = 0; i < n; ++i) b[i] = s[i].get<2>();
        // Gaussian elimination with full pivoting (Eigen::FullPivLU in the reference)
        std::vector<int> colperm(size);
        for (int i = 0; i < size; ++i) colperm[i] = i;
        for (int k = 0; k < size; ++k) {
            int pr = k, pc = k;
            double best = -1;
            for (int r = k; r < size; ++r)
                for (int c = k; c < size; ++c)
                    if (std::fabs(a(r, c)) > best) { best = std::fabs(a(r, c)); pr = r; pc = c; }
            if (pr != k) { for (int c = 0; c < size; ++c) std::swap(a(k, c), a(pr, c)); std::swap(b[k], b[pr]); }
            if (pc != k) { for (int r = 0; r < size; ++r) std::swap(a(r, k), a(r, pc)); std::swap(colperm[k], colperm[pc]); }
            for (int r = k + 1; r < size; ++r) {
                const double f = a(r, k) / a(k, k);
                for (int c = k; c < size; ++c) a(r, c) -= f * a(k, c);
                b[r] -= f * b[k];
            }
        }
        std::vector<double> y(size);
        for (int k = size - 1; k >= 0; --k) {
            double v = b[k];
            for (int c = k + 1; c < size; ++c) v -= a(k, c) * y[c];
            y[k] = v / a(k, k);
        }
        for (int k = 0; k < size; ++k) x[colperm[k]] = y[k];
        double z0 = x[0];
        const double ex = q.get<0>(), ey = q.get<1>();
        for (int i = 1; i < size; ++i) {
            const double xd = s[i - 1].get<0>() - ex, yd = s[i - 1].get<1>() - ey;
            z0 = z0 + x[i] * chmref_tps::basis(std::sqrt(xd * xd + yd * yd));
        }
        return z0;
    }
private:
    interp_alg _ia = interp_alg::tpspline;
};

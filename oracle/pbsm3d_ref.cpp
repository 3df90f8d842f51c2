// CPU restatement (C++17 + OpenMP) of the reference PBSM3D timestep.   TEST INFRASTRUCTURE / CPU BASELINE ONLY.
//
// A restatement written independently of oracle/pbsm3d_oracle.py.  Its assembly is pinned, through that oracle, to the reference's
// own PBSM3D.cpp compiled unmodified (oracle/_ref/libchmref.so, tests/test_reference_pin.py + tests/test_oracle_cpp.py: 1e-12);
// its solver (own GMRES(30) + thread-block-local ILUT) is pinned only through the contract ||b - Ax|| <= 1e-8 ||b||.  Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
// may load it; nothing under chm_b200/ does.
//
// What it follows:
//   assembly + saltation      src/modules/PBSM3D.cpp:417-1410      (OpenMP parallel-for over faces, as the reference)
//   matrix ordering/pattern   src/math/LinearAlgebra.cpp:31-155    (row = layer*G + cell_global_id, <= 6 entries)
//   solver                    src/math/LinearAlgebra.cpp:164-195,228-252: GMRES, 30 blocks, <= 1000 iterations,
//                             tol 1e-8, RIGHT preconditioner ILUT(level-of-fill 3.0, drop 1e-4) re-factored every
//                             solve and LOCAL to each rank (no overlap).  Here a "rank" is one OpenMP thread's
//                             contiguous block of faces, which is how the reference uses all the cores of a box
//                             (mpirun -np cores; its USE_OMP build option only threads the assembly loops).
//                             Trilinos is not vendored: ILUT is Saad's dual-threshold row algorithm with Ifpack2's
//                             fill rule (extra entries per row and per factor = ceil((lof-1)*nnz/(2n))).
//   flux integration          PBSM3D.cpp:1467-1503
//   deposition system/solve   PBSM3D.cpp:1516-1745
//   helpers                   physics/Atmosphere.cpp:32-38,62-80; math/coordinates.cpp:112-131; module_base.hpp:471-479
//   stdDryAirDensity          MeteoIO (not vendored) — restated from its published source, see oracle/pbsm3d_oracle.py
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr double kKappa = 0.4, kRhoIce = 917.0, kZUR = 50.0, kZ0 = 0.01;

inline bool chm_is_nan(double v) { return std::fabs(v - -9999.0) < 1e-5 || std::isnan(v); }
inline double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Csr {
    int n = 0;
    std::vector<int> ptr, col;
    std::vector<double> val;
};

// ------------------------------------------------------------------------------------------------ ILUT
struct Ilut {
    int n = 0;
    std::vector<int> lptr, lcol, uptr, ucol;  // L strictly lower (unit diagonal), U upper incl. diagonal first
    std::vector<double> lval, uval;

    // Saad, "ILUT: a dual threshold incomplete LU factorization" (1994), row version.
    void factor(const Csr& A, double droptol, double lof) {
        n = A.n;
        const double nnz = (double)A.ptr[n];
        const int fill = (int)std::ceil(((lof - 1.0) * nnz) / (2.0 * n));  // Ifpack2::ILUT::compute
        lptr.assign(1, 0); uptr.assign(1, 0);
        lcol.clear(); lval.clear(); ucol.clear(); uval.clear();
        lcol.reserve((size_t)nnz * 2); lval.reserve((size_t)nnz * 2);
        ucol.reserve((size_t)nnz * 2); uval.reserve((size_t)nnz * 2);
        std::vector<double> w(n, 0.0);
        std::vector<char> mark(n, 0);
        std::vector<int> lidx, uidx;  // columns < i (kept sorted), columns >= i
        std::vector<std::pair<double, int>> keep;
        for (int i = 0; i < n; ++i) {
            lidx.clear(); uidx.clear();
            double tn = 0.0;
            int nl = 0, nu = 0;
            for (int p = A.ptr[i]; p < A.ptr[i + 1]; ++p) {
                int j = A.col[p];
                w[j] = A.val[p]; mark[j] = 1;
                tn += A.val[p] * A.val[p];
                if (j < i) { lidx.push_back(j); ++nl; } else { uidx.push_back(j); ++nu; }
            }
            if (!mark[i]) { uidx.push_back(i); mark[i] = 1; w[i] = 0.0; }
            tn = std::sqrt(tn);
            const double thr = droptol * tn;
            std::sort(lidx.begin(), lidx.end());
            for (size_t q = 0; q < lidx.size(); ++q) {  // lidx grows with fill, stays sorted
                int k = lidx[q];
                double piv = w[k] / uval[uptr[k]];  // U row k stores its diagonal first
                if (std::fabs(piv) < thr) { w[k] = 0.0; continue; }
                w[k] = piv;
                for (int p = uptr[k] + 1; p < uptr[k + 1]; ++p) {
                    int j = ucol[p];
                    if (!mark[j]) {
                        mark[j] = 1; w[j] = 0.0;
                        if (j < i) lidx.insert(std::upper_bound(lidx.begin() + q + 1, lidx.end(), j), j);
                        else uidx.push_back(j);
                    }
                    w[j] -= piv * uval[p];
                }
            }
            // L part: the (nl + fill) largest entries above the threshold
            keep.clear();
            for (int k : lidx) if (w[k] != 0.0 && std::fabs(w[k]) >= thr) keep.push_back({std::fabs(w[k]), k});
            size_t maxl = (size_t)(nl + fill);
            if (keep.size() > maxl) { std::nth_element(keep.begin(), keep.begin() + maxl, keep.end(), std::greater<>()); keep.resize(maxl); }
            std::sort(keep.begin(), keep.end(), [](auto& a, auto& b) { return a.second < b.second; });
            for (auto& e : keep) { lcol.push_back(e.second); lval.push_back(w[e.second]); }
            lptr.push_back((int)lcol.size());
            // U part: diagonal always, then the (nu + fill - 1) largest off-diagonals above the threshold
            double d = w[i];
            if (d == 0.0) d = (1e-4 + droptol) * (tn > 0 ? tn : 1.0);  // Saad's zero-pivot guard
            ucol.push_back(i); uval.push_back(d);
            keep.clear();
            for (int k : uidx) if (k != i && std::fabs(w[k]) >= thr && w[k] != 0.0) keep.push_back({std::fabs(w[k]), k});
            size_t maxu = (size_t)std::max(0, nu + fill - 1);
            if (keep.size() > maxu) { std::nth_element(keep.begin(), keep.begin() + maxu, keep.end(), std::greater<>()); keep.resize(maxu); }
            std::sort(keep.begin(), keep.end(), [](auto& a, auto& b) { return a.second < b.second; });
            for (auto& e : keep) { ucol.push_back(e.second); uval.push_back(w[e.second]); }
            uptr.push_back((int)ucol.size());
            for (int k : lidx) { mark[k] = 0; w[k] = 0.0; }
            for (int k : uidx) { mark[k] = 0; w[k] = 0.0; }
        }
    }
    // y = (LU)^{-1} v  (in place on y)
    void solve(double* y) const {
        for (int i = 0; i < n; ++i) {
            double s = y[i];
            for (int p = lptr[i]; p < lptr[i + 1]; ++p) s -= lval[p] * y[lcol[p]];
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int p = uptr[i] + 1; p < uptr[i + 1]; ++p) s -= uval[p] * y[ucol[p]];
            y[i] = s / uval[uptr[i]];
        }
    }
};

// Block-local (rank-local) ILUT: block b owns faces [fs[b], fs[b+1]) in every layer.
struct BlockPrec {
    int T = 0, L = 0;
    std::vector<int> fs;
    std::vector<Ilut> ilu;
    std::vector<std::vector<double>> work;

    void setup(const Csr& A, int T_, int L_, int nblocks, double droptol, double lof) {
        T = T_; L = L_;
        nblocks = std::max(1, std::min(nblocks, T));
        fs.resize(nblocks + 1);
        for (int b = 0; b <= nblocks; ++b) fs[b] = (int)((int64_t)T * b / nblocks);
        ilu.resize(nblocks); work.resize(nblocks);
#pragma omp parallel for schedule(static, 1)
        for (int b = 0; b < nblocks; ++b) {
            const int f0 = fs[b], nb = fs[b + 1] - f0;
            Csr B;
            B.n = nb * L;
            B.ptr.assign(1, 0);
            for (int z = 0; z < L; ++z)
                for (int i = 0; i < nb; ++i) {
                    int r = z * T + f0 + i;
                    for (int p = A.ptr[r]; p < A.ptr[r + 1]; ++p) {
                        int c = A.col[p], cz = c / T, cf = c - cz * T;
                        if (cf < f0 || cf >= f0 + nb) continue;  // coupling to another rank: dropped, as Ifpack2 local filter
                        B.col.push_back(cz * nb + (cf - f0));
                        B.val.push_back(A.val[p]);
                    }
                    // sort the row by column (ILUT wants no particular order, but keep it tidy)
                    B.ptr.push_back((int)B.col.size());
                }
            ilu[b].factor(B, droptol, lof);
            work[b].resize(B.n);
        }
    }
    void apply(const double* v, double* y) {
        const int nblocks = (int)ilu.size();
#pragma omp parallel for schedule(static, 1)
        for (int b = 0; b < nblocks; ++b) {
            const int f0 = fs[b], nb = fs[b + 1] - f0;
            double* w = work[b].data();
            for (int z = 0; z < L; ++z) std::memcpy(w + (size_t)z * nb, v + (size_t)z * T + f0, nb * sizeof(double));
            ilu[b].solve(w);
            for (int z = 0; z < L; ++z) std::memcpy(y + (size_t)z * T + f0, w + (size_t)z * nb, nb * sizeof(double));
        }
    }
};

void spmv(const Csr& A, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < A.n; ++i) {
        double s = 0.0;
        for (int p = A.ptr[i]; p < A.ptr[i + 1]; ++p) s += A.val[p] * x[A.col[p]];
        y[i] = s;
    }
}
double dot(int n, const double* a, const double* b) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

// Restarted right-preconditioned GMRES(m), x0 = 0, stop on the implicit residual relative to ||b||.
template <class Prec>
int gmres(const Csr& A, const double* b, double* x, Prec&& prec, double tol, int m, int maxit, double* achieved) {
    const int n = A.n;
    std::fill(x, x + n, 0.0);
    const double bn = std::sqrt(dot(n, b, b));
    *achieved = 0.0;
    if (bn == 0.0) return 0;
    std::vector<std::vector<double>> V(m + 1, std::vector<double>(n));
    std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), y(m), w(n), z(n), r(n);
    int its = 0;
    while (its < maxit) {
        spmv(A, x, r.data());
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) r[i] = b[i] - r[i];
        double beta = std::sqrt(dot(n, r.data(), r.data()));
        *achieved = beta / bn;
        if (beta / bn <= tol) break;
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) V[0][i] = r[i] / beta;
        int k_used = 0;
        for (int k = 0; k < m; ++k) {
            prec(V[k].data(), z.data());
            spmv(A, z.data(), w.data());
            for (int i = 0; i <= k; ++i) H[(size_t)i * m + k] = 0.0;
            for (int pass = 0; pass < 2; ++pass)  // ICGS, Belos' default orthogonalisation (2 passes)
                for (int i = 0; i <= k; ++i) {
                    double h = dot(n, w.data(), V[i].data());
                    H[(size_t)i * m + k] += h;
                    const double* vi = V[i].data();
#pragma omp parallel for schedule(static)
                    for (int q = 0; q < n; ++q) w[q] -= h * vi[q];
                }
            double hn = std::sqrt(dot(n, w.data(), w.data()));
            H[(size_t)(k + 1) * m + k] = hn;
            if (hn > 0) {
#pragma omp parallel for schedule(static)
                for (int q = 0; q < n; ++q) V[k + 1][q] = w[q] / hn;
            }
            for (int i = 0; i < k; ++i) {
                double t = cs[i] * H[(size_t)i * m + k] + sn[i] * H[(size_t)(i + 1) * m + k];
                H[(size_t)(i + 1) * m + k] = -sn[i] * H[(size_t)i * m + k] + cs[i] * H[(size_t)(i + 1) * m + k];
                H[(size_t)i * m + k] = t;
            }
            double den = std::hypot(H[(size_t)k * m + k], hn);
            cs[k] = H[(size_t)k * m + k] / den; sn[k] = hn / den;
            H[(size_t)k * m + k] = den; H[(size_t)(k + 1) * m + k] = 0.0;
            g[k + 1] = -sn[k] * g[k];
            g[k] = cs[k] * g[k];
            ++its; k_used = k + 1;
            *achieved = std::fabs(g[k + 1]) / bn;
            if (*achieved <= tol || its >= maxit) break;
        }
        for (int i = k_used - 1; i >= 0; --i) {
            double s = g[i];
            for (int j = i + 1; j < k_used; ++j) s -= H[(size_t)i * m + j] * y[j];
            y[i] = s / H[(size_t)i * m + i];
        }
        std::fill(w.begin(), w.end(), 0.0);
        for (int j = 0; j < k_used; ++j) {
            const double* vj = V[j].data();
            const double yj = y[j];
#pragma omp parallel for schedule(static)
            for (int q = 0; q < n; ++q) w[q] += yj * vj[q];
        }
        prec(w.data(), z.data());
#pragma omp parallel for schedule(static)
        for (int q = 0; q < n; ++q) x[q] += z[q];
        if (*achieved <= tol) {  // confirm with the true residual on the next pass of the while loop
            spmv(A, x, r.data());
            double rr = 0.0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
            for (int i = 0; i < n; ++i) { double d = b[i] - r[i]; rr += d * d; }
            *achieved = std::sqrt(rr) / bn;
            if (*achieved <= tol * 1.0000001) break;
        }
    }
    return its;
}

}  // namespace

extern "C" {

struct RefConfig {
    int nLayer, do_fixed_settling, do_sublimation, do_lateral_diff, rouault, enable_veg, use_exp_fetch, use_tanh_fetch,
        use_R94_lambda, n_threads;
    double settling_velocity, smooth_coeff, min_sd_trans, cutoff, snow_diffusion_const, tolerance, ilut_drop, ilut_fill;
    int gmres_restart, max_iterations;
};

struct RefMesh {  // global (single-process) mesh with precomputed face geometry
    int T;
    const int* neigh;  // [T][3], -1 none
    const double *nx, *ny, *elen, *dx;  // [3][T]
    const double *area, *zc;            // [T]
    const double *canopy, *lai, *stalk_n, *stalk_dv;  // [T] or null
    const unsigned char* water;                       // [T] or null
};

struct RefForcing { const double *U_R, *u2, *sd, *swe, *t, *rh, *vw_dir, *fetch; };

struct RefOut {
    double *c, *Qsusp, *Qsalt, *Qsubl;                       // [L][T], [T]...
    double *sum_drift, *sum_subl, *drift_mass, *more_avail;  // state, in/out [T]
    double *diag, *lat, *below, *above, *rhs0, *u_z, *csubl, *c_salt;  // optional ELL dump (null to skip)
    unsigned char* salt;                                                // optional
    int susp_present, dep_present, susp_iters, dep_iters, n_threads, pad;
    double susp_resid, dep_resid, s_assembly, s_factor, s_solve, s_deposition, s_total;
};

int pbsm3d_ref_threads(void) { return omp_get_max_threads(); }

int pbsm3d_ref_step(const RefConfig* c, const RefMesh* m, const RefForcing* f, double dt, RefOut* o) {
    const int T = m->T, L = c->nLayer;
    if (c->n_threads > 0) omp_set_num_threads(c->n_threads);
    const int nth = omp_get_max_threads();
    o->n_threads = nth;
    const double dz = 5.0 / (double)L, l_max = 40.0;
    const double t_start = now();
    // ---- CSR pattern in the reference ordering: self, lateral neighbours in neighbor(0..2) order, below, above
    const int N = T * L;
    Csr A;
    A.n = N;
    A.ptr.resize(N + 1);
    A.ptr[0] = 0;
    for (int z = 0; z < L; ++z)
        for (int i = 0; i < T; ++i) {
            int cnt = 1;
            for (int j = 0; j < 3; ++j) cnt += m->neigh[i * 3 + j] >= 0;
            cnt += (z > 0) + (z < L - 1);
            A.ptr[z * T + i + 1] = cnt;
        }
    for (int r = 0; r < N; ++r) A.ptr[r + 1] += A.ptr[r];
    A.col.resize(A.ptr[N]);
    A.val.assign(A.ptr[N], 0.0);
    std::vector<double> rhs(N, 0.0), u_z(N), csubl_v(N), Qsalt(T), c_salt_v(T), hs_v(T);
    std::vector<unsigned char> salt_v(T);

    // ---- HOT LOOP 1 (PBSM3D.cpp:417-1410)
#pragma omp parallel for schedule(static)
    for (int i = 0; i < T; ++i) {
        double fetch = 1000.0;
        if ((c->use_exp_fetch || c->use_tanh_fetch) && f->fetch) fetch = f->fetch[i];
        const double uref = f->U_R[i];
        double sd = f->sd[i]; sd = chm_is_nan(sd) ? 0.0 : sd;
        const double u2 = f->u2[i];
        double swe = f->swe[i]; swe = chm_is_nan(swe) ? 0.0 : swe;
        const double Tc = f->t[i];
        double hd = 0.0, LAI = 0.0, Ns = 0.0, dv = 0.0;
        if (c->enable_veg) {
            hd = std::max(0.0, m->canopy[i] - sd);
            if (c->use_R94_lambda) LAI = m->lai[i];
            else { Ns = m->stalk_n ? m->stalk_n[i] : 1.0; dv = m->stalk_dv ? m->stalk_dv[i] : 0.8; }
        }
        const bool water = m->water && m->water[i];
        const double ust_th = 0.35 + (1.0 / 150.0) * Tc + (1.0 / 8200.0) * Tc * Tc;
        bool salt = false;
        double lambda = 0.0, ustar = 1.3;
        if (hd <= c->cutoff && sd >= c->min_sd_trans && !water) {
            lambda = c->use_R94_lambda ? 0.5 * LAI * hd : Ns * dv * hd;
            ustar = u2 * kKappa / std::log(2.0 / 0.0002);
            if (ustar >= ust_th) salt = true;
        }
        double z0 = kZ0;
        if (!salt) ustar = std::max(0.01, kKappa * uref / std::log(kZUR / z0));
        z0 = std::max(kZ0, z0);
        ustar = std::max(0.01, ustar);
        const double hs = salt ? 0.08436 * std::pow(ustar, 1.27) : 0.0;
        const double t = Tc + 273.15;
        double h = 450.0 - f->vw_dir[i];
        if (h > 360.0) h -= 360.0;
        const double th = h * M_PI / 180.0;
        const double vx = -std::cos(th), vy = -std::sin(th);
        double E[3], nxj[3], nyj[3];
        int nb[3];
        for (int j = 0; j < 3; ++j) {
            E[j] = m->elen[(size_t)j * T + i]; nxj[j] = m->nx[(size_t)j * T + i]; nyj[j] = m->ny[(size_t)j * T + i];
            nb[j] = m->neigh[i * 3 + j];
        }
        const double area = m->area[i];
        double Qs = 0.0, c_salt = 0.0;
        if (salt) {
            const double R0 = 6356766.0, gE = 9.80665, Rd = 287.058;
            const double zc = m->zc[i];
            const double p = 101325.0 * std::pow(1.0 - ((0.0065 * R0 * zc) / (288.15 * (R0 + zc))), gE / (0.0065 * Rd));
            const double rho_f = p / (Rd * t);
            const double mB = 0.16 * 202.0;
            const double tau = (mB * lambda) / (1.0 + mB * lambda);
            c_salt = rho_f / (3.29 * ustar) * (1.0 - tau - (ust_th * ust_th) / (ustar * ustar));
            if (c_salt < 0 || std::isnan(c_salt)) { c_salt = 0; salt = false; }
            if (c->use_exp_fetch && fetch < 500.0) c_salt *= 1.0 - std::exp(-3.0 * fetch / 500.0);
            else if (c->use_tanh_fetch && fetch <= 300.0) c_salt *= 0.5 * std::tanh(0.1333333333e-1 * 300.0 - 2.0) + 0.5;
            const double uhs = 2.8 * ust_th;
            Qs = c_salt * uhs * hs;
            double mass = 0.0;
            for (int j = 0; j < 3; ++j) mass += -E[j] * Qs * (vx * nxj[j] + vy * nyj[j]);
            mass = mass / area * dt;
            if (mass < 0 && std::fabs(mass) > swe) { c_salt = 0; Qs = 0; }
        }
        Qsalt[i] = Qs; c_salt_v[i] = c_salt; salt_v[i] = salt; hs_v[i] = hs;
        const double rh = f->rh[i] / 100.0;
        const double TA = t - 273.15;
        const double es = 611.21 * std::exp((17.502 * TA) / (240.97 + TA));
        const double nrm = std::sqrt(vx * vx + vy * vy);
        for (int z = 0; z < L; ++z) {
            const int r = z * T + i;
            const double cz = z * dz + hs + dz / 2.0, hz = cz + sd;
            double uz;
            if (salt && cz < hd) uz = 2.8 * ust_th;
            else if (cz < hd) uz = 0.01;
            else if (hz < kZUR) uz = std::max(0.01, uref * std::log((hz - (sd + z0)) / z0) / std::log((kZUR - (sd + z0)) / z0));
            else uz = std::max(0.01, uref);
            u_z[r] = uz;
            const double rm = 4.6e-5 * std::pow(cz, -0.258);
            const double ma = 4.08 + 12.6 * cz;
            const double mm = 4.0 / 3.0 * M_PI * kRhoIce * rm * rm * rm * (1.0 + 3.0 / ma + 2.0 / (ma * ma));
            const double r_z = std::pow((3.0 * mm) / (4 * M_PI * kRhoIce), 0.3333333);
            const double xrz = 0.005 * std::pow(uz, 1.36);
            const double omega = c->do_fixed_settling ? c->settling_velocity : 1.1e7 * std::pow(r_z, 1.8);
            const double Vr = omega + 3.0 * xrz * std::cos(M_PI / 4.0);
            const double Re = 2.0 * r_z * Vr / 1.88e-5;
            const double Nu = 1.79 + 0.606 * std::pow(Re, 0.5), Sh = Nu;
            const double D = 2.06e-5 * std::pow(t / 273.15, 1.75);
            const double lam_t = 0.000063 * t + 0.00673;
            const double Ls = 2.838e6, Mw = 18.01, Rg = 8313.0;
            const double sigma = (rh - 1.0) * (1.019 + 0.027 * std::log(cz));
            const double rho = (Mw * es) / (Rg * t);
            const double Qr = 0.9 * M_PI * rm * rm * 120.0;
            const double dmdt = Sh * rho * D * (6.283185308 * Nu * Rg * r_z * sigma * t * t * lam_t - Ls * Mw * Qr + Qr * Rg * t) /
                                (D * Ls * Sh * (Ls * Mw - Rg * t) * rho + lam_t * t * t * Nu * Rg);
            double csubl = c->do_sublimation ? dmdt / mm : 0.0;
            csubl_v[r] = csubl;
            double Aj[3], al[3];
            for (int j = 0; j < 3; ++j) { Aj[j] = E[j] * dz; al[j] = c->do_lateral_diff ? Aj[j] * 0.00001 : 0.0; }
            const double lmix = kKappa * (cz + z0) * l_max / (kKappa * (cz + z0) + l_max);
            double dcoef = c->snow_diffusion_const;
            if (c->rouault) dcoef = 1.0 / (1.0 + (1.0 * omega * omega) / (1.56 * ustar * ustar));
            const double K = dcoef * ustar * lmix;
            const double a3 = area * K / dz, a4 = area * K / dz;
            const double sc = uz / nrm, ux = vx * sc, uy = vy * sc;
            const double ud3 = -omega, ud4 = omega;
            const double Vc = (area * dz / 5.0) * csubl;
            // row entries in pattern order
            int p = A.ptr[r];
            const int pd = p++;
            A.col[pd] = r;
            double d = 0.0;
            for (int j = 0; j < 3; ++j) {
                const double ud = ux * nxj[j] + uy * nyj[j];
                if (ud > 0) {
                    if (nb[j] >= 0) { d += Vc - Aj[j] * ud - al[j]; A.col[p] = z * T + nb[j]; A.val[p++] = al[j]; }
                    else d += -0.1e-1 * al[j] - 1.0 * Aj[j] * ud + Vc;
                } else {
                    if (nb[j] >= 0) { d += Vc - al[j]; A.col[p] = z * T + nb[j]; A.val[p++] = -Aj[j] * ud + al[j]; }
                    else d += -0.1e-1 * al[j] - 0.99 * Aj[j] * ud + Vc;
                }
            }
            double lo = 0, up = 0;
            if (z == 0) {
                const double a4p = area * K / (hs / 2.0 + dz / 2.0);
                d += Vc - area * ud4 - a4p;
                rhs[r] = -a4p * c_salt;
                if (ud3 > 0) { d += Vc - area * ud3 - a3; up = a3; } else { d += Vc - a3; up = -area * ud3 + a3; }
            } else if (z == L - 1) {
                if (ud3 > 0) d += Vc - area * ud3 - a3; else d += Vc - a3;
                if (ud4 > 0) { d += Vc - area * ud4 - a4; lo = a4; } else { d += Vc - a4; lo = -area * ud4 + a4; }
            } else {
                if (ud3 > 0) { d += Vc - area * ud3 - a3; up = a3; } else { d += Vc - a3; up = -area * ud3 + a3; }
                if (ud4 > 0) { d += Vc - area * ud4 - a4; lo = a4; } else { d += Vc - a4; lo = -area * ud4 + a4; }
            }
            if (z > 0) { A.col[p] = r - T; A.val[p++] = lo; }
            if (z < L - 1) { A.col[p] = r + T; A.val[p++] = up; }
            A.val[pd] = d;
            if (o->diag) {
                o->diag[r] = d; o->below[r] = lo; o->above[r] = up;
                int q = pd + 1;
                for (int j = 0; j < 3; ++j) o->lat[((size_t)j * L + z) * T + i] = (nb[j] >= 0) ? A.val[q++] : 0.0;
            }
        }
    }
    if (o->rhs0) std::memcpy(o->rhs0, rhs.data(), T * sizeof(double));
    if (o->u_z) std::memcpy(o->u_z, u_z.data(), N * sizeof(double));
    if (o->csubl) std::memcpy(o->csubl, csubl_v.data(), N * sizeof(double));
    if (o->c_salt) std::memcpy(o->c_salt, c_salt_v.data(), T * sizeof(double));
    if (o->salt) std::memcpy(o->salt, salt_v.data(), T);
    std::memcpy(o->Qsalt, Qsalt.data(), T * sizeof(double));
    const double t_asm = now();
    o->s_assembly = t_asm - t_start;

    // ---- C/D: suspension solve
    double rmax = 0.0;
    for (int i = 0; i < T; ++i) rmax = std::max(rmax, std::fabs(rhs[i]));
    o->susp_present = rmax > 1e-12;
    o->susp_iters = 0; o->susp_resid = 0; o->s_factor = 0; o->s_solve = 0;
    std::vector<double> x(N, 0.0);
    if (o->susp_present) {
        BlockPrec P;
        P.setup(A, T, L, nth, c->ilut_drop, c->ilut_fill);
        const double t_f = now();
        o->s_factor = t_f - t_asm;
        o->susp_iters = gmres(A, rhs.data(), x.data(), [&](const double* v, double* y) { P.apply(v, y); }, c->tolerance,
                              c->gmres_restart, c->max_iterations, &o->susp_resid);
        o->s_solve = now() - t_f;
    }
    if (o->c) std::memcpy(o->c, x.data(), N * sizeof(double));
    const double t_sus = now();
    // ---- E: flux integration (PBSM3D.cpp:1467-1503)
    std::vector<double> Qsusp(T);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < T; ++i) {
        double qs = 0, ql = 0;
        for (int z = 0; z < L; ++z) {
            double cc = x[z * T + i];
            cc = (cc < 0 || chm_is_nan(cc)) ? 0.0 : cc;
            qs += cc * u_z[z * T + i] * dz;
            ql += csubl_v[z * T + i] * cc * dz;
        }
        Qsusp[i] = qs; o->Qsusp[i] = qs; o->Qsubl[i] = ql;
        o->sum_subl[i] += ql * dt;
    }
    // ---- G: deposition system (PBSM3D.cpp:1516-1658), H: solve, I: drift update
    Csr Dm;
    Dm.n = T;
    Dm.ptr.resize(T + 1);
    Dm.ptr[0] = 0;
    for (int i = 0; i < T; ++i) {
        int cnt = 1;
        for (int j = 0; j < 3; ++j) cnt += m->neigh[i * 3 + j] >= 0;
        Dm.ptr[i + 1] = Dm.ptr[i] + cnt;
    }
    Dm.col.resize(Dm.ptr[T]); Dm.val.resize(Dm.ptr[T]);
    std::vector<double> drhs(T), q(T, 0.0);
    const double eps = c->smooth_coeff;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < T; ++i) {
        double h = 450.0 - f->vw_dir[i];
        if (h > 360.0) h -= 360.0;
        const double th = h * M_PI / 180.0;
        const double vx = -std::cos(th), vy = -std::sin(th);
        int p = Dm.ptr[i];
        const int pd = p++;
        Dm.col[pd] = i;
        double d = m->area[i], rv = 0.0;
        for (int j = 0; j < 3; ++j) {
            const double ud = vx * m->nx[(size_t)j * T + i] + vy * m->ny[(size_t)j * T + i];
            const double Ej = m->elen[(size_t)j * T + i];
            const int n = m->neigh[i * 3 + j];
            double Qt = Qsusp[i], Qs = Qsalt[i];
            if (!(ud > 0) && n >= 0) { Qt = Qsusp[n]; Qs = Qsalt[n]; if (chm_is_nan(Qs)) Qs = 0; }
            if (n >= 0) {
                const double cf = eps * Ej / m->dx[(size_t)j * T + i];
                d += cf;
                Dm.col[p] = n; Dm.val[p++] = -cf;
            }
            rv += -Ej * (Qt + Qs) * ud;
        }
        Dm.val[pd] = d;
        drhs[i] = rv;
    }
    double dmax = 0.0;
    for (int i = 0; i < T; ++i) dmax = std::max(dmax, std::fabs(drhs[i]));
    o->dep_present = o->susp_present && dmax > 1e-12;
    o->dep_iters = 0; o->dep_resid = 0;
    if (o->dep_present) {
        BlockPrec P;
        P.setup(Dm, T, 1, nth, c->ilut_drop, c->ilut_fill);
        o->dep_iters = gmres(Dm, drhs.data(), q.data(), [&](const double* v, double* y) { P.apply(v, y); }, c->tolerance,
                             c->gmres_restart, c->max_iterations, &o->dep_resid);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < T; ++i) {
            double qd = chm_is_nan(q[i]) ? 0.0 : q[i];
            double mass = qd * dt;
            double swe = f->swe[i]; swe = chm_is_nan(swe) ? 0.0 : swe;
            if (mass < 0 && std::fabs(mass) > swe) { o->more_avail[i] = 1; mass = -swe; }
            if (mass < 0 && !salt_v[i]) mass = 0;
            o->drift_mass[i] = mass;
            o->sum_drift[i] += mass;
        }
    }
    o->s_deposition = now() - t_sus;
    o->s_total = now() - t_start;
    return 0;
}

}  // extern "C"

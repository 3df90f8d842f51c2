// sm_100a kernels of the PBSM3D hot path.  All arithmetic is fp64; the path is sparse and HBM-bound
// (≈0.15 flop/B), so there are no tensor-core instructions here: the rules that matter are coalesced
// layer-major SoA streams, one pass over each array, L2-resident neighbour gathers and grids that fill
// the 148 SMs.
//
// Data layout (all device arrays, T = owned faces of this rank, L = nLayer):
//   per face        a[i]                      i in [0,T)
//   per face-edge   a[j*T + i]                j in 0..2   (edge j is shared with neighbour j)
//   per row         a[z*T + i]                the reference's local unknown numbering (LinearAlgebra.cpp:81)
//   ghosts          xg[nl*gstart[g] + z*gcnt[g] + (g - gstart[g])]   g = neigh - T; ghosts are sorted by global id, so
//                   each owner's ghosts are one contiguous block [gstart, gstart+gcnt), received as [nl][gcnt]
#pragma once
#include <cuda_runtime.h>
#include "pbsm3d_physics.cuh"

namespace pbsm3d {

struct DevConfig {
    int L;
    int do_fixed_settling, do_sublimation, do_lateral_diff, rouault, enable_veg;
    int use_exp_fetch, use_tanh_fetch, use_R94_lambda;
    double settling_velocity, eps, min_sd_trans, cutoff, snow_diffusion_const;
    double dz;     // v_edge_height = susp_depth / nLayer (PBSM3D.cpp:225-226)
    double l_max;  // 40 (PBSM3D.cpp:227)
};

struct DevMesh {
    int T, n_ghost;
    const int* neigh;      // [3][T] local ids, -1 none, >=T ghost
    const double* nx;      // [3][T]
    const double* ny;      // [3][T]
    const double* elen;    // [3][T]
    const double* area;    // [T]
    const double* zc;      // [T] centroid elevation (face->get_z())
    const double* canopy;  // [T] or null
    const double* lai;     // [T] or null
    const double* stalk_n; // [T] or null
    const double* stalk_dv;
    const unsigned char* water;  // [T] or null
    const int* gstart;     // [n_ghost] first ghost of the owner block this ghost belongs to
    const int* gcnt;       // [n_ghost] size of that block
};

struct DevForcing {
    const double *U_R, *u2, *sd, *swe, *t, *rh, *vw_dir, *fetch;
};

struct SuspSystem {
    double *diag, *below, *above;  // [L][T]
    double* lat;                   // [3][L][T]
    double *cp, *inv;              // [L][T] Thomas factors of the column blocks
    double* rhs0;                  // [T]
    double *u_z, *csubl;           // [L][T]
    double *Qsalt, *c_salt;        // [T]
    unsigned char* salt;           // [T]
};

// ---------------------------------------------------------------------------------------------- setup
// Face geometry from the three vertices of each face (reference: mesh/triangulation.hpp:1443-1475
// edge_unit_normal/edge, :1491-1498 edge_length, :1577-1589 center, :1830-1856 get_area).
// Runs once in pbsm3d_create.  Products and sums use the __d*_rn intrinsics, which nvcc never contracts
// into FMAs, so every value is bit-identical to the plain IEEE fp64 evaluation the reference (and numpy) do.
__global__ void geometry_kernel(int T, int Tall, const double* __restrict__ verts, const double* __restrict__ area_param,
                                double* __restrict__ nx, double* __restrict__ ny, double* __restrict__ elen,
                                double* __restrict__ area, double* __restrict__ cx, double* __restrict__ cy,
                                double* __restrict__ cz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Tall) return;
    double px[3], py[3], pz[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        px[k] = verts[(size_t)i * 9 + k * 3 + 0];
        py[k] = verts[(size_t)i * 9 + k * 3 + 1];
        pz[k] = verts[(size_t)i * 9 + k * 3 + 2];
    }
    cx[i] = __ddiv_rn(__dadd_rn(__dadd_rn(px[0], px[1]), px[2]), 3.0);
    cy[i] = __ddiv_rn(__dadd_rn(__dadd_rn(py[0], py[1]), py[2]), 3.0);
    cz[i] = __ddiv_rn(__dadd_rn(__dadd_rn(pz[0], pz[1]), pz[2]), 3.0);
    if (i >= T) return;  // ghosts only need a centroid
    double ex[3], ey[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {  // edge(k) = v[cw(k)] - v[ccw(k)]
        int a = (k + 1) % 3, b = (k + 2) % 3;
        ex[k] = __dsub_rn(px[b], px[a]);
        ey[k] = __dsub_rn(py[b], py[a]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int k1 = (k + 1) % 3;
        double n_x = ey[k], n_y = -ex[k];
        double D = __dadd_rn(__dmul_rn(ex[k1], n_x), __dmul_rn(ey[k1], n_y));
        if (D > 0) { n_x = -n_x; n_y = -n_y; }
        double nrm = __dsqrt_rn(__dadd_rn(__dmul_rn(n_x, n_x), __dmul_rn(n_y, n_y)));
        nx[(size_t)k * T + i] = __ddiv_rn(n_x, nrm);
        ny[(size_t)k * T + i] = __ddiv_rn(n_y, nrm);
        elen[(size_t)k * T + i] = __dsqrt_rn(__dadd_rn(__dmul_rn(ex[k], ex[k]), __dmul_rn(ey[k], ey[k])));
    }
    if (area_param) {
        area[i] = area_param[i];
    } else {
        double v1x = __dsub_rn(px[1], px[0]), v1y = __dsub_rn(py[1], py[0]);
        double v2x = __dsub_rn(px[2], px[0]), v2y = __dsub_rn(py[2], py[0]);
        area[i] = __ddiv_rn(__dsub_rn(__dmul_rn(v1x, v2y), __dmul_rn(v1y, v2x)), 2.0);
    }
}

// Static part of the deposition system (reference re-derives it every step, PBSM3D.cpp:1546,1609-1628):
// diag = area + sum eps*E_j/dx_j, off_j = -eps*E_j/dx_j, dx_j = 2-D centroid distance (coordinates.cpp:100-106).
// cx/cy are [T + n_ghost], so a neighbour that is a ghost face resolves like any other.
__global__ void deposition_matrix_kernel(int T, double eps, const int* __restrict__ neigh, const double* __restrict__ elen,
                                         const double* __restrict__ area, const double* __restrict__ cx,
                                         const double* __restrict__ cy, double* __restrict__ dx, double* __restrict__ ddiag,
                                         double* __restrict__ doff, double* __restrict__ dinv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    double d = area[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        int n = neigh[(size_t)j * T + i];
        double dist = 2.0, c = 0.0;  // dx[] default 2.0, PBSM3D.cpp:1534
        if (n >= 0) {
            double ddx = __dsub_rn(cx[i], cx[n]), ddy = __dsub_rn(cy[i], cy[n]);
            dist = __dsqrt_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)));
            c = __ddiv_rn(__dmul_rn(eps, elen[(size_t)j * T + i]), dist);
        }
        dx[(size_t)j * T + i] = dist;
        d = __dadd_rn(d, c);
        doff[(size_t)j * T + i] = -c;
    }
    ddiag[i] = d;
    dinv[i] = 1.0 / d;
}

// ------------------------------------------------------------------------------------------- assembly
// HOT LOOP 1: saltation + every layer of one face column (reference PBSM3D.cpp:436-1406), fused with the
// forward elimination of that column's tridiagonal block (the preconditioner / line solver factor).
// One thread per face; for each layer the 32 lanes of a warp write 32 consecutive doubles of every
// output stream.  Replaces ≈11 Tpetra sumIntoGlobalValues hash lookups per row by direct ELL stores.
__global__ void __launch_bounds__(128) assemble_kernel(DevConfig c, DevMesh m, DevForcing f, SuspSystem s, double dt) {
    const int T = m.T;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;

    double fetch = 1000.0;
    if ((c.use_exp_fetch || c.use_tanh_fetch) && f.fetch) fetch = f.fetch[i];
    const double uref = f.U_R[i];
    double sd = f.sd[i];
    sd = chm_is_nan(sd) ? 0.0 : sd;
    const double u2 = f.u2[i];
    double swe = f.swe[i];
    swe = chm_is_nan(swe) ? 0.0 : swe;
    const double Tc = f.t[i];
    const double phi = f.vw_dir[i];
    const double area = m.area[i];
    double nxj[3], nyj[3], Ej[3];
    int nb[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        nxj[j] = m.nx[(size_t)j * T + i];
        nyj[j] = m.ny[(size_t)j * T + i];
        Ej[j] = m.elen[(size_t)j * T + i];
        nb[j] = m.neigh[(size_t)j * T + i];
    }

    double height_diff = 0.0, LAI = 0.0, Nst = 0.0, dv = 0.0;
    if (c.enable_veg) {
        height_diff = fmax(0.0, m.canopy[i] - sd);
        if (c.use_R94_lambda) LAI = m.lai[i];
        else { Nst = m.stalk_n ? m.stalk_n[i] : 1.0; dv = m.stalk_dv ? m.stalk_dv[i] : 0.8; }
    }
    const bool water = m.water ? (m.water[i] != 0) : false;
    const double ust_th = 0.35 + (1.0 / 150.0) * Tc + (1.0 / 8200.0) * Tc * Tc;

    bool salt = false;
    double lambda = 0.0, ustar = 1.3;
    if (height_diff <= c.cutoff && sd >= c.min_sd_trans && !water) {
        lambda = c.use_R94_lambda ? 0.5 * LAI * height_diff : Nst * dv * height_diff;
        ustar = u2 * kKappa / log(2.0 / 0.0002);
        if (ustar >= ust_th) salt = true;
    }
    double z0 = kZ0Snow;
    if (!salt) ustar = fmax(0.01, kKappa * uref / log(kZUR / z0));
    z0 = fmax(kZ0Snow, z0);
    ustar = fmax(0.01, ustar);
    const double hs = salt ? 0.08436 * pow(ustar, 1.27) : 0.0;

    const double t = Tc + 273.15;
    double vx, vy;
    wind_unit_vector(phi, vx, vy);
    double Qsalt = 0.0, c_salt = 0.0;
    if (salt) {
        const double rho_f = std_dry_air_density(m.zc[i], t);
        const double mB = 0.16 * 202.0;
        const double tau_n_ratio = (mB * lambda) / (1.0 + mB * lambda);
        c_salt = rho_f / (3.29 * ustar) * (1.0 - tau_n_ratio - (ust_th * ust_th) / (ustar * ustar));
        if (c_salt < 0 || isnan(c_salt)) { c_salt = 0.0; salt = false; }
        if (c.use_exp_fetch && fetch < 500.0) {
            c_salt *= 1.0 - exp(-3.0 * fetch / 500.0);
        } else if (c.use_tanh_fetch && fetch <= 300.0) {
            const double Lc = 0.5 * tanh(0.1333333333e-1 * 300.0 - 2.0) + 0.5;  // fetch_ref inside tanh, as the reference
            c_salt *= Lc;
        }
        const double uhs = 2.8 * ust_th;
        Qsalt = c_salt * uhs * hs;
        double mass = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double udotm = vx * nxj[j] + vy * nyj[j];
            mass += -Ej[j] * Qsalt * udotm;
        }
        mass = mass / area * dt;
        if (mass < 0 && fabs(mass) > swe) { c_salt = 0.0; Qsalt = c_salt * uhs * hs; }  // saltation flag survives
    }
    s.Qsalt[i] = Qsalt;
    s.c_salt[i] = c_salt;
    s.salt[i] = salt ? 1 : 0;

    const double rh = f.rh[i] / 100.0;
    const double es = saturated_vapour_pressure(t);
    const double dz = c.dz;
    const double nrm = sqrt(vx * vx + vy * vy);
    // layer-independent pieces of the sublimation model (same expressions as inside the reference's z loop)
    const double D = 2.06e-5 * pow(t / 273.15, 1.75);
    const double lambda_t = 0.000063 * t + 0.00673;
    const double Ls = 2.838e6, Mw = 18.01, Rg = 8313.0;
    const double rho_sat = (Mw * es) / (Rg * t);
    const double ulog_den = log((kZUR - (sd + z0)) / z0);
    double Aj[3], alphaj[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        Aj[j] = Ej[j] * dz;
        alphaj[j] = c.do_lateral_diff ? Aj[j] * 0.00001 : 0.0;
    }

    double cp_prev = 0.0;
    const int L = c.L;
    for (int z = 0; z < L; ++z) {
        const size_t r = (size_t)z * T + i;
        const double cz = z * dz + hs + dz / 2.0;
        const double hz = cz + sd;
        double u_z;
        if (salt && cz < height_diff) u_z = 2.8 * ust_th;
        else if (cz < height_diff) u_z = 0.01;
        else if (hz < kZUR) u_z = fmax(0.01, uref * log((hz - (sd + z0)) / z0) / ulog_den);
        else u_z = fmax(0.01, uref);
        s.u_z[r] = u_z;

        const double rm = 4.6e-5 * pow(cz, -0.258);
        const double mm_alpha = 4.08 + 12.6 * cz;
        const double mm = 4.0 / 3.0 * kPi * kRhoIce * rm * rm * rm * (1.0 + 3.0 / mm_alpha + 2.0 / (mm_alpha * mm_alpha));
        const double r_z = pow((3.0 * mm) / (4 * kPi * kRhoIce), 0.3333333);
        const double xrz = 0.005 * pow(u_z, 1.36);
        const double omega = c.do_fixed_settling ? c.settling_velocity : 1.1e7 * pow(r_z, 1.8);
        const double Vr = omega + 3.0 * xrz * cos(kPi / 4.0);
        const double Re = 2.0 * r_z * Vr / 1.88e-5;
        const double Nu = 1.79 + 0.606 * sqrt(Re);
        const double Sh = Nu;
        const double sigma = (rh - 1.0) * (1.019 + 0.027 * log(cz));
        const double Qr = 0.9 * kPi * rm * rm * 120.0;
        const double dmdtz = Sh * rho_sat * D * (6.283185308 * Nu * Rg * r_z * sigma * t * t * lambda_t - Ls * Mw * Qr + Qr * Rg * t) /
                             (D * Ls * Sh * (Ls * Mw - Rg * t) * rho_sat + lambda_t * t * t * Nu * Rg);
        double csubl = dmdtz / mm;
        if (!c.do_sublimation) csubl = 0.0;
        s.csubl[r] = csubl;

        const double lmix = kKappa * (cz + z0) * c.l_max / (kKappa * (cz + z0) + c.l_max);
        const double w = omega;
        double diffusion_coeff = c.snow_diffusion_const;
        if (c.rouault) diffusion_coeff = 1.0 / (1.0 + (1.0 * w * w) / (1.56 * ustar * ustar));
        const double K = diffusion_coeff * ustar * lmix;
        const double alpha3 = area * K / dz;
        const double alpha4 = area * K / dz;
        const double sc = u_z / nrm;
        const double ux = vx * sc, uy = vy * sc;
        const double udotm3 = -w, udotm4 = w;
        const double Vc = (area * dz / 5.0) * csubl;

        double d = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double udotm = ux * nxj[j] + uy * nyj[j];
            double off = 0.0;
            if (udotm > 0) {
                if (nb[j] >= 0) { d += Vc - Aj[j] * udotm - alphaj[j]; off = alphaj[j]; }
                else d += -0.1e-1 * alphaj[j] - 1.0 * Aj[j] * udotm + Vc;
            } else {
                if (nb[j] >= 0) { d += Vc - alphaj[j]; off = -Aj[j] * udotm + alphaj[j]; }
                else d += -0.1e-1 * alphaj[j] - 0.99 * Aj[j] * udotm + Vc;
            }
            s.lat[((size_t)j * L + z) * T + i] = off;
        }
        double lo = 0.0, up = 0.0;
        if (z == 0) {
            const double alpha4p = area * K / (hs / 2.0 + dz / 2.0);
            d += Vc - area * udotm4 - alpha4p;
            s.rhs0[i] = -alpha4p * c_salt;
            if (udotm3 > 0) { d += Vc - area * udotm3 - alpha3; up = alpha3; }
            else { d += Vc - alpha3; up = -area * udotm3 + alpha3; }
        } else if (z == L - 1) {
            if (udotm3 > 0) d += Vc - area * udotm3 - alpha3;
            else d += Vc - alpha3;
            if (udotm4 > 0) { d += Vc - area * udotm4 - alpha4; lo = alpha4; }
            else { d += Vc - alpha4; lo = -area * udotm4 + alpha4; }
        } else {
            if (udotm3 > 0) { d += Vc - area * udotm3 - alpha3; up = alpha3; }
            else { d += Vc - alpha3; up = -area * udotm3 + alpha3; }
            if (udotm4 > 0) { d += Vc - area * udotm4 - alpha4; lo = alpha4; }
            else { d += Vc - alpha4; lo = -area * udotm4 + alpha4; }
        }
        s.diag[r] = d;
        s.below[r] = lo;
        s.above[r] = up;
        // forward elimination of the column block (Thomas): den_z = d_z - lo_z * cp_{z-1}
        const double inv = 1.0 / (d - lo * cp_prev);
        cp_prev = up * inv;
        s.inv[r] = inv;
        s.cp[r] = cp_prev;
    }
}

// -------------------------------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Block-wide sum staged through shared memory; result valid in thread 0.
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sm[32];
    __syncthreads();  // protect sm reuse across consecutive calls
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v;
}
__device__ __forceinline__ double block_max(double v) {
    __shared__ double smx[32];
    __syncthreads();
    v = warp_max(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smx[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? smx[threadIdx.x] : 0.0;
    if (w == 0) v = warp_max(v);
    return v;
}

constexpr int kRedBlocks = 148 * 4;  // partial-sum slots: a multiple of the SM count
constexpr int kRedThreads = 256;

// ||v||_inf partials (NearestNeighborProblem::getRhsMax, LinearAlgebra.cpp:264-270)
__global__ void __launch_bounds__(kRedThreads) absmax_kernel(size_t n, const double* __restrict__ v, double* __restrict__ partial) {
    double m = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        m = fmax(m, fabs(v[k]));
    m = block_max(m);
    if (threadIdx.x == 0) partial[blockIdx.x] = m;
}

// sum v^2 partials (||b||_2^2 for the relative-residual stopping rule)
__global__ void __launch_bounds__(kRedThreads) sumsq_kernel(size_t n, const double* __restrict__ v, double* __restrict__ partial) {
    double a = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) a += v[k] * v[k];
    a = block_sum(a);
    if (threadIdx.x == 0) partial[blockIdx.x] = a;
}

// Scalar slots shared by the solver kernels (device resident; the host only reads `status` now and then).
struct Scalars {
    double rho, alpha, omega, beta;   // BiCGStab / CG recurrences
    double rr, bnorm2;                // ||r||^2, ||b||^2
    double tmp[4];
    int done;                         // 1 = converged, 2 = breakdown
    int iters;
};

// Final stage of every fused reduction: one block folds `nvals` interleaved partial arrays
// (partial[v*stride + b]) into out[v].  op 0 = sum, 1 = max.
__global__ void __launch_bounds__(256) fold_kernel(int nblocks, int nvals, int stride, const double* __restrict__ partial,
                                                   double* __restrict__ out, int op) {
    for (int v = 0; v < nvals; ++v) {
        double a = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
            double p = partial[(size_t)v * stride + b];
            a = op ? fmax(a, p) : a + p;
        }
        a = op ? block_max(a) : block_sum(a);
        if (threadIdx.x == 0) out[v] = a;
    }
}

// ------------------------------------------------------------------------------------ SpMV / residual
__device__ __forceinline__ double gather_x(const double* __restrict__ x, const double* __restrict__ xg, const DevMesh& m,
                                           int L, int n, int z) {
    if (n < m.T) return x[(size_t)z * m.T + n];
    const int g = n - m.T, gs = m.gstart[g];
    return xg[(size_t)L * gs + (size_t)z * m.gcnt[g] + (g - gs)];
}

// Row of A·x in the extruded-ELL layout: lateral gathers x[z*T + neigh_j], vertical x[(z±1)*T + i].
__device__ __forceinline__ double spmv_row(const SuspSystem& s, const DevMesh& m, int L, const double* __restrict__ x,
                                           const double* __restrict__ xg, int z, int i) {
    const int T = m.T;
    const size_t r = (size_t)z * T + i;
    double acc = s.diag[r] * x[r];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        int n = m.neigh[(size_t)j * T + i];
        if (n >= 0) acc += s.lat[((size_t)j * L + z) * T + i] * gather_x(x, xg, m, L, n, z);
    }
    if (z > 0) acc += s.below[r] * x[r - T];
    if (z < L - 1) acc += s.above[r] * x[r + T];
    return acc;
}

// mode 0: y = A x.   mode 1: y = b - A x (b is non-zero only in layer 0).  Optionally accumulates up to two
// dot products with the freshly produced y: partial[b] = <y, d0>, partial[stride+b] = <y, d1 or y>.
// Grid-stride over rows with a grid that is a multiple of the SM count; partials make the sums deterministic.
template <int MODE>
__global__ void __launch_bounds__(256) spmv_kernel(SuspSystem s, DevMesh m, int L, const double* __restrict__ x,
                                                   const double* __restrict__ xg, double* __restrict__ y,
                                                   const double* __restrict__ d0, const double* __restrict__ d1,
                                                   int self_dot, double* __restrict__ partial, int stride,
                                                   const int* __restrict__ done) {
    if (done && *done) return;
    const int T = m.T;
    const size_t N = (size_t)L * T;
    double a0 = 0.0, a1 = 0.0;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (size_t)gridDim.x * blockDim.x) {
        int z = (int)(r / T), i = (int)(r - (size_t)z * T);
        double v = spmv_row(s, m, L, x, xg, z, i);
        if (MODE == 1) v = ((z == 0) ? s.rhs0[i] : 0.0) - v;
        if (y) y[r] = v;
        if (d0) a0 += v * d0[r];
        if (d1) a1 += v * d1[r];
        else if (self_dot) a1 += v * v;
    }
    if (partial) {
        if (d0) { a0 = block_sum(a0); if (threadIdx.x == 0) partial[blockIdx.x] = a0; }
        if (d1 || self_dot) { a1 = block_sum(a1); if (threadIdx.x == 0) partial[stride + blockIdx.x] = a1; }
    }
}

// ---------------------------------------------------------------------------------- line relaxation
// One sweep of the stationary column-block-Jacobi iteration  x_new = T^{-1} (b - A_lat x_old):
// T = the vertical tridiagonal blocks (factored in assemble_kernel), A_lat = the three lateral couplings.
// One thread per face column; every stream is read exactly once, coalesced; x_old gathers hit L2.
// LT > 0: compile-time layer count (the forward-substitution column stays in registers);
// LT == 0: any L, the column is staged through x_new.
template <int LT>
__global__ void __launch_bounds__(128) line_sweep_kernel(SuspSystem s, DevMesh m, int Lrt, const double* __restrict__ x_old,
                                                         const double* __restrict__ xg_old, double* __restrict__ x_new,
                                                         const int* __restrict__ done) {
    if (done && *done) return;
    const int T = m.T;
    const int L = LT > 0 ? LT : Lrt;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    int nb[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) nb[j] = m.neigh[(size_t)j * T + i];
    double dp[LT > 0 ? LT : 1];
    double prev = 0.0;
#pragma unroll
    for (int z = 0; z < L; ++z) {
        const size_t r = (size_t)z * T + i;
        double g = (z == 0) ? s.rhs0[i] : 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (nb[j] >= 0) g -= s.lat[((size_t)j * L + z) * T + i] * gather_x(x_old, xg_old, m, L, nb[j], z);
        prev = (g - s.below[r] * prev) * s.inv[r];
        if (LT > 0) dp[z] = prev;
        else x_new[r] = prev;
    }
    double xn = prev;
    x_new[(size_t)(L - 1) * T + i] = xn;
#pragma unroll
    for (int z = L - 2; z >= 0; --z) {
        const size_t r = (size_t)z * T + i;
        const double d = (LT > 0) ? dp[z] : x_new[r];
        xn = d - s.cp[r] * xn;
        x_new[r] = xn;
    }
}

// Column-tridiagonal preconditioner apply y = T^{-1} v (right preconditioner of the Krylov path).
__global__ void __launch_bounds__(128) thomas_kernel(SuspSystem s, int T, int L, const double* __restrict__ v,
                                                     double* __restrict__ y, const int* __restrict__ done) {
    if (done && *done) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    double prev = 0.0;
    for (int z = 0; z < L; ++z) {
        const size_t r = (size_t)z * T + i;
        prev = (v[r] - s.below[r] * prev) * s.inv[r];
        y[r] = prev;
    }
    double xn = prev;
    for (int z = L - 2; z >= 0; --z) {
        const size_t r = (size_t)z * T + i;
        xn = y[r] - s.cp[r] * xn;
        y[r] = xn;
    }
}

// ------------------------------------------------------------------------------------------ BiCGStab
// Right-preconditioned BiCGStab, all recurrence scalars on the device (no host round trip per iteration).
// init: r = b (x0 = 0), rhat = r, p = v = 0, rho = alpha = omega = 1; partial <- ||b||^2.
__global__ void __launch_bounds__(256) bicg_init_kernel(int T, int L, const double* __restrict__ rhs0, double* __restrict__ x,
                                                        double* __restrict__ r, double* __restrict__ rhat, double* __restrict__ p,
                                                        double* __restrict__ v, double* __restrict__ partial) {
    const size_t N = (size_t)L * T;
    double a = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (size_t)gridDim.x * blockDim.x) {
        double b = (k < (size_t)T) ? rhs0[k] : 0.0;
        x[k] = 0.0; r[k] = b; rhat[k] = b; p[k] = 0.0; v[k] = 0.0;
        a += b * b;
    }
    a = block_sum(a);
    if (threadIdx.x == 0) partial[blockIdx.x] = a;
}
// p = r + beta (p - omega v)
__global__ void __launch_bounds__(256) bicg_p_kernel(size_t N, const Scalars* __restrict__ sc, const double* __restrict__ r,
                                                     const double* __restrict__ v, double* __restrict__ p) {
    if (sc->done) return;
    const double beta = sc->beta, omega = sc->omega;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (size_t)gridDim.x * blockDim.x)
        p[k] = r[k] + beta * (p[k] - omega * v[k]);
}
// s = r - alpha v (in place into r); partial <- ||s||^2
__global__ void __launch_bounds__(256) bicg_s_kernel(size_t N, const Scalars* __restrict__ sc, double* __restrict__ r,
                                                     const double* __restrict__ v, double* __restrict__ partial) {
    if (sc->done) return;
    const double alpha = sc->alpha;
    double a = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (size_t)gridDim.x * blockDim.x) {
        double sv = r[k] - alpha * v[k];
        r[k] = sv;
        a += sv * sv;
    }
    a = block_sum(a);
    if (threadIdx.x == 0) partial[blockIdx.x] = a;
}
// x += alpha ph + omega sh ; r = s - omega t ; partials <- <rhat,r>, <r,r>
__global__ void __launch_bounds__(256) bicg_xr_kernel(size_t N, const Scalars* __restrict__ sc, double* __restrict__ x,
                                                      double* __restrict__ r, const double* __restrict__ ph,
                                                      const double* __restrict__ sh, const double* __restrict__ tt,
                                                      const double* __restrict__ rhat, double* __restrict__ partial, int stride) {
    if (sc->done) return;
    const double alpha = sc->alpha, omega = sc->omega;
    double a0 = 0.0, a1 = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (size_t)gridDim.x * blockDim.x) {
        x[k] += alpha * ph[k] + omega * sh[k];
        double rv = r[k] - omega * tt[k];
        r[k] = rv;
        a0 += rhat[k] * rv;
        a1 += rv * rv;
    }
    a0 = block_sum(a0);
    if (threadIdx.x == 0) partial[blockIdx.x] = a0;
    a1 = block_sum(a1);
    if (threadIdx.x == 0) partial[stride + blockIdx.x] = a1;
}
// x += alpha ph (early exit of the half step)
__global__ void __launch_bounds__(256) axpy_scalar_kernel(size_t N, const double* __restrict__ a, const double* __restrict__ p,
                                                          double* __restrict__ x) {
    const double al = *a;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (size_t)gridDim.x * blockDim.x) x[k] += al * p[k];
}

// Scalar updates (one thread).  `red` holds the globally reduced dot products of the preceding kernel.
// stage 0: after init            red[0] = ||b||^2
// stage 1: after v = A ph        red[0] = <rhat, v>           -> alpha
// stage 2: after s               red[0] = ||s||^2             -> half-step convergence
// stage 3: after t = A sh        red[0] = <t,s>, red[1]=<t,t> -> omega
// stage 4: after x,r update      red[0] = <rhat,r>, red[1] = ||r||^2 -> beta, rho, convergence
__global__ void bicg_scalar_kernel(int stage, Scalars* sc, const double* __restrict__ red, double tol2) {
    if (stage != 0 && sc->done) return;
    switch (stage) {
        case 0:
            sc->bnorm2 = red[0]; sc->rr = red[0];
            sc->rho = red[0];  // <rhat, r0> = ||b||^2 since rhat = r0 = b
            sc->alpha = 1.0; sc->omega = 1.0; sc->beta = 0.0;
            sc->done = (red[0] == 0.0) ? 1 : 0; sc->iters = 0;
            break;
        case 1: {
            double den = red[0];
            if (den == 0.0 || isnan(den)) { sc->done = 2; break; }
            sc->alpha = sc->rho / den;
        } break;
        case 2:
            sc->tmp[0] = red[0];  // ||s||^2 (reported if we stop at the half step)
            break;
        case 3: {
            double tt = red[1];
            if (tt == 0.0 || isnan(tt)) { sc->omega = 0.0; break; }  // s == 0: stage 4 then sees ||r|| = ||s||
            sc->omega = red[0] / tt;
        } break;
        case 4: {
            double rho_new = red[0];
            sc->rr = red[1];
            sc->iters += 1;
            if (red[1] <= tol2 * sc->bnorm2) { sc->done = 1; break; }
            if (sc->omega == 0.0 || sc->rho == 0.0 || isnan(rho_new)) { sc->done = 2; break; }
            sc->beta = (rho_new / sc->rho) * (sc->alpha / sc->omega);
            sc->rho = rho_new;
        } break;
    }
}

// ------------------------------------------------------------------------------------ flux integration
// reference PBSM3D.cpp:1467-1503: c = max(0,x) (NaN -> 0); Qsusp = sum c u_z dz; Qsubl = sum csubl c dz.
__global__ void __launch_bounds__(256) flux_kernel(int T, int L, double dz, double dt, const double* __restrict__ x,
                                                   const double* __restrict__ u_z, const double* __restrict__ csubl,
                                                   double* __restrict__ Qsusp, double* __restrict__ Qsubl,
                                                   double* __restrict__ Qsubl_mass, double* __restrict__ sum_subl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    double qs = 0.0, ql = 0.0;
    for (int z = 0; z < L; ++z) {
        const size_t r = (size_t)z * T + i;
        double c = x[r];
        c = (c < 0 || chm_is_nan(c)) ? 0.0 : c;
        qs += c * u_z[r] * dz;
        ql += csubl[r] * c * dz;
    }
    Qsusp[i] = qs;
    Qsubl[i] = ql;
    const double qm = ql * dt;
    Qsubl_mass[i] = qm;
    sum_subl[i] += qm;
}

// --------------------------------------------------------------------------------------- deposition
// RHS of the deposition system with the upwind donor rule (PBSM3D.cpp:1523-1656); the matrix is static.
// Qsusp/Qsalt of ghost faces come from qg (per owner block [2][gcnt]: Qsusp then Qsalt) after the halo exchange.
__global__ void __launch_bounds__(256) deposition_rhs_kernel(DevMesh m, const double* __restrict__ vw_dir,
                                                             const double* __restrict__ Qsusp, const double* __restrict__ Qsalt,
                                                             const double* __restrict__ qg, double* __restrict__ rhs,
                                                             double* __restrict__ partial) {
    const int T = m.T;
    double val_abs = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
        double vx, vy;
        wind_unit_vector(vw_dir[i], vx, vy);
        const double own_t = Qsusp[i], own_s = Qsalt[i];
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double udotm = vx * m.nx[(size_t)j * T + i] + vy * m.ny[(size_t)j * T + i];
            const double E = m.elen[(size_t)j * T + i];
            const int n = m.neigh[(size_t)j * T + i];
            double Qt = own_t, Qs = own_s;
            if (!(udotm > 0) && n >= 0) {
                if (n < T) { Qt = Qsusp[n]; Qs = Qsalt[n]; }
                else {
                    const int g = n - T, gs = m.gstart[g], gc = m.gcnt[g];
                    Qt = qg[(size_t)2 * gs + (g - gs)];
                    Qs = qg[(size_t)2 * gs + gc + (g - gs)];
                }
                if (chm_is_nan(Qs)) Qs = 0.0;
            }
            acc += -E * (Qt + Qs) * udotm;
        }
        rhs[i] = acc;
        val_abs = fmax(val_abs, fabs(acc));
    }
    val_abs = block_max(val_abs);
    if (threadIdx.x == 0) partial[blockIdx.x] = val_abs;
}

__device__ __forceinline__ double dep_row(const DevMesh& m, const double* __restrict__ ddiag, const double* __restrict__ doff,
                                          const double* __restrict__ p, const double* __restrict__ pg, int i) {
    const int T = m.T;
    double acc = ddiag[i] * p[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        int n = m.neigh[(size_t)j * T + i];
        if (n >= 0) acc += doff[(size_t)j * T + i] * (n < T ? p[n] : pg[n - T]);
    }
    return acc;
}

// Jacobi-preconditioned CG on the (SPD) deposition system, scalars on the device.
// init: x = 0, r = b, z = r/diag, p = z; partials <- <r,z>, <r,r>
__global__ void __launch_bounds__(256) cg_init_kernel(int T, const double* __restrict__ b, const double* __restrict__ dinv,
                                                      double* __restrict__ x, double* __restrict__ r, double* __restrict__ p,
                                                      double* __restrict__ partial, int stride) {
    double a0 = 0.0, a1 = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < T; k += gridDim.x * blockDim.x) {
        double rv = b[k], zv = rv * dinv[k];
        x[k] = 0.0; r[k] = rv; p[k] = zv;
        a0 += rv * zv; a1 += rv * rv;
    }
    a0 = block_sum(a0);
    if (threadIdx.x == 0) partial[blockIdx.x] = a0;
    a1 = block_sum(a1);
    if (threadIdx.x == 0) partial[stride + blockIdx.x] = a1;
}
// Ap = A p ; partial <- <p, Ap>
__global__ void __launch_bounds__(256) cg_spmv_kernel(DevMesh m, const double* __restrict__ ddiag, const double* __restrict__ doff,
                                                      const double* __restrict__ p, const double* __restrict__ pg,
                                                      double* __restrict__ Ap, double* __restrict__ partial,
                                                      const Scalars* __restrict__ sc) {
    if (sc && sc->done) return;
    double a = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.T; i += gridDim.x * blockDim.x) {
        double v = dep_row(m, ddiag, doff, p, pg, i);
        Ap[i] = v;
        a += v * p[i];
    }
    if (partial) { a = block_sum(a); if (threadIdx.x == 0) partial[blockIdx.x] = a; }
}
// x += alpha p ; r -= alpha Ap ; partials <- <r, r/diag>, <r,r>
__global__ void __launch_bounds__(256) cg_update_kernel(int T, const Scalars* __restrict__ sc, const double* __restrict__ dinv,
                                                        const double* __restrict__ p, const double* __restrict__ Ap,
                                                        double* __restrict__ x, double* __restrict__ r,
                                                        double* __restrict__ partial, int stride) {
    if (sc->done) return;
    const double alpha = sc->alpha;
    double a0 = 0.0, a1 = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < T; k += gridDim.x * blockDim.x) {
        x[k] += alpha * p[k];
        double rv = r[k] - alpha * Ap[k];
        r[k] = rv;
        a0 += rv * rv * dinv[k];
        a1 += rv * rv;
    }
    a0 = block_sum(a0);
    if (threadIdx.x == 0) partial[blockIdx.x] = a0;
    a1 = block_sum(a1);
    if (threadIdx.x == 0) partial[stride + blockIdx.x] = a1;
}
// p = r/diag + beta p
__global__ void __launch_bounds__(256) cg_p_kernel(int T, const Scalars* __restrict__ sc, const double* __restrict__ dinv,
                                                   const double* __restrict__ r, double* __restrict__ p) {
    if (sc->done) return;
    const double beta = sc->beta;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < T; k += gridDim.x * blockDim.x) p[k] = r[k] * dinv[k] + beta * p[k];
}
// stage 0: after init red = {<r,z>, <r,r>}; stage 1: after spmv red[0] = <p,Ap>; stage 2: after update red = {<r,z>,<r,r>}
__global__ void cg_scalar_kernel(int stage, Scalars* sc, const double* __restrict__ red, double tol2) {
    if (stage != 0 && sc->done) return;
    switch (stage) {
        case 0:
            sc->rho = red[0]; sc->bnorm2 = red[1]; sc->rr = red[1];
            sc->done = (red[1] == 0.0) ? 1 : 0; sc->iters = 0; sc->beta = 0.0; sc->alpha = 0.0;
            break;
        case 1: {
            double den = red[0];
            if (den == 0.0 || isnan(den)) { sc->done = 2; break; }
            sc->alpha = sc->rho / den;
        } break;
        case 2:
            sc->iters += 1;
            sc->rr = red[1];
            if (red[1] <= tol2 * sc->bnorm2) { sc->done = 1; break; }
            if (isnan(red[0])) { sc->done = 2; break; }
            sc->beta = red[0] / sc->rho;
            sc->rho = red[0];
            break;
    }
}

// Drift update (PBSM3D.cpp:1710-1740).
__global__ void __launch_bounds__(256) drift_kernel(int T, double dt, const double* __restrict__ q, const double* __restrict__ swe_in,
                                                    const unsigned char* __restrict__ salt, double* __restrict__ drift_mass,
                                                    double* __restrict__ sum_drift, double* __restrict__ more_than_avail) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    double qdep = q[i];
    qdep = chm_is_nan(qdep) ? 0.0 : qdep;
    double mass = qdep * dt;
    double swe = swe_in[i];
    swe = chm_is_nan(swe) ? 0.0 : swe;
    if (mass < 0 && fabs(mass) > swe) { more_than_avail[i] = 1.0; mass = -swe; }
    if (mass < 0 && !salt[i]) mass = 0.0;
    drift_mass[i] = mass;
    sum_drift[i] += mass;
}

// ---------------------------------------------------------------------------------------------- halo
// Pack the rows of the faces a partner needs into its send block: buf[off_p*nl + z*cnt_p + k] = v[z*T + idx[k]].
// seg[] maps each packed entry to (partner block offset, count, position) so one launch serves all partners.
__global__ void __launch_bounds__(256) halo_pack_kernel(int n_send, int nl, int T, const int* __restrict__ send_idx,
                                                        const int* __restrict__ send_boff, const int* __restrict__ send_cnt,
                                                        const int* __restrict__ send_pos, const double* __restrict__ v,
                                                        double* __restrict__ buf) {
    size_t total = (size_t)n_send * nl;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int z = (int)(e / n_send), k = (int)(e - (size_t)z * n_send);
        buf[(size_t)send_boff[k] * nl + (size_t)z * send_cnt[k] + send_pos[k]] = v[(size_t)z * T + send_idx[k]];
    }
}

__global__ void fill_kernel(size_t n, double* __restrict__ p, double v) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) p[k] = v;
}

}  // namespace pbsm3d

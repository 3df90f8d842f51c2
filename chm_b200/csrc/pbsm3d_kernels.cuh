// sm_100a kernels of the PBSM3D hot path.  All arithmetic is fp64; the path is sparse and HBM-bound
// (≈0.15 flop/B), so there are no tensor-core instructions here: the rules that matter are coalesced
// layer-major SoA streams, one pass over each array, L2-resident neighbour gathers, enough independent
// loads in flight per thread to cover HBM latency, and grids that fill the 148 SMs.
//
// INTERNAL FACE ORDER.  CHM's face order (ascending cell_global_id) is the order of every array that crosses
// the C-ABI.  Inside the library faces are stored colour-major: the owned faces are coloured so that no two
// edge-neighbours share a colour (2 colours when the dual graph is bipartite, else greedy ≤ 4), each colour
// class keeps CHM's order internally and starts on a 32-face (256-byte) boundary.  slot p -> CHM face perm[p]
// (-1 = padding slot, assembled as the identity row), CHM face i -> slot iperm[i].
//
// Data layout (device arrays; Tp = padded slots, S = Tp + padded ghost count, L = nLayer):
//   per slot        a[p]                      p in [0,Tp)
//   per slot-edge   a[j*Tp + p]               j in 0..2   (edge j is shared with neighbour j)
//   coefficients    a[z*Tp + p]               one stream per coefficient, layer-major
//   vectors         v[z*S + p]                ghost-extended: v[z*S + Tp + g] is ghost face g of layer z,
//                                             so a neighbour gather is x[z*S + nbs] with no owned/ghost branch
#pragma once
#include <cuda_runtime.h>
#include <type_traits>
#include "pbsm3d_physics.cuh"
#include "pbsm3d_math.cuh"

namespace pbsm3d {

struct DevConfig {
    int L;
    int do_fixed_settling, do_sublimation, do_lateral_diff, rouault, enable_veg;
    int use_exp_fetch, use_tanh_fetch, use_R94_lambda, use_PomLi;
    double settling_velocity, eps, min_sd_trans, cutoff, snow_diffusion_const;
    double dz;     // v_edge_height = susp_depth / nLayer (PBSM3D.cpp:225-226)
    double inv_dz;
    double l_max;  // 40 (PBSM3D.cpp:227)
};

struct DevMesh {
    int T, Tp, S, nG;
    const int* perm;       // [Tp] slot -> CHM local face, -1 = pad
    const int* iperm;      // [T]  CHM local face -> slot
    const int* nbs;        // [3][Tp] neighbour slot (ghost g -> Tp+g); own slot when there is no neighbour
    const double* nx;      // [3][Tp]
    const double* ny;      // [3][Tp]
    const double* elen;    // [3][Tp]
    const double* area;    // [Tp]
    const double* zc;      // [Tp] centroid elevation (face->get_z())
    const double* canopy;  // [Tp] or null
    const double* lai;     // [Tp] or null
    const double* stalk_n; // [Tp] or null
    const double* stalk_dv;
    const unsigned char* water;  // [Tp] or null
};

struct DevForcing {  // CHM order, [T]
    const double *U_R, *u2, *sd, *swe, *t, *rh, *vw_dir, *fetch, *psh /* p_snow_hours: use_PomLi_probability only */;
};

struct SuspSystem {
    // The assembled suspension system in the form the line solver uses (see "assembly" below): Thomas pivots and row-scaled
    // coefficients; the reference's diag/lat/below/above follow from them (reconstruct_rows_kernel).
    double* den;                   // [L][Tp]   den_z = d_z - lo_z cp_{z-1}
    double* cp;                    // [L][Tp]   up_z / den_z
    double* latS;                  // [3][L][Tp] lat_j / den
    double* belowS;                // [L][Tp]   lo / den
    float4* pack32;                // [L][Tp] {latS_0, latS_1, latS_2, belowS} rounded to fp32, one 16-byte load per row: what the
    float* cp32;                   // [L][Tp] sweeps far from convergence stream instead of the five fp64 arrays above
    double* rhs0;                  // [Tp]      b of layer 0 (all other layers are 0)
    double* rhsS0;                 // [Tp]      rhs0 / den[0]
    double *u_z, *csubl;           // [L][Tp]
    double *Qsalt, *c_salt;        // [Tp]
    unsigned char* salt;           // [Tp]
    unsigned char* live;           // [S] active set of the line solver (gs_persistent_kernel): a superset of the columns whose
                                   //     right-hand side or iterate is non-zero; the assembly seeds it with b_p != 0
    double* prob;                  // [Tp] blowingsnow_probability (face variable; written on saltating faces with use_PomLi_probability)
    const double* ltab;            // [kTabN][L] per-layer constants of a column with hs = 0 (layer_table_kernel)
};

// Device-resident control block: recurrence scalars, convergence flags and the tickets of the fused
// "last block folds" reductions.  The host reads it once per step (or per batch while a solve is still open).
struct Scalars {
    double rho, alpha, omega, beta;   // BiCGStab / CG recurrences
    double rr, bnorm2;                // ||r||^2, ||b||^2 of the solve in flight
    double tmp[4];
    double susp_rhs_max, dep_rhs_max;
    double susp_bnorm2, susp_rr;      // line solver: ||b||^2 and the last checked ||b-Ax||^2
    double rr_hist[16];               // ||b-Ax||^2 at the residual checks of this step
    int it_hist[16];
    int n_checks;
    int done;                         // Krylov solve in flight: 1 = converged, 2 = breakdown
    int iters;
    int susp_present, dep_present;
    int susp_done, susp_iters;        // line solver: converged flag and the sweep count at detection
    int susp_ok;                      // suspension phase finished (converged, or nothing to solve)
    int susp_sweeps, susp_stalled;    // persistent solve: sweeps executed; 1 = it stopped contracting (hand over to BiCGStab)
    int dep_sweeps;                   // persistent deposition solve: sweeps executed
    int dep_ok;                       // deposition solve finished (converged)
    int tail_done;                    // flux/deposition-rhs ran on a finished suspension solve
    int drift_done;
    int dep_buf;                      // Chebyshev: which of the two q buffers holds the converged iterate
    int log_n, log_cap;               // setup only: CG recurrence log for the Lanczos spectrum estimate
    double *log_alpha, *log_beta;
    unsigned ticket[4];
    unsigned long long col_updates[4];  // persistent line solver: face-column updates executed in each storage phase
                                        // (fp32 x / fp32 coefficients / fp64) and columns evaluated by the residual checks
    unsigned n_seeds;                   // faces with a non-zero right-hand side, counted by the row kernel (zeroed by the host
    unsigned n_seeds_step;              // at the start of a step); its value when the step's solve starts (FLAGS_SUSP)
    int peer_error;                   // sticky: a peer-memory wait timed out (see PeerTable)
};

// ---------------------------------------------------------------------------------------------- setup
// Neighbour table in slot numbering.  neigh_chm is [T][3] with -1 = none and >= T = ghost (T + g).
__global__ void neighbour_slots_kernel(int T, int Tp, const int* __restrict__ perm, const int* __restrict__ iperm,
                                       const int* __restrict__ neigh_chm, int* __restrict__ nbs) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp) return;
    const int i = perm[p];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        int s = p;
        if (i >= 0) {
            const int n = neigh_chm[(size_t)i * 3 + j];
            if (n >= T) s = Tp + (n - T);
            else if (n >= 0) s = iperm[n];
        }
        nbs[(size_t)j * Tp + p] = s;
    }
}

// dst[p] = src[perm[p]] (pad -> fill): brings a CHM-ordered per-face array into slot order.
template <typename U>
__global__ void to_slots_kernel(int Tp, const int* __restrict__ perm, const U* __restrict__ src, U* __restrict__ dst, U fill) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp) return;
    const int i = perm[p];
    dst[p] = i >= 0 ? src[i] : fill;
}
// dst[r*T + i] = src[r*src_stride + iperm[i]]: slot order -> CHM order for `rows` stacked arrays.
__global__ void to_chm_kernel(int rows, int T, size_t src_stride, const int* __restrict__ iperm, const double* __restrict__ src,
                              double* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const int p = iperm[i];
    for (int r = 0; r < rows; ++r) dst[(size_t)r * T + i] = src[(size_t)r * src_stride + p];
}
__global__ void to_chm_u8_kernel(int T, const int* __restrict__ iperm, const unsigned char* __restrict__ src,
                                 unsigned char* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T) dst[i] = src[iperm[i]];
}

// Face geometry from the three vertices of each face (reference: mesh/triangulation.hpp:1443-1475
// edge_unit_normal/edge, :1491-1498 edge_length, :1577-1589 center, :1830-1856 get_area).
// Runs once in pbsm3d_create.  Products and sums use the __d*_rn intrinsics, which nvcc never contracts
// into FMAs, so every value is bit-identical to the plain IEEE fp64 evaluation the reference (and numpy) do.
// Slots [0,Tp) are owned faces (or pads), slots [Tp, Tp+nG) the ghosts (centroid only).
__global__ void geometry_kernel(int T, int Tp, int nG, const int* __restrict__ perm, const double* __restrict__ verts,
                                const double* __restrict__ area_param, double* __restrict__ nx, double* __restrict__ ny,
                                double* __restrict__ elen, double* __restrict__ area, double* __restrict__ cx,
                                double* __restrict__ cy, double* __restrict__ cz, double* __restrict__ slope) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp + nG) return;
    const int i = p < Tp ? perm[p] : T + (p - Tp);
    if (i < 0) {  // pad: a harmless unit triangle
        cx[p] = cy[p] = cz[p] = 0.0;
        slope[p] = 0.0;
        for (int k = 0; k < 3; ++k) { nx[(size_t)k * Tp + p] = 1.0; ny[(size_t)k * Tp + p] = 0.0; elen[(size_t)k * Tp + p] = 1.0; }
        area[p] = 1.0;
        return;
    }
    double px[3], py[3], pz[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        px[k] = verts[(size_t)i * 9 + k * 3 + 0];
        py[k] = verts[(size_t)i * 9 + k * 3 + 1];
        pz[k] = verts[(size_t)i * 9 + k * 3 + 2];
    }
    cx[p] = __ddiv_rn(__dadd_rn(__dadd_rn(px[0], px[1]), px[2]), 3.0);
    cy[p] = __ddiv_rn(__dadd_rn(__dadd_rn(py[0], py[1]), py[2]), 3.0);
    cz[p] = __ddiv_rn(__dadd_rn(__dadd_rn(pz[0], pz[1]), pz[2]), 3.0);
    if (p >= Tp) return;  // ghosts only need a centroid
    {   // face::slope (triangulation.hpp:1501-1523): acos(norm_dot(CGAL::unit_normal(v0, v1, v2), (0,0,1))); snow_slide reads it
        const double ax = __dsub_rn(px[1], px[0]), ay = __dsub_rn(py[1], py[0]), az = __dsub_rn(pz[1], pz[0]);
        const double bx = __dsub_rn(px[2], px[0]), by = __dsub_rn(py[2], py[0]), bz = __dsub_rn(pz[2], pz[0]);
        double n_x = __dsub_rn(__dmul_rn(ay, bz), __dmul_rn(az, by)), n_y = __dsub_rn(__dmul_rn(az, bx), __dmul_rn(ax, bz)),
               n_z = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
        const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(n_x, n_x), __dmul_rn(n_y, n_y)), __dmul_rn(n_z, n_z)));
        n_x = __ddiv_rn(n_x, len); n_y = __ddiv_rn(n_y, len); n_z = __ddiv_rn(n_z, len);
        const double na = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(n_x, n_x), __dmul_rn(n_y, n_y)), __dmul_rn(n_z, n_z)));
        slope[p] = acos(__ddiv_rn(n_z, na));
    }
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
        nx[(size_t)k * Tp + p] = __ddiv_rn(n_x, nrm);
        ny[(size_t)k * Tp + p] = __ddiv_rn(n_y, nrm);
        elen[(size_t)k * Tp + p] = __dsqrt_rn(__dadd_rn(__dmul_rn(ex[k], ex[k]), __dmul_rn(ey[k], ey[k])));
    }
    if (area_param) {
        area[p] = area_param[i];
    } else {
        double v1x = __dsub_rn(px[1], px[0]), v1y = __dsub_rn(py[1], py[0]);
        double v2x = __dsub_rn(px[2], px[0]), v2y = __dsub_rn(py[2], py[0]);
        area[p] = __ddiv_rn(__dsub_rn(__dmul_rn(v1x, v2y), __dmul_rn(v1y, v2x)), 2.0);
    }
}

// Static part of the deposition system (reference re-derives it every step, PBSM3D.cpp:1546,1609-1628):
// diag = area + sum eps*E_j/dx_j, off_j = -eps*E_j/dx_j, dx_j = math::gis::distance of the two centroids: 2-D Euclidean on a
// projected mesh (distance_UTM, coordinates.cpp:94-100), haversine on a geographic one (distance_latlong, :68-92).
// cx/cy are [Tp + nG], so a neighbour that is a ghost face resolves like any other.
// math::gis::distance_latlong (coordinates.cpp:68-92): haversine on a sphere of radius 6378137 m, x = longitude, y = latitude
// in degrees.  What core.cpp:809-821 installs as math::gis::distance on a geographic mesh.
__device__ __forceinline__ double distance_latlong(double x1, double y1, double x2, double y2) {
    const double d2r = kPi / 180.0;
    const double lat1 = y1 * d2r, lon1 = x1 * d2r, lat2 = y2 * d2r, lon2 = x2 * d2r;
    const double dphi = lat2 - lat1, dlon = lon2 - lon1;
    const double sp = sin(dphi / 2.), sl = sin(dlon / 2.);
    const double a = sp * sp + cos(lat1) * cos(lat2) * sl * sl;
    return 6378137.0 * (2. * atan2(sqrt(a), sqrt(1. - a)));
}
__global__ void deposition_matrix_kernel(int Tp, double eps, int geographic, const int* __restrict__ perm, const int* __restrict__ nbs,
                                         const double* __restrict__ elen, const double* __restrict__ area,
                                         const double* __restrict__ cx, const double* __restrict__ cy, double* __restrict__ dx,
                                         double* __restrict__ ddiag, double* __restrict__ doff, double* __restrict__ dinv) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp) return;
    double d = area[p];
    const bool pad = perm[p] < 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        int n = nbs[(size_t)j * Tp + p];
        double dist = 2.0, c = 0.0;  // dx[] default 2.0, PBSM3D.cpp:1534
        if (n != p && !pad) {
            if (geographic) {
                dist = distance_latlong(cx[p], cy[p], cx[n], cy[n]);
            } else {
                double ddx = __dsub_rn(cx[p], cx[n]), ddy = __dsub_rn(cy[p], cy[n]);
                dist = __dsqrt_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)));
            }
            c = __ddiv_rn(__dmul_rn(eps, elen[(size_t)j * Tp + p]), dist);
        }
        dx[(size_t)j * Tp + p] = dist;
        d = __dadd_rn(d, c);
        doff[(size_t)j * Tp + p] = -c;
    }
    ddiag[p] = d;
    dinv[p] = 1.0 / d;
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
// Block-wide reductions staged through shared memory; result valid in thread 0 (warp 0).
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

constexpr int kRedBlocks = 148 * 8;  // partial-sum slots: a multiple of the SM count
constexpr int kRedThreads = 256;

// Fused grid reduction.  Every block publishes up to two partials; the block that draws the last ticket folds
// them IN INDEX ORDER (so the result does not depend on which block finished last) and returns true in all of
// its threads with the totals in out0/out1 (thread 0).  op: 0 = sum, 1 = max.
template <int NV>
__device__ __forceinline__ bool grid_fold(double v0, double v1, int op0, int op1, double* __restrict__ partial, int stride,
                                          unsigned* ticket, double& out0, double& out1) {
    __shared__ bool last;
    v0 = op0 ? block_max(v0) : block_sum(v0);
    if (NV > 1) v1 = op1 ? block_max(v1) : block_sum(v1);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = v0;
        if (NV > 1) partial[stride + blockIdx.x] = v1;
        __threadfence();
        unsigned t = atomicInc(ticket, gridDim.x - 1);  // wraps to 0 after the last block: self-resetting
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    double a0 = 0.0, a1 = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
        double q0 = __ldcg(&partial[b]);
        a0 = op0 ? fmax(a0, q0) : a0 + q0;
        if (NV > 1) {
            double q1 = __ldcg(&partial[stride + b]);
            a1 = op1 ? fmax(a1, q1) : a1 + q1;
        }
    }
    out0 = op0 ? block_max(a0) : block_sum(a0);
    if (NV > 1) out1 = op1 ? block_max(a1) : block_sum(a1);
    return true;
}

// ------------------------------------------------------------------------------------------- assembly
// HOT LOOP 1: saltation + suspension assembly (reference PBSM3D.cpp:436-1406) fused with the forward elimination of every
// column's tridiagonal block (the line-solver factor), the row scaling the sweep streams, and the two global facts the step
// needs next: max|b| (suspension_present, PBSM3D.cpp:1424-1427) and ||b||^2 (the stopping rule of LinearAlgebra.cpp:168).
//
// What is stored per row (84 B; the reference's own values are recovered on demand, see row_of_A below):
//     den  = d - lo*cp_{z-1}   (Thomas pivot)          cp     = up / den
//     latS = lat_j / den  (3)                          belowS = lo / den
//     pack32 = {latS_0..2, belowS} and cp32 rounded to fp32 (what the sweeps far from convergence stream)
//     u_z, csubl (flux integration)
// with  diag = den (1 + belowS cp_{z-1}),  below = belowS den,  above = cp den,  lat_j = latS_j den.
//
// Padding slots are identity rows with a zero right-hand side; written once in pbsm3d_create.
__global__ void assemble_pads_kernel(DevMesh m, SuspSystem s, int L) {
    const int Tp = m.Tp;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp || m.perm[p] >= 0) return;
    s.Qsalt[p] = 0.0; s.c_salt[p] = 0.0; s.salt[p] = 0; s.rhs0[p] = 0.0; s.rhsS0[p] = 0.0; s.live[p] = 0;
    for (int z = 0; z < L; ++z) {
        const size_t r = (size_t)z * Tp + p;
        s.den[r] = 1.0; s.cp[r] = 0.0; s.belowS[r] = 0.0;
        s.cp32[r] = 0.f; s.pack32[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        s.u_z[r] = 0.0; s.csubl[r] = 0.0;
        for (int j = 0; j < 3; ++j) s.latS[((size_t)j * L + z) * Tp + p] = 0.0;
    }
}

// ---- what depends on the height above the saltation layer only (PBSM3D.cpp:1003-1039, 1156-1158)
struct LayerConsts {
    double cz;       // z*dz + hs + dz/2
    double inv_mm;   // 1 / mean particle mass
    double r_z;      // mean particle radius
    double omega;    // settling velocity
    double Qr;       // radiative term of the sublimation model
    double sig;      // 1.019 + 0.027 ln cz
    double lmix;     // mixing length
    double ulog;     // ln((cz - z0)/z0): the log-profile factor of u_z
    double ulog136;  // ulog^1.36 (table rows only: xrz = 0.005 u_z^1.36 without a log/exp pair)
};
constexpr int kTabN = 9;  // doubles per layer in SuspSystem::ltab (LayerConsts, field order)

// x^y on this path is exp(y ln x) through pbsm3d_math.cuh (flog/fexp: ~1 ulp, branch-free): |y ln x| < 20 on every use, so
// the result is within ~2e-15 relative of pow(); the parity bar on coefficients is 1e-12.
__device__ __forceinline__ LayerConsts layer_consts(const DevConfig& c, double cz, double sd) {
    LayerConsts o;
    o.cz = cz;
    const double lcz = flog(cz);
    const double lrm = -0.258 * lcz;                        // rm = 4.6e-5 cz^-0.258 (:1009)
    const double rm = 4.6e-5 * fexp(lrm);
    const double ia = frcp(4.08 + 12.6 * cz);               // mm_alpha (:1012)
    const double P = 1.0 + ia * (3.0 + 2.0 * ia);           // 1 + 3/a + 2/a^2
    const double mm = 4.0 / 3.0 * kPi * kRhoIce * rm * rm * rm * P;  // :1013-1014
    o.inv_mm = frcp(mm);
    // r_z = (3 mm / (4 pi rho_p))^0.3333333 (:1017, the literal exponent) = (rm^3 P)^(1/3 - e), e = 1/3 - 0.3333333:
    //     = rm cbrt(P) exp(-e ln(rm^3 P)); the last factor is 1 + O(1e-6), so ln(rm^3 P) is needed to ~1e-9 only.
    {
        const double e = 1.0 / 3.0 - 0.3333333;
        const double l3 = 3.0 * (-9.986869161475179 /* ln 4.6e-5 */ + lrm) + (double)__logf((float)P);
        const double t = -e * l3;  // |t| < 2e-6: three terms of exp() are exact to 1e-19
        o.r_z = rm * fcbrt_1_15(P) * (1.0 + t * (1.0 + t * (0.5 + t * (1.0 / 6.0))));  // P in (1, 1.46]
    }
    o.omega = c.do_fixed_settling ? c.settling_velocity : 1.1e7 * fexp(1.8 * flog(o.r_z));  // :1024-1033
    o.Qr = 0.9 * kPi * rm * rm * 120.0;                     // :1097
    o.sig = 1.019 + 0.027 * lcz;                            // :1095
    const double kz = kKappa * (cz + kZ0Snow);
    o.lmix = kz * c.l_max * frcp(kz + c.l_max);             // :1156
    o.ulog = flog(((cz + sd) - (sd + kZ0Snow)) * (1.0 / kZ0Snow));  // :970-976 with hz = cz + sd
    o.ulog136 = 0.0;
    return o;
}
// Table of LayerConsts for hs = 0 (every non-saltating face: hs = 0, z0 = Z0_SNOW), built once in pbsm3d_create by
// the same device code the saltating faces run per row.
__global__ void layer_table_kernel(DevConfig c, double* __restrict__ tab) {
    const int z = threadIdx.x;
    if (z >= c.L) return;
    const LayerConsts o = layer_consts(c, z * c.dz + 0.0 + c.dz / 2.0, 0.0);
    const double v[kTabN] = {o.cz, o.inv_mm, o.r_z, o.omega, o.Qr, o.sig, o.lmix, o.ulog, fpow(o.ulog, 1.36)};
    for (int k = 0; k < kTabN; ++k) tab[k * c.L + z] = v[k];
}

// ---- per-face part: saltation (PBSM3D.cpp:436-925) and the factors of the layer loop that do not depend on z
struct FaceConsts {
    double hs, height_diff, u28 /* 2.8 u*_t */, UQ /* U_R / ln((Z_UR - (sd+z0))/z0) */, UQ136 /* 0.005 UQ^1.36 */, uref, sd;
    double C1p, C2;       // dm/dt = C1p sig Nu r_z + C2 Qr   (the reference's expression :1105-1123 with Sh = Nu cancelled)
    double ustar, area, aod /* area / dz */, a4p /* area / (hs/2 + dz/2) */, v5 /* area dz / 5 */, c_salt;
    double Aj[3], g[3];   // lateral face areas E_j dz; unit wind . edge normal
    int flags;            // bit 0 saltation, bits 1-3 neighbour j present, bit 4 active (a real face)
    int p;                // slot
};
__device__ __forceinline__ FaceConsts face_prelude(const DevConfig& c, const DevMesh& m, const DevForcing& f, const SuspSystem& s,
                                                   double dt, int p, int i) {
    const int Tp = m.Tp;
    FaceConsts o;
    o.p = p;
    double fetch = 1000.0;
    if (c.use_exp_fetch || c.use_tanh_fetch) fetch = f.fetch[i];  // depends("fetch"), PBSM3D.cpp:181-184: the host insists on it
    const double uref = f.U_R[i];
    double sd = f.sd[i];
    sd = chm_is_nan(sd) ? 0.0 : sd;
    const double u2 = f.u2[i];
    double swe = f.swe[i];
    swe = chm_is_nan(swe) ? 0.0 : swe;
    const double Tc = f.t[i];
    const double phi = f.vw_dir[i];
    const double area = m.area[p];
    double nxj[3], nyj[3], Ej[3];
    int flags = 16;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        nxj[j] = m.nx[(size_t)j * Tp + p];
        nyj[j] = m.ny[(size_t)j * Tp + p];
        Ej[j] = m.elen[(size_t)j * Tp + p];
        if (m.nbs[(size_t)j * Tp + p] != p) flags |= 2 << j;
    }
    double height_diff = 0.0, LAI = 0.0, Nst = 0.0, dv = 0.0;
    if (c.enable_veg) {
        height_diff = fmax(0.0, m.canopy[p] - sd);
        if (c.use_R94_lambda) LAI = m.lai[p];
        else { Nst = m.stalk_n ? m.stalk_n[p] : 1.0; dv = m.stalk_dv ? m.stalk_dv[p] : 0.8; }
    }
    const bool water = m.water ? (m.water[p] != 0) : false;
    const double ust_th = 0.35 + (1.0 / 150.0) * Tc + (1.0 / 8200.0) * Tc * Tc;

    bool salt = false;
    double lambda = 0.0, ustar = 1.3;
    if (height_diff <= c.cutoff && sd >= c.min_sd_trans && !water) {
        lambda = c.use_R94_lambda ? 0.5 * LAI * height_diff : Nst * dv * height_diff;
        ustar = u2 * kKappa / 9.210340371976184 /* ln(2/0.0002) */;
        if (ustar >= ust_th) salt = true;
    }
    const double z0 = kZ0Snow;
    if (!salt) ustar = fmax(0.01, kKappa * uref / 8.517193191416238 /* ln(Z_UR/z0) */);
    ustar = fmax(0.01, ustar);
    const double hs = salt ? 0.08436 * fpow(ustar, 1.27) : 0.0;  // ustar^1.27 (:752); see layer_consts on exp(y ln x)

    const double t = Tc + 273.15;
    double vx, vy;
    wind_unit_vector(phi, vx, vy);
    double Qsalt = 0.0, c_salt = 0.0;
    if (salt) {
        const double rho_f = std_dry_air_density_fast(m.zc[p], t);
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
        if (c.use_PomLi) {  // Pomeroy & Li 2000 upscaled probability of blowing snow (PBSM3D.cpp:848-866)
            const double A = f.psh[i];  // hours since the last snowfall
            const double z10 = 10.0 + sd;
            const double u10 = z10 < kZUR ? log_scale_wind(uref, kZUR, z10, sd, kZ0Snow) : uref;  // :451-463
            const double u_mean = 11.2 + 0.365 * Tc + 0.00706 * Tc * Tc + 0.9 * log(A);
            const double delta = 0.145 * Tc + 0.00196 * Tc * Tc + 4.3;
            const double z0v = (Nst * dv * height_diff) / 2.0;
            const double us = u10 / sqrt((1 + 340.0 * z0v));
            const double Pu10 = 1.0 / (1.0 + exp((sqrt(kPi) * (u_mean - us)) / delta));
            s.prob[p] = Pu10;
            c_salt *= Pu10;
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
    s.Qsalt[p] = Qsalt;
    s.c_salt[p] = c_salt;
    s.salt[p] = salt ? 1 : 0;
    if (salt) flags |= 1;

    // layer-independent pieces of the sublimation model: with Sh = Nu (:1046) the reference's dm/dt (:1105-1123)
    //     Sh rho_sat D (2 pi' Nu Rg r_z sigma t^2 lambda_t - Ls Mw Qr + Qr Rg t) / (D Ls Sh (Ls Mw - Rg t) rho_sat + lambda_t t^2 Nu Rg)
    // is  C1 sigma Nu r_z + C2 Qr  with the two per-face constants below (pi' = 6.283185308 / 2 as written there)
    {
        const double rh = f.rh[i] * 0.01;
        const double es = 611.21 * fexp((17.502 * Tc) * frcp(240.97 + Tc));  // saturatedVapourPressure, Kelvin >= 0 branch (Atmosphere.cpp:62-80)
        const double D = 2.06e-5 * fpow(t * (1.0 / 273.15), 1.75);
        const double lambda_t = 0.000063 * t + 0.00673;
        const double Ls = 2.838e6, Mw = 18.01, Rg = 8313.0;
        const double rho_sat = (Mw * es) * frcp(Rg * t);
        const double inv_den = frcp(D * Ls * (Ls * Mw - Rg * t) * rho_sat + lambda_t * t * t * Rg);
        o.C1p = rho_sat * D * (6.283185308 * Rg * t * t * lambda_t) * inv_den * (rh - 1.0);
        o.C2 = rho_sat * D * (Rg * t - Ls * Mw) * inv_den;
    }
    o.hs = hs;
    o.height_diff = height_diff;
    o.u28 = 2.8 * ust_th;
    o.UQ = uref * frcp(flog((kZUR - (sd + z0)) * (1.0 / kZ0Snow)));
    o.UQ136 = o.UQ > 0 ? 0.005 * fpow(o.UQ, 1.36) : 0.0;
    o.uref = uref;
    o.sd = sd;
    o.ustar = ustar;
    o.area = area;
    o.aod = area * c.inv_dz;
    o.a4p = area * frcp(hs / 2.0 + c.dz / 2.0);
    o.v5 = area * c.dz / 5.0;
    o.c_salt = c_salt;
    const double inrm = frcp(fsqrt(vx * vx + vy * vy));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        o.Aj[j] = Ej[j] * c.dz;
        o.g[j] = (vx * nxj[j] + vy * nyj[j]) * inrm;  // u . m_j = u_z g_j  (uvw = u_z v / |v|, :1189-1200)
    }
    o.flags = flags;
    return o;
}

// ---- one row: everything of PBSM3D.cpp:946-1348 for (face, z) except the column recurrence
struct RowCoef64 {
    double d, lo, up, off[3], rhs;
};
__device__ __forceinline__ RowCoef64 assemble_row(const DevConfig& c, const FaceConsts& fc, const LayerConsts& lc, int z, int L,
                                                  double& u_z_out, double& csubl_out) {
    const bool salt = fc.flags & 1;
    const double cz = lc.cz;
    double u_z, xrz;  // xrz = 0.005 u_z^1.36 (:1020)
    if (cz >= fc.height_diff && cz + fc.sd < kZUR && fc.UQ * lc.ulog >= 0.01) {  // the log profile, not clamped: the common case
        u_z = fc.UQ * lc.ulog;
        xrz = fc.UQ136 * (lc.ulog136 != 0.0 ? lc.ulog136 : fpow(lc.ulog, 1.36));
    } else {
        if (salt && cz < fc.height_diff) u_z = fc.u28;
        else if (cz < fc.height_diff) u_z = 0.01;
        else if (cz + fc.sd < kZUR) u_z = fmax(0.01, fc.UQ * lc.ulog);
        else u_z = fmax(0.01, fc.uref);
        xrz = 0.005 * fpow(u_z, 1.36);
    }
    u_z_out = u_z;

    const double omega = lc.omega;
    const double Vr = omega + 3.0 * xrz * 0.70710678118654757 /* cos(pi/4) */;  // :1039
    const double Re = 2.0 * lc.r_z * Vr * (1.0 / 1.88e-5);
    const double Nu = 1.79 + 0.606 * fsqrt(Re);
    const double dmdtz = fc.C1p * lc.sig * Nu * lc.r_z + fc.C2 * lc.Qr;
    const double csubl = c.do_sublimation ? dmdtz * lc.inv_mm : 0.0;
    csubl_out = csubl;

    double diffusion_coeff = c.snow_diffusion_const;
    if (c.rouault) diffusion_coeff = 1.0 / (1.0 + (1.0 * omega * omega) / (1.56 * fc.ustar * fc.ustar));
    const double K = diffusion_coeff * fc.ustar * lc.lmix;
    const double area = fc.area;
    const double alpha3 = fc.aod * K;  // area K / dz
    const double alpha4 = alpha3;
    const double udotm3 = -omega, udotm4 = omega;
    const double Vc = fc.v5 * csubl;

    RowCoef64 o;
    double d = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double udotm = u_z * fc.g[j];
        const double alphaj = c.do_lateral_diff ? fc.Aj[j] * 0.00001 : 0.0;
        const bool has = fc.flags & (2 << j);
        double off = 0.0;
        if (udotm > 0) {
            if (has) { d += Vc - fc.Aj[j] * udotm - alphaj; off = alphaj; }
            else d += -0.1e-1 * alphaj - 1.0 * fc.Aj[j] * udotm + Vc;
        } else {
            if (has) { d += Vc - alphaj; off = -fc.Aj[j] * udotm + alphaj; }
            else d += -0.1e-1 * alphaj - 0.99 * fc.Aj[j] * udotm + Vc;
        }
        o.off[j] = off;
    }
    double lo = 0.0, up = 0.0, rhs = 0.0;
    if (z == 0) {
        const double alpha4p = fc.a4p * K;  // area K / (hs/2 + dz/2)
        d += Vc - area * udotm4 - alpha4p;
        rhs = -alpha4p * fc.c_salt;
        // the coupling to layer 1; with nLayer == 1 the reference sums it into a column outside the matrix (dropped)
        if (udotm3 > 0) { d += Vc - area * udotm3 - alpha3; up = alpha3; }
        else { d += Vc - alpha3; up = -area * udotm3 + alpha3; }
        if (L == 1) up = 0.0;
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
    o.d = d; o.lo = lo; o.up = up; o.rhs = rhs;
    return o;
}
__device__ __forceinline__ LayerConsts layer_lookup(const DevConfig& c, const FaceConsts& fc, const double* __restrict__ tab, int z) {
    // hs > 0 shifts every height (hs is set before the negative-c_salt clamp may clear the saltation flag, so test hs itself)
    if (fc.hs != 0.0) return layer_consts(c, z * c.dz + fc.hs + c.dz / 2.0, fc.sd);
    LayerConsts o;
    const int L = c.L;
    o.cz = __ldg(tab + z); o.inv_mm = __ldg(tab + L + z); o.r_z = __ldg(tab + 2 * L + z); o.omega = __ldg(tab + 3 * L + z);
    o.Qr = __ldg(tab + 4 * L + z); o.sig = __ldg(tab + 5 * L + z); o.lmix = __ldg(tab + 6 * L + z); o.ulog = __ldg(tab + 7 * L + z);
    o.ulog136 = __ldg(tab + 8 * L + z);
    return o;
}
__device__ __forceinline__ void store_row(const SuspSystem& s, size_t r, size_t LTp, const RowCoef64& rc, double den, double inv,
                                          double cp) {
    s.den[r] = den;
    s.cp[r] = cp;
    const double bS = rc.lo * inv, l0 = rc.off[0] * inv, l1 = rc.off[1] * inv, l2 = rc.off[2] * inv;
    s.belowS[r] = bS;
    s.latS[r] = l0; s.latS[LTp + r] = l1; s.latS[2 * LTp + r] = l2;
    s.cp32[r] = (float)cp;
    s.pack32[r] = make_float4((float)l0, (float)l1, (float)l2, (float)bS);
}

// Faces with b != 0 (the seeds of the line solver's active set), counted per warp into Scalars::n_seeds; whole warps call it.
__device__ __forceinline__ void count_seeds(Scalars* sc, unsigned n) {
    const unsigned tot = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(&sc->n_seeds, tot);
}

// Per-face records between the prelude kernel and the row kernels: SoA [kRecD][Tp] doubles + [Tp] ints, SLOT order (so the row
// kernels read records and write every stream fully coalesced whatever the number of colours).
constexpr int kRecD = 21;
struct FaceRecs {
    double* d;   // [kRecD][Tp]
    int* i;      // [Tp] flags (0 on padding slots: never written by the prelude, zeroed at create)
    int T;       // = Tp, the stride
};
__device__ __forceinline__ void store_rec(const FaceRecs& R, int i, const FaceConsts& fc) {
    const double v[kRecD] = {fc.hs, fc.height_diff, fc.u28, fc.UQ, fc.uref, fc.sd, fc.C1p, fc.C2, fc.ustar, fc.area, fc.c_salt,
                             fc.Aj[0], fc.Aj[1], fc.Aj[2], fc.g[0], fc.g[1], fc.g[2], fc.UQ136, fc.aod, fc.a4p, fc.v5};
#pragma unroll
    for (int k = 0; k < kRecD; ++k) R.d[(size_t)k * R.T + i] = v[k];
    R.i[i] = fc.flags;
}
__device__ __forceinline__ FaceConsts load_rec(const FaceRecs& R, int i) {
    FaceConsts fc;
    const double* d = R.d + i;
    const size_t T = R.T;
    fc.hs = d[0]; fc.height_diff = d[T]; fc.u28 = d[2 * T]; fc.UQ = d[3 * T]; fc.uref = d[4 * T]; fc.sd = d[5 * T];
    fc.C1p = d[6 * T]; fc.C2 = d[7 * T]; fc.ustar = d[8 * T]; fc.area = d[9 * T]; fc.c_salt = d[10 * T];
#pragma unroll
    for (int j = 0; j < 3; ++j) { fc.Aj[j] = d[(11 + j) * T]; fc.g[j] = d[(14 + j) * T]; }
    fc.UQ136 = d[17 * T]; fc.aod = d[18 * T]; fc.a4p = d[19 * T]; fc.v5 = d[20 * T];
    fc.flags = R.i[i];
    fc.p = i;
    return fc;
}

// Kernel 1 of the assembly: one thread per CHM face (coalesced forcing reads; runs per forcing chunk as the chunks land from the
// host): saltation, Qsalt/c_salt and the per-face factors of the layer loop, written to the face's SLOT (the partial sectors of
// neighbouring faces merge in L2).
__global__ void __launch_bounds__(128) face_prelude_kernel(DevConfig c, DevMesh m, DevForcing f, SuspSystem s, double dt, int i0, int i1,
                                                          FaceRecs R) {
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    const int p = m.iperm[i];
    store_rec(R, p, face_prelude(c, m, f, s, dt, p, i));
}

// Kernel 2, layer-parallel.  A block of NW warps owns tiles of 32 consecutive slots; warp z computes the rows of layer z, so
// the 32 lanes write 32 consecutive faces of every stream of that layer and no thread walks a column; the per-face record
// is read by every warp of the block (one DRAM read, L1 hits after that).  The one serial piece, the Thomas recurrence
// den_z = d_z - lo_z cp_{z-1} over the tile's 32 columns, runs on warp 0 out of shared memory between two barriers.
// Faces with hs = 0 (all non-saltating ones) take everything that depends on the height only from a per-layer table
// (SuspSystem::ltab): their rows cost one log/exp pair, a square root and the coefficient arithmetic.
// red[0] = max|b|, red[1] = sum b^2 over the chunk (this rank).
template <int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
assemble_tile_kernel(DevConfig c, DevMesh m, FaceRecs R, SuspSystem s, int i0, int i1, double* __restrict__ partial, int pstride,
                     Scalars* sc, double* __restrict__ red) {
    extern __shared__ double sm[];
    const int L = c.L;
    double* colA = sm;                         // [L][32] d   -> den
    double* colB = colA + (size_t)L * 32;      // [L][32] lo  -> 1/den
    double* colC = colB + (size_t)L * 32;      // [L][32] up  -> cp
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ntiles = (i1 - i0 + 31) / 32;
    const size_t LTp = (size_t)L * m.Tp;
    double mx = 0.0, ss = 0.0;
    unsigned nseed = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = i0 + tile * 32 + lane;  // a slot
        const bool act = i < i1 && w < L && (R.i[i] & 16);
        RowCoef64 rc;
        int p = 0;
        if (act) {
            const FaceConsts fc = load_rec(R, i);
            p = fc.p;
            const LayerConsts lc = layer_lookup(c, fc, s.ltab, w);
            double u_z, csubl;
            rc = assemble_row(c, fc, lc, w, L, u_z, csubl);
            const size_t r = (size_t)w * m.Tp + p;
            s.u_z[r] = u_z;
            s.csubl[r] = csubl;
            colA[w * 32 + lane] = rc.d;
            colB[w * 32 + lane] = rc.lo;
            colC[w * 32 + lane] = rc.up;
            if (w == 0) {
                mx = fmax(mx, fabs(rc.rhs));
                ss += rc.rhs * rc.rhs;
            }
        }
        __syncthreads();
        if (w == 0 && act) {
            double cp_prev = 0.0;
            for (int z = 0; z < L; ++z) {
                const double den = colA[z * 32 + lane] - colB[z * 32 + lane] * cp_prev;
                const double inv = frcp(den);
                cp_prev = colC[z * 32 + lane] * inv;
                colA[z * 32 + lane] = den;
                colB[z * 32 + lane] = inv;
                colC[z * 32 + lane] = cp_prev;
            }
        }
        __syncthreads();
        if (act) {
            const double den = colA[w * 32 + lane], inv = colB[w * 32 + lane], cp = colC[w * 32 + lane];
            store_row(s, (size_t)w * m.Tp + p, LTp, rc, den, inv, cp);
            if (w == 0) {
                s.rhs0[p] = rc.rhs; s.rhsS0[p] = rc.rhs * inv;
                s.live[p] = rc.rhs != 0.0 ? 1 : 0;  // seed of the line solver's active set
                nseed += rc.rhs != 0.0 ? 1u : 0u;
            }
        }
        // no barrier: in the next iteration warp w writes only row w of colA..C before the next barrier
    }
    count_seeds(sc, nseed);
    double o0, o1;
    if (grid_fold<2>(mx, ss, 1, 0, partial, pstride, &sc->ticket[0], o0, o1)) {
        if (threadIdx.x == 0) { red[0] = o0; red[1] = o1; }
    }
}

// Kernel 2, column-walking variant (one thread per face column): any nLayer, and the cross-check of the tile kernel in the
// tests (PBSM3D_ASSEMBLY=column).  Same arithmetic, same streams.
template <int MINB, int UNR>
__global__ void __launch_bounds__(128, MINB) assemble_kernel(DevConfig c, DevMesh m, FaceRecs R, SuspSystem s, int i0, int i1,
                                                             double* __restrict__ partial, int pstride, Scalars* sc,
                                                             double* __restrict__ red) {
    double mx = 0.0, ss = 0.0;
    unsigned nseed = 0;
    const int ntiles = (i1 - i0 + 127) / 128;
    const int L = c.L;
    const size_t LTp = (size_t)L * m.Tp;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = i0 + tile * 128 + threadIdx.x;  // a slot
        if (i < i1 && (R.i[i] & 16)) {
            const FaceConsts fc = load_rec(R, i);
            const int p = fc.p;
            double cp_prev = 0.0;
#pragma unroll UNR
            for (int z = 0; z < L; ++z) {
                const LayerConsts lc = layer_lookup(c, fc, s.ltab, z);
                double u_z, csubl;
                const RowCoef64 rc = assemble_row(c, fc, lc, z, L, u_z, csubl);
                const size_t r = (size_t)z * m.Tp + p;
                s.u_z[r] = u_z;
                s.csubl[r] = csubl;
                const double den = rc.d - rc.lo * cp_prev;
                const double inv = frcp(den);
                cp_prev = rc.up * inv;
                store_row(s, r, LTp, rc, den, inv, cp_prev);
                if (z == 0) {
                    s.rhs0[p] = rc.rhs;
                    s.rhsS0[p] = rc.rhs * inv;
                    s.live[p] = rc.rhs != 0.0 ? 1 : 0;  // seed of the line solver's active set
                    nseed += rc.rhs != 0.0 ? 1u : 0u;
                    mx = fmax(mx, fabs(rc.rhs));
                    ss += rc.rhs * rc.rhs;
                }
            }
        }
    }
    count_seeds(sc, nseed);
    double o0, o1;
    if (grid_fold<2>(mx, ss, 1, 0, partial, pstride, &sc->ticket[0], o0, o1)) {
        if (threadIdx.x == 0) { red[0] = o0; red[1] = o1; }
    }
}

// The reference's own coefficients from what is stored (inspection: pbsm3d_get_suspension_system).
__global__ void reconstruct_rows_kernel(SuspSystem s, int Tp, int L, double* __restrict__ diag, double* __restrict__ lat,
                                        double* __restrict__ below, double* __restrict__ above) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp) return;
    const size_t LTp = (size_t)L * Tp;
    double cp_prev = 0.0;
    for (int z = 0; z < L; ++z) {
        const size_t r = (size_t)z * Tp + p;
        const double den = s.den[r], bS = s.belowS[r], cp = s.cp[r];
        if (diag) diag[r] = den * (1.0 + bS * cp_prev);
        if (below) below[r] = bS * den;
        if (above) above[r] = cp * den;
        if (lat)
            for (int j = 0; j < 3; ++j) lat[j * LTp + r] = s.latS[j * LTp + r] * den;
        cp_prev = cp;
    }
}

// -------------------------------------------------------------------------------------- step control
// One-thread bookkeeping between phases.  `red` holds the (already globally reduced) values of the kernel before.
enum { FLAGS_SUSP = 0, FLAGS_SUSP_CHECK = 1, FLAGS_DEP = 2, FLAGS_FORCE_SUSP_OK = 3, FLAGS_CHEB_CHECK = 4, FLAGS_SETUP_CG = 5,
       FLAGS_DEP_RESTART = 6, FLAGS_COMBINE = 7, FLAGS_SOR_CHECK = 8 };

// Chebyshev stopping rule: iteration k measured ||b - A q_k||^2 of the iterate it READ (buffer k&1), which stays
// intact in that buffer, so detection keeps exactly that iterate.
__device__ __forceinline__ void cheb_check(Scalars* sc, double rr, int k, double tol2) {
    sc->rr = rr;
    if (rr <= tol2 * sc->bnorm2) { sc->done = 1; sc->dep_ok = 1; sc->iters = k; sc->dep_buf = k & 1; }
    else if (!(rr == rr) || rr > 1e60 * sc->bnorm2) sc->done = 2;  // diverging: spectrum bounds were wrong
}

// SOR keeps its iterate in one buffer (qA)
__device__ __forceinline__ void sor_check(Scalars* sc, double rr, int k, double tol2) {
    sc->rr = rr;
    if (rr <= tol2 * sc->bnorm2) { sc->done = 1; sc->dep_ok = 1; sc->iters = k; sc->dep_buf = 0; }
    else if (!(rr == rr) || rr > 1e60 * sc->bnorm2) sc->done = 2;
}

__device__ __forceinline__ void susp_check(Scalars* sc, double rr, int it_now, double tol2) {
    sc->susp_rr = rr;
    const int k = sc->n_checks;
    if (k < 16) { sc->rr_hist[k] = rr; sc->it_hist[k] = it_now; }
    sc->n_checks = k + 1;
    if (rr <= tol2 * sc->susp_bnorm2) { sc->susp_done = 1; sc->susp_iters = it_now; sc->susp_ok = 1; }
}

__global__ void flags_kernel(int stage, Scalars* sc, const double* __restrict__ red, int it_now, double tol2) {
    switch (stage) {
        case FLAGS_SUSP: {  // after assembly: suspension_present = ||b||_inf > 1e-12 (PBSM3D.cpp:1424-1427)
            // single rank: red holds it_now per-chunk pairs {max|b|, sum b^2}; multi rank: one reduced pair
            double mx = 0.0, ss = 0.0;
            for (int k = 0; k < it_now; ++k) { mx = fmax(mx, red[2 * k]); ss += red[2 * k + 1]; }
            sc->susp_rhs_max = mx;
            sc->susp_bnorm2 = ss;
            const int present = mx > 1e-12;
            sc->susp_present = present;
            sc->susp_done = present ? 0 : 1;  // nothing to solve: the solution stays the zero vector (:1461-1465)
            sc->susp_ok = present ? 0 : 1;
            sc->susp_iters = 0;
            sc->susp_rr = 0.0;
            sc->n_checks = 0;
            sc->susp_sweeps = 0; sc->susp_stalled = 0; sc->dep_sweeps = 0;
            sc->col_updates[0] = sc->col_updates[1] = sc->col_updates[2] = sc->col_updates[3] = 0ull;
            sc->n_seeds_step = sc->n_seeds;
            sc->dep_present = 0; sc->dep_ok = 0; sc->tail_done = 0; sc->drift_done = 0;
            sc->done = 0; sc->iters = 0; sc->dep_rhs_max = 0.0; sc->rr = 0.0; sc->bnorm2 = 0.0;
        } break;
        case FLAGS_SUSP_CHECK:
            if (!sc->susp_done) susp_check(sc, red[0], it_now, tol2);
            break;
        case FLAGS_COMBINE: {  // fold it_now per-chunk pairs {max, sum} into red[0..1] (ahead of the all-reduce)
            double mx = 0.0, ss = 0.0;
            for (int k = 0; k < it_now; ++k) { mx = fmax(mx, red[2 * k]); ss += red[2 * k + 1]; }
            double* w = const_cast<double*>(red);
            w[0] = mx; w[1] = ss;
        } break;
        case FLAGS_DEP:  // deposition solve iff suspension_present && ||rhs||_inf > 1e-12 (PBSM3D.cpp:1661-1664)
            if (!sc->susp_ok || sc->tail_done) break;
            sc->dep_rhs_max = red[0];
            sc->dep_present = (sc->susp_present && red[0] > 1e-12) ? 1 : 0;
            sc->bnorm2 = red[1];
            sc->rr = red[1];
            sc->done = 0; sc->iters = 0; sc->dep_buf = 0;
            sc->tail_done = 1;
            break;
        case FLAGS_CHEB_CHECK:
            if (sc->tail_done && sc->dep_present && !sc->done) cheb_check(sc, red[0], it_now, tol2);
            break;
        case FLAGS_SOR_CHECK:
            if (sc->tail_done && sc->dep_present && !sc->done) sor_check(sc, red[0], it_now, tol2);
            break;
        case FLAGS_SETUP_CG:  // pbsm3d_create: open the CG kernels for the spectrum estimate
            sc->tail_done = 1; sc->dep_present = 1; sc->done = 0; sc->dep_ok = 0; sc->iters = 0; sc->log_n = 0;
            break;
        case FLAGS_DEP_RESTART:  // Chebyshev gave up: hand the same right-hand side to CG
            sc->done = 0; sc->dep_ok = 0; sc->iters = 0; sc->dep_buf = 0;
            break;
        case FLAGS_FORCE_SUSP_OK:  // the host-driven Krylov path converged
            sc->susp_ok = 1; sc->susp_done = 1;
            break;
    }
}

// ---------------------------------------------------------------------------------- line relaxation
// HOT LOOP 2: one colour pass of the multicolour line Gauss–Seidel iteration
//     x_c  <-  T_c^{-1} (b_c - A_lat[c,:] x)          (in place; T = vertical tridiagonal blocks, factored at assembly)
// Faces of one colour are never edge-neighbours, so a pass only reads columns of other colours (or ghosts) and
// writes its own: no races, no second x array, and the result is independent of scheduling.
// One thread per face column.  Every coefficient stream (3 scaled lateral, scaled sub-diagonal, cp) is read
// exactly once, coalesced, with the evict-first hint so the streams do not push x out of the 126 MB L2; the
// 3·L gathers x[z*S + nbs] are branch-free and hit L1/L2.  All loads of a column are independent of the
// Thomas recurrence, so they are issued up front and the dependent chain runs on registers.
// LT > 0: compile-time layer count; LT == 0: any L (the own column of x is the scratch for the forward pass).
// The sweep streams in fp64 or rounded to fp32 (CT).  x, the right-hand side and all arithmetic stay fp64: the fp32 streams
// only perturb the operator by <= 6e-8 relative per coefficient, so the iteration they drive has its fixed point within a
// few 1e-7 of the true solution -- good for every sweep until the residual is down to ~1e-6; the last sweeps, and every
// residual check, use the fp64 coefficients.  40 -> 20 B of the 58 B a sweep moves per row.
template <typename CT> struct RowCoef { CT l0, l1, l2, bl; };
template <typename CT> __device__ __forceinline__ RowCoef<CT> load_row(const SuspSystem& s, size_t r, size_t LTp);
template <> __device__ __forceinline__ RowCoef<double> load_row<double>(const SuspSystem& s, size_t r, size_t LTp) {
    return {__ldcs(s.latS + r), __ldcs(s.latS + LTp + r), __ldcs(s.latS + 2 * LTp + r), __ldcs(s.belowS + r)};
}
template <> __device__ __forceinline__ RowCoef<float> load_row<float>(const SuspSystem& s, size_t r, size_t) {
    const float4 v = __ldcs(s.pack32 + r);  // one 16-byte load (512 B per warp) instead of four 4-byte ones
    return {v.x, v.y, v.z, v.w};
}
template <typename CT> __device__ __forceinline__ CT load_cp(const SuspSystem& s, size_t r);
template <> __device__ __forceinline__ double load_cp<double>(const SuspSystem& s, size_t r) { return __ldcs(s.cp + r); }
template <> __device__ __forceinline__ float load_cp<float>(const SuspSystem& s, size_t r) { return __ldcs(s.cp32 + r); }

// One face column of a colour pass: x_p <- T_p^{-1} (b_p - A_lat[p,:] x), in place.  XT is the STORAGE type of x: double, or
// float for the sweeps furthest from convergence (gs_persistent_kernel); the arithmetic is fp64 either way.
template <typename XT>
__device__ __forceinline__ double halo_gather(const XT* x, const double* ghost, size_t zS, size_t zG, int n, int Tp) {
    return n < Tp ? (double)x[zS + n] : ghost[zG + (n - Tp)];
}
// HALO: neighbour slots >= Tp are ghost faces, read from `ghost` [L][nGp] (fp64 whatever XT is); else every neighbour is in x.
// Returns whether the new column has a non-zero entry (used by the boundary columns of the active set; dead code elsewhere).
struct Nbs { int n0, n1, n2; };  // the three neighbour slots of a column
__device__ __forceinline__ Nbs load_nbs(const DevMesh& m, int p) {
    return {m.nbs[p], m.nbs[(size_t)m.Tp + p], m.nbs[(size_t)2 * m.Tp + p]};
}
template <int LT, typename CT, typename XT, bool HALO = false>
__device__ __forceinline__ bool gs_column(const SuspSystem& s, const DevMesh& m, int Lrt, int p, XT* x, const Nbs nb,
                                          const double* ghost = nullptr, int nGp = 0) {
    const int Tp = m.Tp, S = m.S;
    const int L = LT > 0 ? LT : Lrt;
    const int n0 = nb.n0, n1 = nb.n1, n2 = nb.n2;
    const size_t LTp = (size_t)L * Tp;
    if (LT > 0) {
        double g[LT > 0 ? LT : 1];
        CT bl[LT > 0 ? LT : 1], cu[LT > 0 ? LT : 1];
#pragma unroll
        for (int z = 0; z < LT; ++z) {
            const size_t r = (size_t)z * Tp + p, xr = (size_t)z * S, zG = (size_t)z * nGp;
            const RowCoef<CT> c = load_row<CT>(s, r, LTp);
            if (HALO)
                g[z] = -((double)c.l0 * halo_gather(x, ghost, xr, zG, n0, Tp) + (double)c.l1 * halo_gather(x, ghost, xr, zG, n1, Tp) +
                         (double)c.l2 * halo_gather(x, ghost, xr, zG, n2, Tp));
            else
                g[z] = -((double)c.l0 * (double)x[xr + n0] + (double)c.l1 * (double)x[xr + n1] + (double)c.l2 * (double)x[xr + n2]);
            bl[z] = c.bl;
        }
#pragma unroll
        for (int z = 0; z < LT; ++z) cu[z] = load_cp<CT>(s, (size_t)z * Tp + p);
        double y = g[0] + s.rhsS0[p];
        g[0] = y;
#pragma unroll
        for (int z = 1; z < LT; ++z) { y = g[z] - (double)bl[z] * y; g[z] = y; }
        x[(size_t)(LT - 1) * S + p] = (XT)y;
        bool nz = y != 0.0;
#pragma unroll
        for (int z = LT - 2; z >= 0; --z) { y = g[z] - (double)cu[z] * y; x[(size_t)z * S + p] = (XT)y; nz = nz || y != 0.0; }
        return nz;
    } else {
        // runtime layer count: the own column of x is the scratch of the forward pass (fp64 x only)
        double y = 0.0;
        for (int z = 0; z < L; ++z) {
            const size_t r = (size_t)z * Tp + p, xr = (size_t)z * S, zG = (size_t)z * nGp;
            const RowCoef<CT> c = load_row<CT>(s, r, LTp);
            double g;
            if (HALO)
                g = -((double)c.l0 * halo_gather(x, ghost, xr, zG, n0, Tp) + (double)c.l1 * halo_gather(x, ghost, xr, zG, n1, Tp) +
                      (double)c.l2 * halo_gather(x, ghost, xr, zG, n2, Tp));
            else
                g = -((double)c.l0 * (double)x[xr + n0] + (double)c.l1 * (double)x[xr + n1] + (double)c.l2 * (double)x[xr + n2]);
            if (z == 0) g += s.rhsS0[p];
            y = g - (double)c.bl * y;
            x[xr + p] = (XT)y;
        }
        bool nz = y != 0.0;
        for (int z = L - 2; z >= 0; --z) {
            y = (double)x[(size_t)z * S + p] - (double)load_cp<CT>(s, (size_t)z * Tp + p) * y;
            x[(size_t)z * S + p] = (XT)y;
            nz = nz || y != 0.0;
        }
        return nz;
    }
}

template <int LT, typename CT>
__global__ void __launch_bounds__(128, 4) gs_sweep_kernel(SuspSystem s, DevMesh m, int Lrt, int p0, int p1, double* x,
                                                          const Scalars* __restrict__ sc) {
    if (sc->susp_done) return;
    const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    gs_column<LT, CT, double>(s, m, Lrt, p, x, load_nbs(m, p));
}

__device__ __forceinline__ double spmv_row(const SuspSystem& s, const DevMesh& m, int L, const double* __restrict__ x, int z, int p);

// ------------------------------------------------------------------- persistent (cooperative) solver kernels
// On one rank a whole solve is ONE cooperative launch: the grid is sized to be co-resident (occupancy x SM count), every
// block walks its share of each colour class, and a grid-wide barrier (one atomic arrive + an acquire spin per block)
// separates the colour passes.  The stopping rule ||b - A x||_2 <= tol ||b||_2 is evaluated inside the kernel at the sweeps
// the host's schedule names (first check at the predicted sweep count, then every `check_every` sweeps): every block folds
// the same per-block partial sums in the same order, so all blocks take the same decision and leave the loop together.
// This replaces ~60 (suspension) and ~150 (deposition) launches per step, and the guarded no-op launches of a calm step.
struct ColourRanges {
    int n;
    int start[8], end[8];
};
struct SolvePlan {
    int nx32;         // leading sweeps that also keep the iterate in fp32 storage (suspension only; <= n32)
    int n32;          // leading sweeps that stream the fp32-rounded coefficient copies (suspension only)
    int check_first;  // first residual check after this many sweeps
    int check_every;
    int maxit;
    int use_live;     // suspension only: skip the columns outside the active set (SuspSystem::live) ...
    int live_max_seeds;  // ... when at most this many faces have a non-zero right-hand side (Scalars::n_seeds_step)
    int prefetch;     // suspension only: load a thread's next neighbour slots before it works on the current column (gs_pass)
    double tol2;
};

// ---- the active set of the line solver.
// Every solve starts from x0 = 0 and b is non-zero only on layer 0 of the saltating faces.  A column update is
// x_p <- T_p^-1 (b_p - A_lat[p,:] x): with b_p = 0 and three neighbour columns that are still identically zero it yields exactly
// zero, i.e. what x_p already holds, so it can be skipped without changing one bit of any iterate.  live[] (one byte per slot) is a
// superset of the columns whose right-hand side or iterate is non-zero: seeded by the assembly with b_p != 0, set when a column is
// updated while a neighbour is live (the set grows by one ring of faces per colour pass), never cleared within a solve.  A column
// is updated iff it or one of its neighbours is live; the flags it reads belong to itself and to faces of OTHER colours, which the
// current pass does not write, so there is no race and the set of updated columns does not depend on scheduling.  The residual
// check skips the same columns (their residual is exactly zero).  A skipped column costs its flag byte, or its three neighbour
// slots + their flags while it is not live, instead of 30-58 B per layer.  tests/models/active_set_model.py is the numpy statement.
// Ghost neighbours (slot >= Tp, boundary columns across ranks) count as live: those columns are always updated, and flagged live
// only when their new values are not all zero, so the always-updated set does not spread inwards from the partition edges.
// 0: outside the active set (skip); 1: live already; 2: not live yet, but a neighbour is (update it and flag it).
__device__ __forceinline__ int column_state(const unsigned char* live, const DevMesh& m, int p, const Nbs nb) {
    if (live[p]) return 1;
    const int Tp = m.Tp;
    if (nb.n0 >= Tp || nb.n1 >= Tp || nb.n2 >= Tp) return 2;
    return (live[nb.n0] | live[nb.n1] | live[nb.n2]) ? 2 : 0;
}
__device__ __forceinline__ int column_state(const unsigned char* live, const DevMesh& m, int p) {
    return column_state(live, m, p, load_nbs(m, p));
}
// Adds a warp's count of column updates to Scalars::col_updates[k] and clears it (called by whole warps, outside divergent code).
__device__ __forceinline__ void flush_col_count(Scalars* sc, int k, unsigned& cnt) {
    const unsigned tot = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(&sc->col_updates[k], (unsigned long long)tot);
    cnt = 0;
}
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// `ctr` is zeroed by the host before the launch; `target` is the block's private running count.
// grid_barrier: arrival is one `red.release.gpu` by thread 0 — the release is cumulative over the block's writes that the preceding
// bar.sync ordered before it — and departure an `ld.acquire.gpu` spin followed by bar.sync, which orders every later load of the
// block after the writes of all blocks that arrived.  No separate fences.  Measured against the fenced form below (c2, one B200,
// profiles/r2m_summary.md): the latency-bound solves (deposition SOR: 148 barriers around 4 us passes; snow_slide) gain 0.35 us per
// barrier; the bandwidth-bound suspension solve (60 barriers around 30 us passes that stream 200-400 MB each) LOSES 0.4 us per
// barrier, so it keeps the fenced form.
__device__ __forceinline__ void atom_add_release_gpu_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        atom_add_release_gpu_u32(ctr, 1u);
        while (ld_acquire_gpu_u32(ctr) < target) {}
    }
    __syncthreads();
}
__device__ __forceinline__ void grid_barrier_fenced(unsigned* ctr, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();  // release: every write of this block (ordered before thread 0 by the barrier above)
        atomicAdd(ctr, 1u);
        while (ld_acquire_gpu_u32(ctr) < target) {}
        __threadfence();
    }
    __syncthreads();
}
// Sum of the per-block partials in index order: identical bits in every block.
__device__ __forceinline__ double fold_partials(const double* partial, int n) {
    double a = 0.0;
    for (int b = threadIdx.x; b < n; b += blockDim.x) a += __ldcg(partial + b);
    __shared__ double bcast;
    a = block_sum(a);
    if (threadIdx.x == 0) bcast = a;
    __syncthreads();
    return bcast;
}

template <int LT>
__device__ __forceinline__ double residual_column(const SuspSystem& s, const DevMesh& m, const double* x, int p) {
    const int Tp = m.Tp, S = m.S;
    const int n0 = m.nbs[p], n1 = m.nbs[(size_t)Tp + p], n2 = m.nbs[(size_t)2 * Tp + p];
    double xo[LT], g[LT], a = 0.0;
#pragma unroll
    for (int z = 0; z < LT; ++z) xo[z] = x[(size_t)z * S + p];
#pragma unroll
    for (int z = 0; z < LT; ++z) {
        const size_t r = (size_t)z * Tp + p, xr = (size_t)z * S;
        g[z] = __ldcs(s.latS + r) * x[xr + n0] + __ldcs(s.latS + (size_t)LT * Tp + r) * x[xr + n1] +
               __ldcs(s.latS + (size_t)2 * LT * Tp + r) * x[xr + n2];
    }
    double cp_prev = 0.0;
#pragma unroll
    for (int z = 0; z < LT; ++z) {
        const size_t r = (size_t)z * Tp + p;
        const double bS = __ldcs(s.belowS + r), cp = __ldcs(s.cp + r);
        double v = g[z] + xo[z];
        if (z > 0) v += bS * (cp_prev * xo[z] + xo[z - 1]);
        if (z < LT - 1) v += cp * xo[z + 1];
        v = (((z == 0) ? s.rhsS0[p] : 0.0) - v) * __ldcs(s.den + r);
        a += v * v;
        cp_prev = cp;
    }
    return a;
}

// The suspension solve: multicolour line Gauss-Seidel sweeps + residual checks, one launch.
// Writes the same control-block fields as the per-pass path (susp_done / susp_iters / susp_rr / rr_hist) plus
// susp_stalled when the sweeps stop contracting (rate > 0.97 after 64 sweeps): the host then hands over to BiCGStab.
// Three phases of one solve (sweep numbers from the host's schedule, which carries iteration COUNTS only from the previous step):
//     [0, nx32)     x stored in fp32 (xf), fp32-rounded coefficient copies   30 B/row
//     [nx32, n32)   x in fp64, fp32-rounded coefficient copies               38 B/row
//     [n32, ...)    x in fp64, fp64 coefficients                             58 B/row, and every residual check
// fp32 storage of the iterate perturbs each sweep's result by <= 6e-8 relative; the iteration contracts (rho ~ 0.5), so the
// error it leaves when the iterate moves to fp64 (one conversion pass) is ~1e-7 ||x|| and the following fp64-x sweeps remove
// it at the usual rate (tests/models/fp32_x_model.py: the sweep count does not change when the switch is >= 10 sweeps before
// the end).  All arithmetic, the right-hand side and the stopping rule are fp64 throughout.
constexpr int kGsThreads = 512;
// A thread's share of one colour pass: columns p0, p0 + stride, ... < p1, those outside the active set skipped (live != nullptr).
template <int LT, typename CT, typename XT>
__device__ __forceinline__ void gs_pass(const SuspSystem& s, const DevMesh& m, int L, int p0, int p1, int stride, XT* x,
                                        unsigned char* live, unsigned& cnt, bool prefetch) {
    // The neighbour slots of a thread's NEXT column are loaded before it works on the current one: the gathers of a column then
    // start together with its coefficient loads instead of one round trip later.
    if (p0 >= p1) return;
    Nbs cur = load_nbs(m, p0);
    for (int p = p0; p < p1; p += stride) {
        const int pn = p + stride;
        Nbs nxt = cur;
        if (prefetch && pn < p1) nxt = load_nbs(m, pn);
        bool go = true;
        if (live) {
            const int st = column_state(live, m, p, cur);
            go = st != 0;
            if (st == 2) live[p] = 1;
        }
        if (go) {
            ++cnt;
            gs_column<LT, CT, XT>(s, m, L, p, x, cur);
        }
        if (!prefetch && pn < p1) nxt = load_nbs(m, pn);  // PBSM3D_NBS_PREFETCH=0: the slots are loaded when they are needed
        cur = nxt;
    }
}
template <int LT>
__global__ void __launch_bounds__(kGsThreads, 1) gs_persistent_kernel(SuspSystem s, DevMesh m, int Lrt, ColourRanges cr, double* x, float* xf,
                                                                      Scalars* sc, double* __restrict__ partial, SolvePlan pl, unsigned* bar) {
    if (sc->susp_done) return;  // uniform: nobody writes it before the first grid barrier
    const int L = LT > 0 ? LT : Lrt;
    const double bnorm2 = sc->susp_bnorm2;
    unsigned target = 0;
    int it = 0, n_checks = 0, converged = 0, stalled = 0;
    double rr = 0.0, prev_rr = bnorm2;
    int prev_it = 0;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int nx32 = (LT > 0 && xf) ? pl.nx32 : 0;
    unsigned char* const live = (pl.use_live && sc->n_seeds_step <= (unsigned)pl.live_max_seeds) ? s.live : nullptr;
    unsigned cnt = 0;
    if (nx32 > 0) {  // x0 = 0 in the fp32 copy (ghost tails included)
        const size_t NS = (size_t)L * m.S;
        for (size_t k = t0; k < NS; k += stride) xf[k] = 0.f;
        grid_barrier_fenced(bar, target);
    }
    while (it < pl.maxit) {
        const int phase = it < nx32 ? 0 : (it < pl.n32 ? 1 : 2);
        for (int c = 0; c < cr.n; ++c) {
            if (phase == 0) gs_pass<LT, float, float>(s, m, L, cr.start[c] + t0, cr.end[c], stride, xf, live, cnt, pl.prefetch != 0);
            else if (phase == 1) gs_pass<LT, float, double>(s, m, L, cr.start[c] + t0, cr.end[c], stride, x, live, cnt, pl.prefetch != 0);
            else gs_pass<LT, double, double>(s, m, L, cr.start[c] + t0, cr.end[c], stride, x, live, cnt, pl.prefetch != 0);
            grid_barrier_fenced(bar, target);
        }
        ++it;
        if (it == nx32) {  // the iterate moves to fp64
            const size_t NS = (size_t)L * m.S;
            for (size_t k = t0; k < NS; k += stride) x[k] = (double)xf[k];
            grid_barrier_fenced(bar, target);
        }
        const bool check = (it >= pl.check_first && (it - pl.check_first) % pl.check_every == 0) || it >= pl.maxit;
        if (check || it == nx32 || it == pl.n32) flush_col_count(sc, phase, cnt);
        if (!check) continue;
        double a = 0.0;
        if (LT > 0) {
            for (int p = t0; p < m.Tp; p += stride) {
                if (live && !column_state(live, m, p)) continue;  // b_p = 0, x = 0 on p and its neighbours: residual exactly 0
                ++cnt;
                a += residual_column<(LT > 0 ? LT : 1)>(s, m, x, p);
            }
        } else {
            for (int p = t0; p < m.Tp; p += stride) {
                if (live && !column_state(live, m, p)) continue;
                ++cnt;
                for (int z = 0; z < L; ++z) {
                    const double v = ((z == 0) ? s.rhs0[p] : 0.0) - spmv_row(s, m, L, x, z, p);
                    a += v * v;
                }
            }
        }
        flush_col_count(sc, 3, cnt);
        a = block_sum(a);
        if (threadIdx.x == 0) partial[blockIdx.x] = a;
        grid_barrier_fenced(bar, target);
        rr = fold_partials(partial, gridDim.x);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (n_checks < 16) { sc->rr_hist[n_checks] = rr; sc->it_hist[n_checks] = it; }
        }
        ++n_checks;
        if (rr <= pl.tol2 * bnorm2) { converged = 1; break; }
        if (!(rr == rr) || rr > 1e60 * bnorm2) { stalled = 1; break; }
        if (it >= 64 && it > prev_it && rr > 0.0 && prev_rr > 0.0) {  // contraction per sweep of ||r|| worse than 0.97: crawling
            const double lim = exp(2.0 * (it - prev_it) * log(0.97));
            if (rr > lim * prev_rr) { stalled = 1; break; }
        }
        prev_rr = rr;
        prev_it = it;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->susp_rr = rr;
        sc->n_checks = n_checks;
        sc->susp_sweeps = it;
        sc->susp_stalled = stalled;
        if (converged) { sc->susp_done = 1; sc->susp_iters = it; sc->susp_ok = 1; }
    }
}

// ------------------------------------------------------------------------------------ SpMV / residual
// Row of A·x in the extruded-ELL layout: lateral gathers x[z*S + nbs_j], vertical x[(z±1)*S + p].  With the stored form
// (A x)_z = den_z [ (1 + belowS_z cp_{z-1}) x_z + belowS_z x_{z-1} + sum_j latS_j x_nb_j + cp_z x_{z+1} ].
__device__ __forceinline__ double spmv_row(const SuspSystem& s, const DevMesh& m, int L, const double* __restrict__ x, int z, int p) {
    const int Tp = m.Tp, S = m.S;
    const size_t r = (size_t)z * Tp + p, xr = (size_t)z * S;
    const double bS = s.belowS[r];
    double acc = x[xr + p];
    if (z > 0) acc += bS * (s.cp[r - Tp] * x[xr + p] + x[xr - S + p]);
#pragma unroll
    for (int j = 0; j < 3; ++j) acc += s.latS[((size_t)j * L + z) * Tp + p] * x[xr + m.nbs[(size_t)j * Tp + p]];
    if (z < L - 1) acc += s.cp[r] * x[xr + S + p];
    return acc * s.den[r];
}

// True residual of the line solver: ||b - A x||_2^2 of this rank into red[0]; with `fused` (single rank) the
// last block also applies the stopping rule ||b-Ax|| <= tol ||b|| (LinearAlgebra.cpp:168, x0 = 0).
__global__ void __launch_bounds__(kRedThreads) residual_kernel(SuspSystem s, DevMesh m, int L, const double* __restrict__ x,
                                                               double* __restrict__ partial, int pstride, Scalars* sc,
                                                               double* __restrict__ red, int it_now, double tol2, int fused) {
    if (sc->susp_done) return;
    const int Tp = m.Tp;
    const size_t N = (size_t)L * Tp;
    double a = 0.0;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(r / Tp), p = (int)(r - (size_t)z * Tp);
        const double v = ((z == 0) ? s.rhs0[p] : 0.0) - spmv_row(s, m, L, x, z, p);
        a += v * v;
    }
    double rr, unused;
    if (grid_fold<1>(a, 0.0, 0, 0, partial, pstride, &sc->ticket[1], rr, unused)) {
        if (threadIdx.x == 0) {
            red[0] = rr;
            if (fused) susp_check(sc, rr, it_now, tol2);
        }
    }
}

// Same quantity, column-structured like the sweep (one thread per face column, every load issued before the
// arithmetic, compile-time layer count): the row-wise kernel above re-reads the neighbour slots per row and
// divides per row, and reaches only about half the bandwidth.
template <int LT>
__global__ void __launch_bounds__(128) residual_col_kernel(SuspSystem s, DevMesh m, const double* __restrict__ x,
                                                           double* __restrict__ partial, int pstride, Scalars* sc,
                                                           double* __restrict__ red, int it_now, double tol2, int fused) {
    if (sc->susp_done) return;
    const int Tp = m.Tp, S = m.S;
    double a = 0.0;
    const int ntiles = (Tp + 127) / 128;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p = tile * 128 + threadIdx.x;
        if (p >= Tp) continue;
        const int n0 = m.nbs[p], n1 = m.nbs[(size_t)Tp + p], n2 = m.nbs[(size_t)2 * Tp + p];
        double xo[LT], g[LT];
#pragma unroll
        for (int z = 0; z < LT; ++z) xo[z] = x[(size_t)z * S + p];
#pragma unroll
        for (int z = 0; z < LT; ++z) {
            const size_t r = (size_t)z * Tp + p, xr = (size_t)z * S;
            g[z] = __ldcs(s.latS + r) * x[xr + n0] + __ldcs(s.latS + (size_t)LT * Tp + r) * x[xr + n1] +
                   __ldcs(s.latS + (size_t)2 * LT * Tp + r) * x[xr + n2];
        }
        double cp_prev = 0.0;
#pragma unroll
        for (int z = 0; z < LT; ++z) {
            const size_t r = (size_t)z * Tp + p;
            const double bS = __ldcs(s.belowS + r), cp = __ldcs(s.cp + r);
            double v = g[z] + xo[z];
            if (z > 0) v += bS * (cp_prev * xo[z] + xo[z - 1]);
            if (z < LT - 1) v += cp * xo[z + 1];
            v = (((z == 0) ? s.rhsS0[p] : 0.0) - v) * __ldcs(s.den + r);
            a += v * v;
            cp_prev = cp;
        }
    }
    double rr, unused;
    if (grid_fold<1>(a, 0.0, 0, 0, partial, pstride, &sc->ticket[1], rr, unused)) {
        if (threadIdx.x == 0) {
            red[0] = rr;
            if (fused) susp_check(sc, rr, it_now, tol2);
        }
    }
}

// y = A x on ghost-extended vectors (Krylov path), optionally with <y,d0> and <y,d1|y> partials.
__global__ void __launch_bounds__(kRedThreads) spmv_kernel(SuspSystem s, DevMesh m, int L, const double* __restrict__ x,
                                                           double* __restrict__ y, const double* __restrict__ d0,
                                                           const double* __restrict__ d1, int self_dot,
                                                           double* __restrict__ partial, int stride, const int* __restrict__ done) {
    if (done && *done) return;
    const int Tp = m.Tp, S = m.S;
    const size_t N = (size_t)L * Tp;
    double a0 = 0.0, a1 = 0.0;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(r / Tp), p = (int)(r - (size_t)z * Tp);
        const size_t e = (size_t)z * S + p;
        const double v = spmv_row(s, m, L, x, z, p);
        if (y) y[e] = v;
        if (d0) a0 += v * d0[e];
        if (d1) a1 += v * d1[e];
        else if (self_dot) a1 += v * v;
    }
    if (partial) {
        if (d0) { a0 = block_sum(a0); if (threadIdx.x == 0) partial[blockIdx.x] = a0; }
        if (d1 || self_dot) { a1 = block_sum(a1); if (threadIdx.x == 0) partial[stride + blockIdx.x] = a1; }
    }
}

// Final stage of the unfused reductions (Krylov path): one block folds `nvals` partial arrays into out[v].
__global__ void __launch_bounds__(256) fold_kernel(int nblocks, int nvals, int stride, const double* __restrict__ partial,
                                                   double* __restrict__ out, int op) {
    for (int v = 0; v < nvals; ++v) {
        double a = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
            double q = partial[(size_t)v * stride + b];
            a = op ? fmax(a, q) : a + q;
        }
        a = op ? block_max(a) : block_sum(a);
        if (threadIdx.x == 0) out[v] = a;
    }
}

// Column-tridiagonal preconditioner apply y = T^{-1} v (right preconditioner of the Krylov path).
__global__ void __launch_bounds__(128) thomas_kernel(SuspSystem s, int Tp, int S, int L, const double* __restrict__ v,
                                                     double* __restrict__ y, const int* __restrict__ done) {
    if (done && *done) return;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp) return;
    double prev = 0.0;
    for (int z = 0; z < L; ++z) {
        const size_t r = (size_t)z * Tp + p;
        prev = v[(size_t)z * S + p] / s.den[r] - s.belowS[r] * prev;
        y[(size_t)z * S + p] = prev;
    }
    double xn = prev;
    for (int z = L - 2; z >= 0; --z) {
        xn = y[(size_t)z * S + p] - s.cp[(size_t)z * Tp + p] * xn;
        y[(size_t)z * S + p] = xn;
    }
}

// ------------------------------------------------------------------------------------------ BiCGStab
// Right-preconditioned BiCGStab (fallback / cross-check solver), recurrence scalars on the device.  Vectors are
// ghost-extended [L][S]; r, rhat, p, v, t keep zero ghost tails, so flat dot products over L*S are exact.
__global__ void __launch_bounds__(256) bicg_init_kernel(int Tp, int S, int L, const double* __restrict__ rhs0, double* __restrict__ x,
                                                        double* __restrict__ r, double* __restrict__ rhat, double* __restrict__ p,
                                                        double* __restrict__ v, double* __restrict__ partial) {
    const size_t N = (size_t)L * S;
    double a = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (size_t)gridDim.x * blockDim.x) {
        double b = (k < (size_t)Tp) ? rhs0[k] : 0.0;
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

// Scalar updates (one thread).  `red` holds the globally reduced dot products of the preceding kernel.
// stage 0: after init            red[0] = ||b||^2
// stage 1: after v = A ph        red[0] = <rhat, v>           -> alpha
// stage 2: after s               red[0] = ||s||^2             -> half-step residual
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
            sc->tmp[0] = red[0];
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
__global__ void __launch_bounds__(256) flux_kernel(int Tp, int S, int L, double dz, double dt, const int* __restrict__ perm,
                                                   const double* __restrict__ x, const double* __restrict__ u_z,
                                                   const double* __restrict__ csubl, double* __restrict__ Qsusp,
                                                   double* __restrict__ Qsubl, double* __restrict__ Qsubl_mass,
                                                   double* __restrict__ sum_subl, const Scalars* __restrict__ sc) {
    if (!sc->susp_ok || sc->tail_done) return;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp) return;
    double qs = 0.0, ql = 0.0;
    for (int z = 0; z < L; ++z) {
        const size_t r = (size_t)z * Tp + p;
        double c = x[(size_t)z * S + p];
        c = (c < 0 || chm_is_nan(c)) ? 0.0 : c;
        qs += c * u_z[r] * dz;
        ql += csubl[r] * c * dz;
    }
    Qsusp[p] = qs;
    Qsubl[p] = ql;
    const double qm = ql * dt;
    Qsubl_mass[p] = qm;
    sum_subl[p] += qm;
}

// --------------------------------------------------------------------------------------- deposition
// RHS of the deposition system with the upwind donor rule (PBSM3D.cpp:1523-1656); the matrix is static.
// Qsusp/Qsalt are ghost-extended [S] (tails filled by the halo exchange).  red[0] = max|rhs| of this rank.
__global__ void __launch_bounds__(256) deposition_rhs_kernel(DevMesh m, const double* __restrict__ vw_dir,
                                                             const double* __restrict__ Qsusp, const double* __restrict__ Qsalt,
                                                             const double* __restrict__ dinv, double* __restrict__ rhs,
                                                             double* __restrict__ rhsS, double* __restrict__ partial, int pstride,
                                                             Scalars* sc, double* __restrict__ red) {
    if (!sc->susp_ok || sc->tail_done) return;
    const int Tp = m.Tp;
    double val_abs = 0.0, sumsq = 0.0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < Tp; p += gridDim.x * blockDim.x) {
        const int i = m.perm[p];
        double acc = 0.0;
        if (i >= 0) {
            double vx, vy;
            wind_unit_vector(vw_dir[i], vx, vy);
            const double own_t = Qsusp[p], own_s = Qsalt[p];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double udotm = vx * m.nx[(size_t)j * Tp + p] + vy * m.ny[(size_t)j * Tp + p];
                const double E = m.elen[(size_t)j * Tp + p];
                const int n = m.nbs[(size_t)j * Tp + p];
                double Qt = own_t, Qs = own_s;
                if (!(udotm > 0) && n != p) {
                    Qt = Qsusp[n];
                    Qs = Qsalt[n];
                    if (chm_is_nan(Qs)) Qs = 0.0;
                }
                acc += -E * (Qt + Qs) * udotm;
            }
        }
        rhs[p] = acc;
        rhsS[p] = acc * dinv[p];
        val_abs = fmax(val_abs, fabs(acc));
        sumsq += acc * acc;
    }
    double mx, ss;
    if (grid_fold<2>(val_abs, sumsq, 1, 0, partial, pstride, &sc->ticket[2], mx, ss)) {
        if (threadIdx.x == 0) { red[0] = mx; red[1] = ss; }
    }
}

// Jacobi-scaled off-diagonals of the (static) deposition matrix: offS_j = off_j / diag.
__global__ void deposition_scale_kernel(int Tp, const double* __restrict__ doff, const double* __restrict__ dinv,
                                        double* __restrict__ offS) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp) return;
#pragma unroll
    for (int j = 0; j < 3; ++j) offS[(size_t)j * Tp + p] = doff[(size_t)j * Tp + p] * dinv[p];
}
// Deterministic pseudo-random right-hand side for the setup-time spectrum estimate (depends on the global face id
// only, so every partitioning of a mesh probes the same vector).
__global__ void probe_rhs_kernel(int Tp, const int* __restrict__ perm, long long gid0, double* __restrict__ b) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp) return;
    const int i = perm[p];
    if (i < 0) { b[p] = 0.0; return; }
    unsigned long long h = (unsigned long long)(gid0 + i) * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
    b[p] = (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// HOT LOOP 3: one iteration of the Jacobi-preconditioned Chebyshev iteration on the deposition system, in its
// three-term form
//     z = D^{-1}(b - A q_k),   q_{k+1} = q_k + a_k (q_k - q_{k-1}) + c_k z
// (a_k, c_k from the spectrum bounds of D^{-1}A estimated once in pbsm3d_create: the matrix is static; a_0 = 0).  One
// launch per iteration, no dot products, no global reduction.  q ping-pongs between two ghost-extended buffers:
// iteration k reads q_k from one, and OVERWRITES q_{k-1} with q_{k+1} in the other (each thread touches only its own
// element of it), so no separate direction vector is streamed: per face 3 scaled off-diagonals + 3 neighbour slots +
// scaled rhs + q_k (+ gathers) + q_{k-1} in, q_{k+1} out = 68 B.  The iterate an iteration read is still intact when
// a CHECK iteration finds it converged.
template <int CHECK>
__global__ void __launch_bounds__(kRedThreads) cheb_iter_kernel(DevMesh m, const double* __restrict__ offS,
                                                                const double* __restrict__ bS, const double* __restrict__ ddiag,
                                                                const double* __restrict__ qin, double* __restrict__ qout, double ak,
                                                                double ck, int k, double* __restrict__ partial, int pstride,
                                                                Scalars* sc, double* __restrict__ red, double tol2, int fused) {
    if (sc && (!sc->tail_done || !sc->dep_present || sc->done)) return;  // sc == null: stand-alone timing launch
    const int Tp = m.Tp;
    double rr = 0.0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < Tp; p += gridDim.x * blockDim.x) {
        const double qp = qin[p];
        double z = bS[p] - qp;
#pragma unroll
        for (int j = 0; j < 3; ++j) z -= offS[(size_t)j * Tp + p] * qin[m.nbs[(size_t)j * Tp + p]];
        if (CHECK) { const double r = z * ddiag[p]; rr += r * r; }
        qout[p] = qp + (ak * (qp - qout[p]) + ck * z);
    }
    if (CHECK && sc) {
        double o0, unused;
        if (grid_fold<1>(rr, 0.0, 0, 0, partial, pstride, &sc->ticket[3], o0, unused)) {
            if (threadIdx.x == 0) { red[0] = o0; if (fused) cheb_check(sc, o0, k, tol2); }
        }
    }
}

// HOT LOOP 3b: multicolour SOR on the Jacobi-scaled deposition system,  q_c <- q_c + w (D^-1 b - q - D^-1 A_off q)_c
// for one colour class c, IN PLACE (faces of one colour are never neighbours).  With Young's w = 2 / (1 + sqrt(1 - rho^2)),
// rho = 1 - lambda_min(D^-1 A) from the setup-time spectrum estimate, a sweep over all colours converges about twice as
// fast as a Chebyshev-accelerated Jacobi iteration (measured 74 vs 125 sweeps on the uniform mesh, 124 vs 198 on the
// variable-resolution one) and streams less: 3 scaled off-diagonals + 3 slots + scaled rhs + q in/out = 60 B per face.
__global__ void __launch_bounds__(256) sor_pass_kernel(DevMesh m, const double* __restrict__ offS, const double* __restrict__ bS,
                                                       double* q, double omega, int p0, int p1, const Scalars* __restrict__ sc) {
    if (sc && (!sc->tail_done || !sc->dep_present || sc->done)) return;
    const int Tp = m.Tp;
    const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    const double qp = q[p];
    double z = bS[p] - qp;
#pragma unroll
    for (int j = 0; j < 3; ++j) z -= __ldcs(offS + (size_t)j * Tp + p) * q[m.nbs[(size_t)j * Tp + p]];
    q[p] = qp + omega * z;
}
// ||b - A q||^2 of this rank (the stopping rule of the deposition solve) for an iterate that lives in one buffer.
__global__ void __launch_bounds__(kRedThreads) dep_residual_kernel(DevMesh m, const double* __restrict__ offS,
                                                                   const double* __restrict__ bS, const double* __restrict__ ddiag,
                                                                   const double* __restrict__ q, int k, double* __restrict__ partial,
                                                                   int pstride, Scalars* sc, double* __restrict__ red, double tol2,
                                                                   int fused) {
    if (!sc->tail_done || !sc->dep_present || sc->done) return;
    const int Tp = m.Tp;
    double rr = 0.0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < Tp; p += gridDim.x * blockDim.x) {
        double z = bS[p] - q[p];
#pragma unroll
        for (int j = 0; j < 3; ++j) z -= offS[(size_t)j * Tp + p] * q[m.nbs[(size_t)j * Tp + p]];
        const double r = z * ddiag[p];
        rr += r * r;
    }
    double o0, unused;
    if (grid_fold<1>(rr, 0.0, 0, 0, partial, pstride, &sc->ticket[3], o0, unused)) {
        if (threadIdx.x == 0) { red[0] = o0; if (fused) sor_check(sc, o0, k, tol2); }
    }
}

// The deposition solve as one cooperative launch: multicolour SOR sweeps with a grid barrier between the colour passes and the
// stopping rule ||b - A q|| <= tol ||b|| evaluated in the kernel (see gs_persistent_kernel).  STREAM: the working set does not fit
// the L2, coefficient streams are read with the evict-first hint; otherwise they stay resident in the 126 MB L2 across the sweeps.
template <bool STREAM, int NT, int B>  // NT threads per block (one block per SM), B faces per thread in flight
__global__ void __launch_bounds__(NT, 1) sor_persistent_kernel(DevMesh m, const double* __restrict__ offS, const double* __restrict__ bS,
                                                                        const double* __restrict__ ddiag, double* q, double omega,
                                                                        ColourRanges cr, Scalars* sc, double* __restrict__ partial,
                                                                        SolvePlan pl, unsigned* bar) {
    if (!sc->tail_done || !sc->dep_present || sc->done) return;  // uniform: written only after the last grid barrier
    const int Tp = m.Tp;
    const double bnorm2 = sc->bnorm2;
    unsigned target = 0;
    int it = 0, done = 0;
    double rr = bnorm2;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    // a pass is latency-bound (two dependent L2 round trips per face): B independent faces per thread are in flight at once
    while (it < pl.maxit) {
        for (int c = 0; c < cr.n; ++c) {
            const int end = cr.end[c];
            for (int base = cr.start[c] + t0; base < end; base += B * stride) {
                double qp[B], z[B], o[B][3];
                int n[B][3];
#pragma unroll
                for (int k = 0; k < B; ++k) {
                    const int p = base + k * stride;
                    if (p < end) {
                        qp[k] = q[p];
                        z[k] = bS[p] - qp[k];
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            n[k][j] = m.nbs[(size_t)j * Tp + p];
                            o[k][j] = STREAM ? __ldcs(offS + (size_t)j * Tp + p) : offS[(size_t)j * Tp + p];
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < B; ++k)
                    if (base + k * stride < end) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) z[k] -= o[k][j] * q[n[k][j]];
                    }
#pragma unroll
                for (int k = 0; k < B; ++k)
                    if (base + k * stride < end) q[base + k * stride] = qp[k] + omega * z[k];
            }
            grid_barrier(bar, target);
        }
        ++it;
        const bool check = (it >= pl.check_first && (it - pl.check_first) % pl.check_every == 0) || it >= pl.maxit;
        if (!check) continue;
        double a = 0.0;
        for (int p = t0; p < Tp; p += stride) {
            double z = bS[p] - q[p];
#pragma unroll
            for (int j = 0; j < 3; ++j) z -= offS[(size_t)j * Tp + p] * q[m.nbs[(size_t)j * Tp + p]];
            const double r = z * ddiag[p];
            a += r * r;
        }
        a = block_sum(a);
        if (threadIdx.x == 0) partial[blockIdx.x] = a;
        grid_barrier(bar, target);
        rr = fold_partials(partial, gridDim.x);
        if (rr <= pl.tol2 * bnorm2) { done = 1; break; }
        if (!(rr == rr) || rr > 1e60 * bnorm2) { done = 2; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->rr = rr;
        sc->dep_sweeps = it;
        sc->iters = it;
        if (done == 1) { sc->done = 1; sc->dep_ok = 1; sc->dep_buf = 0; }
        else if (done == 2) sc->done = 2;
    }
}

// The same solve with the STATIC part of every face resident on chip for the whole launch.  A persistent thread updates the same
// faces in every sweep, so their neighbour slots and two of the three scaled off-diagonals live in shared memory (28 B per face,
// laid out [colour][k][field][thread]: conflict-free), the third off-diagonal in registers; a colour pass then costs ONE L2 round
// trip (the q gathers + bS, all issued at once for the thread's K faces of that colour) plus the grid barrier, instead of two
// dependent round trips (slots, then gathers) over 84 B per face.  Two colour classes (every structured split mesh), at most K faces
// per thread and colour; anything else runs sor_persistent_kernel.  Same operations in the same order: identical iterates.
template <int K, int NT>
__global__ void __launch_bounds__(NT, 1) sor_resident_kernel(DevMesh m, const double* __restrict__ offS, const double* __restrict__ bS,
                                                                         const double* __restrict__ ddiag, double* q, double omega,
                                                                         ColourRanges cr, Scalars* sc, double* __restrict__ partial,
                                                                         SolvePlan pl, unsigned* bar) {
    if (!sc->tail_done || !sc->dep_present || sc->done) return;  // uniform: written only after the last grid barrier
    extern __shared__ unsigned char sor_smem[];
    int* s_nb = reinterpret_cast<int*>(sor_smem);                                        // [2][K][3][NT]
    double* s_o = reinterpret_cast<double*>(sor_smem + (size_t)2 * K * 3 * NT * sizeof(int));  // [2][K][2][NT]
    const int Tp = m.Tp, tid = threadIdx.x;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + tid;
    double o2[2][K];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int p = (c < cr.n ? cr.start[c] : 0) + t0 + k * stride;
            const bool ok = c < cr.n && p < cr.end[c];
#pragma unroll
            for (int j = 0; j < 3; ++j) s_nb[((c * K + k) * 3 + j) * NT + tid] = ok ? m.nbs[(size_t)j * Tp + p] : 0;
            s_o[((c * K + k) * 2 + 0) * NT + tid] = ok ? offS[p] : 0.0;
            s_o[((c * K + k) * 2 + 1) * NT + tid] = ok ? offS[(size_t)Tp + p] : 0.0;
            o2[c][k] = ok ? offS[(size_t)2 * Tp + p] : 0.0;
        }
    const double bnorm2 = sc->bnorm2;
    unsigned target = 0;
    int it = 0, done = 0;
    double rr = bnorm2;
    while (it < pl.maxit) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (c >= cr.n) break;
            const int start = cr.start[c], end = cr.end[c];
            double qp[K], bs[K], g[K][3];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int p = start + t0 + k * stride;
                if (p < end) {
                    qp[k] = __ldcg(q + p);
                    bs[k] = bS[p];
#pragma unroll
                    for (int j = 0; j < 3; ++j) g[k][j] = __ldcg(q + s_nb[((c * K + k) * 3 + j) * NT + tid]);
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int p = start + t0 + k * stride;
                if (p < end) {
                    double z = bs[k] - qp[k];
                    z -= s_o[((c * K + k) * 2 + 0) * NT + tid] * g[k][0];
                    z -= s_o[((c * K + k) * 2 + 1) * NT + tid] * g[k][1];
                    z -= o2[c][k] * g[k][2];
                    q[p] = qp[k] + omega * z;
                }
            }
            grid_barrier(bar, target);
        }
        ++it;
        const bool check = (it >= pl.check_first && (it - pl.check_first) % pl.check_every == 0) || it >= pl.maxit;
        if (!check) continue;
        double a = 0.0;
        for (int p = t0; p < Tp; p += stride) {
            double z = bS[p] - __ldcg(q + p);
#pragma unroll
            for (int j = 0; j < 3; ++j) z -= offS[(size_t)j * Tp + p] * __ldcg(q + m.nbs[(size_t)j * Tp + p]);
            const double r = z * ddiag[p];
            a += r * r;
        }
        a = block_sum(a);
        if (threadIdx.x == 0) partial[blockIdx.x] = a;
        grid_barrier(bar, target);
        rr = fold_partials(partial, gridDim.x);
        if (rr <= pl.tol2 * bnorm2) { done = 1; break; }
        if (!(rr == rr) || rr > 1e60 * bnorm2) { done = 2; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->rr = rr;
        sc->dep_sweeps = it;
        sc->iters = it;
        if (done == 1) { sc->done = 1; sc->dep_ok = 1; sc->dep_buf = 0; }
        else if (done == 2) sc->done = 2;
    }
}

// Jacobi-preconditioned CG on the (SPD) deposition system; q, r, Ap are [Tp], the search direction pv is
// ghost-extended [S].  Recurrence scalars live on the device; on a single rank the block that folds a dot
// product also updates them (FUSED), so one iteration is three launches and no host round trip.
__device__ __forceinline__ void cg_scalars(int stage, Scalars* sc, double r0, double r1, double tol2) {
    switch (stage) {
        case 0:
            sc->rho = r0; sc->bnorm2 = r1; sc->rr = r1;
            sc->done = (r1 == 0.0) ? 1 : 0; sc->iters = 0; sc->beta = 0.0; sc->alpha = 0.0;
            if (r1 == 0.0) sc->dep_ok = 1;
            break;
        case 1:
            if (r0 == 0.0 || isnan(r0)) { sc->done = 2; break; }
            sc->alpha = sc->rho / r0;
            if (sc->log_alpha && sc->log_n < sc->log_cap) sc->log_alpha[sc->log_n] = sc->alpha;
            break;
        case 2:
            if (sc->log_beta && sc->log_n < sc->log_cap) { sc->log_beta[sc->log_n] = r0 / sc->rho; sc->log_n += 1; }
            sc->iters += 1;
            sc->rr = r1;
            if (r1 <= tol2 * sc->bnorm2) { sc->done = 1; sc->dep_ok = 1; break; }
            if (isnan(r0)) { sc->done = 2; break; }
            sc->beta = r0 / sc->rho;
            sc->rho = r0;
            break;
    }
}
__global__ void cg_scalar_kernel(int stage, Scalars* sc, const double* __restrict__ red, double tol2) {
    if (!sc->tail_done || !sc->dep_present) return;
    if (stage != 0 && sc->done) return;
    cg_scalars(stage, sc, red[0], red[1], tol2);
}
// init: q = 0, r = b, z = r/diag, p = z; red <- <r,z>, <r,r>
__global__ void __launch_bounds__(kRedThreads) cg_init_kernel(int Tp, const double* __restrict__ b, const double* __restrict__ dinv,
                                                              double* __restrict__ q, double* __restrict__ r, double* __restrict__ pv,
                                                              double* __restrict__ partial, int pstride, Scalars* sc,
                                                              double* __restrict__ red, double tol2, int fused) {
    if (!sc->tail_done || !sc->dep_present) return;
    double a0 = 0.0, a1 = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Tp; k += gridDim.x * blockDim.x) {
        double rv = b[k], zv = rv * dinv[k];
        q[k] = 0.0; r[k] = rv; pv[k] = zv;
        a0 += rv * zv; a1 += rv * rv;
    }
    double o0, o1;
    if (grid_fold<2>(a0, a1, 0, 0, partial, pstride, &sc->ticket[3], o0, o1)) {
        if (threadIdx.x == 0) { red[0] = o0; red[1] = o1; if (fused) cg_scalars(0, sc, o0, o1, tol2); }
    }
}
__device__ __forceinline__ bool cg_idle(const Scalars* sc) { return !sc->tail_done || !sc->dep_present || sc->done; }
// Ap = A p ; red[0] <- <p, Ap>
__global__ void __launch_bounds__(kRedThreads) cg_spmv_kernel(DevMesh m, const double* __restrict__ ddiag,
                                                              const double* __restrict__ doff, const double* __restrict__ pv,
                                                              double* __restrict__ Ap, double* __restrict__ partial, int pstride,
                                                              Scalars* sc, double* __restrict__ red, double tol2, int fused) {
    if (sc && cg_idle(sc)) return;
    const int Tp = m.Tp;
    double a = 0.0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < Tp; p += gridDim.x * blockDim.x) {
        const double pp = pv[p];
        double v = ddiag[p] * pp;
#pragma unroll
        for (int j = 0; j < 3; ++j) v += doff[(size_t)j * Tp + p] * pv[m.nbs[(size_t)j * Tp + p]];
        Ap[p] = v;
        a += v * pp;
    }
    if (!sc) return;
    double o0, unused;
    if (grid_fold<1>(a, 0.0, 0, 0, partial, pstride, &sc->ticket[3], o0, unused)) {
        if (threadIdx.x == 0) { red[0] = o0; red[1] = 0.0; if (fused) cg_scalars(1, sc, o0, 0.0, tol2); }
    }
}
// q += alpha p ; r -= alpha Ap ; red <- <r, r/diag>, <r,r>
__global__ void __launch_bounds__(kRedThreads) cg_update_kernel(int Tp, Scalars* sc, const double* __restrict__ dinv,
                                                                const double* __restrict__ pv, const double* __restrict__ Ap,
                                                                double* __restrict__ q, double* __restrict__ r,
                                                                double* __restrict__ partial, int pstride, double* __restrict__ red,
                                                                double tol2, int fused) {
    if (cg_idle(sc)) return;
    const double alpha = sc->alpha;
    double a0 = 0.0, a1 = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Tp; k += gridDim.x * blockDim.x) {
        q[k] += alpha * pv[k];
        double rv = r[k] - alpha * Ap[k];
        r[k] = rv;
        a0 += rv * rv * dinv[k];
        a1 += rv * rv;
    }
    double o0, o1;
    if (grid_fold<2>(a0, a1, 0, 0, partial, pstride, &sc->ticket[3], o0, o1)) {
        if (threadIdx.x == 0) { red[0] = o0; red[1] = o1; if (fused) cg_scalars(2, sc, o0, o1, tol2); }
    }
}
// p = r/diag + beta p
__global__ void __launch_bounds__(kRedThreads) cg_p_kernel(int Tp, const Scalars* __restrict__ sc, const double* __restrict__ dinv,
                                                           const double* __restrict__ r, double* __restrict__ pv) {
    if (cg_idle(sc)) return;
    const double beta = sc->beta;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Tp; k += gridDim.x * blockDim.x) pv[k] = r[k] * dinv[k] + beta * pv[k];
}

// Drift update (PBSM3D.cpp:1710-1740); runs only when the deposition solve happened and converged, otherwise
// drift_mass keeps its previous value as in the reference.  swe is in CHM order.
__global__ void __launch_bounds__(256) drift_kernel(int Tp, double dt, const int* __restrict__ perm, const double* __restrict__ q0,
                                                    const double* __restrict__ q1, const double* __restrict__ swe_in,
                                                    const unsigned char* __restrict__ salt,
                                                    double* __restrict__ drift_mass, double* __restrict__ sum_drift,
                                                    double* __restrict__ more_than_avail, const Scalars* __restrict__ sc) {
    if (!sc->dep_ok || sc->drift_done) return;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp) return;
    const int i = perm[p];
    if (i < 0) return;
    double qdep = sc->dep_buf ? q1[p] : q0[p];
    qdep = chm_is_nan(qdep) ? 0.0 : qdep;
    double mass = qdep * dt;
    double swe = swe_in[i];
    swe = chm_is_nan(swe) ? 0.0 : swe;
    if (mass < 0 && fabs(mass) > swe) { more_than_avail[p] = 1.0; mass = -swe; }
    if (mass < 0 && !salt[p]) mass = 0.0;
    drift_mass[p] = mass;
    sum_drift[p] += mass;
}
__global__ void drift_done_kernel(Scalars* sc) {
    if (sc->dep_ok) sc->drift_done = 1;
}

// Outputs back to CHM face order (the order of every array that crosses the C-ABI).
struct ExportPtrs {
    const double* src[9];
    double* dst[9];
};
__global__ void __launch_bounds__(256) export_kernel(int T, const int* __restrict__ iperm, ExportPtrs e) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const int p = iperm[i];
#pragma unroll
    for (int k = 0; k < 9; ++k)
        if (e.dst[k]) e.dst[k][i] = e.src[k][p];
}
// dst[iperm[i]] = src[i] (state restore)
__global__ void from_chm_kernel(int T, const int* __restrict__ iperm, const double* __restrict__ src, double* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T) dst[iperm[i]] = src[i];
}

// ---------------------------------------------------------------------------------------------- halo
// Pack the rows of the faces a partner needs into its send block: buf[off_p*nl + z*cnt_p + k] = v[z*S + slot[k]].
__global__ void __launch_bounds__(256) halo_pack_kernel(int n_send, int nl, int S, const int* __restrict__ send_slot,
                                                        const int* __restrict__ send_boff, const int* __restrict__ send_cnt,
                                                        const int* __restrict__ send_pos, const double* __restrict__ v,
                                                        double* __restrict__ buf) {
    size_t total = (size_t)n_send * nl;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int z = (int)(e / n_send), k = (int)(e - (size_t)z * n_send);
        buf[(size_t)send_boff[k] * nl + (size_t)z * send_cnt[k] + send_pos[k]] = v[(size_t)z * S + send_slot[k]];
    }
}
// Received blocks (per owner [nl][cnt]) into the ghost tails: v[z*S + Tp + g] = buf[gstart*nl + z*gcnt + (g-gstart)].
__global__ void __launch_bounds__(256) halo_unpack_kernel(int nG, int nl, int Tp, int S, const int* __restrict__ gstart,
                                                          const int* __restrict__ gcnt, const double* __restrict__ buf,
                                                          double* __restrict__ v) {
    size_t total = (size_t)nG * nl;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int z = (int)(e / nG), g = (int)(e - (size_t)z * nG);
        const int gs = gstart[g];
        v[(size_t)z * S + Tp + g] = buf[(size_t)gs * nl + (size_t)z * gcnt[g] + (g - gs)];
    }
}

// ------------------------------------------------------------------------------ halo over peer memory
// One process per GPU; every rank exports ONE arena (cudaIpc) holding its flags, all-reduce slots and two halo
// staging buffers, and maps the arenas of the other ranks of the NVSwitch box.  A halo is then
//     push : each rank stores the rows its partners need straight into THEIR staging buffer over NVLink, fences,
//            and the last block raises the partner's flag to this exchange's epoch (st.release.sys)
//     wait : the consumer spins on its own flags (ld.acquire.sys) until every partner's epoch arrived, then
//            copies the staged rows into the ghost tails of the vector
// with no host involvement, no NCCL proxy and no rendez-vous: two short launches, or none when the producer /
// consumer kernels carry the push / wait themselves (cheb_iter_kernel).  Epochs count the halo operations of the
// handle; every rank enqueues the same sequence (control flow depends only on globally reduced scalars), so
// partners agree on them.  Staging is double-buffered on the epoch's parity: a rank can be at most one exchange
// ahead of a partner (its next push needs the partner's previous one), so a buffer is never overwritten before
// its owner has unpacked it.
constexpr int kMaxRanks = 16;

struct PeerTable {
    int n_ranks, me;
    int n_partners;                                   // halo partners (ranks that share a partition boundary with me)
    int partner_rank[kMaxRanks];
    unsigned long long* halo_flag_remote[kMaxRanks];  // [partner] -> that rank's halo_flag[me]
    unsigned long long* halo_flag_local;              // [kMaxRanks] indexed by source rank (my arena)
    double* ar_slots_remote[kMaxRanks];               // [rank] -> that rank's ar_slots
    unsigned long long* ar_flag_remote[kMaxRanks];    // [rank] -> that rank's ar_flag[me]
    double* ar_slots_local;                           // [2][kMaxRanks][4]
    unsigned long long* ar_flag_local;                // [kMaxRanks]
    int* error;                                       // sticky: 1 = a wait timed out (peer died / protocol bug)
    unsigned long long* ar_epoch;                     // device-resident count of the all-reduces done so far
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spin until *flag >= epoch.  Bounded: a peer that never arrives sets the sticky error instead of hanging the GPU.
__device__ __forceinline__ void peer_wait(const unsigned long long* flag, unsigned long long epoch, const PeerTable* pt) {
    if (ld_acquire_sys(flag) >= epoch) return;
    if (*(volatile int*)pt->error) return;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flag) < epoch) {
        __nanosleep(40);
        if (global_timer_ns() - t0 > pt->timeout_ns) { *(volatile int*)pt->error = 1; return; }
    }
}
// every partner's halo of `epoch` has landed in my arena (call from all threads of a block; ends with a barrier)
__device__ __forceinline__ void halo_wait_block(const PeerTable* pt, unsigned long long epoch) {
    if ((int)threadIdx.x < pt->n_partners) peer_wait(pt->halo_flag_local + pt->partner_rank[threadIdx.x], epoch, pt);
    __syncthreads();
}
// all remote stores of this grid are done (each thread fenced its own): the last block tells the partners
__device__ __forceinline__ void halo_signal_grid(const PeerTable* pt, unsigned long long epoch, unsigned* ticket) {
    __shared__ bool last_block;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicInc(ticket, gridDim.x - 1);
        last_block = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last_block && (int)threadIdx.x < pt->n_partners) {
        __threadfence_system();
        st_release_sys(pt->halo_flag_remote[threadIdx.x], epoch);
    }
}

// push: stage_remote[k] points at (partner's staging buffer of this parity) + (my block there) + send_pos[k]
__global__ void __launch_bounds__(256) halo_push_kernel(int n_send, int nl, int S, const int* __restrict__ send_slot,
                                                        const int* __restrict__ send_cnt, double* const* __restrict__ stage_remote,
                                                        const double* __restrict__ v, const PeerTable* __restrict__ pt,
                                                        unsigned long long epoch, unsigned* ticket) {
    const size_t total = (size_t)n_send * nl;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(e / n_send), k = (int)(e - (size_t)z * n_send);
        stage_remote[k][(size_t)z * send_cnt[k]] = __ldcg(v + (size_t)z * S + send_slot[k]);
    }
    halo_signal_grid(pt, epoch, ticket);
}
// wait + unpack: v[z*S + Tp + g] = stage[gstart*Lmax + z*gcnt + (g-gstart)]   (block of an owner: [nl][cnt] at gstart*Lmax)
__global__ void __launch_bounds__(256) halo_wait_unpack_kernel(int nG, int nl, int Lmax, int Tp, int S, const int* __restrict__ gstart,
                                                               const int* __restrict__ gcnt, const double* stage, double* __restrict__ v,
                                                               const PeerTable* __restrict__ pt, unsigned long long epoch) {
    halo_wait_block(pt, epoch);
    const size_t total = (size_t)nG * nl;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(e / nG), g = (int)(e - (size_t)z * nG);
        const int gs = gstart[g];
        v[(size_t)z * S + Tp + g] = __ldcg(stage + (size_t)gs * Lmax + (size_t)z * gcnt[g] + (g - gs));
    }
}

// All-reduce of n <= 4 doubles over all ranks through peer memory, in place in red[]: lane q stores my values into
// rank q's slot [parity][me], raises rank q's flag, then waits for rank q's values; lane 0 folds the P slots in
// RANK ORDER, so every rank gets the same bits and the result does not depend on arrival order.  One warp.
// The epoch is a device-resident counter (every rank runs the same sequence of all-reduces, so the counters agree):
// kernels that decide on the device how many reductions they need (the persistent solves) need no host bookkeeping.
__device__ __forceinline__ void peer_allreduce_warp(double* red, int n, int is_max, const PeerTable* __restrict__ pt) {
    const int q = threadIdx.x & 31, P = pt->n_ranks, me = pt->me;
    unsigned long long epoch = 0;
    if (q == 0) epoch = *pt->ar_epoch + 1;
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    const int parity = (int)(epoch & 1ull);
    if (q < P) {
        double* dst = pt->ar_slots_remote[q] + ((size_t)parity * kMaxRanks + me) * 4;
        for (int i = 0; i < n; ++i) dst[i] = __ldcg(red + i);
        __threadfence_system();
        st_release_sys(pt->ar_flag_remote[q], epoch);
        peer_wait(pt->ar_flag_local + q, epoch, pt);
    }
    __syncwarp();
    if (q == 0) {
        const double* base = pt->ar_slots_local + (size_t)parity * kMaxRanks * 4;
        for (int i = 0; i < n; ++i) {
            double acc = __ldcg(base + i);
            for (int r = 1; r < P; ++r) {
                const double w = __ldcg(base + (size_t)r * 4 + i);
                acc = is_max ? fmax(acc, w) : acc + w;
            }
            red[i] = acc;
        }
        *pt->ar_epoch = epoch;
    }
    __syncwarp();
}
__global__ void peer_allreduce_kernel(double* red, int n, int is_max, const PeerTable* __restrict__ pt) {
    peer_allreduce_warp(red, n, is_max, pt);
}

// ------------------------------------------------------- halos carried by the solver kernels themselves
// The two iterations that dominate a step exchange one halo per iteration (x of the line sweeps: L x n_shared
// doubles; q of the Chebyshev iteration: n_shared doubles).  As separate push / wait-unpack launches each exchange
// costs ~17 us on two B200s -- more than the Chebyshev iteration it serves.  Here the producer kernel itself stores
// the values its partners need straight into THEIR ghost buffer over NVLink and the consumer kernel reads its
// ghosts from that buffer: no pack, no unpack, no extra launch, and the transfer overlaps the interior work.
//   * Slots are ordered boundary-first inside every colour class (boundary = a partner needs the face), so the
//     boundary columns are the first blocks of a launch: they compute, store locally AND remotely, fence, and the
//     last of them raises the partners' flags while the rest of the grid is still working on interior faces.
//   * Ghost values live in THREE buffers per channel in the arena, indexed by the iteration number e (monotonic
//     over the life of the handle): iteration e reads buffer e%3 (filled by the partners during their iteration
//     e-1) and writes the partners' buffer (e+1)%3.  A rank starts iteration e only when every partner's flag is
//     >= e, so it is never more than one iteration ahead, and the partner it is ahead of reads (e-1)%3: no buffer
//     is written while it can still be read.  The first iteration of a solve reads ghosts as 0 (x0 = 0).
//   * A launch that is a no-op because the solve has already converged still raises the flags (every rank takes
//     the same decision from the same globally reduced scalars), so the epochs stay in step.
struct HaloLink {
    const PeerTable* pt;
    unsigned long long* const* flag_remote;  // [n_partners] -> the partner's flag of this channel, slot [me]
    const unsigned long long* flag_local;    // my flags of this channel, indexed by source rank
    const int* bptr;                         // [n_boundary + 1] boundary face (in slot order) -> its send entries
    double* const* remote;                   // [n_entries] where each entry lands in the partner's buffer written now
    const int* rstride;                      // [n_entries] layer stride (the partner's padded ghost count)
    const double* ghost;                     // [nl][nGp] my ghost buffer read by this launch (a zero buffer for x0 = 0)
    int nGp;
    unsigned long long wait_epoch, signal_epoch;  // signal_epoch == 0: push only, a later launch signals
};
struct BndRanges {  // boundary slots of colour c: [start[c], start[c] + count[c]), numbered off[c].. in bptr;
    int n_colours, total;  // interior (and padding) slots of the class: [start[c] + count[c], end[c])
    int start[8], count[8], off[8], end[8];
};

// Flags of the fused channels are COUNTERS: every launch that signals adds exactly kLinkUnits to each partner's flag,
// spread over its boundary blocks (block 0 adds what the others do not), so no grid-wide ticket and no second
// fence sit between a block's last remote store and the partner seeing it; iteration e is complete at the partner
// when the counter reaches kLinkUnits * (e + 1).
constexpr unsigned long long kLinkUnits = 1ull << 20;  // > the largest number of boundary blocks a launch may use

__device__ __forceinline__ void link_wait(const HaloLink& hl) {
    if ((int)threadIdx.x < hl.pt->n_partners)
        peer_wait(hl.flag_local + hl.pt->partner_rank[threadIdx.x], hl.wait_epoch * kLinkUnits, hl.pt);
    __syncthreads();
}
__device__ __forceinline__ void link_add(unsigned long long* flag, unsigned long long v) {
    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(flag), "l"(v) : "memory");
}
// the launch is a no-op on every rank: keep the partners' counters moving
__device__ __forceinline__ void link_signal_idle(const HaloLink& hl) {
    if (hl.signal_epoch && blockIdx.x == 0 && (int)threadIdx.x < hl.pt->n_partners) link_add(hl.flag_remote[threadIdx.x], kLinkUnits);
}
// Called by all threads of each of the first `nbb` blocks after their remote stores.  The barrier orders the
// block's stores before the release (the PTX memory model's release is cumulative over barrier synchronisation).
__device__ __forceinline__ void link_signal(const HaloLink& hl, int nbb) {
    __syncthreads();
    if ((int)threadIdx.x < hl.pt->n_partners) {
        // push-only launch: fence, so its stores are ordered before the later launch's release; else the release itself
        if (hl.signal_epoch) link_add(hl.flag_remote[threadIdx.x], blockIdx.x == 0 ? kLinkUnits - (unsigned long long)(nbb - 1) : 1ull);
        else __threadfence_system();
    }
}
// Branch-free gather: one address select, one load, so a column's loads can all be issued up front as in the
// single-rank kernel.  Ghost lines are first touched after link_wait()'s acquire + barrier, and only by the
// boundary blocks, so they cannot be stale in L1.
__device__ __forceinline__ double link_gather(const HaloLink& hl, const double* v, size_t zS, size_t zG, int n, int Tp) {
    const double* src = n < Tp ? v + zS + n : hl.ghost + zG + (n - Tp);
    return *src;
}

// One colour pass of the line Gauss-Seidel sweep (see gs_sweep_kernel) with the halo inside.
template <int LT, typename CT>
__global__ void __launch_bounds__(128, 4) gs_sweep_halo_kernel(SuspSystem s, DevMesh m, int Lrt, int p0, int p1, double* x,
                                                               const Scalars* __restrict__ sc, HaloLink hl, int bnd_n, int bnd_off) {
    if (sc->susp_done) { link_signal_idle(hl); return; }
    const int Tp = m.Tp, S = m.S;
    const int L = LT > 0 ? LT : Lrt;
    const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int nbb = max(1, (bnd_n + (int)blockDim.x - 1) / (int)blockDim.x);
    // Only boundary faces have ghost neighbours (edge adjacency is symmetric: a face next to a ghost is a face the
    // ghost's owner needs), and they sit in the first nbb blocks: nobody else waits or ever takes the ghost branch.
    if ((int)blockIdx.x < nbb) link_wait(hl);
    const bool act = p < p1;
    double g[LT > 0 ? LT : 1];
    if (act) {
        const int n0 = m.nbs[p], n1 = m.nbs[(size_t)Tp + p], n2 = m.nbs[(size_t)2 * Tp + p];
        const size_t LTp = (size_t)L * Tp;
        if (LT > 0) {
            CT bl[LT > 0 ? LT : 1], cu[LT > 0 ? LT : 1];
#pragma unroll
            for (int z = 0; z < LT; ++z) {
                const size_t r = (size_t)z * Tp + p;
                const size_t zS = (size_t)z * S, zG = (size_t)z * hl.nGp;
                const RowCoef<CT> c = load_row<CT>(s, r, LTp);
                g[z] = -((double)c.l0 * link_gather(hl, x, zS, zG, n0, Tp) + (double)c.l1 * link_gather(hl, x, zS, zG, n1, Tp) +
                         (double)c.l2 * link_gather(hl, x, zS, zG, n2, Tp));
                bl[z] = c.bl;
            }
#pragma unroll
            for (int z = 0; z < LT; ++z) cu[z] = load_cp<CT>(s, (size_t)z * Tp + p);
            double y = g[0] + s.rhsS0[p];
            g[0] = y;
#pragma unroll
            for (int z = 1; z < LT; ++z) { y = g[z] - (double)bl[z] * y; g[z] = y; }
            x[(size_t)(LT - 1) * S + p] = y;
#pragma unroll
            for (int z = LT - 2; z >= 0; --z) { y = g[z] - (double)cu[z] * y; g[z] = y; x[(size_t)z * S + p] = y; }
        } else {
            double y = 0.0;
            for (int z = 0; z < L; ++z) {
                const size_t r = (size_t)z * Tp + p;
                const size_t zS = (size_t)z * S, zG = (size_t)z * hl.nGp;
                const RowCoef<CT> c = load_row<CT>(s, r, LTp);
                double gg = -((double)c.l0 * link_gather(hl, x, zS, zG, n0, Tp) + (double)c.l1 * link_gather(hl, x, zS, zG, n1, Tp) +
                              (double)c.l2 * link_gather(hl, x, zS, zG, n2, Tp));
                if (z == 0) gg += s.rhsS0[p];
                y = gg - (double)c.bl * y;
                x[zS + p] = y;
            }
            for (int z = L - 2; z >= 0; --z) {
                y = x[(size_t)z * S + p] - (double)load_cp<CT>(s, (size_t)z * Tp + p) * y;
                x[(size_t)z * S + p] = y;
            }
        }
    }
    if ((int)blockIdx.x < nbb) {
        const int i = p - p0;
        if (act && i < bnd_n) {
            const int e0 = hl.bptr[bnd_off + i], e1 = hl.bptr[bnd_off + i + 1];
            for (int e = e0; e < e1; ++e) {
                double* dst = hl.remote[e];
                const int st = hl.rstride[e];
                if (LT > 0) {
#pragma unroll
                    for (int z = 0; z < LT; ++z) dst[(size_t)z * st] = g[z];
                } else {
                    for (int z = 0; z < L; ++z) dst[(size_t)z * st] = x[(size_t)z * S + p];
                }
            }
        }
        link_signal(hl, nbb);
    }
}

// ||b - A x||^2 of this rank with the ghosts read from the halo buffer the NEXT sweep would read (residual_col_kernel).
template <int LT>
__global__ void __launch_bounds__(128) residual_halo_kernel(SuspSystem s, DevMesh m, int Lrt, const double* __restrict__ x,
                                                            double* __restrict__ partial, int pstride, Scalars* sc,
                                                            double* __restrict__ red, HaloLink hl) {
    if (sc->susp_done) return;
    link_wait(hl);
    const int Tp = m.Tp, S = m.S;
    const int L = LT > 0 ? LT : Lrt;
    double a = 0.0;
    const int ntiles = (Tp + 127) / 128;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p = tile * 128 + threadIdx.x;
        if (p >= Tp) continue;
        const int n0 = m.nbs[p], n1 = m.nbs[(size_t)Tp + p], n2 = m.nbs[(size_t)2 * Tp + p];
        double xm = 0.0, xc = x[p], cp_prev = 0.0;
        for (int z = 0; z < L; ++z) {
            const size_t r = (size_t)z * Tp + p;
            const size_t zS = (size_t)z * S, zG = (size_t)z * hl.nGp;
            const double xp = z < L - 1 ? x[zS + S + p] : 0.0;
            double v = __ldcs(s.latS + r) * link_gather(hl, x, zS, zG, n0, Tp) +
                       __ldcs(s.latS + (size_t)L * Tp + r) * link_gather(hl, x, zS, zG, n1, Tp) +
                       __ldcs(s.latS + (size_t)2 * L * Tp + r) * link_gather(hl, x, zS, zG, n2, Tp);
            const double bS = __ldcs(s.belowS + r), cp = __ldcs(s.cp + r);
            v += xc + bS * (cp_prev * xc + xm) + cp * xp;  // belowS = 0 in layer 0, cp = 0 in the top layer
            v = (((z == 0) ? s.rhsS0[p] : 0.0) - v) * __ldcs(s.den + r);
            a += v * v;
            xm = xc;
            xc = xp;
            cp_prev = cp;
        }
    }
    double rr, unused;
    if (grid_fold<1>(a, 0.0, 0, 0, partial, pstride, &sc->ticket[1], rr, unused)) {
        if (threadIdx.x == 0) red[0] = rr;
    }
}

// One Jacobi-Chebyshev iteration of the deposition solve (see cheb_iter_kernel) with the halo inside.
// An iteration is ~8 us on 10^6 faces, so even one flag round trip per iteration shows.  The q channel therefore
// carries its synchronisation IN the data (the LL idea of NCCL): a ghost entry is a 16-byte {value, tag} pair written
// with one 16-byte store, tag = number of the iteration that will read it.  The consumer's boundary thread spins on
// exactly the entries it gathers until their tag is its own iteration number: no flags, no fences, no signal, one
// NVLink hop between a partner's store and its use.  Buffers rotate over three generations as for the x channel
// (a writer reaches generation e+3 only after an iteration that consumed what the reader produced AFTER reading e).
struct TaggedLink {
    const PeerTable* pt;            // timeout + sticky error only
    const int* bptr;                // [n_boundary + 1] boundary face -> its send entries
    ulonglong2* const* remote;      // [n_entries] the partner's entry for this face in the generation written now
    const ulonglong2* ghost;        // [nGp] my entries of the generation read now; null: q_0 = 0, ghosts read as 0
    unsigned long long read_tag;    // = this iteration's number
    unsigned long long write_tag;   // = read_tag + 1
};
__device__ __forceinline__ double tagged_read(const ulonglong2* e, unsigned long long tag, const PeerTable* pt) {
    unsigned long long v, t;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v), "=l"(t) : "l"(e) : "memory");
    if (t != tag && !*(volatile int*)pt->error) {
        const unsigned long long t0 = global_timer_ns();
        do {
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v), "=l"(t) : "l"(e) : "memory");
            if (t != tag && global_timer_ns() - t0 > pt->timeout_ns) { *(volatile int*)pt->error = 1; break; }
        } while (t != tag);
    }
    return __longlong_as_double((long long)v);
}
__device__ __forceinline__ void tagged_write(ulonglong2* e, double v, unsigned long long tag) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(e), "l"((unsigned long long)__double_as_longlong(v)), "l"(tag) : "memory");
}
template <bool BND>
__device__ __forceinline__ double cheb_face(const DevMesh& m, const double* __restrict__ offS, const double* __restrict__ bS,
                                            const double* __restrict__ qin, double* __restrict__ qout,
                                            double ak, double ck, int p, const TaggedLink& tl, double& z_out) {
    const int Tp = m.Tp;
    // everything that does not depend on the partners is loaded before the first ghost is waited for
    const double qp = qin[p], bp = bS[p], qm = qout[p];  // qout still holds q_{k-1}
    int n[3];
    double o[3], qn[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { n[j] = m.nbs[(size_t)j * Tp + p]; o[j] = offS[(size_t)j * Tp + p]; }
#pragma unroll
    for (int j = 0; j < 3; ++j) qn[j] = (!BND || n[j] < Tp) ? qin[n[j]] : 0.0;
    if (BND && tl.ghost) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (n[j] >= Tp) qn[j] = tagged_read(tl.ghost + (n[j] - Tp), tl.read_tag, tl.pt);
    }
    double z = bp - qp;
#pragma unroll
    for (int j = 0; j < 3; ++j) z -= o[j] * qn[j];
    const double qnew = qp + (ak * (qp - qm) + ck * z);
    qout[p] = qnew;
    z_out = z;
    return qnew;
}
// The boundary faces of one iteration (a function of its own so that the spin loop's registers are not the interior
// loop's): update, then one tagged 16-byte store per partner that needs the face.
template <int CHECK>
__device__ __noinline__ double cheb_boundary(const DevMesh& m, const double* __restrict__ offS, const double* __restrict__ bS,
                                             const double* __restrict__ ddiag, const double* __restrict__ qin,
                                             double* __restrict__ qout, double ak, double ck, const TaggedLink& tl, const BndRanges& br,
                                             int nbb) {
    double rr = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < br.total; i += nbb * blockDim.x) {
        int p = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c < br.n_colours && i >= br.off[c] && i < br.off[c] + br.count[c]) p = br.start[c] + (i - br.off[c]);
        const int e0 = tl.bptr[i], e1 = tl.bptr[i + 1];
        ulonglong2* dst0 = e1 > e0 ? tl.remote[e0] : nullptr;
        double z;
        const double qn = cheb_face<true>(m, offS, bS, qin, qout, ak, ck, p, tl, z);
        if (dst0) tagged_write(dst0, qn, tl.write_tag);
        for (int e = e0 + 1; e < e1; ++e) tagged_write(tl.remote[e], qn, tl.write_tag);
        if (CHECK) { const double r = z * ddiag[p]; rr += r * r; }
    }
    return rr;
}
// Grid = nbb boundary blocks + the interior blocks, all co-resident (the host sizes the grid to one wave).  The
// boundary blocks only do the boundary faces and leave; the interior blocks never read a ghost.
template <int CHECK>
__global__ void __launch_bounds__(kRedThreads, 8) cheb_iter_halo_kernel(const __grid_constant__ DevMesh m, const double* __restrict__ offS,
                                                                        const double* __restrict__ bS, const double* __restrict__ ddiag,
                                                                        const double* __restrict__ qin,
                                                                        double* __restrict__ qout, double ak, double ck,
                                                                        double* __restrict__ partial, int pstride, Scalars* sc,
                                                                        double* __restrict__ red, const __grid_constant__ TaggedLink tl,
                                                                        const __grid_constant__ BndRanges br, int nbb) {
    if (!sc->tail_done || !sc->dep_present || sc->done) return;
    double rr = 0.0;
    if ((int)blockIdx.x < nbb) {
        rr = cheb_boundary<CHECK>(m, offS, bS, ddiag, qin, qout, ak, ck, tl, br, nbb);
    } else {
        const int nmain = gridDim.x - nbb;
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // the interior slots of each colour class (static indices into br)
            if (c >= br.n_colours) break;
            for (int p = br.start[c] + br.count[c] + (blockIdx.x - nbb) * blockDim.x + threadIdx.x; p < br.end[c]; p += nmain * blockDim.x) {
                double z;
                cheb_face<false>(m, offS, bS, qin, qout, ak, ck, p, tl, z);
                if (CHECK) { const double r = z * ddiag[p]; rr += r * r; }
            }
        }
    }
    if (CHECK) {
        double o0, unused;
        if (grid_fold<1>(rr, 0.0, 0, 0, partial, pstride, &sc->ticket[3], o0, unused)) {
            if (threadIdx.x == 0) red[0] = o0;
        }
    }
}

// Multicolour SOR across ranks.  Colours are rank-local, so the sweep order is made global by the key
// (colour, rank): a face is updated after every neighbour with a smaller key and before every neighbour with a larger one
// -- a proper sequential ordering (two faces with the same key are on the same rank and have the same colour, hence are not
// neighbours), i.e. true SOR on the global system, not a block-hybrid.  Ghost values travel as tagged 16-byte entries
// written by the owner's boundary thread the moment it has updated the face (tag = sweep number): a boundary thread of
// sweep e reads a ghost with a smaller key at tag e (spinning until it lands) and one with a larger key at tag e-1 (0 in
// the first sweep of a solve: q0 = 0).  One buffer suffices: every reader of an entry is a neighbour of the face behind it,
// and that face's next update needs that neighbour's next value, which is produced after the read.
struct SorLink {
    TaggedLink tl;          // ghost = my tagged ghost entries, remote/bptr = where my boundary faces go
    const int* ghost_key;   // [nG] colour * n_ranks + rank of each ghost face
};
__device__ __forceinline__ double sor_ghost(const SorLink& sl, int g, int my_key, unsigned long long e, int first) {
    if (sl.ghost_key[g] < my_key) return tagged_read(sl.tl.ghost + g, e, sl.tl.pt);
    return first ? 0.0 : tagged_read(sl.tl.ghost + g, e - 1, sl.tl.pt);
}
__global__ void __launch_bounds__(256) sor_pass_halo_kernel(const __grid_constant__ DevMesh m, const double* __restrict__ offS,
                                                            const double* __restrict__ bS, double* q, double omega, int p0, int p1,
                                                            int bnd_n, int bnd_off, const Scalars* __restrict__ sc,
                                                            const __grid_constant__ SorLink sl, int my_key, unsigned long long e,
                                                            int first) {
    if (!sc->tail_done || !sc->dep_present || sc->done) return;
    const int Tp = m.Tp;
    const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    const double qp = q[p];
    double z = bS[p] - qp;
    const int i = p - p0;
    if (i < bnd_n) {  // boundary faces (first in their colour class): the only ones with ghost neighbours
        int n[3];
        double o[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { n[j] = m.nbs[(size_t)j * Tp + p]; o[j] = offS[(size_t)j * Tp + p]; }
        const int e0 = sl.tl.bptr[bnd_off + i], e1 = sl.tl.bptr[bnd_off + i + 1];
#pragma unroll
        for (int j = 0; j < 3; ++j) z -= o[j] * (n[j] < Tp ? q[n[j]] : sor_ghost(sl, n[j] - Tp, my_key, e, first));
        const double qn = qp + omega * z;
        q[p] = qn;
        for (int k = e0; k < e1; ++k) tagged_write(sl.tl.remote[k], qn, e);
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) z -= __ldcs(offS + (size_t)j * Tp + p) * q[m.nbs[(size_t)j * Tp + p]];
        q[p] = qp + omega * z;
    }
}
// ||b - A q||^2 of this rank after sweep e (every ghost entry then carries tag e)
__global__ void __launch_bounds__(kRedThreads) dep_residual_halo_kernel(const __grid_constant__ DevMesh m, const double* __restrict__ offS,
                                                                        const double* __restrict__ bS, const double* __restrict__ ddiag,
                                                                        const double* __restrict__ q, double* __restrict__ partial,
                                                                        int pstride, Scalars* sc, double* __restrict__ red,
                                                                        const __grid_constant__ SorLink sl, unsigned long long e) {
    if (!sc->tail_done || !sc->dep_present || sc->done) return;
    const int Tp = m.Tp;
    double rr = 0.0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < Tp; p += gridDim.x * blockDim.x) {
        double z = bS[p] - q[p];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int n = m.nbs[(size_t)j * Tp + p];
            z -= offS[(size_t)j * Tp + p] * (n < Tp ? q[n] : tagged_read(sl.tl.ghost + (n - Tp), e, sl.tl.pt));
        }
        const double r = z * ddiag[p];
        rr += r * r;
    }
    double o0, unused;
    if (grid_fold<1>(rr, 0.0, 0, 0, partial, pstride, &sc->ticket[3], o0, unused)) {
        if (threadIdx.x == 0) red[0] = o0;
    }
}

// ---------------------------------------------------------------- persistent solver kernels across ranks
// gs_persistent_kernel / sor_persistent_kernel with the halos of §"halos carried by the solver kernels themselves" inside the
// loop: per colour pass the blocks that own boundary columns wait for the partners' counters, read ghosts from the generation
// buffer of this sweep, and store their new values straight into the partners' next generation; the residual checks all-reduce
// through peer memory from inside the kernel (peer_allreduce_warp), so every rank takes the same decision at the same sweep.
// Sweep numbers (generation / tag arithmetic) start from a base the host advances by the executed count after the step.
struct XHalo {
    const PeerTable* pt;
    unsigned long long* const* flag_remote;
    const unsigned long long* flag_local;
    const int* bptr;
    const int* rstride;
    double* const* remote[3];   // [generation][entry]
    const double* ghost[3];     // my ghost buffers
    const double* g_zero;
    int nGp;
    unsigned long long e0;      // number of the first sweep of this solve
    int nb[8], boff[8];         // boundary columns of each colour range (first in the range) and their offset in bptr
};
template <int LT>
__global__ void __launch_bounds__(kGsThreads, 1) gs_persistent_halo_kernel(SuspSystem s, DevMesh m, int Lrt, ColourRanges cr, double* x, float* xf,
                                                                           Scalars* sc, double* __restrict__ partial, double* red,
                                                                           SolvePlan pl, unsigned* bar, const __grid_constant__ XHalo xh) {
    if (sc->susp_done) return;
    const int L = LT > 0 ? LT : Lrt;
    const int Tp = m.Tp, S = m.S;
    const double bnorm2 = sc->susp_bnorm2;
    unsigned target = 0;
    int it = 0, n_checks = 0, converged = 0, stalled = 0;
    double rr = 0.0, prev_rr = bnorm2;
    int prev_it = 0;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int nx32 = (LT > 0 && xf) ? pl.nx32 : 0;
    unsigned char* const live = (pl.use_live && sc->n_seeds_step <= (unsigned)pl.live_max_seeds) ? s.live : nullptr;  // column_state
    unsigned cnt = 0;
    if (nx32 > 0) {
        const size_t NS = (size_t)L * S;
        for (size_t k = t0; k < NS; k += stride) xf[k] = 0.f;
        grid_barrier_fenced(bar, target);
    }
    while (it < pl.maxit) {
        const int phase = it < nx32 ? 0 : (it < pl.n32 ? 1 : 2);
        const unsigned long long e = xh.e0 + (unsigned long long)it;
        HaloLink hl;
        hl.pt = xh.pt; hl.flag_remote = xh.flag_remote; hl.flag_local = xh.flag_local; hl.bptr = xh.bptr; hl.rstride = xh.rstride;
        hl.remote = xh.remote[(e + 1) % 3];
        hl.ghost = it == 0 ? xh.g_zero : xh.ghost[e % 3];
        hl.nGp = xh.nGp;
        hl.wait_epoch = e;
        for (int c = 0; c < cr.n; ++c) {
            hl.signal_epoch = (c == cr.n - 1) ? e + 1 : 0;
            const int nbc = xh.nb[c];
            // The first nbb blocks own the boundary columns and nothing else: they wait for the partners, update, store to the
            // partners and release -- a few microseconds of NVLink latency that the other blocks, which split the interior columns
            // among themselves, never see (with a static share of the interior on top they would hold every pass back).
            const int nbb = max(1, (nbc + (int)blockDim.x - 1) / (int)blockDim.x);
            const bool bblock = (int)blockIdx.x < nbb;
            if (bblock) {
                link_wait(hl);
                for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nbc; i += nbb * blockDim.x) {
                    const int p = cr.start[c] + i;
                    bool nz;  // boundary columns are always updated (their ghosts may have become non-zero) and always sent
                    if (phase == 0) nz = gs_column<LT, float, float, true>(s, m, L, p, xf, load_nbs(m, p), hl.ghost, hl.nGp);
                    else if (phase == 1) nz = gs_column<LT, float, double, true>(s, m, L, p, x, load_nbs(m, p), hl.ghost, hl.nGp);
                    else nz = gs_column<LT, double, double, true>(s, m, L, p, x, load_nbs(m, p), hl.ghost, hl.nGp);
                    if (live && nz) live[p] = 1;
                    ++cnt;
                    const int e0 = hl.bptr[xh.boff[c] + i], e1 = hl.bptr[xh.boff[c] + i + 1];
                    for (int k = e0; k < e1; ++k) {
                        double* dst = hl.remote[k];
                        const int st = hl.rstride[k];
                        for (int z = 0; z < L; ++z) dst[(size_t)z * st] = phase == 0 ? (double)xf[(size_t)z * S + p] : x[(size_t)z * S + p];
                    }
                }
                link_signal(hl, nbb);
            } else {
                const int strideI = ((int)gridDim.x - nbb) * blockDim.x;
                const int pI = cr.start[c] + nbc + ((int)blockIdx.x - nbb) * (int)blockDim.x + (int)threadIdx.x;
                if (phase == 0) gs_pass<LT, float, float>(s, m, L, pI, cr.end[c], strideI, xf, live, cnt, pl.prefetch != 0);
                else if (phase == 1) gs_pass<LT, float, double>(s, m, L, pI, cr.end[c], strideI, x, live, cnt, pl.prefetch != 0);
                else gs_pass<LT, double, double>(s, m, L, pI, cr.end[c], strideI, x, live, cnt, pl.prefetch != 0);
            }
            grid_barrier_fenced(bar, target);
        }
        ++it;
        if (it == nx32) {
            const size_t NS = (size_t)L * S;
            for (size_t k = t0; k < NS; k += stride) x[k] = (double)xf[k];
            grid_barrier_fenced(bar, target);
        }
        const bool check = (it >= pl.check_first && (it - pl.check_first) % pl.check_every == 0) || it >= pl.maxit;
        if (check || it == nx32 || it == pl.n32) flush_col_count(sc, phase, cnt);
        if (!check) continue;
        // ---- ||b - A x||^2 with the ghosts the next sweep would read
        {
            const unsigned long long en = xh.e0 + (unsigned long long)it;
            hl.ghost = xh.ghost[en % 3];
            hl.wait_epoch = en;
            link_wait(hl);
        }
        double a = 0.0;
        for (int p = t0; p < Tp; p += stride) {
            if (live && !column_state(live, m, p)) continue;  // b_p = 0, x = 0 on p and its neighbours: residual exactly 0
            ++cnt;
            const int n0 = m.nbs[p], n1 = m.nbs[(size_t)Tp + p], n2 = m.nbs[(size_t)2 * Tp + p];
            double xm = 0.0, xc = x[p], cp_prev = 0.0;
            for (int z = 0; z < L; ++z) {
                const size_t r = (size_t)z * Tp + p;
                const size_t zS = (size_t)z * S, zG = (size_t)z * hl.nGp;
                const double xp = z < L - 1 ? x[zS + S + p] : 0.0;
                double v = __ldcs(s.latS + r) * halo_gather(x, hl.ghost, zS, zG, n0, Tp) +
                           __ldcs(s.latS + (size_t)L * Tp + r) * halo_gather(x, hl.ghost, zS, zG, n1, Tp) +
                           __ldcs(s.latS + (size_t)2 * L * Tp + r) * halo_gather(x, hl.ghost, zS, zG, n2, Tp);
                const double bS = __ldcs(s.belowS + r), cp = __ldcs(s.cp + r);
                v += xc + bS * (cp_prev * xc + xm) + cp * xp;
                v = (((z == 0) ? s.rhsS0[p] : 0.0) - v) * __ldcs(s.den + r);
                a += v * v;
                xm = xc;
                xc = xp;
                cp_prev = cp;
            }
        }
        flush_col_count(sc, 3, cnt);
        a = block_sum(a);
        if (threadIdx.x == 0) partial[blockIdx.x] = a;
        grid_barrier_fenced(bar, target);
        const double rloc = fold_partials(partial, gridDim.x);
        if (blockIdx.x == 0 && threadIdx.x < 32) {  // the sum over ranks
            if (threadIdx.x == 0) red[0] = rloc;
            __syncwarp();
            peer_allreduce_warp(red, 1, 0, xh.pt);
            __threadfence();
        }
        grid_barrier_fenced(bar, target);
        rr = __ldcg(red);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (n_checks < 16) { sc->rr_hist[n_checks] = rr; sc->it_hist[n_checks] = it; }
        }
        ++n_checks;
        if (rr <= pl.tol2 * bnorm2) { converged = 1; break; }
        if (!(rr == rr) || rr > 1e60 * bnorm2) { stalled = 1; break; }
        if (it >= 64 && it > prev_it && rr > 0.0 && prev_rr > 0.0) {
            const double lim = exp(2.0 * (it - prev_it) * log(0.97));
            if (rr > lim * prev_rr) { stalled = 1; break; }
        }
        prev_rr = rr;
        prev_it = it;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->susp_rr = rr;
        sc->n_checks = n_checks;
        sc->susp_sweeps = it;
        sc->susp_stalled = stalled;
        if (converged) { sc->susp_done = 1; sc->susp_iters = it; sc->susp_ok = 1; }
    }
}

// Multicolour SOR across ranks, one launch per solve: sor_pass_halo_kernel's (colour, rank) order and tagged ghost entries inside
// the persistent loop.  Sweep k of this solve carries tag e0 + k + 1.
struct QHalo {
    SorLink sl;
    unsigned long long e0;
    int n_ranks, rank;
    int nb[8], boff[8], colour_of[8];  // boundary faces of each (non-empty) colour range; its colour number (for the key)
};
template <bool STREAM, int NT, int B>
__global__ void __launch_bounds__(NT, 1) sor_persistent_halo_kernel(const __grid_constant__ DevMesh m, const double* __restrict__ offS,
                                                                    const double* __restrict__ bS, const double* __restrict__ ddiag, double* q,
                                                                    double omega, ColourRanges cr, Scalars* sc, double* __restrict__ partial,
                                                                    double* red, SolvePlan pl, unsigned* bar, const __grid_constant__ QHalo qh) {
    if (!sc->tail_done || !sc->dep_present || sc->done) return;
    const int Tp = m.Tp;
    const double bnorm2 = sc->bnorm2;
    unsigned target = 0;
    int it = 0, done = 0;
    double rr = bnorm2;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    while (it < pl.maxit) {
        const unsigned long long e = qh.e0 + (unsigned long long)it + 1ull;
        const int first = it == 0 ? 1 : 0;
        for (int c = 0; c < cr.n; ++c) {
            const int start = cr.start[c], end = cr.end[c], nbc = qh.nb[c];
            const int my_key = qh.colour_of[c] * qh.n_ranks + qh.rank;
            // boundary faces (first in the class, the only ones with ghost neighbours) belong to the first nbb blocks, which do
            // nothing else in this pass; the other blocks split the interior
            const int nbb = max(1, (nbc + NT - 1) / NT);
            if ((int)blockIdx.x < nbb)
            for (int i = blockIdx.x * NT + threadIdx.x; i < nbc; i += nbb * NT) {
                const int p = start + i;
                const double qp = q[p];
                double z = bS[p] - qp;
                int n[3];
                double o[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) { n[j] = m.nbs[(size_t)j * Tp + p]; o[j] = offS[(size_t)j * Tp + p]; }
                const int k0 = qh.sl.tl.bptr[qh.boff[c] + i], k1 = qh.sl.tl.bptr[qh.boff[c] + i + 1];
#pragma unroll
                for (int j = 0; j < 3; ++j) z -= o[j] * (n[j] < Tp ? q[n[j]] : sor_ghost(qh.sl, n[j] - Tp, my_key, e, first));
                const double qn = qp + omega * z;
                q[p] = qn;
                for (int k = k0; k < k1; ++k) tagged_write(qh.sl.tl.remote[k], qn, e);
            }
            const int strideI = ((int)gridDim.x - nbb) * NT;
            if ((int)blockIdx.x >= nbb)
            for (int base = start + nbc + ((int)blockIdx.x - nbb) * NT + (int)threadIdx.x; base < end; base += B * strideI) {
                double qp[B], z[B], o[B][3];
                int n[B][3];
#pragma unroll
                for (int k = 0; k < B; ++k) {
                    const int p = base + k * strideI;
                    if (p < end) {
                        qp[k] = q[p];
                        z[k] = bS[p] - qp[k];
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            n[k][j] = m.nbs[(size_t)j * Tp + p];
                            o[k][j] = STREAM ? __ldcs(offS + (size_t)j * Tp + p) : offS[(size_t)j * Tp + p];
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < B; ++k)
                    if (base + k * strideI < end) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) z[k] -= o[k][j] * q[n[k][j]];
                    }
#pragma unroll
                for (int k = 0; k < B; ++k)
                    if (base + k * strideI < end) q[base + k * strideI] = qp[k] + omega * z[k];
            }
            grid_barrier(bar, target);
        }
        ++it;
        const bool check = (it >= pl.check_first && (it - pl.check_first) % pl.check_every == 0) || it >= pl.maxit;
        if (!check) continue;
        double a = 0.0;
        for (int p = t0; p < Tp; p += stride) {  // every ghost entry carries tag e after sweep e
            double z = bS[p] - q[p];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int n = m.nbs[(size_t)j * Tp + p];
                z -= offS[(size_t)j * Tp + p] * (n < Tp ? q[n] : tagged_read(qh.sl.tl.ghost + (n - Tp), e, qh.sl.tl.pt));
            }
            const double r = z * ddiag[p];
            a += r * r;
        }
        a = block_sum(a);
        if (threadIdx.x == 0) partial[blockIdx.x] = a;
        grid_barrier(bar, target);
        const double rloc = fold_partials(partial, gridDim.x);
        if (blockIdx.x == 0 && threadIdx.x < 32) {
            if (threadIdx.x == 0) red[0] = rloc;
            __syncwarp();
            peer_allreduce_warp(red, 1, 0, qh.sl.tl.pt);
            __threadfence();
        }
        grid_barrier(bar, target);
        rr = __ldcg(red);
        if (rr <= pl.tol2 * bnorm2) { done = 1; break; }
        if (!(rr == rr) || rr > 1e60 * bnorm2) { done = 2; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->rr = rr;
        sc->dep_sweeps = it;
        sc->iters = it;
        if (done == 1) { sc->done = 1; sc->dep_ok = 1; sc->dep_buf = 0; }
        else if (done == 2) sc->done = 2;
    }
}

// sor_persistent_halo_kernel with the interior blocks' static face data resident on chip (sor_resident_kernel): two colour classes,
// at most K interior faces per thread and colour.
template <int K, int NT>
__global__ void __launch_bounds__(NT, 1) sor_resident_halo_kernel(const __grid_constant__ DevMesh m, const double* __restrict__ offS,
                                                                    const double* __restrict__ bS, const double* __restrict__ ddiag, double* q,
                                                                    double omega, ColourRanges cr, Scalars* sc, double* __restrict__ partial,
                                                                    double* red, SolvePlan pl, unsigned* bar, const __grid_constant__ QHalo qh) {
    if (!sc->tail_done || !sc->dep_present || sc->done) return;
    const int Tp = m.Tp;
    // static data of this block's INTERIOR faces on chip (see sor_resident_kernel); boundary faces keep the global-memory path
    extern __shared__ unsigned char sor_smem[];
    int* s_nb = reinterpret_cast<int*>(sor_smem);                                               // [2][K][3][NT]
    double* s_o = reinterpret_cast<double*>(sor_smem + (size_t)2 * K * 3 * NT * sizeof(int));   // [2][K][2][NT]
    const int tid = threadIdx.x;
    double o2[2][K];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const bool have = c < cr.n;
        const int nbc = have ? qh.nb[c] : 0, nbb = max(1, (nbc + NT - 1) / NT);
        const int strideI = ((int)gridDim.x - nbb) * NT;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int p = (have ? cr.start[c] : 0) + nbc + ((int)blockIdx.x - nbb) * NT + tid + k * strideI;
            const bool ok = have && (int)blockIdx.x >= nbb && p < cr.end[c];
#pragma unroll
            for (int j = 0; j < 3; ++j) s_nb[((c * K + k) * 3 + j) * NT + tid] = ok ? m.nbs[(size_t)j * Tp + p] : 0;
            s_o[((c * K + k) * 2 + 0) * NT + tid] = ok ? offS[p] : 0.0;
            s_o[((c * K + k) * 2 + 1) * NT + tid] = ok ? offS[(size_t)Tp + p] : 0.0;
            o2[c][k] = ok ? offS[(size_t)2 * Tp + p] : 0.0;
        }
    }
    const double bnorm2 = sc->bnorm2;
    unsigned target = 0;
    int it = 0, done = 0;
    double rr = bnorm2;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    while (it < pl.maxit) {
        const unsigned long long e = qh.e0 + (unsigned long long)it + 1ull;
        const int first = it == 0 ? 1 : 0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (c >= cr.n) break;
            const int start = cr.start[c], end = cr.end[c], nbc = qh.nb[c];
            const int my_key = qh.colour_of[c] * qh.n_ranks + qh.rank;
            // boundary faces (first in the class, the only ones with ghost neighbours) belong to the first nbb blocks, which do
            // nothing else in this pass; the other blocks split the interior
            const int nbb = max(1, (nbc + NT - 1) / NT);
            if ((int)blockIdx.x < nbb)
            for (int i = blockIdx.x * NT + threadIdx.x; i < nbc; i += nbb * NT) {
                const int p = start + i;
                const double qp = q[p];
                double z = bS[p] - qp;
                int n[3];
                double o[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) { n[j] = m.nbs[(size_t)j * Tp + p]; o[j] = offS[(size_t)j * Tp + p]; }
                const int k0 = qh.sl.tl.bptr[qh.boff[c] + i], k1 = qh.sl.tl.bptr[qh.boff[c] + i + 1];
#pragma unroll
                for (int j = 0; j < 3; ++j) z -= o[j] * (n[j] < Tp ? q[n[j]] : sor_ghost(qh.sl, n[j] - Tp, my_key, e, first));
                const double qn = qp + omega * z;
                q[p] = qn;
                for (int k = k0; k < k1; ++k) tagged_write(qh.sl.tl.remote[k], qn, e);
            }
            const int strideI = ((int)gridDim.x - nbb) * NT;
            if ((int)blockIdx.x >= nbb) {
                const int base = start + nbc + ((int)blockIdx.x - nbb) * NT + tid;
                double qp[K], bs[K], g[K][3];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int p = base + k * strideI;
                    if (p < end) {
                        qp[k] = __ldcg(q + p);
                        bs[k] = bS[p];
#pragma unroll
                        for (int j = 0; j < 3; ++j) g[k][j] = __ldcg(q + s_nb[((c * K + k) * 3 + j) * NT + tid]);
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int p = base + k * strideI;
                    if (p < end) {
                        double z = bs[k] - qp[k];
                        z -= s_o[((c * K + k) * 2 + 0) * NT + tid] * g[k][0];
                        z -= s_o[((c * K + k) * 2 + 1) * NT + tid] * g[k][1];
                        z -= o2[c][k] * g[k][2];
                        q[p] = qp[k] + omega * z;
                    }
                }
            }
            grid_barrier(bar, target);
        }
        ++it;
        const bool check = (it >= pl.check_first && (it - pl.check_first) % pl.check_every == 0) || it >= pl.maxit;
        if (!check) continue;
        double a = 0.0;
        for (int p = t0; p < Tp; p += stride) {  // every ghost entry carries tag e after sweep e
            double z = bS[p] - q[p];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int n = m.nbs[(size_t)j * Tp + p];
                z -= offS[(size_t)j * Tp + p] * (n < Tp ? q[n] : tagged_read(qh.sl.tl.ghost + (n - Tp), e, qh.sl.tl.pt));
            }
            const double r = z * ddiag[p];
            a += r * r;
        }
        a = block_sum(a);
        if (threadIdx.x == 0) partial[blockIdx.x] = a;
        grid_barrier(bar, target);
        const double rloc = fold_partials(partial, gridDim.x);
        if (blockIdx.x == 0 && threadIdx.x < 32) {
            if (threadIdx.x == 0) red[0] = rloc;
            __syncwarp();
            peer_allreduce_warp(red, 1, 0, qh.sl.tl.pt);
            __threadfence();
        }
        grid_barrier(bar, target);
        rr = __ldcg(red);
        if (rr <= pl.tol2 * bnorm2) { done = 1; break; }
        if (!(rr == rr) || rr > 1e60 * bnorm2) { done = 2; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->rr = rr;
        sc->dep_sweeps = it;
        sc->iters = it;
        if (done == 1) { sc->done = 1; sc->dep_ok = 1; sc->dep_buf = 0; }
        else if (done == 2) sc->done = 2;
    }
}

__global__ void fill_slots_kernel(int Tp, const int* __restrict__ perm, double* __restrict__ a, double v) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < Tp) a[p] = perm[p] >= 0 ? v : 0.0;
}

}  // namespace pbsm3d

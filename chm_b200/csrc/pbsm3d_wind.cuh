// sm_100a kernels of the two per-face providers of PBSM3D inputs:
//   scale_wind_vert  U_R (50 m) -> U_2m_above_srf     reference src/modules/scale_wind_vert.cpp:48-229
//   fetchr           vw_dir     -> fetch              reference src/modules/fetchr.cpp:54-119
// Both are one thread per face.  scale_wind_vert's domain mode blends every face with its (<= 3) edge neighbours
// through a thin plate spline (src/interpolation/TPSpline.cpp:40-173): a 4x4 dense solve per face, done in
// registers.  fetchr walks `steps` points up-wind and asks for the nearest face CENTRE to each
// (triangulation.cpp:170-186, a CGAL kd-tree there): here a uniform cell grid over the centres, built once, searched
// ring by ring with an exact termination bound.
#pragma once
#include <cuda_runtime.h>
#include "pbsm3d_kernels.cuh"

namespace pbsm3d {

// Atmosphere::exp_scale_wind (physics/Atmosphere.cpp:41-46)
__device__ __forceinline__ double exp_scale_wind(double u, double Z_in, double Z_out, double alpha) {
    return u * exp(alpha * (Z_out / Z_in - 1.0));
}

// scale_wind_vert::point_scale (scale_wind_vert.cpp:48-136) for one face.
__device__ __forceinline__ double wind_point_scale(double U_R, double sd_in, bool have_sd, double Z_CanTop, double LAI,
                                                   bool canopy_on) {
    const double Z_R = kZUR;
    if (!canopy_on) Z_CanTop = 0.0;
    const double Z_CanBot = Z_CanTop / 2.0;
    double sd = have_sd ? sd_in : 0.0;
    if (chm_is_nan(sd)) sd = 0.0;
    const double Z_2m = sd + 2.0;
    if (Z_2m >= Z_R) return U_R;  // snow above the reference height: wind taken as constant
    double u2;
    if (canopy_on && Z_CanTop > 0.0 && Z_2m < Z_CanTop) {
        const double alpha = LAI;
        if (sd < Z_CanTop) {
            const double U_CanTop = log_scale_wind(U_R, Z_R, Z_CanTop, sd, kZ0Snow);
            const double U_CanBot = exp_scale_wind(U_CanTop, Z_CanTop, Z_CanBot, alpha);
            if (Z_2m < Z_CanBot) u2 = log_scale_wind(U_CanBot, Z_CanBot, Z_2m, sd, kZ0Snow);
            else u2 = exp_scale_wind(U_CanTop, Z_CanTop, Z_2m, alpha);
        } else {
            u2 = log_scale_wind(U_R, Z_R, Z_2m, sd, kZ0Snow);
        }
    } else {
        u2 = log_scale_wind(U_R, Z_R, Z_2m, sd, kZ0Snow);
    }
    return u2 > 0.1 ? u2 : 0.1;  // std::max(0.1, u2): a NaN u2 gives 0.1
}

// One thread per CHM face i.  u_slot is ghost-extended [S] in slot order (the spline gathers neighbours from it);
// in point mode the value also goes straight to out[i] (CHM order).
__global__ void __launch_bounds__(256) wind_point_kernel(int T, const int* __restrict__ iperm, const double* __restrict__ U_R,
                                                         const double* __restrict__ sd, const double* __restrict__ canopy,
                                                         const double* __restrict__ lai, int ignore_canopy,
                                                         double* __restrict__ u_slot, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const int p = iperm[i];
    const bool canopy_on = !ignore_canopy && canopy != nullptr;
    const double u = wind_point_scale(U_R[i], sd ? sd[i] : 0.0, sd != nullptr, canopy_on ? canopy[p] : 0.0,
                                      (canopy_on && lai) ? lai[p] : 0.0, canopy_on);
    u_slot[p] = u;
    if (out) out[i] = u;
}

// E1(x), x > 0: the series of its definition for x <= 1, the Lentz continued fraction above (GSL's gsl_sf_expint_E1
// in the reference; both are accurate to a few ulp, and the spline is compared at 1e-10).
__device__ __forceinline__ double expint_E1(double x) {
    if (x <= 1.0) {
        double sum = 0.0, term = 1.0;
        for (int k = 1; k < 40; ++k) {
            term *= -x / k;
            const double add = -term / k;
            sum += add;
            if (fabs(add) < 1e-18 * fabs(sum)) break;
        }
        return -0.57721566490153286061 - log(x) + sum;
    }
    double b = x + 1.0, c = 1e300, d = 1.0 / b, h = d;
    for (int i = 1; i < 200; ++i) {
        const double an = -1.0 * i * i;
        b += 2.0;
        d = 1.0 / (an * d + b);
        c = b + an / c;
        const double del = c * d;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return h * exp(-x);
}
// Rd = -(log x + c + E1(x)), x = (d * weight / 2)^2, c = 0.577215, weight = 0.01 (TPSpline.cpp:83-94,196-198)
__device__ __forceinline__ double tps_basis(double dist) {
    const double x = (dist * 0.01 / 2.0) * (dist * 0.01 / 2.0);
    return -(log(x) + 0.577215 + expint_E1(x));
}

// thin_plate_spline::operator() for n <= 3 samples: A = [1 | Rd(i,j)] with the zero-sum row, full-pivot elimination
// (Eigen::FullPivLU in the reference), evaluation at the query.
__device__ __forceinline__ double tps3(int n, const double* sx, const double* sy, const double* sv, double qx, double qy) {
    const int size = n + 1;
    double A[4][4], b[4], y[4], x[4];
    int colperm[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        b[r] = 0.0; colperm[r] = r;
#pragma unroll
        for (int c = 0; c < 4; ++c) A[r][c] = 0.0;
    }
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            const double xd = sx[i] - sx[j], yd = sy[i] - sy[j];
            if (xd == 0. && yd == 0.) continue;
            const double Rd = tps_basis(sqrt(xd * xd + yd * yd));
            A[i][j + 1] = Rd;
            A[j][i + 1] = Rd;
        }
    for (int i = 0; i < size; ++i) { A[i][0] = 1.0; A[size - 1][i] = 1.0; }
    A[size - 1][0] = 0.0;
    for (int i = 0; i < n; ++i) b[i] = sv[i];
    for (int k = 0; k < size; ++k) {
        int pr = k, pc = k;
        double best = -1.0;
        for (int r = k; r < size; ++r)
            for (int c = k; c < size; ++c)
                if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); pr = r; pc = c; }
        if (pr != k) {
            for (int c = 0; c < size; ++c) { const double t = A[k][c]; A[k][c] = A[pr][c]; A[pr][c] = t; }
            const double t = b[k]; b[k] = b[pr]; b[pr] = t;
        }
        if (pc != k) {
            for (int r = 0; r < size; ++r) { const double t = A[r][k]; A[r][k] = A[r][pc]; A[r][pc] = t; }
            const int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t;
        }
        for (int r = k + 1; r < size; ++r) {
            const double f = A[r][k] / A[k][k];
            for (int c = k; c < size; ++c) A[r][c] -= f * A[k][c];
            b[r] -= f * b[k];
        }
    }
    for (int k = size - 1; k >= 0; --k) {
        double v = b[k];
        for (int c = k + 1; c < size; ++c) v -= A[k][c] * y[c];
        y[k] = v / A[k][k];
    }
    for (int k = 0; k < size; ++k) x[colperm[k]] = y[k];
    double z0 = x[0];
    for (int i = 1; i < size; ++i) {
        const double xd = sx[i - 1] - qx, yd = sy[i - 1] - qy;
        z0 = z0 + x[i] * tps_basis(sqrt(xd * xd + yd * yd));
    }
    return z0;
}

// Second half of scale_wind_vert::run(mesh&) (scale_wind_vert.cpp:180-227): every face takes the spline of its
// neighbours' point-scaled values (ghosts included: u_slot's ghost tail is filled by the halo) at its own centre.
__global__ void __launch_bounds__(128) wind_spline_kernel(int T, DevMesh m, const double* __restrict__ cx, const double* __restrict__ cy,
                                                          const double* __restrict__ u_slot, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const int Tp = m.Tp, p = m.iperm[i];
    double sx[3], sy[3], sv[3];
    int n = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int nb = m.nbs[(size_t)j * Tp + p];
        if (nb == p) continue;  // no neighbour on this edge
        sx[n] = cx[nb]; sy[n] = cy[nb]; sv[n] = u_slot[nb];  // cx/cy are [Tp + nG], u_slot [S]: ghost g sits at Tp + g in both
        ++n;
    }
    double u;
    if (n > 0) u = tps3(n, sx, sy, sv, cx[p], cy[p]);
    else u = u_slot[p];
    out[i] = u > 0.1 ? u : 0.1;
}

// ---------------------------------------------------------------------------------------------- fetchr
// Uniform cell grid over the face centres.  Faces are stored SORTED BY CELL (row-major), so the faces of a run of
// x-adjacent cells are one contiguous range: a (2r+1)^2 block is 2r+1 contiguous reads of {x, y} pairs, and what a
// query needs from the face it finds (centre elevation, canopy height) sits at the same index in `zc`.
struct CellGrid {
    double x0, y0, h, inv_h;
    int ncx, ncy;
    const int* cell_start;  // [ncx*ncy + 1]
    const double2* xy;      // [n faces] centre (x, y), cell order
    const double2* zc;      // [n faces] (centre z, CanopyHeight or 0), cell order
};

// Index (cell order) of the nearest face centre to (qx, qy); exact, ties broken by scan order.
// A query outside the grid (up-wind of the domain edge) is searched from its projection q' onto the grid's box:
// for every centre p in the box |q-p|^2 >= |q-q'|^2 + |q'-p|^2, so rings around q' bound what is left, and the
// search still ends after a ring or two instead of growing to the distance between q and the domain.
__device__ __forceinline__ int nearest_centre(const CellGrid& g, double qx, double qy) {
    const double bx = fmin(fmax(qx, g.x0), g.x0 + g.ncx * g.h), by = fmin(fmax(qy, g.y0), g.y0 + g.ncy * g.h);
    const double out2 = (qx - bx) * (qx - bx) + (qy - by) * (qy - by);  // 0 for a query inside the box
    int qcx = (int)floor((bx - g.x0) * g.inv_h), qcy = (int)floor((by - g.y0) * g.inv_h);
    qcx = min(max(qcx, 0), g.ncx - 1);
    qcy = min(max(qcy, 0), g.ncy - 1);
    double best = 1e300;
    int bi = -1;
    auto scan_run = [&](int yy, int xa, int xb) {  // cells [xa, xb] of row yy, clipped to the grid
        if (yy < 0 || yy >= g.ncy) return;
        xa = max(xa, 0);
        xb = min(xb, g.ncx - 1);
        if (xa > xb) return;
        const int k1 = g.cell_start[yy * g.ncx + xb + 1];
        for (int k = g.cell_start[yy * g.ncx + xa]; k < k1; ++k) {
            const double2 c = g.xy[k];
            const double dx = c.x - qx, dy = c.y - qy;
            const double d2 = dx * dx + dy * dy;
            if (d2 < best) { best = d2; bi = k; }
        }
    };
    const int rmax = max(g.ncx, g.ncy);
    for (int r = 1; r <= rmax; ++r) {
        const int y0 = qcy - r, y1 = qcy + r, x0 = qcx - r, x1 = qcx + r;
        if (r == 1) {  // the 3x3 block around the cell: three contiguous runs
            scan_run(y0, x0, x1); scan_run(qcy, x0, x1); scan_run(y1, x0, x1);
        } else {       // the border of the (2r+1)^2 block
            scan_run(y0, x0, x1); scan_run(y1, x0, x1);
            for (int yy = y0 + 1; yy < y1; ++yy) { scan_run(yy, x0, x0); scan_run(yy, x1, x1); }
        }
        // every centre not yet examined lies beyond one of the sides of the examined block that is inside the grid
        double lb = 1e300;
        if (x0 > 0) lb = fmin(lb, fmax(0.0, bx - (g.x0 + x0 * g.h)));
        if (x1 < g.ncx - 1) lb = fmin(lb, fmax(0.0, (g.x0 + (x1 + 1) * g.h) - bx));
        if (y0 > 0) lb = fmin(lb, fmax(0.0, by - (g.y0 + y0 * g.h)));
        if (y1 < g.ncy - 1) lb = fmin(lb, fmax(0.0, (g.y0 + (y1 + 1) * g.h) - by));
        if (lb >= 1e300) break;                        // the whole grid has been examined
        if (bi >= 0 && best <= out2 + lb * lb) break;  // nothing closer can be left
    }
    return bi;
}

// fetchr::run (fetchr.cpp:54-119) for CHM face i; canopy == null: the mesh has no vegetation parameters.
__global__ void __launch_bounds__(128) fetchr_kernel(int T, const int* __restrict__ iperm, CellGrid g, const double* __restrict__ cx,
                                                     const double* __restrict__ cy, const double* __restrict__ cz,
                                                     const double* __restrict__ canopy, const double* __restrict__ vw_dir,
                                                     int steps, double max_distance, double I, int incl_veg,
                                                     double* __restrict__ fetch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const int p = iperm[i];
    double out = max_distance;
    const bool veg = incl_veg && canopy != nullptr;
    if (veg && canopy[p] > 1.0) { fetch[i] = 0.0; return; }
    const double size_of_step = max_distance / steps;
    const double bearing = vw_dir[i] * (M_PI / 180.0);  // point_from_bearing_UTM (coordinates.cpp:60-71)
    double sb, cb;
    sincos(bearing, &sb, &cb);
    const double mx = cx[p], my = cy[p], mz = cz[p];
    for (int j = 1; j <= steps; ++j) {
        const double distance = j * size_of_step;
        // no FMA contraction: the query point and Z_core are the reference's two-rounding expressions
        const int k = nearest_centre(g, __dadd_rn(mx, __dmul_rn(distance, sb)), __dadd_rn(my, __dmul_rn(distance, cb)));
        const double2 zc = g.zc[k];
        const double Z_CanTop = veg ? zc.y : 0.0;
        const double Z_test = __dadd_rn(zc.x, Z_CanTop);
        const double Z_core = __dadd_rn(mz, __dmul_rn(distance, I));
        const double z0_1 = 0.12 * Z_CanTop, z0_2 = 0.001, n = 1.0 / 0.8, h = 5.0;
        // x_sss = 0 without a canopy on the face found (log(0) = -inf, pow(-0, n) = 0): skip the transcendental there
        const double x_sss = Z_CanTop == 0.0 ? 0.0 : pow(((33.33333333 * h - 25. * z0_2) / (log(z0_1 / z0_2) * z0_2)), n) * z0_2;
        if (Z_test >= Z_core || (incl_veg && distance < x_sss)) { out = distance; break; }
    }
    fetch[i] = out;
}

}  // namespace pbsm3d

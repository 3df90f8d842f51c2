// snow_slide on the device: gravitational redistribution of snow that exceeds a slope-dependent holding depth.
//   src/modules/snow_slide.cpp:406-446   init: maxDepth = max(mult * slopeDeg^pow, CanopyHeight) * max(0.001, cos(slope))
//                             :94-404    run: faces sorted by (centre elevation + vertical snow depth), descending; ONE SEQUENTIAL
//                                        sweep in that order — a face above its holding depth sheds the excess to its lower
//                                        neighbours (weights = surface height differences), out of the domain at a mesh edge, into
//                                        ghost accumulators at a partition edge (sent back to the owner afterwards)
//   src/mesh/triangulation.hpp:1501-1523,1549-1574   slope = acos(z component of the unit normal)
//
// The reference sweep is sequential, and its result depends on the order (a face reads the surface of its neighbours, which earlier
// faces raise).  It is reproduced EXACTLY, not approximated, by running the same updates as a wavefront over the dependency order:
//   * the turn order is the reference's sort: key = centre z + vertical depth at the start of the sweep, descending (ties, which
//     tbb::parallel_sort leaves undefined, by CHM face index);
//   * face f touches itself and its edge neighbours, so two faces commute unless they are within two edges of each other; f may
//     take its turn once every EARLIER face within distance 2 has had its turn;
//   * only faces that can ever exceed their holding depth take part: the start candidates (depth > maxDepth) and, transitively,
//     their later-ordered neighbours (a deposit can only matter to a face whose turn is still to come) — the "live" set, found by a
//     frontier expansion; everything else is settled from the start (on most hours of a winter the live set is empty and the call
//     costs two passes over the faces).
// One persistent cooperative kernel: candidates -> frontier expansion -> wavefront rounds over a compacted work list, a grid barrier
// per round; a face waiting for an up-slope neighbour is parked and woken by that neighbour's turn.  Every load of data another
// block may have written goes through L2 (__ldcg).  Concurrently firing faces are more than
// two edges apart, so plain stores suffice; the only atomics are the work-list cursors and the ghost accumulators.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include "pbsm3d_kernels.cuh"

namespace pbsm3d {

struct SlideArrays {
    int T, Tp, S;
    const int* perm;      // [Tp] slot -> CHM face (-1 pad)
    const int* nbs;       // [3][Tp] neighbour slot; own slot = no neighbour; >= Tp = ghost
    const double* cz;     // [Tp + nG] centre elevation (ghosts at Tp + g)
    const double* area;   // [S] face area, ghost-extended
    const double* maxD;   // [Tp]
    const double* cosf;   // [Tp] max(0.001, cos(slope))
    double* sd;           // [Tp] snowdepthavg_copy
    double* sdv;          // [S]  snowdepthavg_vert_copy, ghost-extended (ghost values arrive by the forward halo)
    double* swe;          // [Tp] swe_copy (m)
    double* dsd;          // [Tp] delta_avalanche_snowdepth (m^3), running over the outer iterations of one run
    double* dmass;        // [Tp] delta_avalanche_mass (m^3 of water)
    double* key;          // [Tp] sort key of this sweep
    double* gacc;         // [nG][4] ghost accumulators {snowdepth_to_xfer, swe_to_xfer, delta_snowdepth, delta_swe}
    int* stamp;           // [Tp] 1 = not live; 0 = live, turn still to come; r >= 2: took its turn in wavefront round r - 2
    int* queued;          // [Tp] last round stamp for which the face was put on a work list (one entry per face and round)
    int* list[3];         // [Tp] work lists: the live list (built segment by segment by the frontier expansion), two alternating round lists
    int* cnt;             // [8] cursors: 0..2 rotating round cursors, 3 = live count, 4 = wavefront rounds, 5 = fired faces, 6 = mass error flag, 7 = frontier rounds
};

// Append to a work list with ONE atomic per group of converged lanes (the cursors are single addresses: per-element atomics
// serialise at the L2 and dominate rounds that append ~1e5 faces).
__device__ __forceinline__ int slide_reserve(int* cursor) {
    namespace cg = cooperative_groups;
    auto g = cg::coalesced_threads();
    int base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(cursor, (int)g.size());
    return g.shfl(base, 0) + (int)g.thread_rank();
}

__device__ __forceinline__ bool slide_earlier(double kg, int ig, double kf, int i_f) { return kg > kf || (kg == kf && ig < i_f); }

// One face's turn (snow_slide.cpp:186-328).  Returns false when the mass check of :322 fails.
__device__ __forceinline__ bool slide_fire(const SlideArrays& a, int f) {
    const int Tp = a.Tp;
    const double maxD = a.maxD[f], area_f = a.area[f], cf = a.cosf[f];
    const double snow = __ldcg(a.sd + f), snow_v = __ldcg(a.sdv + f), swe = __ldcg(a.swe + f);
    const double del_depth = __dsub_rn(snow, maxD);
    const double del_swe = __dmul_rn(swe, __dsub_rn(1.0, __ddiv_rn(maxD, snow)));
    const double orig_mass = __dmul_rn(del_swe, area_f);
    const double z_s = __dadd_rn(a.cz[f], snow_v);
    int nb[3];
    double w[3], w_dem = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int n = a.nbs[(size_t)j * Tp + f];
        nb[j] = n;
        // a missing neighbour is treated as a snow-free face at our own elevation; a ghost reads the owner's vertical depth
        const double zn = n == f ? a.cz[f] : __dadd_rn(a.cz[n], __ldcg(a.sdv + n));
        w[j] = fmax(0.0, __dsub_rn(z_s, zn));
        w_dem = __dadd_rn(w_dem, w[j]);
    }
    if (w_dem == 0.0) return true;  // a sink: nothing is routed, the face keeps its snow
    double out_mass = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double wj = __ddiv_rn(w[j], w_dem);
        const int n = nb[j];
        const double m3 = __dmul_rn(__dmul_rn(del_swe, area_f), wj);
        if (n != f) {
            const double ratio = __ddiv_rn(area_f, a.area[n]);
            const double d_sd = __dmul_rn(__dmul_rn(del_depth, ratio), wj);
            const double d_swe = __dmul_rn(__dmul_rn(del_swe, ratio), wj);
            const double d_sd_m3 = __dmul_rn(__dmul_rn(del_depth, area_f), wj);
            if (n >= Tp) {  // partition edge: the share waits in the ghost accumulators for the reverse exchange
                double* g = a.gacc + (size_t)(n - Tp) * 4;
                atomicAdd(g + 0, d_sd);
                atomicAdd(g + 1, d_swe);
                atomicAdd(g + 2, d_sd_m3);
                atomicAdd(g + 3, m3);
            } else {
                const double nsd = __dadd_rn(__ldcg(a.sd + n), d_sd);
                a.sd[n] = nsd;
                a.swe[n] = __dadd_rn(__ldcg(a.swe + n), d_swe);
                a.sdv[n] = __ddiv_rn(nsd, cf);  // the DONOR's slope, as written (snow_slide.cpp:297)
                a.dsd[n] = __dadd_rn(__ldcg(a.dsd + n), d_sd_m3);
                a.dmass[n] = __dadd_rn(__ldcg(a.dmass + n), m3);
            }
        }
        out_mass = __dadd_rn(out_mass, m3);
    }
    a.sd[f] = maxD;
    a.sdv[f] = __ddiv_rn(maxD, cf);
    a.swe[f] = __ddiv_rn(__dmul_rn(swe, maxD), snow);
    a.dsd[f] = __dsub_rn(__ldcg(a.dsd + f), __dmul_rn(del_depth, area_f));
    a.dmass[f] = __dsub_rn(__ldcg(a.dmass + f), __dmul_rn(del_swe, area_f));
    return !(fabs(__dsub_rn(orig_mass, out_mass)) > 0.0001);
}

constexpr int kSlideThreads = 256;

__global__ void __launch_bounds__(kSlideThreads, 1) slide_sweep_kernel(SlideArrays a, unsigned* bar) {
    unsigned target = 0;
    const int Tp = a.Tp;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    int* live = a.list[0];
    // ---- start candidates; every other face is settled until a deposit can reach it
    for (int p = tid; p < Tp; p += nth) {
        const bool real = a.perm[p] >= 0;
        const double k = real ? __dadd_rn(a.cz[p], a.sdv[p]) : 0.0;
        a.key[p] = k;
        const bool cand = real && a.sd[p] > a.maxD[p];
        a.stamp[p] = cand ? 0 : 1;
        a.queued[p] = 0;
        if (cand) live[slide_reserve(a.cnt + 3)] = p;
    }
    grid_barrier(bar, target);
    // ---- frontier expansion: the live set = candidates and, transitively, their later-ordered neighbours.  The frontier of a
    // round is the segment of the live list the previous round appended, so one cursor serves both.
    int r = 0;
    for (int seg0 = 0, seg1 = __ldcg(a.cnt + 3); seg1 > seg0; ++r) {
        for (int idx = seg0 + tid; idx < seg1; idx += nth) {
            const int g = __ldcg(live + idx);
            const double kg = a.key[g];
            const int ig = a.perm[g];
            int f[3];
            bool take[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) f[j] = a.nbs[(size_t)j * Tp + g];
#pragma unroll
            for (int j = 0; j < 3; ++j) take[j] = f[j] != g && f[j] < Tp && slide_earlier(kg, ig, a.key[f[j] < Tp ? f[j] : g], a.perm[f[j] < Tp ? f[j] : g]);
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (take[j] && atomicCAS(a.stamp + f[j], 1, 0) == 1) live[slide_reserve(a.cnt + 3)] = f[j];
        }
        grid_barrier(bar, target);
        seg0 = seg1;
        seg1 = __ldcg(a.cnt + 3);
    }
    // ---- wavefront rounds over the live faces whose turn is still to come.  Event-driven: round 0 examines every live face; after
    // that a face is examined again only when a face within two edges of it has just had its turn (the only event that can unblock
    // it).  A round costs a few dependent L2 round trips plus the grid barrier, so everything a decision needs is loaded up front,
    // level by level (the face, its neighbours, their neighbours), and the wake-up atomics are issued together.
    int n_in = __ldcg(a.cnt + 3);
    const int* in = live;
    int fired = 0;
    bool bad = false;
    int w = 0;
    for (; n_in > 0; ++w) {
        const int stampv = w + 2;
        if (tid == 0) a.cnt[(w + 2) % 3] = 0;
        int* out = a.list[1 + w % 2];
        int* cur = a.cnt + (w + 1) % 3;
        for (int idx = tid; idx < n_in; idx += nth) {
            const int f = __ldcg(in + idx);
            // level 1: indexed by the face
            const int sf = __ldcg(a.stamp + f);
            int nb[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) nb[j] = a.nbs[(size_t)j * Tp + f];
            const double kf = a.key[f];
            const int i_f = a.perm[f];
            const bool active = __ldcg(a.sd + f) > a.maxD[f];
            if (sf != 0) continue;  // woken twice and already done
            // level 2: the edge neighbours (a missing / ghost neighbour reads the face itself and is masked out)
            bool v1[3];
            int s1[3], p1[3], m2[3][3];
            double k1[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                v1[j] = nb[j] != f && nb[j] < Tp;
                const int n = v1[j] ? nb[j] : f;
                s1[j] = __ldcg(a.stamp + n);
                k1[j] = a.key[n];
                p1[j] = a.perm[n];
#pragma unroll
                for (int k = 0; k < 3; ++k) m2[j][k] = a.nbs[(size_t)k * Tp + n];
            }
            // level 3: the faces two edges away
            bool v2[3][3];
            int s2[3][3], p2[3][3];
            double k2[3][3];
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int m = m2[j][k];
                    v2[j][k] = v1[j] && m != nb[j] && m != f && m < Tp;
                    const int mm = v2[j][k] ? m : f;
                    s2[j][k] = __ldcg(a.stamp + mm);
                    k2[j][k] = a.key[mm];
                    p2[j][k] = a.perm[mm];
                }
            bool wait = false;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (v1[j] && (s1[j] == 0 || s1[j] == stampv) && slide_earlier(k1[j], p1[j], kf, i_f)) wait = true;
            if (!wait && active) {  // a firing face also needs the earlier faces two edges away to be done
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (v2[j][k] && (s2[j][k] == 0 || s2[j][k] == stampv) && slide_earlier(k2[j][k], p2[j][k], kf, i_f)) wait = true;
            }
            if (wait) continue;  // parked: the blocking face wakes it when it has had its turn
            if (active) {
                if (!slide_fire(a, f)) bad = true;
                ++fired;
            }  // else every possible donor has had its turn: this face never fires
            a.stamp[f] = stampv;
            // wake the later-ordered live faces within two edges whose turn is still to come.  Their stamps were read above: such
            // a face cannot have had its turn in this round (this face was in its way), so 0 is still 0.
            bool w1[3], w2[3][3];
            int o1[3], o2[3][3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                w1[j] = v1[j] && s1[j] == 0 && slide_earlier(kf, i_f, k1[j], p1[j]);
#pragma unroll
                for (int k = 0; k < 3; ++k) w2[j][k] = v2[j][k] && s2[j][k] == 0 && slide_earlier(kf, i_f, k2[j][k], p2[j][k]);
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                o1[j] = w1[j] ? atomicExch(a.queued + nb[j], stampv) : stampv;
#pragma unroll
                for (int k = 0; k < 3; ++k) o2[j][k] = w2[j][k] ? atomicExch(a.queued + m2[j][k], stampv) : stampv;
            }
            int n_new = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                n_new += o1[j] != stampv;
#pragma unroll
                for (int k = 0; k < 3; ++k) n_new += o2[j][k] != stampv;
            }
            if (n_new) {
                int pos = atomicAdd(cur, n_new);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (o1[j] != stampv) out[pos++] = nb[j];
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (o2[j][k] != stampv) out[pos++] = m2[j][k];
                }
            }
        }
        grid_barrier(bar, target);
        n_in = __ldcg(cur);
        in = out;
    }
    if (fired) atomicAdd(a.cnt + 5, fired);
    if (bad) a.cnt[6] = 1;
    if (tid == 0) { a.cnt[4] = w; a.cnt[7] = r; }
}

// snow_slide::init per face (snow_slide.cpp:414-444)
__global__ void slide_init_kernel(int Tp, const int* __restrict__ perm, const double* __restrict__ slope, const double* __restrict__ canopy,
                                  double mult, double power, double* __restrict__ maxD, double* __restrict__ cosf) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp) return;
    if (perm[p] < 0) { maxD[p] = 1e300; cosf[p] = 1.0; return; }
    const double sl = slope[p];
    const double zc = canopy ? canopy[p] : 0.0;
    const double slope_deg = fmax(10.0, __ddiv_rn(__dmul_rn(sl, 180.0), kPi));
    const double c = fmax(0.001, cos(sl));
    maxD[p] = __dmul_rn(fmax(__dmul_rn(mult, pow(slope_deg, power)), zc), c);
    cosf[p] = c;
}

// first outer iteration of a run: private copies of the inputs (snow_slide.cpp:111-129); CHM order in, slot order out
__global__ void slide_begin_kernel(int Tp, const int* __restrict__ perm, const double* __restrict__ snowdepthavg,
                                   const double* __restrict__ snowdepthavg_vert, const double* __restrict__ swe_mm, SlideArrays a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Tp) return;
    const int i = perm[p];
    a.sd[p] = i < 0 ? 0.0 : snowdepthavg[i];
    a.sdv[p] = i < 0 ? 0.0 : snowdepthavg_vert[i];
    a.swe[p] = i < 0 ? 0.0 : __ddiv_rn(swe_mm[i], 1000.0);
    a.dsd[p] = 0.0;
    a.dmass[p] = 0.0;
}

// the reverse exchange's unpack for ONE partner (ghost_to_neighbors_communicate_variable, triangulation.cpp:2172-2182): the owner's
// four variables are SET to what the partner accumulated on its ghost copy; the host launches the partners in ascending rank order,
// so when two ranks hold the same face as a ghost the later one stays, as in the reference
__global__ void slide_unpack_kernel(int n, const int* __restrict__ send_slot, const double* __restrict__ rev, double* __restrict__ xfer, int Tp) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int p = send_slot[e];
#pragma unroll
    for (int k = 0; k < 4; ++k) xfer[(size_t)k * Tp + p] = rev[(size_t)e * 4 + k];
}

// after the exchange (snow_slide.cpp:341-360): absorb what the partners routed to my faces, publish, accumulate the sums; counts
// the faces that received transport (another outer iteration follows while any rank has one)
__global__ void slide_absorb_kernel(SlideArrays a, const double* __restrict__ xfer, double* __restrict__ sum_sd, double* __restrict__ sum_mass,
                                    int* __restrict__ moved) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.Tp || a.perm[p] < 0) return;
    const int Tp = a.Tp;
    double x_sd = 0.0, x_swe = 0.0, x_dsd = 0.0, x_dswe = 0.0;
    if (xfer) { x_sd = xfer[p]; x_swe = xfer[(size_t)Tp + p]; x_dsd = xfer[(size_t)2 * Tp + p]; x_dswe = xfer[(size_t)3 * Tp + p]; }
    a.sd[p] = __dadd_rn(a.sd[p], x_sd);
    a.sdv[p] = __dadd_rn(a.sdv[p], __ddiv_rn(x_sd, a.cosf[p]));
    a.swe[p] = __dadd_rn(a.swe[p], x_swe);
    const double d1 = __dadd_rn(a.dsd[p], x_dsd), d2 = __dadd_rn(a.dmass[p], x_dswe);
    a.dsd[p] = d1;
    a.dmass[p] = d2;
    sum_sd[p] = __dadd_rn(sum_sd[p], d1);
    sum_mass[p] = __dadd_rn(sum_mass[p], d2);
    if (x_dsd > 0.0) atomicAdd(moved, 1);
}

}  // namespace pbsm3d

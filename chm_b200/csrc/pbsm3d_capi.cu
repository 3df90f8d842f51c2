// libpbsm3d_b200.so — C-ABI (include/pbsm3d.h) and host orchestration of one PBSM3D timestep on one B200.
// One process per GPU.  Ghost-face halos and global reductions travel through peer memory over NVLink (cudaIpc arenas;
// the halos of the two solves are written by the solver kernels themselves); NCCL bootstraps and is the fallback
// transport.  No CPU fallback.
//
// A step is enqueued optimistically: assembly, a predicted number of line Gauss-Seidel sweeps with residual
// checks, flux integration, the deposition right-hand side, a predicted number of SOR sweeps (or Chebyshev / CG
// iterations), the drift update and the export to CHM order all go onto one stream, each kernel guarded by
// device-resident flags (suspension converged? deposition present? deposition converged?), and the host synchronises
// ONCE at the end.  Only when a prediction was too short does the host add more sweeps / iterations and re-enqueue the tail.
#include "../../include/pbsm3d.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <string>
#include <vector>

#include "pbsm3d_kernels.cuh"
#include "pbsm3d_wind.cuh"
#include "pbsm3d_snobal.cuh"
#include "pbsm3d_slide.cuh"

using namespace pbsm3d;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CU(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(PBSM3D_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                             std::to_string(__LINE__) + ")");                                \
    } while (0)
#define NC(call)                                                                                             \
    do {                                                                                                     \
        ncclResult_t e_ = (call);                                                                            \
        if (e_ != ncclSuccess)                                                                               \
            return fail(PBSM3D_ERR_NCCL, std::string(#call) + ": " + ncclGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                             std::to_string(__LINE__) + ")");                                \
    } while (0)
#define TRY(call)              \
    do {                       \
        int rc_ = (call);      \
        if (rc_) return rc_;   \
    } while (0)

inline int cdiv(size_t a, int b) { return (int)((a + b - 1) / b); }
inline int align_up(int a, int b) { return (a + b - 1) / b * b; }

// every kernel launch goes through here so the step can report how many of our kernels it launched
#define LAUNCH(h_, kernel_, grid_, block_, ...)                                  \
    do {                                                                         \
        ++(h_)->n_launch;                                                        \
        kernel_<<<(grid_), (block_), 0, (h_)->stream>>>(__VA_ARGS__);            \
    } while (0)
#define LAUNCH_ON(h_, stream_, kernel_, grid_, block_, ...)                      \
    do {                                                                         \
        ++(h_)->n_launch;                                                        \
        kernel_<<<(grid_), (block_), 0, (stream_)>>>(__VA_ARGS__);               \
    } while (0)

struct Partner {
    int rank;
    int send_off, send_cnt;  // into the concatenated send list
    int recv_off, recv_cnt;  // ghost block [recv_off, recv_off+recv_cnt)
};

constexpr int kMaxColours = 8;
constexpr int kIn = 9, kOut = 9;  // forcing arrays / output arrays that cross the C-ABI (pbsm3d_forcing, pbsm3d_outputs)
constexpr int kChunks = 4;  // forcing chunks of the host-buffer entry point (H2D of chunk c+1 overlaps assembly of chunk c)

}  // namespace

struct pbsm3d_handle {
    pbsm3d_config cfg;
    DevConfig dc;
    DevMesh dm;
    SuspSystem ss;
    int device = 0;
    int T = 0, Tp = 0, S = 0, nG = 0, L = 0;
    int64_t G = 0, gstart_id = 0;
    size_t N = 0;   // L * Tp coefficient rows
    size_t NS = 0;  // L * S   ghost-extended vector length
    int n_colours = 0;
    int cstart[kMaxColours] = {0}, ccount[kMaxColours] = {0};
    cudaStream_t stream = nullptr;                  // compute
    cudaStream_t s_in = nullptr, s_out = nullptr;   // H2D of the forcing / export + D2H of the outputs
    cudaEvent_t ev_in[kChunks] = {nullptr}, ev_asm = nullptr, ev_flux = nullptr, ev_out = nullptr, ev_prov = nullptr;
    std::vector<void*> allocs;

    // mesh / static (slot order)
    int *perm = nullptr, *iperm = nullptr, *nbs = nullptr, *gstart = nullptr, *gcnt = nullptr;
    double *nx = nullptr, *ny = nullptr, *elen = nullptr, *area = nullptr, *cx = nullptr, *cy = nullptr, *cz = nullptr, *dx = nullptr;
    double *canopy = nullptr, *lai = nullptr, *stalk_n = nullptr, *stalk_dv = nullptr;
    unsigned char* water = nullptr;
    double *ddiag = nullptr, *doff = nullptr, *dinv = nullptr;
    // forcing (own device copies for the host-pointer entry point), CHM order
    double* forcing_buf[kIn] = {nullptr};
    DevForcing last_forcing{};
    double last_dt = 0.0;
    // solution / work
    double* x = nullptr;         // [L][S] suspended concentration, ghost-extended
    double* kry[7] = {nullptr};  // r, rhat, p, v, ph, sh, t (allocated on first Krylov use), each [L][S]
    // per-face outputs and deposition work (slot order)
    double *Qsusp = nullptr /*[S]*/, *Qsubl = nullptr, *Qsubl_mass = nullptr, *sum_subl = nullptr, *drift_mass = nullptr,
           *sum_drift = nullptr, *more_avail = nullptr;
    double *drhs = nullptr, *drhsS = nullptr, *cg_r = nullptr, *cg_p = nullptr /*[S]*/, *cg_Ap = nullptr;
    double *qA = nullptr, *qB = nullptr;  // [S] deposition iterate (Chebyshev ping-pong; CG uses qA)
    double* offS = nullptr;
    // Chebyshev: spectrum bounds of D^-1 A (static matrix, estimated once) and the coefficient sequence
    bool cheb_ready = false;
    double cheb_lmin = 0.0, cheb_lmax = 0.0;
    std::vector<double> cheb_a, cheb_c;
    int cheb_kest = 0, cheb_enqueued = 0, pred_dep = 0;
    // multicolour SOR on the deposition system: Young's omega from the same spectrum estimate
    double sor_omega = 0.0;
    int sor_kest = 0, sor_enqueued = 0, pred_sor = 0;
    int* ghost_key = nullptr;                       // [nG] (colour, rank) key of every ghost face: the global SOR order
    unsigned long long sor_epoch = 1ull << 40;      // sweep numbers = tags of the SOR ghost entries (disjoint from the Chebyshev tags)
    int lanczos_steps = 0;
    int n_syncs = 0;
    size_t l2_persist_max = 0, l2_window_max = 0;
    bool trace = false;
    double* out_stage = nullptr;  // [kOut][T] CHM-ordered outputs on their way to host buffers
    double* scratch = nullptr;    // inspection getters
    size_t scratch_n = 0;
    // reductions
    double *partial = nullptr, *red = nullptr;
    Scalars* sc = nullptr;
    Scalars* h_sc = nullptr;  // pinned
    // comm
    ncclComm_t comm = nullptr;
    int rank = 0, n_ranks = 1;
    std::vector<Partner> partners;
    int n_send = 0;
    int *send_slot = nullptr, *send_boff = nullptr, *send_cnt = nullptr, *send_pos = nullptr;
    double *sendbuf = nullptr, *recvbuf = nullptr;
    // peer-memory transport (cudaIpc arenas over NVLink); NCCL stays the fallback transport and the setup plumbing
    bool peer = false;
    unsigned char* arena = nullptr;                 // my exported arena: flags | all-reduce slots | 2 halo staging buffers
    std::vector<void*> peer_base;                   // [n_ranks] mapped arenas of the other ranks (nullptr for me)
    PeerTable* d_pt = nullptr;
    double** stage_remote[2] = {nullptr, nullptr};  // [parity][n_send] where each send entry lands on its partner
    double* stage_local[2] = {nullptr, nullptr};
    unsigned* halo_ticket = nullptr;
    unsigned long long halo_epoch = 0;
    unsigned long long* d_ar_epoch = nullptr;       // device-resident all-reduce counter (PeerTable::ar_epoch)
    int halo_ops = 0;                               // halo exchanges enqueued by the step in flight
    // halos carried by the solver kernels themselves (HaloLink): x of the line sweeps, q of the Chebyshev iteration
    bool fused_halo = false;
    int halo_fused_ops = 0;
    std::vector<long long> give_ids;                // handshake: global ids of my faces the partners need, per partner block
    std::vector<int> need_matrix;                   // M[r][q] = ghosts rank r needs from rank q
    std::vector<unsigned char> is_boundary;         // [T] a partner needs this face
    int nb[kMaxColours] = {0}, boff[kMaxColours] = {0}, nb_total = 0, n_entries = 0, nGp = 0;
    int *bptr = nullptr, *rstride = nullptr;
    double** x_remote[3] = {nullptr, nullptr, nullptr};
    ulonglong2** q_remote[3] = {nullptr, nullptr, nullptr};  // tagged 16-byte entries (TaggedLink)
    double* xg[3] = {nullptr, nullptr, nullptr};    // my ghost buffers (in the arena)
    ulonglong2* qg[3] = {nullptr, nullptr, nullptr};
    double* g_zero = nullptr;                       // [L][nGp] zeros: the ghosts of the first iteration of a solve
    unsigned long long **xflag_remote = nullptr, **qflag_remote = nullptr;
    unsigned long long *xflag_local = nullptr, *qflag_local = nullptr;
    unsigned long long xh_epoch = 0, qh_epoch = 0;  // iteration numbers of the two channels (monotonic)
    bool x_first = true;                            // the next sweep is the first of a solve (x = 0: ghosts read as 0)
    // providers of U_2m_above_srf / fetch (scale_wind_vert, fetchr): vegetation in slot order, work vector, centre grid
    double *wv_canopy = nullptr, *wv_lai = nullptr;  // [Tp] or null (the mesh carries no such parameter)
    double* wv_u = nullptr;                           // [S] point-scaled wind, ghost-extended
    bool grid_ready = false;
    CellGrid grid{};
    int* grid_start = nullptr;
    double2 *grid_xy = nullptr, *grid_zc = nullptr;
    bool geographic = false;                          // pbsm3d_mesh.is_geographic: dx of the deposition matrix is a haversine distance
    bool providers_on = false;                        // pbsm3d_set_providers: the step derives missing inputs itself
    pbsm3d_wind_config wind_cfg{};
    float ms_providers = 0.f;
    // predictions carried from step to step (iteration counts only; every solve still starts from x0 = 0)
    int pred_sweeps = 0, pred_cg = 0;
    double sweep_rate2 = 0.0;  // observed per-sweep contraction of ||r||^2
    // timing
    cudaEvent_t ev[6] = {nullptr};
    cudaEvent_t ev_sw[3] = {nullptr};
    int sweeps_timed = 0, sweeps_timed32 = 0;
    // persistent (cooperative) solver kernels: single rank
    bool persistent = false;
    // The persistent line solver's active set (pbsm3d_kernels.cuh): PBSM3D_ACTIVE_SET=0 never, =1 always; default: on the steps
    // where at most kActiveSetMaxSeeded of the faces have a non-zero right-hand side (beyond that the set covers the mesh within a
    // few sweeps and the measured gain is ~1 %, profiles/r2ae_summary.md)
    int active_set = 2;
    bool sor_persistent = false;        // the deposition solve too: only while its working set stays in the L2 (else per-pass launches
                                        // stream better: 17 vs 21 ms on 10 M faces, profiles/r2b)
    unsigned* grid_bar = nullptr;       // [2] grid-barrier counters (suspension, deposition)
    int gs_grid = 0, sor_grid = 0;
    void* gs_fn = nullptr;
    const void* sor_fn = nullptr;
    bool sor_resident = false;          // sor_resident_kernel: static face data on chip for the whole solve (2 colours, <= 7 faces/thread/colour)
    size_t sor_smem = 0;
    int sor_threads = 1024;
    int plan_n32 = 0, plan_nx32 = 0;
    float* xf = nullptr;                // [L][S] fp32 storage of the iterate for the sweeps furthest from convergence
    int asm_nw = 0, asm_grid = 1;  // assembly: warps per block of the tile kernel (0 = column kernel), persistent grid size
    size_t asm_smem = 0;
    void* asm_fn = nullptr;
    int asm_threads = 128;
    FaceRecs recs{};  // per-face records between the prelude kernel and the row kernel
    int pred_n32 = 0;  // leading sweeps of the next solve that may stream fp32 coefficient copies
    bool have_system = false;
    long long n_launch = 0;
    // snow_slide (pbsm3d_slide_init / _run)
    double* slope = nullptr;            // [Tp] face slope (rad), computed with the rest of the geometry
    bool slide_ready = false;
    pbsm3d_slide_config slide_cfg{};
    SlideArrays sl{};
    double *sl_sum_sd = nullptr, *sl_sum_mass = nullptr, *sl_xfer = nullptr, *sl_rev = nullptr, *sl_in = nullptr;
    int* sl_moved = nullptr;
    unsigned* sl_bar = nullptr;
    int sl_grid = 0;
    double* sno_stage = nullptr;  // [23][T] staging of a host-side snowpack (pbsm3d_apply_drift / _avalanche with host buffers)

    template <typename U>
    int alloc(U** p, size_t n) {
        if (n == 0) n = 1;
        void* q_ = nullptr;
        CU(cudaMalloc(&q_, n * sizeof(U)));
        allocs.push_back(q_);
        *p = (U*)q_;
        return 0;
    }
    template <typename U>
    int alloc_zero(U** p, size_t n) {
        TRY(alloc(p, n));
        CU(cudaMemsetAsync(*p, 0, std::max<size_t>(n, 1) * sizeof(U), stream));
        return 0;
    }
    void release(void* p) {
        for (size_t k = 0; k < allocs.size(); ++k)
            if (allocs[k] == p) { allocs.erase(allocs.begin() + k); break; }
        cudaFree(p);
    }
};

namespace {

int upload(pbsm3d_handle* h, void* dst, const void* src, size_t bytes) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return 0;
}
inline int red_grid(size_t n) { return std::max(1, std::min(kRedBlocks, cdiv(n, kRedThreads))); }
inline bool fused(const pbsm3d_handle* h) { return h->n_ranks == 1; }

int allreduce(pbsm3d_handle* h, double* buf, int n, bool is_max) {
    if (h->n_ranks == 1) return 0;
    if (h->peer) {
        LAUNCH(h, peer_allreduce_kernel, 1, 32, buf, n, is_max ? 1 : 0, h->d_pt);
        return 0;
    }
    NC(ncclAllReduce(buf, buf, n, ncclDouble, is_max ? ncclMax : ncclSum, h->comm, h->stream));
    return 0;
}
int sync_stream(pbsm3d_handle* h) {
    ++h->n_syncs;
    CU(cudaStreamSynchronize(h->stream));
    if (h->peer && h->h_sc->peer_error)
        return fail(PBSM3D_ERR_NCCL, "peer-memory halo: a partner rank did not arrive within the time-out");
    return 0;
}
int read_scalars(pbsm3d_handle* h) {
    CU(cudaMemcpyAsync(h->h_sc, h->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
    return sync_stream(h);
}

// L2 residency hint for the phase in flight: accesses to [base, base+bytes) are kept in the persisting part of the
// 126 MB L2 (as much as fits), everything else streams through.  bytes == 0 clears the window.
int l2_window(pbsm3d_handle* h, const void* base, size_t bytes) {
    if (h->l2_persist_max == 0) return 0;
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    if (bytes > 0) {
        const size_t nb = std::min(bytes, h->l2_window_max);
        attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
        attr.accessPolicyWindow.num_bytes = nb;
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, 0.9 * (double)h->l2_persist_max / (double)nb);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    CU(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    return 0;
}

// ---- colouring (host, once): a proper colouring of the owned faces' edge-adjacency graph ---------------------
// 2 colours when the dual graph is bipartite (every structured/alternating-diagonal mesh, many others), else
// greedy first-fit in CHM order, which needs at most 4 colours on a graph of maximum degree 3.
void colour_faces(int T, const int32_t* neigh, std::vector<int>& colour, int& n_colours) {
    colour.assign(T, -1);
    bool bipartite = true;
    std::vector<int> stack;
    for (int s = 0; s < T && bipartite; ++s) {
        if (colour[s] >= 0) continue;
        colour[s] = 0;
        stack.push_back(s);
        while (!stack.empty() && bipartite) {
            int i = stack.back();
            stack.pop_back();
            for (int j = 0; j < 3; ++j) {
                int n = neigh[(size_t)i * 3 + j];
                if (n < 0 || n >= T) continue;
                if (colour[n] < 0) { colour[n] = 1 - colour[i]; stack.push_back(n); }
                else if (colour[n] == colour[i]) { bipartite = false; break; }
            }
        }
    }
    if (bipartite) { n_colours = T > 1 ? 2 : 1; return; }
    colour.assign(T, -1);
    n_colours = 0;
    for (int i = 0; i < T; ++i) {
        unsigned used = 0;
        for (int j = 0; j < 3; ++j) {
            int n = neigh[(size_t)i * 3 + j];
            if (n >= 0 && n < T && colour[n] >= 0) used |= 1u << colour[n];
        }
        int c = 0;
        while (used & (1u << c)) ++c;
        colour[i] = c;
        n_colours = std::max(n_colours, c + 1);
    }
}

// ---- halo (reference: triangulation::ghost_neighbors_communicate_variable, triangulation.cpp:1976-2079) -------
// v is a ghost-extended [nl][S] vector on the device; the ghost tails v[z*S + Tp + g] are refreshed in place.
int halo_exchange(pbsm3d_handle* h, double* v, int nl) {
    if (h->n_ranks == 1 || h->partners.empty()) return 0;
    ++h->halo_ops;
    if (h->peer) {
        const unsigned long long epoch = ++h->halo_epoch;
        const int par = (int)(epoch & 1ull);
        const size_t ns = (size_t)h->n_send * nl, ng = (size_t)h->nG * nl;
        LAUNCH(h, halo_push_kernel, std::max(1, std::min(cdiv(ns, 256), 148 * 4)), 256, h->n_send, nl, h->S, h->send_slot, h->send_cnt,
               h->stage_remote[par], v, h->d_pt, epoch, h->halo_ticket);
        LAUNCH(h, halo_wait_unpack_kernel, std::max(1, std::min(cdiv(ng, 256), 148 * 4)), 256, h->nG, nl, h->L, h->Tp, h->S, h->gstart,
               h->gcnt, h->stage_local[par], v, h->d_pt, epoch);
        return 0;
    }
    if (h->n_send > 0) {
        size_t total = (size_t)h->n_send * nl;
        int blocks = std::min(cdiv(total, 256), 148 * 8);
        LAUNCH(h, halo_pack_kernel, blocks, 256, h->n_send, nl, h->S, h->send_slot, h->send_boff, h->send_cnt, h->send_pos, v,
               h->sendbuf);
    }
    NC(ncclGroupStart());
    for (const Partner& p : h->partners) {
        if (p.send_cnt > 0)
            NC(ncclSend(h->sendbuf + (size_t)p.send_off * nl, (size_t)p.send_cnt * nl, ncclDouble, p.rank, h->comm, h->stream));
        if (p.recv_cnt > 0)
            NC(ncclRecv(h->recvbuf + (size_t)p.recv_off * nl, (size_t)p.recv_cnt * nl, ncclDouble, p.rank, h->comm, h->stream));
    }
    NC(ncclGroupEnd());
    if (h->nG > 0) {
        size_t total = (size_t)h->nG * nl;
        int blocks = std::min(cdiv(total, 256), 148 * 8);
        LAUNCH(h, halo_unpack_kernel, blocks, 256, h->nG, nl, h->Tp, h->S, h->gstart, h->gcnt, h->recvbuf, v);
    }
    return 0;
}

// ---- peer-memory transport ----------------------------------------------------------------------------------
// Arena layout (identical on every rank up to the staging size): halo flags | all-reduce flags | all-reduce slots |
// staging[2][nGp * L].  M[r][q] = ghosts rank r needs from rank q, known to everyone, gives each rank the place of
// its block in every partner's staging buffer without another exchange.
constexpr size_t kArenaHaloFlag = 0, kArenaArFlag = kMaxRanks * 8, kArenaArSlots = 2 * kMaxRanks * 8,
                 kArenaXFlag = kArenaArSlots + 2 * kMaxRanks * 4 * 8, kArenaQFlag = kArenaXFlag + kMaxRanks * 8,
                 kArenaStage = kArenaQFlag + kMaxRanks * 8;
// after the two staging buffers: xg[3][L][gp] and qg[3][gp], the ghost buffers of the fused channels (gp = padded
// ghost count of the arena's owner)
struct PeerHello {
    cudaIpcMemHandle_t mem;
    int ok, pad;
};
void close_peer(pbsm3d_handle* h) {
    for (void* b : h->peer_base)
        if (b) cudaIpcCloseMemHandle(b);
    h->peer_base.clear();
    h->peer = false;
}
int setup_peer(pbsm3d_handle* h, const std::vector<int>& M, const std::vector<int>& sslot) {
    const int P = h->n_ranks, me = h->rank, L = h->L;
    const char* want = getenv("PBSM3D_HALO");
    const bool verbose = getenv("PBSM3D_VERBOSE") != nullptr;
    if (want && std::string(want) == "nccl") return 0;
    if (P > kMaxRanks) return 0;
    auto ghosts_of = [&](int r) { int n = 0; for (int q = 0; q < P; ++q) n += M[(size_t)r * P + q]; return n; };
    auto gp = [&](int r) { return (size_t)align_up(std::max(ghosts_of(r), 1), 32); };
    auto stage_elems = [&](int r) { return gp(r) * L; };
    auto off_xg = [&](int r) { return kArenaStage + 2 * stage_elems(r) * sizeof(double); };
    auto off_qg = [&](int r) { return off_xg(r) + 3 * stage_elems(r) * sizeof(double); };
    const size_t arena_bytes = off_qg(me) + 3 * gp(me) * sizeof(ulonglong2);
    TRY(h->alloc(&h->arena, arena_bytes));
    CU(cudaMemsetAsync(h->arena, 0, arena_bytes, h->stream));
    PeerHello hello;
    std::memset(&hello, 0, sizeof(hello));
    hello.ok = cudaIpcGetMemHandle(&hello.mem, h->arena) == cudaSuccess ? 1 : 0;
    if (!hello.ok) cudaGetLastError();
    unsigned char *d_hello = nullptr, *d_all = nullptr;
    TRY(h->alloc(&d_hello, sizeof(PeerHello)));
    TRY(h->alloc(&d_all, sizeof(PeerHello) * P));
    TRY(upload(h, d_hello, &hello, sizeof(hello)));
    NC(ncclAllGather(d_hello, d_all, sizeof(PeerHello), ncclUint8, h->comm, h->stream));
    std::vector<PeerHello> all(P);
    CU(cudaMemcpyAsync(all.data(), d_all, sizeof(PeerHello) * P, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int ok = 1;
    for (int q = 0; q < P; ++q) ok = ok && all[q].ok;
    h->peer_base.assign(P, nullptr);
    for (int q = 0; q < P && ok; ++q) {
        if (q == me) continue;
        void* b = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&b, all[q].mem, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            if (verbose) fprintf(stderr, "[pbsm3d] rank %d: cudaIpcOpenMemHandle(rank %d): %s\n", me, q, cudaGetErrorString(e));
            cudaGetLastError();
            ok = 0;
        } else {
            h->peer_base[q] = b;
        }
    }
    // everyone or no one: a rank that cannot map a peer sends every rank back to NCCL
    double* d_ok = nullptr;
    TRY(h->alloc(&d_ok, 1));
    double okd = ok;
    TRY(upload(h, d_ok, &okd, sizeof(double)));
    NC(ncclAllReduce(d_ok, d_ok, 1, ncclDouble, ncclMin, h->comm, h->stream));
    CU(cudaMemcpyAsync(&okd, d_ok, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->release(d_hello);
    h->release(d_all);
    h->release(d_ok);
    if (okd < 0.5) {
        if (verbose && me == 0) fprintf(stderr, "[pbsm3d] peer-memory transport unavailable: halos and reductions go over NCCL\n");
        close_peer(h);
        return 0;
    }
    auto base_of = [&](int q) { return q == me ? h->arena : (unsigned char*)h->peer_base[q]; };
    PeerTable pt;
    std::memset(&pt, 0, sizeof(pt));
    pt.n_ranks = P;
    pt.me = me;
    pt.n_partners = (int)h->partners.size();
    const int np = std::max(pt.n_partners, 1);
    std::vector<unsigned long long*> xfr(np, nullptr), qfr(np, nullptr);
    for (int i = 0; i < pt.n_partners; ++i) {
        const int r = h->partners[i].rank;
        pt.partner_rank[i] = r;
        pt.halo_flag_remote[i] = (unsigned long long*)(base_of(r) + kArenaHaloFlag) + me;
        xfr[i] = (unsigned long long*)(base_of(r) + kArenaXFlag) + me;
        qfr[i] = (unsigned long long*)(base_of(r) + kArenaQFlag) + me;
    }
    pt.halo_flag_local = (unsigned long long*)(h->arena + kArenaHaloFlag);
    for (int q = 0; q < P; ++q) {
        pt.ar_slots_remote[q] = (double*)(base_of(q) + kArenaArSlots);
        pt.ar_flag_remote[q] = (unsigned long long*)(base_of(q) + kArenaArFlag) + me;
    }
    pt.ar_slots_local = (double*)(h->arena + kArenaArSlots);
    pt.ar_flag_local = (unsigned long long*)(h->arena + kArenaArFlag);
    pt.error = &h->sc->peer_error;
    TRY(h->alloc_zero(&h->d_ar_epoch, 1));
    pt.ar_epoch = h->d_ar_epoch;
    const char* to = getenv("PBSM3D_PEER_TIMEOUT_MS");
    pt.timeout_ns = (unsigned long long)(to ? std::max(1L, atol(to)) : 20000L) * 1000000ull;
    TRY(h->alloc(&h->d_pt, 1));
    TRY(upload(h, h->d_pt, &pt, sizeof(pt)));
    auto roff_at = [&](int r) {  // where the ghosts rank r needs from me start in its ghost list
        size_t roff = 0;
        for (int q = 0; q < me; ++q) roff += M[(size_t)r * P + q];
        return roff;
    };
    for (int par = 0; par < 2; ++par) {
        h->stage_local[par] = (double*)(h->arena + kArenaStage) + (size_t)par * stage_elems(me);
        std::vector<double*> dst((size_t)std::max(h->n_send, 1), nullptr);
        for (const Partner& p : h->partners) {
            const size_t roff = roff_at(p.rank);
            double* stage = (double*)(base_of(p.rank) + kArenaStage) + (size_t)par * stage_elems(p.rank);
            for (int k = 0; k < p.send_cnt; ++k) dst[p.send_off + k] = stage + roff * L + k;
        }
        TRY(h->alloc(&h->stage_remote[par], dst.size()));
        TRY(upload(h, h->stage_remote[par], dst.data(), dst.size() * sizeof(double*)));
    }
    TRY(h->alloc_zero(&h->halo_ticket, 1));

    // ---- fused channels: boundary CSR + where every send entry lands in the partner's ghost buffers
    h->nGp = (int)gp(me);
    for (int b = 0; b < 3; ++b) {
        h->xg[b] = (double*)(h->arena + off_xg(me)) + (size_t)b * stage_elems(me);
        h->qg[b] = (ulonglong2*)(h->arena + off_qg(me)) + (size_t)b * gp(me);
    }
    h->xflag_local = (unsigned long long*)(h->arena + kArenaXFlag);
    h->qflag_local = (unsigned long long*)(h->arena + kArenaQFlag);
    TRY(h->alloc(&h->xflag_remote, np));
    TRY(h->alloc(&h->qflag_remote, np));
    TRY(upload(h, h->xflag_remote, xfr.data(), np * sizeof(void*)));
    TRY(upload(h, h->qflag_remote, qfr.data(), np * sizeof(void*)));
    {
        const int nbt = h->nb_total, ns = h->n_send;
        std::vector<int> bidx(std::max(ns, 1), 0), bptr((size_t)nbt + 1, 0);
        for (int k = 0; k < ns; ++k) {
            int c = 0;
            while (c + 1 < h->n_colours && sslot[k] >= h->cstart[c + 1]) ++c;
            const int i = sslot[k] - h->cstart[c];
            if (i < 0 || i >= h->nb[c]) return fail(PBSM3D_ERR_INVALID, "internal: a sent face is not in the boundary block of its colour");
            bidx[k] = h->boff[c] + i;
            bptr[(size_t)bidx[k] + 1]++;
        }
        for (int i = 0; i < nbt; ++i) bptr[(size_t)i + 1] += bptr[i];
        std::vector<int> fillp(bptr.begin(), bptr.end() - 1), stride(std::max(ns, 1), 0);
        std::vector<double*> xr[3];
        std::vector<ulonglong2*> qr[3];
        for (int b = 0; b < 3; ++b) { xr[b].assign(std::max(ns, 1), nullptr); qr[b].assign(std::max(ns, 1), nullptr); }
        for (const Partner& p : h->partners) {
            const int r = p.rank;
            const size_t roff = roff_at(r);
            for (int k = 0; k < p.send_cnt; ++k) {
                const int pos = fillp[bidx[p.send_off + k]]++;
                const size_t g = roff + k;
                stride[pos] = (int)gp(r);
                for (int b = 0; b < 3; ++b) {
                    xr[b][pos] = (double*)(base_of(r) + off_xg(r)) + (size_t)b * stage_elems(r) + g;
                    qr[b][pos] = (ulonglong2*)(base_of(r) + off_qg(r)) + (size_t)b * gp(r) + g;
                }
            }
        }
        h->n_entries = ns;
        TRY(h->alloc(&h->bptr, bptr.size()));
        TRY(upload(h, h->bptr, bptr.data(), bptr.size() * sizeof(int)));
        TRY(h->alloc(&h->rstride, stride.size()));
        TRY(upload(h, h->rstride, stride.data(), stride.size() * sizeof(int)));
        for (int b = 0; b < 3; ++b) {
            TRY(h->alloc(&h->x_remote[b], xr[b].size()));
            TRY(upload(h, h->x_remote[b], xr[b].data(), xr[b].size() * sizeof(double*)));
            TRY(h->alloc(&h->q_remote[b], qr[b].size()));
            TRY(upload(h, h->q_remote[b], qr[b].data(), qr[b].size() * sizeof(ulonglong2*)));
        }
        TRY(h->alloc_zero(&h->g_zero, stage_elems(me)));
        CU(cudaStreamSynchronize(h->stream));  // the host vectors above go out of scope
    }
    CU(cudaStreamSynchronize(h->stream));
    h->peer = true;
    h->fused_halo = !(want && std::string(want) == "staged");
    if (verbose && me == 0)
        fprintf(stderr, "[pbsm3d] peer-memory transport: %d ranks, arena %zu KB per rank, halos %s\n", P, arena_bytes >> 10,
                h->fused_halo ? "inside the solver kernels" : "staged (push + wait/unpack launches)");
    return 0;
}

// Communicator + the negotiation of what each rank sends (setup_nearest_neighbor_communication,
// triangulation.cpp:1845-1945).  Runs before the slot order is chosen: the faces a partner needs (h->is_boundary) go
// first inside their colour class.
int comm_handshake(pbsm3d_handle* h, const pbsm3d_mesh* mesh, const pbsm3d_comm* comm) {
    const int P = h->n_ranks, me = h->rank, nG = h->nG;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(&id, comm->nccl_unique_id, sizeof(id));
    NC(ncclCommInitRank(&h->comm, P, id, me));

    // ghost blocks per owner (ghosts are sorted by global id ⇒ contiguous per owner, triangulation.cpp:1784-1829)
    std::vector<int> need_cnt(P, 0), need_off(P, 0);
    for (int g = 0; g < nG; ++g) {
        int o = mesh->ghost_owner[g];
        if (o < 0 || o >= P || o == me) return fail(PBSM3D_ERR_INVALID, "ghost_owner out of range");
        if (g > 0 && mesh->ghost_owner[g] < mesh->ghost_owner[g - 1])
            return fail(PBSM3D_ERR_INVALID, "ghost faces must be sorted by global id (owner blocks contiguous)");
        need_cnt[o]++;
    }
    for (int q = 1; q < P; ++q) need_off[q] = need_off[q - 1] + need_cnt[q - 1];
    // everyone learns the P×P matrix of needs: M[r][q] = ghosts rank r needs from rank q
    int* d_cnt = nullptr;
    int* d_all = nullptr;
    TRY(h->alloc(&d_cnt, P));
    TRY(h->alloc(&d_all, (size_t)P * P));
    TRY(upload(h, d_cnt, need_cnt.data(), P * sizeof(int)));
    NC(ncclAllGather(d_cnt, d_all, P, ncclInt32, h->comm, h->stream));
    std::vector<int>& M = h->need_matrix;
    M.assign((size_t)P * P, 0);
    CU(cudaMemcpyAsync(M.data(), d_all, M.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    // tell each owner which of its faces we need
    int n_send = 0;
    for (int r = 0; r < P; ++r) n_send += M[(size_t)r * P + me];
    h->n_send = n_send;
    long long *d_need = nullptr, *d_give = nullptr;
    TRY(h->alloc(&d_need, (size_t)nG));
    TRY(h->alloc(&d_give, (size_t)n_send));
    std::vector<long long> need_ids(std::max(nG, 1));
    for (int g = 0; g < nG; ++g) need_ids[g] = mesh->global_id[h->T + g];
    TRY(upload(h, d_need, need_ids.data(), nG * sizeof(long long)));
    NC(ncclGroupStart());
    int soff = 0;
    for (int r = 0; r < P; ++r) {
        if (r == me) continue;
        int sc_ = M[(size_t)r * P + me], rc_ = need_cnt[r];
        if (rc_ > 0) NC(ncclSend(d_need + need_off[r], rc_, ncclInt64, r, h->comm, h->stream));
        if (sc_ > 0) NC(ncclRecv(d_give + soff, sc_, ncclInt64, r, h->comm, h->stream));
        if (sc_ > 0 || rc_ > 0) h->partners.push_back({r, soff, sc_, need_off[r], rc_});
        soff += sc_;
    }
    NC(ncclGroupEnd());
    h->give_ids.assign(std::max(n_send, 1), 0);
    CU(cudaMemcpyAsync(h->give_ids.data(), d_give, n_send * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->release(d_cnt);
    h->release(d_all);
    h->release(d_need);
    h->release(d_give);
    h->is_boundary.assign(h->T, 0);
    for (int k = 0; k < n_send; ++k) {
        const long long loc = h->give_ids[k] - h->gstart_id;
        if (loc < 0 || loc >= h->T) return fail(PBSM3D_ERR_INVALID, "a partner asked for a face this rank does not own");
        h->is_boundary[(size_t)loc] = 1;
    }
    return 0;
}

// Send lists in slot numbering + the transport (after the slot order exists).
int setup_comm(pbsm3d_handle* h, const std::vector<int>& iperm) {
    const int n_send = h->n_send, nG = h->nG;
    std::vector<int> sslot(std::max(n_send, 1)), sboff(std::max(n_send, 1)), scnt(std::max(n_send, 1)), spos(std::max(n_send, 1));
    for (const Partner& p : h->partners)
        for (int k = 0; k < p.send_cnt; ++k) {
            const long long loc = h->give_ids[p.send_off + k] - h->gstart_id;
            sslot[p.send_off + k] = iperm[(int)loc];
            sboff[p.send_off + k] = p.send_off;
            scnt[p.send_off + k] = p.send_cnt;
            spos[p.send_off + k] = k;
        }
    TRY(h->alloc(&h->send_slot, n_send));
    TRY(h->alloc(&h->send_boff, n_send));
    TRY(h->alloc(&h->send_cnt, n_send));
    TRY(h->alloc(&h->send_pos, n_send));
    TRY(upload(h, h->send_slot, sslot.data(), n_send * sizeof(int)));
    TRY(upload(h, h->send_boff, sboff.data(), n_send * sizeof(int)));
    TRY(upload(h, h->send_cnt, scnt.data(), n_send * sizeof(int)));
    TRY(upload(h, h->send_pos, spos.data(), n_send * sizeof(int)));
    TRY(h->alloc(&h->sendbuf, (size_t)n_send * h->L));
    TRY(h->alloc(&h->recvbuf, (size_t)std::max(nG, 1) * h->L));
    CU(cudaStreamSynchronize(h->stream));
    TRY(setup_peer(h, h->need_matrix, sslot));
    if (h->peer && h->fused_halo && nG > 0) {
        // (colour, rank) keys of the ghost faces: one halo of the owners' keys, through the staged path
        std::vector<double> key((size_t)h->S, 0.0);
        for (int c = 0; c < h->n_colours; ++c)
            for (int p = h->cstart[c]; p < h->cstart[c] + h->ccount[c]; ++p) key[p] = (double)(c * h->n_ranks + h->rank);
        double* d_key = nullptr;
        TRY(h->alloc(&d_key, (size_t)h->S));
        TRY(upload(h, d_key, key.data(), key.size() * sizeof(double)));
        TRY(halo_exchange(h, d_key, 1));
        CU(cudaMemcpyAsync(key.data(), d_key, key.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        TRY(sync_stream(h));
        std::vector<int> gk((size_t)nG);
        for (int g = 0; g < nG; ++g) gk[g] = (int)key[(size_t)h->Tp + g];
        TRY(h->alloc(&h->ghost_key, (size_t)nG));
        TRY(upload(h, h->ghost_key, gk.data(), gk.size() * sizeof(int)));
        CU(cudaStreamSynchronize(h->stream));
        h->release(d_key);
    }
    return 0;
}

// ---- suspension solve: multicolour line Gauss–Seidel ---------------------------------------------------------
template <int LT, typename CT>
void launch_colour(pbsm3d_handle* h, int c) {
    const int p0 = h->cstart[c], p1 = p0 + h->ccount[c];
    LAUNCH(h, (gs_sweep_kernel<LT, CT>), cdiv(h->ccount[c], 128), 128, h->ss, h->dm, h->L, p0, p1, h->x, h->sc);
}
// the x channel of iteration (sweep) number e: read my buffer e%3, write the partners' (e+1)%3
HaloLink x_link(const pbsm3d_handle* h, unsigned long long e, bool ghosts_zero, bool signal) {
    HaloLink hl;
    hl.pt = h->d_pt;
    hl.flag_remote = h->xflag_remote;
    hl.flag_local = h->xflag_local;
    hl.bptr = h->bptr;
    hl.remote = h->x_remote[(e + 1) % 3];
    hl.rstride = h->rstride;
    hl.ghost = ghosts_zero ? h->g_zero : h->xg[e % 3];
    hl.nGp = h->nGp;
    hl.wait_epoch = e;
    hl.signal_epoch = signal ? e + 1 : 0;
    return hl;
}
template <int LT, typename CT>
void launch_colour_halo(pbsm3d_handle* h, int c, const HaloLink& hl) {
    const int p0 = h->cstart[c], p1 = p0 + h->ccount[c];
    LAUNCH(h, (gs_sweep_halo_kernel<LT, CT>), cdiv(h->ccount[c], 128), 128, h->ss, h->dm, h->L, p0, p1, h->x, h->sc, hl, h->nb[c],
           h->boff[c]);
}
template <typename CT>
int enqueue_sweeps_t(pbsm3d_handle* h, int n);
// n full sweeps; fp32 = stream the fp32-rounded coefficient copies (sweeps far from convergence only)
int enqueue_sweeps(pbsm3d_handle* h, int n, bool fp32 = false) {
    return fp32 ? enqueue_sweeps_t<float>(h, n) : enqueue_sweeps_t<double>(h, n);
}
template <typename CT>
int enqueue_sweeps_t(pbsm3d_handle* h, int n) {
    const bool fh = h->fused_halo && h->n_ranks > 1;
    int c_last = 0;
    for (int c = 0; c < h->n_colours; ++c)
        if (h->ccount[c] > 0) c_last = c;
    for (int k = 0; k < n; ++k) {
        for (int c = 0; c < h->n_colours; ++c) {
            if (h->ccount[c] == 0) continue;
            if (fh) {
                const HaloLink hl = x_link(h, h->xh_epoch, h->x_first, c == c_last);
                switch (h->L) {
                    case 5: launch_colour_halo<5, CT>(h, c, hl); break;
                    case 10: launch_colour_halo<10, CT>(h, c, hl); break;
                    case 15: launch_colour_halo<15, CT>(h, c, hl); break;
                    case 20: launch_colour_halo<20, CT>(h, c, hl); break;
                    default: launch_colour_halo<0, CT>(h, c, hl); break;
                }
                continue;
            }
            switch (h->L) {
                case 5: launch_colour<5, CT>(h, c); break;
                case 10: launch_colour<10, CT>(h, c); break;
                case 15: launch_colour<15, CT>(h, c); break;
                case 20: launch_colour<20, CT>(h, c); break;
                default: launch_colour<0, CT>(h, c); break;
            }
        }
        if (fh) {
            ++h->xh_epoch;
            h->x_first = false;
            ++h->halo_ops;
            ++h->halo_fused_ops;
        } else {
            TRY(halo_exchange(h, h->x, h->L));
        }
    }
    return 0;
}
void launch_residual(pbsm3d_handle* h, int it_now, double tol2, int fuse) {
    const int gcol = std::max(1, std::min(cdiv(h->Tp, 128), kRedBlocks));
    if (h->fused_halo && h->n_ranks > 1) {  // ghosts come from the buffer the next sweep would read
        const HaloLink hl = x_link(h, h->xh_epoch, h->x_first, false);
#define RES_HALO(LT_) LAUNCH(h, residual_halo_kernel<LT_>, gcol, 128, h->ss, h->dm, h->L, h->x, h->partial, kRedBlocks, h->sc, h->red, hl)
        switch (h->L) {
            case 5: RES_HALO(5); break;
            case 10: RES_HALO(10); break;
            case 15: RES_HALO(15); break;
            case 20: RES_HALO(20); break;
            default: RES_HALO(0); break;
        }
#undef RES_HALO
        return;
    }
#define RES_COL(LT_)                                                                                                         \
    LAUNCH(h, residual_col_kernel<LT_>, gcol, 128, h->ss, h->dm, h->x, h->partial, kRedBlocks, h->sc, h->red, it_now, tol2, fuse)
    switch (h->L) {
        case 5: RES_COL(5); break;
        case 10: RES_COL(10); break;
        case 15: RES_COL(15); break;
        case 20: RES_COL(20); break;
        default:
            LAUNCH(h, residual_kernel, red_grid(h->N), kRedThreads, h->ss, h->dm, h->L, h->x, h->partial, kRedBlocks, h->sc, h->red,
                   it_now, tol2, fuse);
    }
#undef RES_COL
}
// ||b - A x||^2 against tol^2 ||b||^2, decided on the device (x's ghost tails must be current)
int enqueue_check(pbsm3d_handle* h, int it_now) {
    const double tol2 = h->cfg.tolerance * h->cfg.tolerance;
    launch_residual(h, it_now, tol2, fused(h) ? 1 : 0);
    if (!fused(h)) {
        TRY(allreduce(h, h->red, 1, false));
        LAUNCH(h, flags_kernel, 1, 1, FLAGS_SUSP_CHECK, h->sc, h->red, it_now, tol2);
    }
    return 0;
}

// ---- a whole solve is one cooperative launch (gs_persistent_kernel / sor_persistent_kernel; across ranks their _halo variants)
SorLink sor_link(const pbsm3d_handle* h);
using GsKernel = void (*)(SuspSystem, DevMesh, int, ColourRanges, double*, float*, Scalars*, double*, SolvePlan, unsigned*);
ColourRanges colour_ranges(const pbsm3d_handle* h) {
    ColourRanges cr;
    std::memset(&cr, 0, sizeof(cr));
    for (int c = 0; c < h->n_colours; ++c)
        if (h->ccount[c] > 0) { cr.start[cr.n] = h->cstart[c]; cr.end[cr.n] = h->cstart[c] + h->ccount[c]; ++cr.n; }
    return cr;
}
using GsHaloKernel = void (*)(SuspSystem, DevMesh, int, ColourRanges, double*, float*, Scalars*, double*, double*, SolvePlan, unsigned*, XHalo);
int setup_persistent(pbsm3d_handle* h) {
    const char* env = getenv("PBSM3D_PERSISTENT");
    const char* env_as = getenv("PBSM3D_ACTIVE_SET");
    h->active_set = env_as ? (atoi(env_as) == 0 ? 0 : 1) : 2;
    h->persistent = false;
    if (env && atoi(env) == 0) return 0;
    const bool multi = h->n_ranks > 1;
    if (multi && !(h->peer && h->fused_halo)) return 0;  // across ranks the persistent kernels live on the in-kernel halos
    int coop = 0;
    CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    if (!coop) return 0;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, h->device));
    const void* fn;
    if (multi) {
        GsHaloKernel f;
        switch (h->L) {
            case 5: f = gs_persistent_halo_kernel<5>; break;
            case 10: f = gs_persistent_halo_kernel<10>; break;
            case 15: f = gs_persistent_halo_kernel<15>; break;
            case 20: f = gs_persistent_halo_kernel<20>; break;
            default: f = gs_persistent_halo_kernel<0>; break;
        }
        fn = (const void*)f;
    } else {
        GsKernel f;
        switch (h->L) {
            case 5: f = gs_persistent_kernel<5>; break;
            case 10: f = gs_persistent_kernel<10>; break;
            case 15: f = gs_persistent_kernel<15>; break;
            case 20: f = gs_persistent_kernel<20>; break;
            default: f = gs_persistent_kernel<0>; break;
        }
        fn = (const void*)f;
    }
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, kGsThreads, 0));
    h->gs_fn = (void*)fn;
    h->gs_grid = std::min(kRedBlocks, std::max(nb, 1) * prop.multiProcessorCount);
    {
        const char* v = getenv("PBSM3D_SOR_VARIANT");  // tuning knob: threads per block / faces in flight per thread
        const int var = v ? atoi(v) : 2;  // measured on c2: 512 threads x 4 faces in flight (profiles/r2b)
        const bool stream = (size_t)h->Tp * 60 > ((size_t)80 << 20);  // working set beyond what stays in the 126 MB L2
        const void* sf;
        if (multi) { h->sor_threads = 512; sf = stream ? (const void*)sor_persistent_halo_kernel<true, 512, 4> : (const void*)sor_persistent_halo_kernel<false, 512, 4>; }
        else if (var == 1) { h->sor_threads = 1024; sf = stream ? (const void*)sor_persistent_kernel<true, 1024, 2> : (const void*)sor_persistent_kernel<false, 1024, 2>; }
        else if (var == 3) { h->sor_threads = 512; sf = stream ? (const void*)sor_persistent_kernel<true, 512, 8> : (const void*)sor_persistent_kernel<false, 512, 8>; }
        else if (var == 4) { h->sor_threads = 1024; sf = stream ? (const void*)sor_persistent_kernel<true, 1024, 4> : (const void*)sor_persistent_kernel<false, 1024, 4>; }
        else { h->sor_threads = 512; sf = stream ? (const void*)sor_persistent_kernel<true, 512, 4> : (const void*)sor_persistent_kernel<false, 512, 4>; }
        h->sor_fn = sf;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sf, h->sor_threads, 0));
        h->sor_grid = std::min(kRedBlocks, std::max(nb, 1) * prop.multiProcessorCount);
        // the on-chip variant: two colour classes whose faces fit K per thread of a one-CTA-per-SM grid
        const char* rs = getenv("PBSM3D_SOR_RESIDENT");
        const char* rv = getenv("PBSM3D_SOR_RES_NT");  // tuning knob: threads per block (K follows: 512 x 7, 384 x 9, 256 x 14)
        const int rnt = rv ? atoi(rv) : 384;  // measured on c2: 0.60 / 0.68 / 0.75 ms for 384 / 512 / 256 (profiles/r2m_summary.md)
        int used = 0, maxc = 0;
        for (int c = 0; c < h->n_colours; ++c)
            if (h->ccount[c] > 0) { ++used; maxc = std::max(maxc, h->ccount[c]); }
        const void* rf = rnt == 256 ? (const void*)sor_resident_kernel<14, 256> : rnt == 384 ? (const void*)sor_resident_kernel<9, 384>
                                                                                                : (const void*)sor_resident_kernel<7, 512>;
        const int resK = rnt == 256 ? 14 : rnt == 384 ? 9 : 7, resNT = rnt == 256 ? 256 : rnt == 384 ? 384 : 512;
        const size_t smem = (size_t)2 * resK * (3 * sizeof(int) + 2 * sizeof(double)) * resNT;
        if (multi) rf = rnt == 512 ? (const void*)sor_resident_halo_kernel<7, 512> : (const void*)sor_resident_halo_kernel<9, 384>;
        if ((!multi || rnt != 256) && !stream && used <= 2 && !(rs && atoi(rs) == 0) && smem <= (size_t)prop.sharedMemPerBlockOptin) {
            CU(cudaFuncSetAttribute(rf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int rb = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rb, rf, resNT, smem));
            const int grid = std::min(kRedBlocks, std::min(rb, 1) * prop.multiProcessorCount);
            bool fits = rb >= 1 && (long long)maxc <= (long long)resK * grid * resNT;
            if (multi && fits)  // interior faces of a colour are shared by the blocks that are not its boundary blocks
                for (int c = 0; c < h->n_colours; ++c) {
                    if (h->ccount[c] <= 0) continue;
                    const int nbb = std::max(1, (h->nb[c] + resNT - 1) / resNT);
                    if (grid - nbb < 1 || (long long)(h->ccount[c] - h->nb[c]) > (long long)resK * (grid - nbb) * resNT) fits = false;
                }
            if (fits) {
                h->sor_resident = true;
                h->sor_fn = rf;
                h->sor_threads = resNT;
                h->sor_grid = grid;
                h->sor_smem = smem;
            }
        }
    }
    if (multi) {  // the boundary columns of a colour must fit the first grid-stride iteration
        for (int c = 0; c < h->n_colours; ++c)
            if (h->nb[c] > h->gs_grid / 2 * kGsThreads || h->nb[c] > h->sor_grid / 2 * h->sor_threads) return 0;
    }
    TRY(h->alloc_zero(&h->grid_bar, 2));
    if (h->cfg.fp32_sweep_streams && (h->L == 5 || h->L == 10 || h->L == 15 || h->L == 20)) TRY(h->alloc_zero(&h->xf, h->NS));
    h->persistent = true;
    {
        const char* sp = getenv("PBSM3D_SOR_PERSISTENT");
        h->sor_persistent = sp ? atoi(sp) != 0 : (size_t)h->Tp * 60 <= ((size_t)80 << 20);
    }
    if (getenv("PBSM3D_VERBOSE"))
        fprintf(stderr, "[pbsm3d] persistent solver kernels%s: sweep grid %d x %d, SOR grid %d x %d\n", multi ? " (halos inside)" : "",
                h->gs_grid, kGsThreads, h->sor_grid, h->sor_threads);
    return 0;
}
// colour ranges in the order colour_ranges() emits them, with the boundary bookkeeping of each
template <typename H>
void fill_boundary(const pbsm3d_handle* h, H& x) {
    int k = 0;
    for (int c = 0; c < h->n_colours; ++c)
        if (h->ccount[c] > 0) { x.nb[k] = h->nb[c]; x.boff[k] = h->boff[c]; ++k; }
}
constexpr double kActiveSetMaxSeeded = 0.4;
int live_max_seeds(const pbsm3d_handle* h) {
    return h->active_set == 1 ? INT_MAX : (int)(kActiveSetMaxSeeded * h->T);
}
int line_enqueue_persistent(pbsm3d_handle* h) {
    const int maxit = h->cfg.max_iterations;
    const bool known = h->pred_sweeps > 0;
    SolvePlan pl;
    pl.check_first = std::max(1, std::min(known ? h->pred_sweeps : 8, maxit));
    pl.check_every = known ? 1 : 4;
    pl.n32 = (known && h->cfg.fp32_sweep_streams) ? std::max(0, std::min(h->pred_n32, pl.check_first - 3)) : 0;
    // the iterate itself stays in fp32 storage until 10 sweeps before the predicted end (never past the fp32-coefficient phase)
    const char* nox = getenv("PBSM3D_FP32_X");
    pl.nx32 = (h->xf && !(nox && atoi(nox) == 0)) ? std::max(0, std::min(pl.n32, pl.check_first - 10)) : 0;
    pl.maxit = maxit;
    pl.use_live = h->active_set ? 1 : 0;
    pl.live_max_seeds = live_max_seeds(h);
    {
        const char* e = getenv("PBSM3D_NBS_PREFETCH");
        pl.prefetch = (e && atoi(e) == 0) ? 0 : 1;
    }
    pl.tol2 = h->cfg.tolerance * h->cfg.tolerance;
    h->plan_n32 = pl.n32;
    h->plan_nx32 = pl.nx32;
    ColourRanges cr = colour_ranges(h);
    unsigned* bar = h->grid_bar;
    CU(cudaMemsetAsync(bar, 0, sizeof(unsigned), h->stream));
    int L = h->L;
    CU(cudaEventRecord(h->ev_sw[0], h->stream));
    ++h->n_launch;
    if (h->n_ranks > 1) {
        XHalo xh;
        std::memset(&xh, 0, sizeof(xh));
        xh.pt = h->d_pt;
        xh.flag_remote = h->xflag_remote;
        xh.flag_local = h->xflag_local;
        xh.bptr = h->bptr;
        xh.rstride = h->rstride;
        for (int b = 0; b < 3; ++b) { xh.remote[b] = h->x_remote[b]; xh.ghost[b] = h->xg[b]; }
        xh.g_zero = h->g_zero;
        xh.nGp = h->nGp;
        xh.e0 = h->xh_epoch;
        fill_boundary(h, xh);
        void* args[] = {&h->ss, &h->dm, &L, &cr, &h->x, &h->xf, &h->sc, &h->partial, &h->red, &pl, &bar, &xh};
        CU(cudaLaunchCooperativeKernel((const void*)h->gs_fn, dim3(h->gs_grid), dim3(kGsThreads), args, 0, h->stream));
        h->x_first = false;
    } else {
        void* args[] = {&h->ss, &h->dm, &L, &cr, &h->x, &h->xf, &h->sc, &h->partial, &pl, &bar};
        CU(cudaLaunchCooperativeKernel((const void*)h->gs_fn, dim3(h->gs_grid), dim3(kGsThreads), args, 0, h->stream));
    }
    CU(cudaEventRecord(h->ev_sw[1], h->stream));
    CU(cudaEventRecord(h->ev_sw[2], h->stream));
    return 0;
}
int sor_enqueue_persistent(pbsm3d_handle* h) {
    const int maxit = std::min(h->cfg.max_iterations, 6 * h->sor_kest + 64);
    const bool known = h->pred_sor > 0;
    SolvePlan pl;
    pl.n32 = pl.nx32 = pl.use_live = pl.live_max_seeds = pl.prefetch = 0;
    pl.check_first = std::max(1, std::min(maxit, known ? h->pred_sor : h->sor_kest));
    pl.check_every = known ? 1 : 4;
    pl.maxit = maxit;
    pl.tol2 = h->cfg.tolerance * h->cfg.tolerance;
    ColourRanges cr = colour_ranges(h);
    unsigned* bar = h->grid_bar + 1;
    CU(cudaMemsetAsync(bar, 0, sizeof(unsigned), h->stream));
    ++h->n_launch;
    if (h->n_ranks > 1) {
        QHalo qh;
        std::memset(&qh, 0, sizeof(qh));
        qh.sl = sor_link(h);
        qh.e0 = h->sor_epoch;
        qh.n_ranks = h->n_ranks;
        qh.rank = h->rank;
        fill_boundary(h, qh);
        int k = 0;
        for (int c = 0; c < h->n_colours; ++c)
            if (h->ccount[c] > 0) qh.colour_of[k++] = c;
        void* args[] = {&h->dm, &h->offS, &h->drhsS, &h->ddiag, &h->qA, &h->sor_omega, &cr, &h->sc, &h->partial, &h->red, &pl, &bar, &qh};
        CU(cudaLaunchCooperativeKernel(h->sor_fn, dim3(h->sor_grid), dim3(h->sor_threads), args, h->sor_resident ? h->sor_smem : 0, h->stream));
    } else {
        void* args[] = {&h->dm, &h->offS, &h->drhsS, &h->ddiag, &h->qA, &h->sor_omega, &cr, &h->sc, &h->partial, &pl, &bar};
        CU(cudaLaunchCooperativeKernel(h->sor_fn, dim3(h->sor_grid), dim3(h->sor_threads), args, h->sor_resident ? h->sor_smem : 0, h->stream));
    }
    h->sor_enqueued = maxit;  // replaced by the executed count (Scalars::dep_sweeps) once the step has synchronised
    return 0;
}

// Optimistic part: a predicted number of sweeps, a check, and a few short speculative rounds (no-ops once converged).
int line_enqueue_initial(pbsm3d_handle* h, int* total_out) {
    const int maxit = h->cfg.max_iterations;
    int total = 0;
    const bool known = h->pred_sweeps > 0;
    int first = std::min(known ? h->pred_sweeps : 8, maxit);
    // the leading sweeps, while the residual is still above ~1e-6 ||b||, may stream the fp32 coefficient copies
    const int n32 = (known && h->cfg.fp32_sweep_streams) ? std::max(0, std::min(h->pred_n32, first - 3)) : 0;
    CU(cudaEventRecord(h->ev_sw[0], h->stream));
    if (n32 > 0) TRY(enqueue_sweeps(h, n32, true));
    CU(cudaEventRecord(h->ev_sw[2], h->stream));
    TRY(enqueue_sweeps(h, first - n32));
    CU(cudaEventRecord(h->ev_sw[1], h->stream));
    h->sweeps_timed = first;
    h->sweeps_timed32 = n32;
    total = first;
    TRY(enqueue_check(h, total));
    const int spec_known[3] = {1, 1, 2}, spec_unknown[3] = {8, 8, 8};
    for (int k = 0; k < 3 && total < maxit; ++k) {
        int step = std::min(known ? spec_known[k] : spec_unknown[k], maxit - total);
        TRY(enqueue_sweeps(h, step));
        total += step;
        TRY(enqueue_check(h, total));
    }
    *total_out = total;
    return 0;
}

// Slow path: the optimistic rounds did not reach the tolerance.  Predict the sweeps still needed from the observed
// geometric rate, run them, look again.  Returns with *converged set; stagnation hands over to the Krylov path.
int line_continue(pbsm3d_handle* h, int total, bool allow_bail, bool* converged) {
    const double tol2 = h->cfg.tolerance * h->cfg.tolerance;
    const int maxit = h->cfg.max_iterations;
    *converged = false;
    const Scalars& s = *h->h_sc;
    double prev_rr = s.susp_bnorm2;
    int prev_it = 0;
    int nh = std::min(s.n_checks, 16);
    if (nh >= 2) { prev_rr = s.rr_hist[nh - 2]; prev_it = s.it_hist[nh - 2]; }
    double rr = s.susp_rr;
    int it = total, slow = 0;
    while (it < maxit) {
        if (!std::isfinite(rr)) return 0;
        int step = 8;
        double rate = (rr > 0 && prev_rr > 0 && it > prev_it) ? std::pow(rr / prev_rr, 0.5 / (double)(it - prev_it)) : 2.0;
        if (rate < 1.0 && rate > 0.0) {
            double need = 0.5 * std::log(tol2 * s.susp_bnorm2 / rr) / std::log(rate);
            step = std::max(1, std::min((int)std::ceil(need), 256));
            if (rate > 0.97 && it >= 64) ++slow;
        } else {
            ++slow;
        }
        if (allow_bail && slow >= 2) return 0;  // not contracting / crawling: the Krylov path is the better tool
        step = std::min(step, maxit - it);
        TRY(enqueue_sweeps(h, step));
        it += step;
        TRY(enqueue_check(h, it));
        prev_rr = rr;
        prev_it = it - step;
        TRY(read_scalars(h));
        rr = h->h_sc->susp_rr;
        if (h->h_sc->susp_done) { *converged = true; return 0; }
    }
    return 0;
}

int ensure_krylov(pbsm3d_handle* h) {
    if (h->kry[0]) return 0;
    for (int k = 0; k < 7; ++k) TRY(h->alloc_zero(&h->kry[k], h->NS));
    return 0;
}

int fold(pbsm3d_handle* h, int nblocks, int nvals) {
    LAUNCH(h, fold_kernel, 1, 256, nblocks, nvals, kRedBlocks, h->partial, h->red, 0);
    return allreduce(h, h->red, nvals, false);
}

// Right-preconditioned BiCGStab with the column-tridiagonal preconditioner (fallback and cross-check solver;
// host-paced in batches of 4 iterations).
int solve_bicgstab(pbsm3d_handle* h, bool* converged) {
    TRY(ensure_krylov(h));
    const double tol2 = h->cfg.tolerance * h->cfg.tolerance;
    double *x = h->x, *r = h->kry[0], *rhat = h->kry[1], *p = h->kry[2], *v = h->kry[3], *ph = h->kry[4], *sh = h->kry[5],
           *t = h->kry[6];
    const size_t NS = h->NS;
    const int g = red_grid(h->N), gv = red_grid(NS), gt = cdiv(h->Tp, 128);
    const int* done = &h->sc->done;
    LAUNCH(h, bicg_init_kernel, gv, 256, h->Tp, h->S, h->L, h->ss.rhs0, x, r, rhat, p, v, h->partial);
    TRY(fold(h, gv, 1));
    LAUNCH(h, bicg_scalar_kernel, 1, 1, 0, h->sc, h->red, tol2);
    *converged = false;
    const int maxit = h->cfg.max_iterations;
    int it = 0;
    while (it < maxit) {
        int target = std::min(it + 4, maxit);
        for (; it < target; ++it) {
            LAUNCH(h, bicg_p_kernel, gv, 256, NS, h->sc, r, v, p);
            LAUNCH(h, thomas_kernel, gt, 128, h->ss, h->Tp, h->S, h->L, p, ph, done);
            TRY(halo_exchange(h, ph, h->L));
            LAUNCH(h, spmv_kernel, g, kRedThreads, h->ss, h->dm, h->L, ph, v, rhat, nullptr, 0, h->partial, kRedBlocks, done);
            TRY(fold(h, g, 1));
            LAUNCH(h, bicg_scalar_kernel, 1, 1, 1, h->sc, h->red, tol2);
            LAUNCH(h, bicg_s_kernel, gv, 256, NS, h->sc, r, v, h->partial);
            LAUNCH(h, thomas_kernel, gt, 128, h->ss, h->Tp, h->S, h->L, r, sh, done);
            TRY(halo_exchange(h, sh, h->L));
            LAUNCH(h, spmv_kernel, g, kRedThreads, h->ss, h->dm, h->L, sh, t, r, nullptr, 1, h->partial, kRedBlocks, done);
            TRY(fold(h, g, 2));
            LAUNCH(h, bicg_scalar_kernel, 1, 1, 3, h->sc, h->red, tol2);
            LAUNCH(h, bicg_xr_kernel, gv, 256, NS, h->sc, x, r, ph, sh, t, rhat, h->partial, kRedBlocks);
            TRY(fold(h, gv, 2));
            LAUNCH(h, bicg_scalar_kernel, 1, 1, 4, h->sc, h->red, tol2);
        }
        TRY(read_scalars(h));
        if (h->h_sc->done) break;
    }
    *converged = (h->h_sc->done == 1);
    return 0;
}

// ---- deposition: Jacobi-preconditioned CG, three launches per iteration on one rank -------------------------
int enqueue_cg_iterations(pbsm3d_handle* h, int n) {
    const double tol2 = h->cfg.tolerance * h->cfg.tolerance;
    const int Tp = h->Tp, g = red_grid(Tp), f = fused(h) ? 1 : 0;
    for (int k = 0; k < n; ++k) {
        LAUNCH(h, cg_spmv_kernel, g, kRedThreads, h->dm, h->ddiag, h->doff, h->cg_p, h->cg_Ap, h->partial, kRedBlocks, h->sc, h->red,
               tol2, f);
        if (!f) {
            TRY(allreduce(h, h->red, 1, false));
            LAUNCH(h, cg_scalar_kernel, 1, 1, 1, h->sc, h->red, tol2);
        }
        LAUNCH(h, cg_update_kernel, g, kRedThreads, Tp, h->sc, h->dinv, h->cg_p, h->cg_Ap, h->qA, h->cg_r, h->partial, kRedBlocks,
               h->red, tol2, f);
        if (!f) {
            TRY(allreduce(h, h->red, 2, false));
            LAUNCH(h, cg_scalar_kernel, 1, 1, 2, h->sc, h->red, tol2);
        }
        LAUNCH(h, cg_p_kernel, g, kRedThreads, Tp, h->sc, h->dinv, h->cg_r, h->cg_p);
        TRY(halo_exchange(h, h->cg_p, 1));
    }
    return 0;
}

// ---- deposition: Jacobi-preconditioned Chebyshev iteration (one launch per iteration, no reductions) ---------
void cheb_coefficients(pbsm3d_handle* h, int upto) {
    const double theta = 0.5 * (h->cheb_lmax + h->cheb_lmin), delta = 0.5 * (h->cheb_lmax - h->cheb_lmin), sigma1 = theta / delta;
    if (h->cheb_a.empty()) {
        h->cheb_a.push_back(0.0);
        h->cheb_c.push_back(1.0 / theta);
    }
    // rho_k is recomputed from the start when the table grows (cheap, and keeps the handle free of hidden state)
    double rho = 1.0 / sigma1;
    for (int k = 1; k <= upto; ++k) {
        const double rho_new = 1.0 / (2.0 * sigma1 - rho);
        if (k >= (int)h->cheb_a.size()) {
            h->cheb_a.push_back(rho_new * rho);
            h->cheb_c.push_back(2.0 * rho_new / delta);
        }
        rho = rho_new;
    }
}
// iterations [k0, k1); iterations >= check_from also measure ||b - A q_k|| and apply the stopping rule
int enqueue_cheb(pbsm3d_handle* h, int k0, int k1, int check_from) {
    const double tol2 = h->cfg.tolerance * h->cfg.tolerance;
    const int g = red_grid(h->Tp), f = fused(h) ? 1 : 0;
    const bool fh = h->fused_halo && h->n_ranks > 1;
    cheb_coefficients(h, k1);
    BndRanges br;
    std::memset(&br, 0, sizeof(br));
    if (fh) {
        br.n_colours = h->n_colours;
        br.total = h->nb_total;
        for (int c = 0; c < h->n_colours; ++c) {
            br.start[c] = h->cstart[c];
            br.count[c] = h->nb[c];
            br.off[c] = h->boff[c];
            br.end[c] = c + 1 < h->n_colours ? h->cstart[c + 1] : h->Tp;
        }
    }
    for (int k = k0; k < k1; ++k) {
        const double* qin = (k & 1) ? h->qB : h->qA;
        double* qout = (k & 1) ? h->qA : h->qB;
        const bool check = k >= check_from;
        if (fh) {
            const unsigned long long e = ++h->qh_epoch;  // tags start at 1: a zeroed arena never matches
            TaggedLink hl;
            hl.pt = h->d_pt;
            hl.bptr = h->bptr;
            hl.remote = h->q_remote[(e + 1) % 3];
            hl.ghost = k == 0 ? nullptr : h->qg[e % 3];  // q_0 = 0
            hl.read_tag = e;
            hl.write_tag = e + 1;
            // one wave: nbb boundary blocks + interior blocks, never more than the kRedBlocks partial-sum slots
            const int nbb = std::max(1, std::min(cdiv(h->nb_total, kRedThreads), 64));
            const int gh = std::max(1, std::min(g, kRedBlocks - nbb)) + nbb;
            if (check)
                LAUNCH(h, cheb_iter_halo_kernel<1>, gh, kRedThreads, h->dm, h->offS, h->drhsS, h->ddiag, qin, qout, h->cheb_a[k],
                       h->cheb_c[k], h->partial, kRedBlocks, h->sc, h->red, hl, br, nbb);
            else
                LAUNCH(h, cheb_iter_halo_kernel<0>, gh, kRedThreads, h->dm, h->offS, h->drhsS, h->ddiag, qin, qout, h->cheb_a[k],
                       h->cheb_c[k], h->partial, kRedBlocks, h->sc, h->red, hl, br, nbb);
            ++h->halo_ops;
            ++h->halo_fused_ops;
        } else if (check) {
            LAUNCH(h, cheb_iter_kernel<1>, g, kRedThreads, h->dm, h->offS, h->drhsS, h->ddiag, qin, qout, h->cheb_a[k],
                   h->cheb_c[k], k, h->partial, kRedBlocks, h->sc, h->red, tol2, f);
        } else {
            LAUNCH(h, cheb_iter_kernel<0>, g, kRedThreads, h->dm, h->offS, h->drhsS, h->ddiag, qin, qout, h->cheb_a[k],
                   h->cheb_c[k], k, h->partial, kRedBlocks, h->sc, h->red, tol2, f);
        }
        if (check && !f) {
            TRY(allreduce(h, h->red, 1, false));
            LAUNCH(h, flags_kernel, 1, 1, FLAGS_CHEB_CHECK, h->sc, h->red, k, tol2);
        }
        if (!fh) TRY(halo_exchange(h, qout, 1));
    }
    h->cheb_enqueued = k1;
    return 0;
}
bool use_chebyshev(const pbsm3d_handle* h) { return h->cheb_ready && h->cfg.deposition_solver != PBSM3D_DEP_CG; }
// SOR: asked for or AUTO.  Across ranks the sweep order is (colour, rank) and every colour pass carries its own halo
// (sor_pass_halo_kernel); with rank-local colours and ghosts one sweep old, over-relaxation would diverge.
bool use_sor(const pbsm3d_handle* h) {
    if (!h->cheb_ready || !(h->sor_omega > 0)) return false;
    if (h->n_ranks > 1 && !(h->peer && h->fused_halo)) return false;  // across ranks it lives on the tagged in-kernel halo
    return h->cfg.deposition_solver == PBSM3D_DEP_SOR || h->cfg.deposition_solver == PBSM3D_DEP_AUTO;
}
SorLink sor_link(const pbsm3d_handle* h) {
    SorLink sl;
    std::memset(&sl, 0, sizeof(sl));
    sl.tl.pt = h->d_pt;
    sl.tl.bptr = h->bptr;
    sl.tl.remote = h->q_remote[0];
    sl.tl.ghost = h->qg[0];
    sl.ghost_key = h->ghost_key;
    return sl;
}
int enqueue_sor_sweeps(pbsm3d_handle* h, int n) {
    const bool multi = h->n_ranks > 1;
    const SorLink sl = multi ? sor_link(h) : SorLink{};
    for (int k = 0; k < n; ++k) {
        const unsigned long long e = ++h->sor_epoch;
        const int first = (h->sor_enqueued + k) == 0 ? 1 : 0;
        for (int c = 0; c < h->n_colours; ++c) {
            if (h->ccount[c] == 0) continue;
            const int p0 = h->cstart[c], p1 = p0 + h->ccount[c];
            if (multi)
                LAUNCH(h, sor_pass_halo_kernel, cdiv(h->ccount[c], 256), 256, h->dm, h->offS, h->drhsS, h->qA, h->sor_omega, p0, p1,
                       h->nb[c], h->boff[c], h->sc, sl, c * h->n_ranks + h->rank, e, first);
            else
                LAUNCH(h, sor_pass_kernel, cdiv(h->ccount[c], 256), 256, h->dm, h->offS, h->drhsS, h->qA, h->sor_omega, p0, p1, h->sc);
        }
        if (multi) { ++h->halo_ops; ++h->halo_fused_ops; }
    }
    h->sor_enqueued += n;
    return 0;
}
int enqueue_sor_check(pbsm3d_handle* h) {
    const double tol2 = h->cfg.tolerance * h->cfg.tolerance;
    const int f = fused(h) ? 1 : 0;
    if (h->n_ranks > 1)
        LAUNCH(h, dep_residual_halo_kernel, red_grid(h->Tp), kRedThreads, h->dm, h->offS, h->drhsS, h->ddiag, h->qA, h->partial, kRedBlocks,
               h->sc, h->red, sor_link(h), h->sor_epoch);
    else
        LAUNCH(h, dep_residual_kernel, red_grid(h->Tp), kRedThreads, h->dm, h->offS, h->drhsS, h->ddiag, h->qA, h->sor_enqueued,
               h->partial, kRedBlocks, h->sc, h->red, tol2, f);
    if (!f) {
        TRY(allreduce(h, h->red, 1, false));
        LAUNCH(h, flags_kernel, 1, 1, FLAGS_SOR_CHECK, h->sc, h->red, h->sor_enqueued, tol2);
    }
    return 0;
}
// predicted number of sweeps, a check, three short speculative rounds (no-ops once converged)
int enqueue_sor_initial(pbsm3d_handle* h) {
    const int maxit = h->cfg.max_iterations;
    const bool known = h->pred_sor > 0;
    h->sor_enqueued = 0;
    TRY(enqueue_sor_sweeps(h, std::min(maxit, known ? h->pred_sor : h->sor_kest)));
    TRY(enqueue_sor_check(h));
    const int spec_known[3] = {1, 1, 2}, spec_unknown[3] = {4, 8, 8};
    for (int k = 0; k < 3 && h->sor_enqueued < maxit; ++k) {
        TRY(enqueue_sor_sweeps(h, std::min(known ? spec_known[k] : spec_unknown[k], maxit - h->sor_enqueued)));
        TRY(enqueue_sor_check(h));
    }
    return 0;
}

int enqueue_cg_start(pbsm3d_handle* h, int n_cg) {
    const double tol2 = h->cfg.tolerance * h->cfg.tolerance;
    LAUNCH(h, cg_init_kernel, red_grid(h->Tp), kRedThreads, h->Tp, h->drhs, h->dinv, h->qA, h->cg_r, h->cg_p, h->partial, kRedBlocks,
           h->sc, h->red, tol2, fused(h) ? 1 : 0);
    if (!fused(h)) {
        TRY(allreduce(h, h->red, 2, false));
        LAUNCH(h, cg_scalar_kernel, 1, 1, 0, h->sc, h->red, tol2);
    }
    TRY(halo_exchange(h, h->cg_p, 1));
    return enqueue_cg_iterations(h, n_cg);
}

// Smallest and largest eigenvalue of a symmetric tridiagonal matrix (diagonal a[0..n), off-diagonal b[0..n-1)) by
// bisection on the Sturm count.
void tridiag_extremes(const std::vector<double>& a, const std::vector<double>& b, double* lo_out, double* hi_out) {
    const int n = (int)a.size();
    double lo = a[0], hi = a[0];
    for (int i = 0; i < n; ++i) {
        double r = (i > 0 ? std::fabs(b[i - 1]) : 0.0) + (i + 1 < n ? std::fabs(b[i]) : 0.0);
        lo = std::min(lo, a[i] - r);
        hi = std::max(hi, a[i] + r);
    }
    auto count_below = [&](double x) {  // eigenvalues < x
        int c = 0;
        double q = a[0] - x;
        if (q < 0) ++c;
        for (int i = 1; i < n; ++i) {
            if (q == 0.0) q = 1e-300;
            q = a[i] - x - b[i - 1] * b[i - 1] / q;
            if (q < 0) ++c;
        }
        return c;
    };
    auto kth = [&](int k) {  // k-th smallest (0-based)
        double l = lo, r = hi;
        for (int it = 0; it < 200 && (r - l) > 1e-14 * std::max(std::fabs(l), std::fabs(r)); ++it) {
            double mid = 0.5 * (l + r);
            if (count_below(mid) > k) r = mid; else l = mid;
        }
        return 0.5 * (l + r);
    };
    *lo_out = kth(0);
    *hi_out = kth(n - 1);
}

// pbsm3d_create: spectrum bounds of D^-1 A for the Chebyshev iteration.  The deposition matrix depends on the mesh
// and smooth_coeff only, so this runs once: CG on a fixed pseudo-random right-hand side, its recurrence
// coefficients give the Lanczos tridiagonal, whose extreme Ritz values converge to the extreme eigenvalues.
int estimate_spectrum(pbsm3d_handle* h) {
    const int cap = std::min(512, std::max(8, h->cfg.max_iterations));
    double *d_la = nullptr, *d_lb = nullptr;
    TRY(h->alloc_zero(&d_la, cap));
    TRY(h->alloc_zero(&d_lb, cap));
    LAUNCH(h, probe_rhs_kernel, cdiv(h->Tp, 256), 256, h->Tp, h->perm, (long long)h->gstart_id, h->drhs);
    LAUNCH(h, flags_kernel, 1, 1, FLAGS_SETUP_CG, h->sc, h->red, 0, 0.0);
    CU(cudaMemcpyAsync(&h->sc->log_alpha, &d_la, sizeof(double*), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(&h->sc->log_beta, &d_lb, sizeof(double*), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(&h->sc->log_cap, &cap, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    const double keep_tol = h->cfg.tolerance;
    h->cfg.tolerance = 1e-12;  // let the Krylov space grow until the extreme Ritz values have settled
    TRY(enqueue_cg_start(h, 0));
    int it = 0;
    double lmin = 0.0, lmax = 0.0, prev_lmin = -1.0;
    bool settled = false;
    std::vector<double> la(cap), lb(cap);
    while (it < cap && !settled) {
        const int n = std::min(48, cap - it);
        TRY(enqueue_cg_iterations(h, n));
        it += n;
        TRY(read_scalars(h));
        const int m = std::min(h->h_sc->log_n, cap);
        if (m < 2) break;
        CU(cudaMemcpy(la.data(), d_la, m * sizeof(double), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(lb.data(), d_lb, m * sizeof(double), cudaMemcpyDeviceToHost));
        std::vector<double> ta(m), tb(m > 1 ? m - 1 : 0);
        for (int j = 0; j < m; ++j) {
            ta[j] = 1.0 / la[j] + (j > 0 ? lb[j - 1] / la[j - 1] : 0.0);
            if (j + 1 < m) tb[j] = std::sqrt(std::max(lb[j], 0.0)) / la[j];
        }
        tridiag_extremes(ta, tb, &lmin, &lmax);
        h->lanczos_steps = m;
        if (h->h_sc->done) { settled = true; break; }
        if (prev_lmin > 0 && std::fabs(lmin - prev_lmin) <= 0.01 * lmin) settled = true;
        prev_lmin = lmin;
    }
    h->cfg.tolerance = keep_tol;
    double* null_ptr = nullptr;
    CU(cudaMemcpyAsync(&h->sc->log_alpha, &null_ptr, sizeof(double*), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(&h->sc->log_beta, &null_ptr, sizeof(double*), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(h->sc, 0, offsetof(Scalars, log_alpha), h->stream));
    CU(cudaMemsetAsync(h->qA, 0, (size_t)h->S * sizeof(double), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->release(d_la);
    h->release(d_lb);
    if (!(lmin > 0) || !(lmax > lmin) || !std::isfinite(lmax)) return 0;  // leave Chebyshev off: CG will be used
    // Ritz values lie inside the spectrum: widen a little.  D^-1 A of this diagonally dominant M-matrix has its
    // spectrum in (0, 2) (Gershgorin), so 2 caps the upper bound; an upper bound that is too small would diverge.
    h->cheb_lmin = (settled ? 0.97 : 0.7) * lmin;
    h->cheb_lmax = std::min(2.0, 1.01 * lmax + 1e-3);
    const double kappa = h->cheb_lmax / h->cheb_lmin;
    const double qf = (std::sqrt(kappa) - 1.0) / (std::sqrt(kappa) + 1.0);
    h->cheb_kest = qf > 0 ? (int)std::ceil(std::log(h->cfg.tolerance / 2.0) / std::log(qf)) + 2 : 4;
    h->cheb_ready = true;
    {   // Young's optimal relaxation factor for the Jacobi matrix J = I - D^-1 A, rho(J) = 1 - lambda_min.  The bound is the
        // widened (smaller) lambda_min: erring towards a larger omega costs a few sweeps, a smaller one many.
        const double rho = std::min(1.0 - h->cheb_lmin, 1.0 - 1e-12);
        h->sor_omega = 2.0 / (1.0 + std::sqrt(std::max(1.0 - rho * rho, 1e-24)));
        const double fac = h->sor_omega - 1.0;  // asymptotic error reduction per sweep
        h->sor_kest = fac > 0 && fac < 1 ? (int)std::ceil(std::log(h->cfg.tolerance) / std::log(fac)) + 8 : 64;
    }
    return 0;
}

struct OutTargets {
    double* dst[kOut];    // where export_kernel writes (device pointers: caller's buffers or out_stage)
    double* host[kOut];   // optional host destinations for a following D2H
};

// Export of the outputs in `mask` to CHM order on `stream` (+ D2H when the caller's buffers are on the host).
int enqueue_export(pbsm3d_handle* h, cudaStream_t stream, const OutTargets* out, unsigned mask) {
    if (!out) return 0;
    const int T = h->T;
    ExportPtrs e;
    const double* src[kOut] = {h->ss.Qsalt, h->Qsusp, h->Qsubl, h->Qsubl_mass, h->sum_subl, h->drift_mass, h->sum_drift, h->more_avail,
                               h->ss.prob};
    bool any = false;
    for (int k = 0; k < kOut; ++k) {
        e.src[k] = src[k];
        e.dst[k] = (mask >> k & 1u) ? out->dst[k] : nullptr;
        any = any || e.dst[k];
    }
    if (!any) return 0;
    LAUNCH_ON(h, stream, export_kernel, cdiv(T, 256), 256, T, h->iperm, e);
    for (int k = 0; k < kOut; ++k)
        if (e.dst[k] && out->host[k])
            CU(cudaMemcpyAsync(out->host[k], out->dst[k], (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, stream));
    return 0;
}
constexpr unsigned kOutQsalt = (1u << 0) | (1u << 8) /* Qsalt and blowingsnow_probability: final after the assembly */, kOutFlux = (1u << 1) | (1u << 2) | (1u << 3) | (1u << 4), kOutDrift = (1u << 5) | (1u << 6) | (1u << 7);

// E..I of SURVEY §3.2 plus the export: everything after the suspension solve.  Safe to enqueue before the host
// knows whether the solve converged (device guards), and safe to enqueue again if it had not.
int enqueue_tail(pbsm3d_handle* h, const DevForcing& f, double dt, const OutTargets* out, int n_cg) {
    cudaStream_t s = h->stream;
    const int Tp = h->Tp;
    const double tol2 = h->cfg.tolerance * h->cfg.tolerance;
    // E: flux integration
    LAUNCH(h, flux_kernel, cdiv(Tp, 256), 256, Tp, h->S, h->L, h->dc.dz, dt, h->perm, h->x, h->ss.u_z, h->ss.csubl, h->Qsusp, h->Qsubl,
           h->Qsubl_mass, h->sum_subl, h->sc);
    if (out) {  // early export: Qsusp, Qsubl, Qsubl_mass, sum_subl are final from here on
        CU(cudaEventRecord(h->ev_flux, s));
        CU(cudaStreamWaitEvent(h->s_out, h->ev_flux, 0));
        TRY(enqueue_export(h, h->s_out, out, kOutFlux));
    }
    // F: halo of Qsusp, Qsalt (PBSM3D.cpp:1509-1510)
    TRY(halo_exchange(h, h->Qsusp, 1));
    TRY(halo_exchange(h, h->ss.Qsalt, 1));
    // G: deposition RHS (the matrix is static) + H: rhs max
    LAUNCH(h, deposition_rhs_kernel, red_grid(Tp), kRedThreads, h->dm, f.vw_dir, h->Qsusp, h->ss.Qsalt, h->dinv, h->drhs, h->drhsS,
           h->partial, kRedBlocks, h->sc, h->red);
    if (h->n_ranks > 1) {
        TRY(allreduce(h, h->red, 1, true));
        TRY(allreduce(h, h->red + 1, 1, false));
    }
    LAUNCH(h, flags_kernel, 1, 1, FLAGS_DEP, h->sc, h->red, 0, tol2);
    CU(cudaEventRecord(h->ev[3], s));
    // deposition solve (x0 = 0)
    CU(cudaMemsetAsync(h->qA, 0, (size_t)h->S * sizeof(double), s));
    if (use_sor(h) && h->persistent && h->sor_persistent) {
        TRY(sor_enqueue_persistent(h));
    } else if (use_sor(h)) {
        TRY(enqueue_sor_initial(h));
    } else if (use_chebyshev(h)) {
        CU(cudaMemsetAsync(h->qB, 0, (size_t)h->S * sizeof(double), s));  // q_{-1}: multiplied by a_0 = 0, must be finite
        const int maxit = h->cfg.max_iterations;
        const bool known = h->pred_dep > 0;
        const int n = known ? h->pred_dep : h->cheb_kest;
        const int check_from = known ? std::max(0, n - 3) : std::max(0, n / 2);
        const int k1 = std::min(maxit, known ? n + 6 : n + n / 4 + 8);
        TRY(enqueue_cheb(h, 0, k1, check_from));
    } else {
        TRY(enqueue_cg_start(h, n_cg));
    }
    return 0;
}

// I: drift update + export of `mask` (+ the control block) on the compute stream
int enqueue_finish(pbsm3d_handle* h, const DevForcing& f, double dt, const OutTargets* out, unsigned mask) {
    cudaStream_t s = h->stream;
    const int Tp = h->Tp;
    LAUNCH(h, drift_kernel, cdiv(Tp, 256), 256, Tp, dt, h->perm, h->qA, h->qB, f.swe, h->ss.salt, h->drift_mass, h->sum_drift,
           h->more_avail, h->sc);
    LAUNCH(h, drift_done_kernel, 1, 1, h->sc);
    CU(cudaEventRecord(h->ev[4], s));
    TRY(enqueue_export(h, s, out, mask));
    CU(cudaMemcpyAsync(h->h_sc, h->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s));
    return 0;
}

// Assembly of CHM faces [i0, i1); its {max|b|, sum b^2} go to red[2*chunk ..].
// The layer-parallel tile kernel runs with one warp per layer (nLayer <= 32; block sizes 4/5/8/10/16/20/32 warps, surplus warps
// only keep the barriers); deeper columns, or PBSM3D_ASSEMBLY=column, take the column-walking kernel.
using AsmKernel = void (*)(DevConfig, DevMesh, FaceRecs, SuspSystem, int, int, double*, int, Scalars*, double*);
int setup_assembly_fn(pbsm3d_handle* h, AsmKernel fn, int nw, int threads) {
    h->asm_nw = nw;
    h->asm_fn = (void*)fn;
    h->asm_threads = threads;
    h->asm_smem = nw ? (size_t)(3 * h->L * 32) * sizeof(double) : 0;
    CU(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->asm_smem));
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)fn, threads, h->asm_smem));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, h->device));
    h->asm_grid = std::max(1, std::min(kRedBlocks, std::max(nb, 1) * prop.multiProcessorCount));
    if (getenv("PBSM3D_VERBOSE"))
        fprintf(stderr, "[pbsm3d] assembly: %s kernel, %d threads per block, %d blocks per SM, grid %d\n", nw ? "tile" : "column", threads,
                nb, h->asm_grid);
    return 0;
}
int setup_assembly(pbsm3d_handle* h) {
    const char* want = getenv("PBSM3D_ASSEMBLY");
    const char* mb = getenv("PBSM3D_ASM_MINB");  // tuning knob: resident blocks per SM the kernel is compiled for
    const int minb = mb ? atoi(mb) : 0;
    const int L = h->L;
    TRY(h->alloc(&h->recs.d, (size_t)kRecD * h->Tp));
    TRY(h->alloc_zero(&h->recs.i, (size_t)h->Tp));
    h->recs.T = h->Tp;
    const bool column = !(want && std::string(want) == "tile") || L > 32;  // column-walking is the faster one (profiles/r2a)
    if (column) {  // measured on c2 (profiles/r2a_assembly.md): 3 resident blocks, layer loop unrolled by 2
        const char* un = getenv("PBSM3D_ASM_UNROLL");
        const int unr = un ? atoi(un) : 2;
        if (unr == 2) {
            if (minb == 4) return setup_assembly_fn(h, assemble_kernel<4, 2>, 0, 128);
            if (minb == 2) return setup_assembly_fn(h, assemble_kernel<2, 2>, 0, 128);
            return setup_assembly_fn(h, assemble_kernel<3, 2>, 0, 128);
        }
        if (minb == 4) return setup_assembly_fn(h, assemble_kernel<4, 1>, 0, 128);
        if (minb == 5) return setup_assembly_fn(h, assemble_kernel<5, 1>, 0, 128);
        if (minb == 6) return setup_assembly_fn(h, assemble_kernel<6, 1>, 0, 128);
        return setup_assembly_fn(h, assemble_kernel<3, 1>, 0, 128);
    }
    if (L <= 4) return setup_assembly_fn(h, assemble_tile_kernel<4, 4>, 4, 128);
    if (L <= 5) return setup_assembly_fn(h, assemble_tile_kernel<5, 4>, 5, 160);
    if (L <= 8) return setup_assembly_fn(h, assemble_tile_kernel<8, 3>, 8, 256);
    if (L <= 10) {
        if (minb == 2) return setup_assembly_fn(h, assemble_tile_kernel<10, 2>, 10, 320);
        if (minb == 4) return setup_assembly_fn(h, assemble_tile_kernel<10, 4>, 10, 320);
        if (minb == 5) return setup_assembly_fn(h, assemble_tile_kernel<10, 5>, 10, 320);
        return setup_assembly_fn(h, assemble_tile_kernel<10, 3>, 10, 320);
    }
    if (L <= 16) return setup_assembly_fn(h, assemble_tile_kernel<16, 2>, 16, 512);
    if (L <= 20) return setup_assembly_fn(h, assemble_tile_kernel<20, 1>, 20, 640);
    return setup_assembly_fn(h, assemble_tile_kernel<32, 1>, 32, 1024);
}
// prelude of CHM faces [i0, i1) (one forcing chunk)
void launch_prelude(pbsm3d_handle* h, const DevForcing& f, double dt, int i0, int i1) {
    LAUNCH(h, face_prelude_kernel, cdiv((size_t)(i1 - i0), 128), 128, h->dc, h->dm, f, h->ss, dt, i0, i1, h->recs);
}
// rows of all slots; {max|b|, sum b^2} of this rank go to red[0..1]
void launch_rows(pbsm3d_handle* h) {
    const int per = h->asm_nw ? 32 : 128;
    const int grid = std::max(1, std::min(cdiv((size_t)h->Tp, per), h->asm_grid));
    ++h->n_launch;
    ((AsmKernel)h->asm_fn)<<<grid, h->asm_threads, h->asm_smem, h->stream>>>(h->dc, h->dm, h->recs, h->ss, 0, h->Tp, h->partial, kRedBlocks,
                                                                            h->sc, h->red);
}
void launch_assembly(pbsm3d_handle* h, const DevForcing& f, double dt) {
    launch_prelude(h, f, dt, 0, h->T);
    launch_rows(h);
}

// One PBSM3D::run with device-resident forcing (reference PBSM3D.cpp:400-1748, phases A–I of SURVEY §3.2).
// ---- providers of two PBSM3D inputs: host side
// U_R, sd: device, CHM order (sd may be null: no module provides snowdepthavg); out: device, CHM order
int enqueue_scale_wind_vert(pbsm3d_handle* h, const pbsm3d_wind_config* wc, const double* U_R, const double* sd, double* out) {
    const int T = h->T;
    const bool canopy_on = !wc->ignore_canopy && h->wv_canopy;
    if (canopy_on && !h->wv_lai) return fail(PBSM3D_ERR_INVALID, "Parameter LAI does not exist.");  // triangulation.hpp:1685-1688
    LAUNCH(h, wind_point_kernel, cdiv(T, 256), 256, T, h->iperm, U_R, sd, h->wv_canopy, h->wv_lai, wc->ignore_canopy, h->wv_u,
           wc->point_mode ? out : nullptr);
    if (wc->point_mode) return 0;
    TRY(halo_exchange(h, h->wv_u, 1));  // domain->ghost_neighbors_communicate_variable("U_2m_above_srf"), scale_wind_vert.cpp:178
    LAUNCH(h, wind_spline_kernel, cdiv(T, 128), 128, T, h->dm, h->cx, h->cy, h->wv_u, out);
    return 0;
}

// Uniform cell grid over the centres of the owned faces (≈4 faces per cell), built once on the host from the
// device-computed centres.  Stands in for the reference's kd-tree of face centres (triangulation.cpp:1037-1058).
int ensure_grid(pbsm3d_handle* h) {
    if (h->grid_ready) return 0;
    const int Tp = h->Tp;
    std::vector<double> cx(Tp), cy(Tp), cz(Tp), can(Tp, 0.0);
    std::vector<int> perm(Tp);
    CU(cudaMemcpyAsync(cx.data(), h->cx, (size_t)Tp * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(cy.data(), h->cy, (size_t)Tp * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(cz.data(), h->cz, (size_t)Tp * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (h->wv_canopy) CU(cudaMemcpyAsync(can.data(), h->wv_canopy, (size_t)Tp * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(perm.data(), h->perm, (size_t)Tp * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int p = 0; p < Tp; ++p)
        if (perm[p] >= 0) { x0 = std::min(x0, cx[p]); x1 = std::max(x1, cx[p]); y0 = std::min(y0, cy[p]); y1 = std::max(y1, cy[p]); }
    const double w = std::max(x1 - x0, 1e-9), hh = std::max(y1 - y0, 1e-9);
    double cell = std::sqrt(2.0 * w * hh / std::max(h->T, 1));  // ≈ 2 faces per cell
    if (!(cell > 0)) cell = 1.0;
    CellGrid g;
    g.x0 = x0; g.y0 = y0; g.h = cell; g.inv_h = 1.0 / cell;
    g.ncx = std::max(1, std::min(1 << 14, (int)(w / cell) + 1));
    g.ncy = std::max(1, std::min(1 << 14, (int)(hh / cell) + 1));
    const size_t nc = (size_t)g.ncx * g.ncy;
    auto cell_of = [&](int p) {
        int ix = (int)std::floor((cx[p] - g.x0) * g.inv_h), iy = (int)std::floor((cy[p] - g.y0) * g.inv_h);
        ix = std::min(std::max(ix, 0), g.ncx - 1);
        iy = std::min(std::max(iy, 0), g.ncy - 1);
        return (size_t)iy * g.ncx + ix;
    };
    const size_t nf = (size_t)std::max(h->T, 1);
    std::vector<int> start(nc + 1, 0);
    std::vector<double2> xy(nf), zc(nf);
    for (int p = 0; p < Tp; ++p)
        if (perm[p] >= 0) start[cell_of(p) + 1]++;
    for (size_t c = 0; c < nc; ++c) start[c + 1] += start[c];
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int p = 0; p < Tp; ++p)
        if (perm[p] >= 0) {
            const int k = fill[cell_of(p)]++;
            xy[k] = make_double2(cx[p], cy[p]);
            zc[k] = make_double2(cz[p], can[p]);
        }
    TRY(h->alloc(&h->grid_start, nc + 1));
    TRY(h->alloc(&h->grid_xy, nf));
    TRY(h->alloc(&h->grid_zc, nf));
    TRY(upload(h, h->grid_start, start.data(), (nc + 1) * sizeof(int)));
    TRY(upload(h, h->grid_xy, xy.data(), nf * sizeof(double2)));
    TRY(upload(h, h->grid_zc, zc.data(), nf * sizeof(double2)));
    CU(cudaStreamSynchronize(h->stream));
    g.cell_start = h->grid_start;
    g.xy = h->grid_xy;
    g.zc = h->grid_zc;
    h->grid = g;
    h->grid_ready = true;
    return 0;
}
int enqueue_fetchr(pbsm3d_handle* h, const pbsm3d_wind_config* wc, const double* vw_dir, double* out) {
    // On a geographic mesh the reference walks up-wind with point_from_bearing_latlong, which returns (lat, lon) where the
    // kd-tree expects (x = lon, y = lat) (coordinates.cpp:33-61): not a behaviour worth reproducing -- refuse, never guess.
    if (h->geographic) return fail(PBSM3D_ERR_UNSUPPORTED, "fetchr on a geographic (lat/long) mesh is not implemented");
    if (wc->fetch_steps < 1 || !(wc->fetch_max_distance > 0)) return fail(PBSM3D_ERR_INVALID, "fetchr: steps and max_distance must be positive");
    TRY(ensure_grid(h));
    LAUNCH(h, fetchr_kernel, cdiv(h->T, 128), 128, h->T, h->iperm, h->grid, h->cx, h->cy, h->cz, h->wv_canopy, vw_dir, wc->fetch_steps,
           wc->fetch_max_distance, wc->fetch_I, wc->fetch_incl_veg, out);
    return 0;
}
// host or device pointers in, host or device pointer out, through the handle's staging buffers
int provider_call(pbsm3d_handle* h, const pbsm3d_wind_config* wc, int which, const double* in0, const double* in1, double* out,
                  int device_ptrs) {
    if (!h || !in0 || !out) return fail(PBSM3D_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    pbsm3d_wind_config local;
    if (!wc) { pbsm3d_wind_config_defaults(&local); wc = &local; }
    const size_t bytes = (size_t)h->T * sizeof(double);
    const double *d0 = in0, *d1 = in1;
    double* dout = out;
    if (!device_ptrs) {
        TRY(upload(h, h->forcing_buf[0], in0, bytes));
        d0 = h->forcing_buf[0];
        if (in1) { TRY(upload(h, h->forcing_buf[2], in1, bytes)); d1 = h->forcing_buf[2]; }
        dout = h->out_stage;
    }
    if (which == 0) TRY(enqueue_scale_wind_vert(h, wc, d0, d1, dout));
    else TRY(enqueue_fetchr(h, wc, d0, dout));
    if (!device_ptrs) CU(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, h->stream));
    TRY(sync_stream(h));
    CU(cudaGetLastError());
    return 0;
}

int step_impl(pbsm3d_handle* h, double dt, const DevForcing& f_in, const double* const* host_in, const OutTargets* out,
              pbsm3d_stats* st) {
    cudaStream_t s = h->stream;
    DevForcing f = f_in;
    // pbsm3d_set_providers: inputs the caller left out are derived on the device (scale_wind_vert, fetchr) before assembly
    const bool derive_u2 = h->providers_on && f.u2 == nullptr;
    const bool derive_fetch = h->providers_on && f.fetch == nullptr && (h->cfg.use_exp_fetch || h->cfg.use_tanh_fetch);
    if (derive_u2) f.u2 = h->forcing_buf[1];
    if (derive_fetch) f.fetch = h->forcing_buf[7];
    const int T = h->T;
    std::memset(st, 0, sizeof(*st));
    const long long launch0 = h->n_launch;
    const double tol2 = h->cfg.tolerance * h->cfg.tolerance;
    const int maxit = h->cfg.max_iterations;
    h->n_syncs = 0;
    h->halo_ops = 0;
    h->halo_fused_ops = 0;
    const auto t_host0 = std::chrono::steady_clock::now();
    h->last_forcing = f;
    h->last_dt = dt;
    const int solver = h->cfg.solver;
    CU(cudaEventRecord(h->ev[0], s));
    // x0 = 0 (Belos starts from the zero vector); ghost tails included
    CU(cudaMemsetAsync(h->x, 0, h->NS * sizeof(double), s));
    CU(cudaMemsetAsync(&h->sc->n_seeds, 0, sizeof(unsigned), s));  // counted by the row kernel
    h->x_first = true;
    // A+B: zeroSystem is implicit (every coefficient is overwritten); saltation + suspension assembly,
    // C: suspension_present = ||rhs||_inf > 1e-12 and ||b||_2^2, reduced inside the assembly kernel
    // With host buffers the forcing crosses PCIe in chunks on its own stream and each chunk is assembled as it lands.
    // With the providers fused in, their inputs (U_R, snowdepthavg, vw_dir) cross first as whole fields and the
    // provider kernels run while the chunks of the other arrays are still on their way.
    const bool derive = derive_u2 || derive_fetch;
    const int nch = host_in ? std::max(1, std::min(kChunks, T / 32768)) : 1;
    auto provider_input = [&](int k) { return derive && (k == 0 || k == 2 || k == 6); };
    if (host_in) {
        if (derive) {
            for (int k = 0; k < kIn; ++k)
                if (host_in[k] && provider_input(k))
                    CU(cudaMemcpyAsync(h->forcing_buf[k], host_in[k], (size_t)T * sizeof(double), cudaMemcpyHostToDevice, h->s_in));
            CU(cudaEventRecord(h->ev_prov, h->s_in));
        }
        for (int c = 0; c < nch; ++c) {
            const size_t i0 = (size_t)T * c / nch, i1 = (size_t)T * (c + 1) / nch;
            for (int k = 0; k < kIn; ++k)
                if (host_in[k] && !provider_input(k))
                    CU(cudaMemcpyAsync(h->forcing_buf[k] + i0, host_in[k] + i0, (i1 - i0) * sizeof(double), cudaMemcpyHostToDevice,
                                       h->s_in));
            CU(cudaEventRecord(h->ev_in[c], h->s_in));
        }
    }
    if (derive) {
        if (host_in) CU(cudaStreamWaitEvent(s, h->ev_prov, 0));
        if (derive_u2) TRY(enqueue_scale_wind_vert(h, &h->wind_cfg, f.U_R, f.sd, h->forcing_buf[1]));
        if (derive_fetch) TRY(enqueue_fetchr(h, &h->wind_cfg, f.vw_dir, h->forcing_buf[7]));
    }
    for (int c = 0; c < nch; ++c) {  // the per-face prelude of each forcing chunk as it lands
        if (host_in) CU(cudaStreamWaitEvent(s, h->ev_in[c], 0));
        launch_prelude(h, f, dt, (int)((size_t)T * c / nch), (int)((size_t)T * (c + 1) / nch));
    }
    launch_rows(h);
    h->have_system = true;
    if (h->n_ranks > 1) {
        TRY(allreduce(h, h->red, 1, true));
        TRY(allreduce(h, h->red + 1, 1, false));
    }
    LAUNCH(h, flags_kernel, 1, 1, FLAGS_SUSP, h->sc, h->red, 1, tol2);
    CU(cudaEventRecord(h->ev[1], s));
    // outputs that are final early leave on their own stream while the solves run (host-buffer entry point)
    const bool early = out && host_in;
    if (early) {
        CU(cudaEventRecord(h->ev_asm, s));
        CU(cudaStreamWaitEvent(h->s_out, h->ev_asm, 0));
        TRY(enqueue_export(h, h->s_out, out, kOutQsalt));
    }

    // D: suspension solve
    int total = 0;
    bool line = (solver == PBSM3D_SOLVER_AUTO || solver == PBSM3D_SOLVER_LINE);
    h->sweeps_timed = 0;
    h->sweeps_timed32 = 0;
    if (line && h->persistent) {
        TRY(line_enqueue_persistent(h));
    } else if (line) {
        TRY(l2_window(h, h->x, h->NS * sizeof(double)));
        TRY(line_enqueue_initial(h, &total));
        TRY(l2_window(h, nullptr, 0));
    } else {
        TRY(read_scalars(h));
        if (h->h_sc->susp_present) {
            bool conv = false;
            TRY(solve_bicgstab(h, &conv));
            if (!conv) return fail(PBSM3D_ERR_NOCONVERGE, "suspension solver failed to converge");
            st->suspension_iterations = h->h_sc->iters;
            st->suspension_residual = std::sqrt(h->h_sc->rr / h->h_sc->bnorm2);
            st->suspension_solver_used = PBSM3D_SOLVER_BICGSTAB;
            LAUNCH(h, flags_kernel, 1, 1, FLAGS_FORCE_SUSP_OK, h->sc, h->red, 0, tol2);
        }
    }
    CU(cudaEventRecord(h->ev[2], s));
    // E..I, optimistic
    int n_cg = std::min(maxit, h->pred_cg > 0 ? h->pred_cg + std::max(4, h->pred_cg / 16) : 64);
    TRY(enqueue_tail(h, f, dt, early ? out : nullptr, n_cg));
    TRY(enqueue_finish(h, f, dt, out, early ? kOutDrift : 0x1ffu));
    if (early) {
        CU(cudaEventRecord(h->ev_out, h->s_out));
        CU(cudaStreamWaitEvent(s, h->ev_out, 0));
    }
    const auto t_enq = std::chrono::steady_clock::now();
    TRY(sync_stream(h));
    if (h->trace) {  // PBSM3D_TRACE=1: where the step's time went (ms since the first event of the step)
        const auto t_end = std::chrono::steady_clock::now();
        auto since = [&](cudaEvent_t e) { float ms = -1.f; cudaEventElapsedTime(&ms, h->ev[0], e); return ms; };
        fprintf(stderr, "[pbsm3d trace] host enqueue %.3f ms, call-to-sync %.3f ms | assembled %.3f line %.3f dep-rhs %.3f drift %.3f",
                std::chrono::duration<double, std::milli>(t_enq - t_host0).count(),
                std::chrono::duration<double, std::milli>(t_end - t_host0).count(), since(h->ev[1]), since(h->ev[2]), since(h->ev[3]),
                since(h->ev[4]));
        if (host_in) {
            fprintf(stderr, " | h2d chunks");
            for (int c = 0; c < nch; ++c) fprintf(stderr, " %.3f", since(h->ev_in[c]));
        }
        if (early) fprintf(stderr, " | qsalt-export-start %.3f flux %.3f early-d2h-done %.3f", since(h->ev_asm), since(h->ev_flux), since(h->ev_out));
        fprintf(stderr, "\n");
    }

    // ---- what actually happened
    bool redo_tail = false;
    bool sor_counted = false;
    auto count_persistent_halos = [&](bool with_x) {  // sweep numbers of the in-kernel halo channels advance by what was executed
        if (!(h->persistent && h->n_ranks > 1)) return;
        if (with_x && line && h->h_sc->susp_sweeps > 0) {
            h->xh_epoch += (unsigned long long)h->h_sc->susp_sweeps;
            h->halo_ops += h->h_sc->susp_sweeps;
            h->halo_fused_ops += h->h_sc->susp_sweeps;
        }
        if (!sor_counted && h->h_sc->dep_sweeps > 0) {
            h->sor_epoch += (unsigned long long)h->h_sc->dep_sweeps;
            h->halo_ops += h->h_sc->dep_sweeps;
            h->halo_fused_ops += h->h_sc->dep_sweeps;
            sor_counted = true;
        }
    };
    count_persistent_halos(true);
    if (line && h->persistent && h->h_sc->susp_present) {
        h->sweeps_timed = h->h_sc->susp_sweeps;
        h->sweeps_timed32 = std::min(h->plan_n32, h->h_sc->susp_sweeps);
    }
    if (line && h->h_sc->susp_present && !h->h_sc->susp_ok) {
        bool conv = false;
        if (!h->persistent) TRY(line_continue(h, total, solver == PBSM3D_SOLVER_AUTO, &conv));
        if (!conv && solver == PBSM3D_SOLVER_AUTO) {
            TRY(solve_bicgstab(h, &conv));
            if (conv) {
                st->suspension_iterations = h->h_sc->iters;
                st->suspension_residual = std::sqrt(h->h_sc->rr / h->h_sc->bnorm2);
                st->suspension_solver_used = PBSM3D_SOLVER_BICGSTAB;
                LAUNCH(h, flags_kernel, 1, 1, FLAGS_FORCE_SUSP_OK, h->sc, h->red, 0, tol2);
            }
        }
        if (!conv) return fail(PBSM3D_ERR_NOCONVERGE, "suspension solver failed to converge");
        CU(cudaEventRecord(h->ev[2], s));
        redo_tail = true;
    }
    if (line && h->h_sc->susp_present && st->suspension_solver_used == 0) {
        const Scalars& c = *h->h_sc;
        st->suspension_iterations = c.susp_iters;
        st->suspension_residual = std::sqrt(c.susp_rr / c.susp_bnorm2);
        st->suspension_solver_used = PBSM3D_SOLVER_LINE;
        // carry the iteration count (not the solution) to the next step's schedule
        int nh = std::min(c.n_checks, 16);
        if (nh >= 2 && c.rr_hist[nh - 1] > 0 && c.rr_hist[nh - 2] > 0 && c.it_hist[nh - 1] > c.it_hist[nh - 2])
            h->sweep_rate2 = std::pow(c.rr_hist[nh - 1] / c.rr_hist[nh - 2], 1.0 / (c.it_hist[nh - 1] - c.it_hist[nh - 2]));
        int pred = c.susp_iters;
        if (c.susp_rr > 0 && h->sweep_rate2 > 0 && h->sweep_rate2 < 1) {
            // sweeps the detection overshot the tolerance by (the check grid is coarser than one sweep)
            double over = std::log(tol2 * c.susp_bnorm2 / c.susp_rr) / std::log(h->sweep_rate2);
            if (over <= -1.0) pred = std::max(1, pred + (int)std::ceil(over + 1e-9));
        }
        h->pred_sweeps = pred;
        // sweeps until the residual is down to 1e-6 ||b|| (||r||^2 contracts by sweep_rate2 per sweep), one in hand
        h->pred_n32 = (h->sweep_rate2 > 0 && h->sweep_rate2 < 1) ? std::max(0, (int)std::floor(std::log(1e-12) / std::log(h->sweep_rate2)) - 1) : 0;
    }
    if (redo_tail) {
        TRY(enqueue_tail(h, f, dt, nullptr, n_cg));
        TRY(enqueue_finish(h, f, dt, out, 0x1ffu));
        TRY(sync_stream(h));
        count_persistent_halos(false);
    }
    // deposition solve still open?
    bool sor = use_sor(h);
    bool cheb = !sor && use_chebyshev(h);
    int dep_used = sor ? PBSM3D_DEP_SOR : (cheb ? PBSM3D_DEP_CHEBYSHEV : PBSM3D_DEP_CG);
    while (h->h_sc->tail_done && h->h_sc->dep_present && !h->h_sc->dep_ok) {
        if (sor && h->persistent && h->sor_persistent) h->sor_enqueued = std::max(h->h_sc->dep_sweeps, std::min(maxit, 6 * h->sor_kest + 64));  // it ran to its bound
        if (sor) {
            const bool gave_up = h->h_sc->done == 2 || h->sor_enqueued >= std::min(maxit, 6 * h->sor_kest + 64);
            if (gave_up && h->cfg.deposition_solver == PBSM3D_DEP_SOR)
                return fail(PBSM3D_ERR_NOCONVERGE, "deposition solver (SOR) failed to converge");
            if (gave_up) {  // the relaxation factor does not suit this mesh: CG needs no parameter
                sor = false;
                dep_used = PBSM3D_DEP_CG;
                h->sor_omega = 0.0;
                LAUNCH(h, flags_kernel, 1, 1, FLAGS_DEP_RESTART, h->sc, h->red, 0, tol2);
                CU(cudaMemsetAsync(h->qA, 0, (size_t)h->S * sizeof(double), s));
                TRY(enqueue_cg_start(h, n_cg));
            } else {
                TRY(enqueue_sor_sweeps(h, std::min(8, maxit - h->sor_enqueued)));
                TRY(enqueue_sor_check(h));
            }
        } else if (cheb) {
            const bool gave_up = h->h_sc->done == 2 || h->cheb_enqueued >= std::min(maxit, 4 * h->cheb_kest + 32);
            if (gave_up && h->cfg.deposition_solver == PBSM3D_DEP_CHEBYSHEV)
                return fail(PBSM3D_ERR_NOCONVERGE, "deposition solver (Chebyshev) failed to converge");
            if (gave_up) {  // the spectrum bounds were not good enough for this mesh: CG needs none
                cheb = false;
                dep_used = PBSM3D_DEP_CG;
                h->cheb_ready = false;
                LAUNCH(h, flags_kernel, 1, 1, FLAGS_DEP_RESTART, h->sc, h->red, 0, tol2);
                CU(cudaMemsetAsync(h->qA, 0, (size_t)h->S * sizeof(double), s));
                TRY(enqueue_cg_start(h, n_cg));
            } else {
                const int k0 = h->cheb_enqueued;
                TRY(enqueue_cheb(h, k0, std::min(maxit, k0 + 16), k0));
            }
        } else {
            if (h->h_sc->done == 2 || h->h_sc->iters >= maxit || !std::isfinite(h->h_sc->rr))
                return fail(PBSM3D_ERR_NOCONVERGE, "deposition solver failed to converge");
            TRY(enqueue_cg_iterations(h, std::min(64, maxit - h->h_sc->iters)));
        }
        TRY(enqueue_finish(h, f, dt, out, 0x1ffu));
        TRY(sync_stream(h));
    }
    const Scalars& c = *h->h_sc;
    st->suspension_present = c.susp_present;
    st->suspension_rhs_max = c.susp_rhs_max;
    st->deposition_present = c.dep_present;
    st->deposition_rhs_max = c.dep_rhs_max;
    if (c.dep_present) {
        st->deposition_iterations = c.iters;
        st->deposition_residual = c.bnorm2 > 0 ? std::sqrt(c.rr / c.bnorm2) : 0.0;
        st->deposition_solver_used = dep_used;
        if (dep_used == PBSM3D_DEP_CHEBYSHEV) h->pred_dep = c.iters;
        else if (dep_used == PBSM3D_DEP_SOR) {
            // sweeps the detection overshot the tolerance by, from the asymptotic rate (omega - 1 per sweep on the error)
            int pred = c.iters;
            const double fac2 = (h->sor_omega - 1.0) * (h->sor_omega - 1.0);
            if (c.rr > 0 && c.bnorm2 > 0 && fac2 > 0 && fac2 < 1) {
                const double over = std::log(tol2 * c.bnorm2 / c.rr) / std::log(fac2);
                if (over <= -2.0) pred = std::max(1, pred + (int)std::ceil(over + 1.0));
            }
            h->pred_sor = pred;
        } else h->pred_cg = c.iters;
    }
    st->host_syncs = h->n_syncs;
    CU(cudaEventElapsedTime(&st->ms_assembly, h->ev[0], h->ev[1]));
    CU(cudaEventElapsedTime(&st->ms_suspension_solve, h->ev[1], h->ev[2]));
    CU(cudaEventElapsedTime(&st->ms_flux_and_halo, h->ev[2], h->ev[3]));
    CU(cudaEventElapsedTime(&st->ms_deposition, h->ev[3], h->ev[4]));
    CU(cudaEventElapsedTime(&st->ms_total, h->ev[0], h->ev[4]));
    if (h->sweeps_timed > 0) {
        CU(cudaEventElapsedTime(&st->ms_line_sweeps, h->ev_sw[0], h->ev_sw[1]));
        CU(cudaEventElapsedTime(&st->ms_line_sweeps_fp32, h->ev_sw[0], h->ev_sw[2]));
    }
    st->residual_checks = h->h_sc->n_checks;
    if (h->persistent && line && h->h_sc->susp_present) {
        st->column_updates_fp32_x = (int64_t)h->h_sc->col_updates[0];
        st->column_updates_fp32 = (int64_t)h->h_sc->col_updates[1];
        st->column_updates_fp64 = (int64_t)h->h_sc->col_updates[2];
        st->columns_checked = (int64_t)h->h_sc->col_updates[3];
        st->active_set = (h->active_set && h->h_sc->n_seeds_step <= (unsigned)live_max_seeds(h)) ? 1 : 0;
        st->faces_with_rhs = (int32_t)h->h_sc->n_seeds_step;
    }
    st->sweeps_fp32_x = h->persistent ? std::min(h->plan_nx32, h->sweeps_timed) : 0;
    st->persistent_kernels = h->persistent ? 1 : 0;
    st->sweeps_timed = h->sweeps_timed;
    st->sweeps_timed_fp32 = h->sweeps_timed32;
    st->n_colours = h->n_colours;
    st->kernel_launches = (int32_t)(h->n_launch - launch0);
    st->halo_exchanges = h->halo_ops;
    st->halo_fused = h->halo_fused_ops;
    st->halo_transport = h->n_ranks == 1 ? PBSM3D_HALO_NONE : (h->peer ? PBSM3D_HALO_PEER : PBSM3D_HALO_NCCL);
    CU(cudaGetLastError());
    return 0;
}

int ensure_scratch(pbsm3d_handle* h, size_t n) {
    if (h->scratch_n >= n) return 0;
    if (h->scratch) h->release(h->scratch);
    h->scratch = nullptr;
    h->scratch_n = 0;
    TRY(h->alloc(&h->scratch, n));
    h->scratch_n = n;
    return 0;
}
// slot-ordered device array(s) -> CHM-ordered host array
int fetch_chm(pbsm3d_handle* h, double* host_dst, const double* dev_src, int rows, size_t src_stride) {
    if (!host_dst) return 0;
    const size_t n = (size_t)rows * h->T;
    TRY(ensure_scratch(h, n));
    LAUNCH(h, to_chm_kernel, cdiv(h->T, 256), 256, rows, h->T, src_stride, h->iperm, dev_src, h->scratch);
    CU(cudaMemcpyAsync(host_dst, h->scratch, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}
int store_chm(pbsm3d_handle* h, double* dev_dst, const double* host_src) {
    if (!host_src) return 0;
    TRY(ensure_scratch(h, h->T));
    TRY(upload(h, h->scratch, host_src, (size_t)h->T * sizeof(double)));
    LAUNCH(h, from_chm_kernel, cdiv(h->T, 256), 256, h->T, h->iperm, h->scratch, dev_dst);
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

template <typename U>
int slots_from_host(pbsm3d_handle* h, U** dst, const U* host_src, U fill) {
    U* tmp = nullptr;
    TRY(h->alloc(&tmp, h->T));
    TRY(upload(h, tmp, host_src, (size_t)h->T * sizeof(U)));
    TRY(h->alloc(dst, h->Tp));
    LAUNCH(h, to_slots_kernel<U>, cdiv(h->Tp, 256), 256, h->Tp, h->perm, tmp, *dst, fill);
    CU(cudaStreamSynchronize(h->stream));
    h->release(tmp);
    return 0;
}

}  // namespace

// ======================================================================================== C ABI
extern "C" {

int pbsm3d_abi_version(void) { return PBSM3D_ABI_VERSION; }
const char* pbsm3d_last_error(void) { return g_last_error.c_str(); }

void pbsm3d_config_defaults(pbsm3d_config* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->nLayer = 10;             // PBSM3D.cpp:223
    c->do_fixed_settling = 0;   // :229
    c->settling_velocity = 0.5; // :233
    c->do_sublimation = 1;      // :244
    c->do_lateral_diff = 1;     // :245
    c->smooth_coeff = 820;      // :246
    c->min_sd_trans = 0.1;      // :247
    c->cutoff = 0.3;            // :249
    c->snow_diffusion_const = 0.3;  // :252
    c->rouault_diffusion_coef = 0;  // :254
    c->enable_veg = 1;          // :256
    c->iterative_subl = 0;      // :258
    c->use_exp_fetch = 0;       // :123
    c->use_tanh_fetch = 1;      // :124
    c->use_PomLi_probability = 0;  // :125
    c->z0_ustar_coupling = 0;   // :126
    c->use_subgrid_topo = 0;    // :129
    c->use_subgrid_topo_V2 = 0; // :130
    c->use_R94_lambda = 1;      // :141
    c->debug_output = 0;        // :145
    c->tolerance = 1e-8;        // LinearAlgebra.cpp:168
    c->max_iterations = 1000;   // LinearAlgebra.cpp:167
    c->solver = PBSM3D_SOLVER_AUTO;
    c->deposition_solver = PBSM3D_DEP_AUTO;
    c->fp32_sweep_streams = 1;
}

int pbsm3d_nccl_unique_id(void* out) {
    if (!out) return fail(PBSM3D_ERR_INVALID, "null output");
    ncclUniqueId id;
    NC(ncclGetUniqueId(&id));
    std::memcpy(out, &id, sizeof(id));
    return 0;
}

void* pbsm3d_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        g_last_error = "cudaMallocHost failed";
        return nullptr;
    }
    return p;
}
void pbsm3d_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

void pbsm3d_destroy(pbsm3d_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->s_in) { cudaStreamSynchronize(h->s_in); cudaStreamDestroy(h->s_in); }
    if (h->s_out) { cudaStreamSynchronize(h->s_out); cudaStreamDestroy(h->s_out); }
    for (auto& e : h->ev_in)
        if (e) cudaEventDestroy(e);
    if (h->ev_asm) cudaEventDestroy(h->ev_asm);
    if (h->ev_flux) cudaEventDestroy(h->ev_flux);
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    if (h->ev_prov) cudaEventDestroy(h->ev_prov);
    for (void* b : h->peer_base)
        if (b) cudaIpcCloseMemHandle(b);
    if (h->comm) ncclCommDestroy(h->comm);
    for (void* p : h->allocs) cudaFree(p);
    if (h->h_sc) cudaFreeHost(h->h_sc);
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : h->ev_sw)
        if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

static int create_impl(pbsm3d_handle* h, const pbsm3d_config* cfg, const pbsm3d_mesh* mesh, int device, const pbsm3d_comm* comm) {
    // ---- config checks: same errors the reference raises, plus "unsupported" for optional paths
    if (cfg->use_exp_fetch && cfg->use_tanh_fetch)
        return fail(PBSM3D_ERR_INVALID, "PBSM3d: Cannot specify both exp_fetch and tanh_fetch");  // PBSM3D.cpp:132-135
    if (cfg->settling_velocity < 0) return fail(PBSM3D_ERR_INVALID, "PBSM3D settling velocity must be positive");  // :239-242
    if (cfg->nLayer < 1) return fail(PBSM3D_ERR_INVALID, "nLayer must be >= 1");  // nLayer == 1: only the z == 0 branch runs (PBSM3D.cpp:1285-1321)
    if (cfg->iterative_subl) return fail(PBSM3D_ERR_UNSUPPORTED, "iterative_subl is not implemented");
    if (cfg->z0_ustar_coupling) return fail(PBSM3D_ERR_UNSUPPORTED, "z0_ustar_coupling is not implemented");
    if (cfg->use_subgrid_topo || cfg->use_subgrid_topo_V2) return fail(PBSM3D_ERR_UNSUPPORTED, "use_subgrid_topo* is not implemented");
    if (cfg->debug_output) return fail(PBSM3D_ERR_UNSUPPORTED, "debug_output is not implemented");
    if (!(cfg->tolerance > 0) || cfg->max_iterations < 1) return fail(PBSM3D_ERR_INVALID, "bad solver controls");
    if (cfg->solver < 0 || cfg->solver > 2 || cfg->deposition_solver < 0 || cfg->deposition_solver > 3)
        return fail(PBSM3D_ERR_INVALID, "unknown solver id");
    if (mesh->n_local < 1 || mesh->n_ghost < 0 || !mesh->neigh || !mesh->vertices || !mesh->global_id)
        return fail(PBSM3D_ERR_INVALID, "mesh arrays missing");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(PBSM3D_ERR_CUDA, "no CUDA device: libpbsm3d_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(PBSM3D_ERR_INVALID, "device index out of range");
    CU(cudaSetDevice(device));
    h->device = device;
    h->cfg = *cfg;
    const int T = h->T = mesh->n_local, nG = h->nG = mesh->n_ghost, L = h->L = cfg->nLayer;
    h->G = mesh->n_global;
    h->geographic = mesh->is_geographic != 0;
    h->rank = comm ? comm->rank : 0;
    h->n_ranks = comm ? comm->n_ranks : 1;
    if (h->n_ranks > 1 && !comm->nccl_unique_id) return fail(PBSM3D_ERR_INVALID, "nccl_unique_id missing");
    if (h->n_ranks == 1 && nG != 0) return fail(PBSM3D_ERR_INVALID, "ghost faces on a single-rank mesh");
    if (nG > 0 && !mesh->ghost_owner) return fail(PBSM3D_ERR_INVALID, "ghost_owner missing");
    // owned faces are one contiguous ascending global range (triangulation.cpp:1482-1531)
    h->gstart_id = mesh->global_id[0];
    for (int i = 0; i < T; ++i)
        if (mesh->global_id[i] != h->gstart_id + i)
            return fail(PBSM3D_ERR_INVALID, "owned faces must be a contiguous ascending range of cell_global_id");
    for (int g = 1; g < nG; ++g)
        if (mesh->global_id[T + g] <= mesh->global_id[T + g - 1])
            return fail(PBSM3D_ERR_INVALID, "ghost faces must be sorted by cell_global_id");
    if (h->gstart_id < 0 || h->gstart_id + T > h->G) return fail(PBSM3D_ERR_INVALID, "global ids exceed n_global");
    for (int i = 0; i < T; ++i)
        for (int j = 0; j < 3; ++j) {
            int n = mesh->neigh[(size_t)i * 3 + j];
            if (n < -1 || n >= T + nG || n == i)
                return fail(PBSM3D_ERR_INVALID, "Face " + std::to_string(i) + " has out of bound neighbors.");
        }

    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (auto& e : h->ev_in) CU(cudaEventCreate(&e));
    CU(cudaEventCreate(&h->ev_asm));
    CU(cudaEventCreate(&h->ev_flux));
    CU(cudaEventCreate(&h->ev_out));
    CU(cudaEventCreate(&h->ev_prov));
    h->trace = getenv("PBSM3D_TRACE") != nullptr;
    {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        h->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
        h->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
        // Off unless asked for: pinning x in a persisting carve-out makes a sweep ≈8 % faster on config c2 but takes the
        // carve-out away from the deposition solve, whose whole working set otherwise lives in L2 (profiles/r1c).
        const char* mb = getenv("PBSM3D_L2_PERSIST_MB");
        h->l2_persist_max = mb ? std::min(h->l2_persist_max, (size_t)atol(mb) << 20) : 0;
        if (getenv("PBSM3D_VERBOSE"))
            fprintf(stderr, "[pbsm3d] L2 %d MB, persisting max %zu MB (using %zu MB), window max %zu MB\n", prop.l2CacheSize >> 20,
                    (size_t)prop.persistingL2CacheMaxSize >> 20, h->l2_persist_max >> 20, h->l2_window_max >> 20);
        if (h->l2_persist_max > 0) CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, h->l2_persist_max));
    }
    for (auto& e : h->ev) CU(cudaEventCreate(&e));
    for (auto& e : h->ev_sw) CU(cudaEventCreate(&e));
    CU(cudaMallocHost((void**)&h->h_sc, sizeof(Scalars)));
    std::memset(h->h_sc, 0, sizeof(Scalars));

    // ---- who needs which of my faces (multi-rank): decides which faces go first inside their colour class
    TRY(h->alloc_zero(&h->sc, 1));
    if (h->n_ranks > 1) TRY(comm_handshake(h, mesh, comm));

    // ---- colour-major slot order; inside a colour: the faces a partner rank needs first (so the blocks that own
    // them run first and their halo is under way while the interior is computed), each group in CHM's order
    std::vector<int> colour;
    colour_faces(T, mesh->neigh, colour, h->n_colours);
    if (h->n_colours > kMaxColours) return fail(PBSM3D_ERR_INVALID, "face colouring needs too many colours");
    std::vector<int> count(h->n_colours, 0), bcount(h->n_colours, 0);
    const bool have_bnd = !h->is_boundary.empty();
    for (int i = 0; i < T; ++i) {
        count[colour[i]]++;
        if (have_bnd && h->is_boundary[i]) bcount[colour[i]]++;
    }
    int Tp = 0;
    h->nb_total = 0;
    for (int c = 0; c < h->n_colours; ++c) {
        h->cstart[c] = Tp;
        h->ccount[c] = count[c];
        h->nb[c] = bcount[c];
        h->boff[c] = h->nb_total;
        h->nb_total += bcount[c];
        Tp = align_up(Tp + count[c], 32);
    }
    h->Tp = Tp;
    h->S = Tp + align_up(nG, 32);
    h->N = (size_t)L * Tp;
    h->NS = (size_t)L * h->S;
    std::vector<int> perm(Tp, -1), iperm(T), fill_b(h->n_colours), fill_i(h->n_colours);
    for (int c = 0; c < h->n_colours; ++c) { fill_b[c] = h->cstart[c]; fill_i[c] = h->cstart[c] + bcount[c]; }
    for (int i = 0; i < T; ++i) {
        const int c = colour[i];
        int p = (have_bnd && h->is_boundary[i]) ? fill_b[c]++ : fill_i[c]++;
        perm[p] = i;
        iperm[i] = p;
    }
    TRY(h->alloc(&h->perm, Tp));
    TRY(h->alloc(&h->iperm, T));
    TRY(upload(h, h->perm, perm.data(), (size_t)Tp * sizeof(int)));
    TRY(upload(h, h->iperm, iperm.data(), (size_t)T * sizeof(int)));
    {
        int* d_neigh = nullptr;
        TRY(h->alloc(&d_neigh, (size_t)3 * T));
        TRY(upload(h, d_neigh, mesh->neigh, (size_t)3 * T * sizeof(int)));
        TRY(h->alloc(&h->nbs, (size_t)3 * Tp));
        LAUNCH(h, neighbour_slots_kernel, cdiv(Tp, 256), 256, T, Tp, h->perm, h->iperm, d_neigh, h->nbs);
        CU(cudaStreamSynchronize(h->stream));
        h->release(d_neigh);
    }
    // ghost owner blocks
    std::vector<int> gs(std::max(nG, 1), 0), gc(std::max(nG, 1), 0);
    for (int g = 0; g < nG;) {
        int e = g;
        while (e < nG && mesh->ghost_owner[e] == mesh->ghost_owner[g]) ++e;
        for (int k = g; k < e; ++k) { gs[k] = g; gc[k] = e - g; }
        g = e;
    }
    TRY(h->alloc(&h->gstart, nG));
    TRY(h->alloc(&h->gcnt, nG));
    TRY(upload(h, h->gstart, gs.data(), std::max(nG, 1) * sizeof(int)));
    TRY(upload(h, h->gcnt, gc.data(), std::max(nG, 1) * sizeof(int)));

    // ---- geometry on the device
    {
        const size_t Tall = (size_t)T + nG;
        double* d_verts = nullptr;
        double* d_area_param = nullptr;
        TRY(h->alloc(&d_verts, Tall * 9));
        TRY(upload(h, d_verts, mesh->vertices, Tall * 9 * sizeof(double)));
        if (mesh->area) {
            TRY(h->alloc(&d_area_param, T));
            TRY(upload(h, d_area_param, mesh->area, T * sizeof(double)));
        }
        TRY(h->alloc(&h->nx, (size_t)3 * Tp));
        TRY(h->alloc(&h->ny, (size_t)3 * Tp));
        TRY(h->alloc(&h->elen, (size_t)3 * Tp));
        TRY(h->alloc(&h->dx, (size_t)3 * Tp));
        TRY(h->alloc(&h->area, Tp));
        TRY(h->alloc(&h->cx, (size_t)Tp + nG));
        TRY(h->alloc(&h->cy, (size_t)Tp + nG));
        TRY(h->alloc(&h->cz, (size_t)Tp + nG));
        TRY(h->alloc(&h->slope, Tp));
        LAUNCH(h, geometry_kernel, cdiv((size_t)Tp + nG, 256), 256, T, Tp, nG, h->perm, d_verts, d_area_param, h->nx, h->ny, h->elen,
               h->area, h->cx, h->cy, h->cz, h->slope);
        TRY(h->alloc(&h->ddiag, Tp));
        TRY(h->alloc(&h->doff, (size_t)3 * Tp));
        TRY(h->alloc(&h->dinv, Tp));
        LAUNCH(h, deposition_matrix_kernel, cdiv(Tp, 256), 256, Tp, cfg->smooth_coeff, h->geographic ? 1 : 0, h->perm, h->nbs, h->elen, h->area, h->cx, h->cy,
               h->dx, h->ddiag, h->doff, h->dinv);
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaGetLastError());
        h->release(d_verts);
        if (d_area_param) h->release(d_area_param);
    }

    // ---- vegetation (PBSM3D.cpp:284-324): no vegetation information => veg off for the whole run
    bool veg = cfg->enable_veg && mesh->canopy_height != nullptr;
    if (veg && cfg->use_R94_lambda && !mesh->lai) return fail(PBSM3D_ERR_INVALID, "Parameter LAI does not exist.");
    if (veg) {
        TRY(slots_from_host(h, &h->canopy, mesh->canopy_height, 0.0));
        if (cfg->use_R94_lambda) {
            TRY(slots_from_host(h, &h->lai, mesh->lai, 0.0));
        } else {
            if (mesh->stalk_number) TRY(slots_from_host(h, &h->stalk_n, mesh->stalk_number, 1.0));
            if (mesh->stalk_diameter) TRY(slots_from_host(h, &h->stalk_dv, mesh->stalk_diameter, 0.8));
        }
    }
    if (mesh->is_water) TRY(slots_from_host<unsigned char>(h, &h->water, mesh->is_water, (unsigned char)0));
    // the providers read vegetation whatever PBSM3D's own enable_veg / use_R94_lambda say
    if (mesh->canopy_height) {
        if (h->canopy) h->wv_canopy = h->canopy; else TRY(slots_from_host(h, &h->wv_canopy, mesh->canopy_height, 0.0));
        if (mesh->lai) { if (h->lai) h->wv_lai = h->lai; else TRY(slots_from_host(h, &h->wv_lai, mesh->lai, 0.0)); }
    }
    TRY(h->alloc_zero(&h->wv_u, h->S));
    pbsm3d_wind_config_defaults(&h->wind_cfg);

    // ---- per-step arrays
    for (auto& b : h->forcing_buf) TRY(h->alloc(&b, T));
    SuspSystem& ss = h->ss;
    const size_t N = h->N;
    TRY(h->alloc(&ss.den, N));
    TRY(h->alloc(&ss.cp, N));
    TRY(h->alloc(&ss.latS, 3 * N));
    TRY(h->alloc(&ss.belowS, N));
    TRY(h->alloc(&ss.pack32, N));
    TRY(h->alloc(&ss.cp32, N));
    TRY(h->alloc(&ss.u_z, N));
    TRY(h->alloc(&ss.csubl, N));
    TRY(h->alloc(&ss.rhs0, Tp));
    TRY(h->alloc(&ss.rhsS0, Tp));
    TRY(h->alloc_zero(&ss.Qsalt, h->S));
    TRY(h->alloc(&ss.c_salt, Tp));
    TRY(h->alloc(&ss.salt, Tp));
    TRY(h->alloc_zero(&ss.live, h->S));
    TRY(h->alloc(&ss.prob, Tp));
    TRY(h->alloc_zero(&h->x, h->NS));
    TRY(h->alloc_zero(&h->Qsusp, h->S));
    TRY(h->alloc_zero(&h->cg_p, h->S));
    TRY(h->alloc_zero(&h->qA, h->S));
    TRY(h->alloc_zero(&h->qB, h->S));
    TRY(h->alloc(&h->offS, (size_t)3 * Tp));
    LAUNCH(h, deposition_scale_kernel, cdiv(Tp, 256), 256, Tp, h->doff, h->dinv, h->offS);
    double** perslot[] = {&h->Qsubl, &h->Qsubl_mass, &h->sum_subl, &h->drift_mass, &h->sum_drift, &h->more_avail,
                          &h->drhs,  &h->drhsS,      &h->cg_r,     &h->cg_Ap};
    for (double** p : perslot) TRY(h->alloc_zero(p, Tp));
    TRY(h->alloc(&h->out_stage, (size_t)kOut * T));
    // drift_mass is a face variable that is -9999 until first written (variablestorage default)
    LAUNCH(h, fill_slots_kernel, cdiv(Tp, 256), 256, Tp, h->perm, h->drift_mass, -9999.0);
    LAUNCH(h, fill_slots_kernel, cdiv(Tp, 256), 256, Tp, h->perm, h->ss.prob, -9999.0);  // blowingsnow_probability: unset face variable
    TRY(h->alloc_zero(&h->partial, (size_t)kRedBlocks * 4));
    TRY(h->alloc_zero(&h->red, 8));

    DevConfig& dc = h->dc;
    dc.L = L;
    dc.do_fixed_settling = cfg->do_fixed_settling;
    dc.do_sublimation = cfg->do_sublimation;
    dc.do_lateral_diff = cfg->do_lateral_diff;
    dc.rouault = cfg->rouault_diffusion_coef;
    dc.enable_veg = veg ? 1 : 0;
    dc.use_exp_fetch = cfg->use_exp_fetch;
    dc.use_tanh_fetch = cfg->use_tanh_fetch;
    dc.use_R94_lambda = cfg->use_R94_lambda;
    dc.use_PomLi = cfg->use_PomLi_probability;
    dc.settling_velocity = cfg->settling_velocity;
    dc.eps = cfg->smooth_coeff;
    dc.min_sd_trans = cfg->min_sd_trans;
    dc.cutoff = cfg->cutoff;
    dc.snow_diffusion_const = cfg->snow_diffusion_const;
    dc.dz = 5.0 / (double)L;  // susp_depth / nLayer
    dc.inv_dz = 1.0 / dc.dz;
    dc.l_max = 40.0;
    DevMesh& dm = h->dm;
    dm.T = T;
    dm.Tp = Tp;
    dm.S = h->S;
    dm.nG = nG;
    dm.perm = h->perm;
    dm.iperm = h->iperm;
    dm.nbs = h->nbs;
    dm.nx = h->nx;
    dm.ny = h->ny;
    dm.elen = h->elen;
    dm.area = h->area;
    dm.zc = h->cz;
    dm.canopy = h->canopy;
    dm.lai = h->lai;
    dm.stalk_n = h->stalk_n;
    dm.stalk_dv = h->stalk_dv;
    dm.water = h->water;

    {   // per-layer constants of a column with hs = 0 (every non-saltating face)
        double* tab = nullptr;
        TRY(h->alloc(&tab, (size_t)kTabN * L));
        LAUNCH(h, layer_table_kernel, 1, std::max(32, align_up(L, 32)), h->dc, tab);
        h->ss.ltab = tab;
    }
    TRY(setup_assembly(h));
    LAUNCH(h, assemble_pads_kernel, cdiv(Tp, 256), 256, h->dm, h->ss, L);
    if (h->n_ranks > 1) TRY(setup_comm(h, iperm));
    TRY(setup_persistent(h));  // after the transport is known: across ranks the persistent kernels need the in-kernel halos
    CU(cudaStreamSynchronize(h->stream));
    if (cfg->deposition_solver != PBSM3D_DEP_CG) TRY(estimate_spectrum(h));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return 0;
}

int pbsm3d_create(const pbsm3d_config* cfg, const pbsm3d_mesh* mesh, int device, const pbsm3d_comm* comm, pbsm3d_handle** out) {
    if (!cfg || !mesh || !out) return fail(PBSM3D_ERR_INVALID, "null argument");
    *out = nullptr;
    pbsm3d_handle* h = new pbsm3d_handle();
    int rc = create_impl(h, cfg, mesh, device, comm);
    if (rc) {
        std::string keep = g_last_error;
        pbsm3d_destroy(h);
        g_last_error = keep;
        return rc;
    }
    *out = h;
    return 0;
}

static bool forcing_complete(const pbsm3d_handle* h, const pbsm3d_forcing* f) {
    const bool fetch_on = h->cfg.use_exp_fetch || h->cfg.use_tanh_fetch;  // depends("fetch"), PBSM3D.cpp:181-184
    return f->U_R && (f->U_2m_above_srf || h->providers_on) && f->snowdepthavg && f->swe && f->t && f->rh && f->vw_dir &&
           (f->fetch || !fetch_on || h->providers_on) && (f->p_snow_hours || !h->cfg.use_PomLi_probability);
}
static void out_pointers(const pbsm3d_outputs* o, double* p[kOut]) {
    p[0] = o->Qsalt; p[1] = o->Qsusp; p[2] = o->Qsubl; p[3] = o->Qsubl_mass; p[4] = o->sum_subl; p[5] = o->drift_mass;
    p[6] = o->sum_drift; p[7] = o->pbsm_more_than_avail; p[8] = o->blowingsnow_probability;
}

int pbsm3d_step_device(pbsm3d_handle* h, double dt, const pbsm3d_forcing* f, const pbsm3d_outputs* out, pbsm3d_stats* stats) {
    if (!h || !f) return fail(PBSM3D_ERR_INVALID, "null argument");
    if (!forcing_complete(h, f)) return fail(PBSM3D_ERR_INVALID, "forcing array missing");
    CU(cudaSetDevice(h->device));
    pbsm3d_stats local;
    if (!stats) stats = &local;
    DevForcing df{f->U_R, f->U_2m_above_srf, f->snowdepthavg, f->swe, f->t, f->rh, f->vw_dir, f->fetch, f->p_snow_hours};
    OutTargets ot{};
    if (out) out_pointers(out, ot.dst);
    return step_impl(h, dt, df, nullptr, out ? &ot : nullptr, stats);
}

int pbsm3d_step(pbsm3d_handle* h, double dt, const pbsm3d_forcing* f, const pbsm3d_outputs* out, pbsm3d_stats* stats) {
    if (!h || !f) return fail(PBSM3D_ERR_INVALID, "null argument");
    if (!forcing_complete(h, f)) return fail(PBSM3D_ERR_INVALID, "forcing array missing");
    CU(cudaSetDevice(h->device));
    pbsm3d_stats local;
    if (!stats) stats = &local;
    const double* src[kIn] = {f->U_R, f->U_2m_above_srf, f->snowdepthavg, f->swe, f->t, f->rh, f->vw_dir, f->fetch, f->p_snow_hours};
    DevForcing df{h->forcing_buf[0], f->U_2m_above_srf ? h->forcing_buf[1] : nullptr, h->forcing_buf[2], h->forcing_buf[3],
                  h->forcing_buf[4], h->forcing_buf[5], h->forcing_buf[6], f->fetch ? h->forcing_buf[7] : nullptr,
                  f->p_snow_hours ? h->forcing_buf[8] : nullptr};
    OutTargets ot{};
    if (out) {
        out_pointers(out, ot.host);
        for (int k = 0; k < kOut; ++k) ot.dst[k] = ot.host[k] ? h->out_stage + (size_t)k * h->T : nullptr;
    }
    return step_impl(h, dt, df, src, out ? &ot : nullptr, stats);
}

int pbsm3d_get_state(pbsm3d_handle* h, double* sum_drift, double* sum_subl, double* drift_mass, double* more) {
    if (!h) return fail(PBSM3D_ERR_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    TRY(fetch_chm(h, sum_drift, h->sum_drift, 1, h->Tp));
    TRY(fetch_chm(h, sum_subl, h->sum_subl, 1, h->Tp));
    TRY(fetch_chm(h, drift_mass, h->drift_mass, 1, h->Tp));
    TRY(fetch_chm(h, more, h->more_avail, 1, h->Tp));
    return 0;
}

int pbsm3d_set_state(pbsm3d_handle* h, const double* sum_drift, const double* sum_subl, const double* drift_mass, const double* more) {
    if (!h) return fail(PBSM3D_ERR_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    TRY(store_chm(h, h->sum_drift, sum_drift));
    TRY(store_chm(h, h->sum_subl, sum_subl));
    TRY(store_chm(h, h->drift_mass, drift_mass));
    TRY(store_chm(h, h->more_avail, more));
    return 0;
}

int pbsm3d_get_geometry(pbsm3d_handle* h, double* nx, double* ny, double* el, double* area, double* dx, double* cx, double* cy, double* cz) {
    if (!h) return fail(PBSM3D_ERR_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    const size_t Tp = h->Tp;
    TRY(fetch_chm(h, nx, h->nx, 3, Tp));
    TRY(fetch_chm(h, ny, h->ny, 3, Tp));
    TRY(fetch_chm(h, el, h->elen, 3, Tp));
    TRY(fetch_chm(h, area, h->area, 1, Tp));
    TRY(fetch_chm(h, dx, h->dx, 3, Tp));
    TRY(fetch_chm(h, cx, h->cx, 1, Tp));
    TRY(fetch_chm(h, cy, h->cy, 1, Tp));
    TRY(fetch_chm(h, cz, h->cz, 1, Tp));
    return 0;
}

int pbsm3d_get_layout(pbsm3d_handle* h, int32_t* n_colours, int32_t* n_slots, int32_t* slot_of_face, int32_t* colour_of_face) {
    if (!h) return fail(PBSM3D_ERR_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    if (n_colours) *n_colours = h->n_colours;
    if (n_slots) *n_slots = h->Tp;
    if (slot_of_face || colour_of_face) {
        std::vector<int> slot(h->T);
        CU(cudaMemcpyAsync(slot.data(), h->iperm, (size_t)h->T * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < h->T; ++i) {
            if (slot_of_face) slot_of_face[i] = slot[i];
            if (colour_of_face) {
                int c = 0;
                while (c + 1 < h->n_colours && slot[i] >= h->cstart[c + 1]) ++c;
                colour_of_face[i] = c;
            }
        }
    }
    return 0;
}

int pbsm3d_get_solution(pbsm3d_handle* h, double* x) {
    if (!h || !x) return fail(PBSM3D_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    return fetch_chm(h, x, h->x, h->L, h->S);
}

int pbsm3d_get_suspension_system(pbsm3d_handle* h, double* diag, double* lat, double* below, double* above, double* rhs0,
                                 double* u_z, double* csubl, double* c_salt, uint8_t* saltation) {
    if (!h) return fail(PBSM3D_ERR_INVALID, "null handle");
    if (!h->have_system) return fail(PBSM3D_ERR_INVALID, "no system assembled yet");
    CU(cudaSetDevice(h->device));
    const size_t Tp = h->Tp;
    const int L = h->L;
    if (diag || lat || below || above) {  // the reference's own coefficients from the stored pivots and scaled rows
        double* tmp = nullptr;
        const size_t N = h->N;
        TRY(h->alloc(&tmp, 6 * N));
        LAUNCH(h, reconstruct_rows_kernel, cdiv(Tp, 128), 128, h->ss, (int)Tp, L, tmp, tmp + N, tmp + 4 * N, tmp + 5 * N);
        int rc = fetch_chm(h, diag, tmp, L, Tp);
        if (!rc) rc = fetch_chm(h, lat, tmp + N, 3 * L, Tp);
        if (!rc) rc = fetch_chm(h, below, tmp + 4 * N, L, Tp);
        if (!rc) rc = fetch_chm(h, above, tmp + 5 * N, L, Tp);
        h->release(tmp);
        if (rc) return rc;
    }
    TRY(fetch_chm(h, rhs0, h->ss.rhs0, 1, Tp));
    TRY(fetch_chm(h, u_z, h->ss.u_z, L, Tp));
    TRY(fetch_chm(h, csubl, h->ss.csubl, L, Tp));
    TRY(fetch_chm(h, c_salt, h->ss.c_salt, 1, Tp));
    if (saltation) {
        TRY(ensure_scratch(h, h->T));
        unsigned char* tmp = (unsigned char*)h->scratch;
        LAUNCH(h, to_chm_u8_kernel, cdiv(h->T, 256), 256, h->T, h->iperm, h->ss.salt, tmp);
        CU(cudaMemcpyAsync(saltation, tmp, h->T, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

int pbsm3d_get_deposition_system(pbsm3d_handle* h, double* diag, double* off, double* rhs, double* q) {
    if (!h) return fail(PBSM3D_ERR_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    const size_t Tp = h->Tp;
    TRY(fetch_chm(h, diag, h->ddiag, 1, Tp));
    TRY(fetch_chm(h, off, h->doff, 3, Tp));
    TRY(fetch_chm(h, rhs, h->drhs, 1, Tp));
    TRY(fetch_chm(h, q, h->h_sc->dep_buf ? h->qB : h->qA, 1, Tp));
    return 0;
}

// ---- providers of two PBSM3D inputs ------------------------------------------------------
void pbsm3d_wind_config_defaults(pbsm3d_wind_config* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->ignore_canopy = 0;          // scale_wind_vert.cpp:161
    c->point_mode = 0;             // domain mode unless CHM runs in point mode (scale_wind_vert.cpp:140-142)
    c->fetch_steps = 10;           // fetchr.cpp:34
    c->fetch_max_distance = 1000;  // :36
    c->fetch_I = 0.06;             // :41
    c->fetch_incl_veg = 1;         // :43
}

int pbsm3d_scale_wind_vert(pbsm3d_handle* h, const pbsm3d_wind_config* wc, const double* U_R, const double* snowdepthavg,
                           double* U_2m_above_srf, int device_ptrs) {
    return provider_call(h, wc, 0, U_R, snowdepthavg, U_2m_above_srf, device_ptrs);
}
int pbsm3d_fetchr(pbsm3d_handle* h, const pbsm3d_wind_config* wc, const double* vw_dir, double* fetch, int device_ptrs) {
    return provider_call(h, wc, 1, vw_dir, nullptr, fetch, device_ptrs);
}
int pbsm3d_set_providers(pbsm3d_handle* h, const pbsm3d_wind_config* wc) {
    if (!h) return fail(PBSM3D_ERR_INVALID, "null handle");
    h->providers_on = wc != nullptr;
    if (wc) h->wind_cfg = *wc;
    return 0;
}

// ---- the consumer of drift_mass: snobal's snowpack mass adjustment -----------------------
void pbsm3d_snobal_config_defaults(pbsm3d_snobal_config* c) {
    if (!c) return;
    c->drift_density = 300.0;    // snobal.cpp:83
    c->threshold = 0.2;          // snobal.cpp:190
    c->max_active_layer = 0.1;   // snobal.cpp:101
}

static int adj_snow_call(pbsm3d_handle* h, const pbsm3d_snobal_config* cfg, const pbsm3d_snowpack* pk, int mode, const double* a,
                         const double* b, double* swe_out, double* depth_out, int device_ptrs) {
    if (!h || !pk) return fail(PBSM3D_ERR_INVALID, "null argument");
    double* const* fields[19] = {&pk->z_s, &pk->m_s, &pk->rho, nullptr, &pk->z_s_0, &pk->z_s_l, &pk->m_s_0, &pk->m_s_l, &pk->cc_s,
                                 &pk->cc_s_0, &pk->cc_s_l, &pk->T_s, &pk->T_s_0, &pk->T_s_l, &pk->h2o_total, &pk->h2o_vol, &pk->h2o,
                                 &pk->h2o_max, &pk->h2o_sat};
    for (int k = 0; k < 19; ++k)
        if (k == 3 ? pk->layer_count == nullptr : *fields[k] == nullptr) return fail(PBSM3D_ERR_INVALID, "snowpack: every state array is required");
    if (mode == 1 && (!a || !b)) return fail(PBSM3D_ERR_INVALID, "apply_avalanche: both delta arrays are required");
    pbsm3d_snobal_config local;
    if (!cfg) { pbsm3d_snobal_config_defaults(&local); cfg = &local; }
    if (!(cfg->drift_density > 0)) return fail(PBSM3D_ERR_INVALID, "drift_density must be positive");
    CU(cudaSetDevice(h->device));
    const int T = h->T;
    const size_t n = (size_t)T, bytes = n * sizeof(double);
    SnowpackPtrs p;
    double* dptr[19];
    const double *da = a, *db = b;
    double *dswe = swe_out, *ddepth = depth_out;
    int a_slots = 0;
    if (mode == 0 && !a) {  // the handle's own drift_mass (slot order), left on the device by the last step
        da = h->drift_mass;
        a_slots = 1;
    }
    if (!device_ptrs) {
        if (!h->sno_stage) TRY(h->alloc(&h->sno_stage, 23 * n));
        for (int k = 0; k < 19; ++k) {
            dptr[k] = h->sno_stage + k * n;
            if (k == 3) CU(cudaMemcpyAsync(dptr[k], pk->layer_count, n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
            else CU(cudaMemcpyAsync(dptr[k], *fields[k], bytes, cudaMemcpyHostToDevice, h->stream));
        }
        if (a) { CU(cudaMemcpyAsync(h->sno_stage + 19 * n, a, bytes, cudaMemcpyHostToDevice, h->stream)); da = h->sno_stage + 19 * n; }
        if (b) { CU(cudaMemcpyAsync(h->sno_stage + 20 * n, b, bytes, cudaMemcpyHostToDevice, h->stream)); db = h->sno_stage + 20 * n; }
        dswe = swe_out ? h->sno_stage + 21 * n : nullptr;
        ddepth = depth_out ? h->sno_stage + 22 * n : nullptr;
    } else {
        for (int k = 0; k < 19; ++k) dptr[k] = k == 3 ? (double*)pk->layer_count : *fields[k];
    }
    p.z_s = dptr[0]; p.m_s = dptr[1]; p.rho = dptr[2]; p.layer_count = (int*)dptr[3];
    p.z_s_0 = dptr[4]; p.z_s_l = dptr[5]; p.m_s_0 = dptr[6]; p.m_s_l = dptr[7];
    p.cc_s = dptr[8]; p.cc_s_0 = dptr[9]; p.cc_s_l = dptr[10];
    p.T_s = dptr[11]; p.T_s_0 = dptr[12]; p.T_s_l = dptr[13];
    p.h2o_total = dptr[14]; p.h2o_vol = dptr[15]; p.h2o = dptr[16]; p.h2o_max = dptr[17]; p.h2o_sat = dptr[18];
    const SnobalConst kc{cfg->drift_density, cfg->threshold, cfg->max_active_layer};
    if (mode == 0)
        LAUNCH(h, snobal_adj_snow_kernel<0>, cdiv(T, 256), 256, T, p, kc, da, db, h->iperm, a_slots, h->area, dswe, ddepth);
    else
        LAUNCH(h, snobal_adj_snow_kernel<1>, cdiv(T, 256), 256, T, p, kc, da, db, h->iperm, 0, h->area, dswe, ddepth);
    if (!device_ptrs) {
        for (int k = 0; k < 19; ++k) {
            if (k == 3) CU(cudaMemcpyAsync(pk->layer_count, dptr[k], n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            else CU(cudaMemcpyAsync(*fields[k], dptr[k], bytes, cudaMemcpyDeviceToHost, h->stream));
        }
        if (swe_out) CU(cudaMemcpyAsync(swe_out, dswe, bytes, cudaMemcpyDeviceToHost, h->stream));
        if (depth_out) CU(cudaMemcpyAsync(depth_out, ddepth, bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    TRY(sync_stream(h));
    CU(cudaGetLastError());
    return 0;
}

int pbsm3d_apply_drift(pbsm3d_handle* h, const pbsm3d_snobal_config* cfg, const pbsm3d_snowpack* pack, const double* drift_mass,
                       double* swe_out, double* snowdepth_out, int device_ptrs) {
    return adj_snow_call(h, cfg, pack, 0, drift_mass, nullptr, swe_out, snowdepth_out, device_ptrs);
}
int pbsm3d_apply_avalanche(pbsm3d_handle* h, const pbsm3d_snobal_config* cfg, const pbsm3d_snowpack* pack,
                           const double* delta_avalanche_snowdepth, const double* delta_avalanche_mass, double* swe_out,
                           double* snowdepth_out, int device_ptrs) {
    return adj_snow_call(h, cfg, pack, 1, delta_avalanche_snowdepth, delta_avalanche_mass, swe_out, snowdepth_out, device_ptrs);
}

// ---- snow_slide (src/modules/snow_slide.cpp) -----------------------------------------------------
void pbsm3d_slide_config_defaults(pbsm3d_slide_config* c) {
    if (!c) return;
    c->avalache_mult = 3178.4;    // snow_slide.cpp:409
    c->avalache_pow = -1.998;     // :410
    c->use_vertical_snow = 1;     // :33 (read by the reference's constructor, not used by its run())
}

int pbsm3d_slide_init(pbsm3d_handle* h, const pbsm3d_slide_config* cfg) {
    if (!h) return fail(PBSM3D_ERR_INVALID, "null handle");
    if (h->geographic) return fail(PBSM3D_ERR_UNSUPPORTED, "snow_slide on a geographic (lat/long) mesh is not implemented");
    CU(cudaSetDevice(h->device));
    pbsm3d_slide_config local;
    if (!cfg) { pbsm3d_slide_config_defaults(&local); cfg = &local; }
    h->slide_cfg = *cfg;
    const int Tp = h->Tp, S = h->S, nG = h->nG;
    SlideArrays& a = h->sl;
    if (!h->slide_ready) {
        double *maxD = nullptr, *cosf = nullptr, *area = nullptr;
        TRY(h->alloc(&maxD, Tp));
        TRY(h->alloc(&cosf, Tp));
        TRY(h->alloc_zero(&area, S));
        TRY(h->alloc_zero(&a.sd, Tp));
        TRY(h->alloc_zero(&a.sdv, S));
        TRY(h->alloc_zero(&a.swe, Tp));
        TRY(h->alloc_zero(&a.dsd, Tp));
        TRY(h->alloc_zero(&a.dmass, Tp));
        TRY(h->alloc(&a.key, Tp));
        TRY(h->alloc_zero(&a.gacc, (size_t)std::max(nG, 1) * 4));
        TRY(h->alloc(&a.stamp, Tp));
        TRY(h->alloc_zero(&a.queued, Tp));
        for (int k = 0; k < 3; ++k) TRY(h->alloc(&a.list[k], Tp));
        TRY(h->alloc_zero(&a.cnt, 8));
        TRY(h->alloc_zero(&h->sl_sum_sd, Tp));
        TRY(h->alloc_zero(&h->sl_sum_mass, Tp));
        TRY(h->alloc_zero(&h->sl_xfer, (size_t)4 * Tp));
        TRY(h->alloc_zero(&h->sl_rev, (size_t)std::max(h->n_send, 1) * 4));
        TRY(h->alloc(&h->sl_in, (size_t)3 * h->T));
        TRY(h->alloc_zero(&h->sl_moved, 2));
        TRY(h->alloc_zero(&h->sl_bar, 1));
        a.T = h->T; a.Tp = Tp; a.S = S;
        a.perm = h->perm; a.nbs = h->nbs; a.cz = h->cz; a.area = area; a.maxD = maxD; a.cosf = cosf;
        // face areas, ghost-extended: a ghost's area is its owner's get_area() (the mesh "area" parameter included)
        CU(cudaMemcpyAsync(area, h->area, (size_t)Tp * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        TRY(halo_exchange(h, area, 1));
        int nb = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, slide_sweep_kernel, kSlideThreads, 0));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, h->device));
        h->sl_grid = std::max(1, std::min(nb, 2)) * prop.multiProcessorCount;
        h->slide_ready = true;
    } else {  // re-init: the reference's init() zeroes the sums
        CU(cudaMemsetAsync(h->sl_sum_sd, 0, (size_t)Tp * sizeof(double), h->stream));
        CU(cudaMemsetAsync(h->sl_sum_mass, 0, (size_t)Tp * sizeof(double), h->stream));
    }
    LAUNCH(h, slide_init_kernel, cdiv(Tp, 256), 256, Tp, h->perm, h->slope, h->wv_canopy, cfg->avalache_mult, cfg->avalache_pow,
           const_cast<double*>(a.maxD), const_cast<double*>(a.cosf));
    TRY(sync_stream(h));
    CU(cudaGetLastError());
    return 0;
}

int pbsm3d_slide_run(pbsm3d_handle* h, const double* snowdepthavg, const double* snowdepthavg_vert, const double* swe,
                     double* delta_avalanche_snowdepth, double* delta_avalanche_mass, double* delta_avalanche_snowdepth_sum,
                     double* delta_avalanche_mass_sum, double* maxDepth, pbsm3d_slide_stats* stats, int device_ptrs) {
    if (!h || !snowdepthavg || !snowdepthavg_vert || !swe) return fail(PBSM3D_ERR_INVALID, "null argument");
    if (!h->slide_ready) return fail(PBSM3D_ERR_INVALID, "call pbsm3d_slide_init first");
    CU(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int T = h->T, Tp = h->Tp, nG = h->nG;
    const size_t bytes = (size_t)T * sizeof(double);
    SlideArrays& a = h->sl;
    const double *d_sd = snowdepthavg, *d_sdv = snowdepthavg_vert, *d_swe = swe;
    if (!device_ptrs) {
        CU(cudaMemcpyAsync(h->sl_in, snowdepthavg, bytes, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->sl_in + T, snowdepthavg_vert, bytes, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->sl_in + 2 * (size_t)T, swe, bytes, cudaMemcpyHostToDevice, s));
        d_sd = h->sl_in; d_sdv = h->sl_in + T; d_swe = h->sl_in + 2 * (size_t)T;
    }
    CU(cudaEventRecord(h->ev[0], s));
    LAUNCH(h, slide_begin_kernel, cdiv(Tp, 256), 256, Tp, h->perm, d_sd, d_sdv, d_swe, a);
    int iterations = 0, rounds = 0, fired = 0, frontier = 0, live = 0;
    int host_cnt[8];
    for (;;) {
        // owners -> ghosts: the vertical depth the weights of a partition-edge face read (snow_slide.cpp:166-169, :398-402)
        TRY(halo_exchange(h, a.sdv, 1));
        CU(cudaMemsetAsync(a.cnt, 0, 8 * sizeof(int), s));
        CU(cudaMemsetAsync(h->sl_bar, 0, sizeof(unsigned), s));
        if (nG > 0) CU(cudaMemsetAsync(a.gacc, 0, (size_t)nG * 4 * sizeof(double), s));
        void* args[] = {&a, &h->sl_bar};
        ++h->n_launch;
        CU(cudaLaunchCooperativeKernel((const void*)slide_sweep_kernel, dim3(h->sl_grid), dim3(kSlideThreads), args, 0, s));
        // ghosts -> owners (ghost_to_neighbors_communicate_variable): what I routed to faces of other ranks
        const double* xfer = nullptr;
        if (h->n_ranks > 1) {
            CU(cudaMemsetAsync(h->sl_xfer, 0, (size_t)4 * Tp * sizeof(double), s));
            NC(ncclGroupStart());
            for (const Partner& p : h->partners) {
                if (p.recv_cnt > 0) NC(ncclSend(a.gacc + (size_t)p.recv_off * 4, (size_t)p.recv_cnt * 4, ncclDouble, p.rank, h->comm, s));
                if (p.send_cnt > 0) NC(ncclRecv(h->sl_rev + (size_t)p.send_off * 4, (size_t)p.send_cnt * 4, ncclDouble, p.rank, h->comm, s));
            }
            NC(ncclGroupEnd());
            for (const Partner& p : h->partners)  // ascending rank: the last partner's value stays
                if (p.send_cnt > 0)
                    LAUNCH(h, slide_unpack_kernel, cdiv(p.send_cnt, 256), 256, p.send_cnt, h->send_slot + p.send_off,
                           h->sl_rev + (size_t)p.send_off * 4, h->sl_xfer, Tp);
            xfer = h->sl_xfer;
        }
        CU(cudaMemsetAsync(h->sl_moved, 0, sizeof(int), s));
        LAUNCH(h, slide_absorb_kernel, cdiv(Tp, 256), 256, a, xfer, h->sl_sum_sd, h->sl_sum_mass, h->sl_moved);
        int moved = 0;
        if (h->n_ranks > 1) {  // done = min over ranks of "nobody received transport" (snow_slide.cpp:381-388)
            NC(ncclAllReduce(h->sl_moved, h->sl_moved + 1, 1, ncclInt, ncclMax, h->comm, s));
            CU(cudaMemcpyAsync(&moved, h->sl_moved + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
        }
        CU(cudaMemcpyAsync(host_cnt, a.cnt, sizeof(host_cnt), cudaMemcpyDeviceToHost, s));
        TRY(sync_stream(h));
        CU(cudaGetLastError());
        rounds += host_cnt[4];
        fired += host_cnt[5];
        frontier += host_cnt[7];
        live += host_cnt[3];
        if (host_cnt[6]) return fail(PBSM3D_ERR_INVALID, "Snowslide did not conserve mass");  // snow_slide.cpp:322-328
        ++iterations;
        bool done = moved == 0;
        if (!done && iterations > 25) done = true;  // snow_slide.cpp:391-396
        if (done) break;
    }
    CU(cudaEventRecord(h->ev[1], s));
    // publish (snow_slide.cpp:352-357, :441): slot order -> CHM order
    const double* src[5] = {a.dsd, a.dmass, h->sl_sum_sd, h->sl_sum_mass, a.maxD};
    double* dst[5] = {delta_avalanche_snowdepth, delta_avalanche_mass, delta_avalanche_snowdepth_sum, delta_avalanche_mass_sum, maxDepth};
    for (int k = 0; k < 5; ++k) {
        if (!dst[k]) continue;
        if (device_ptrs) {
            LAUNCH(h, to_chm_kernel, cdiv(T, 256), 256, 1, T, (size_t)0, h->iperm, src[k], dst[k]);
        } else {
            TRY(fetch_chm(h, dst[k], src[k], 1, 0));
        }
    }
    TRY(sync_stream(h));
    CU(cudaGetLastError());
    if (stats) {
        stats->iterations = iterations;
        stats->wavefront_rounds = rounds;
        stats->faces_fired = fired;
        stats->frontier_rounds = frontier;
        stats->live_faces = live;
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        stats->ms_device = ms;
    }
    return 0;
}

int pbsm3d_slide_get_constants(pbsm3d_handle* h, double* maxDepth, double* cos_slope) {
    if (!h || !h->slide_ready) return fail(PBSM3D_ERR_INVALID, "call pbsm3d_slide_init first");
    CU(cudaSetDevice(h->device));
    TRY(fetch_chm(h, maxDepth, h->sl.maxD, 1, 0));
    TRY(fetch_chm(h, cos_slope, h->sl.cosf, 1, 0));
    return 0;
}
int pbsm3d_slide_get_state(pbsm3d_handle* h, double* delta_avalanche_snowdepth, double* delta_avalanche_mass,
                           double* delta_avalanche_snowdepth_sum, double* delta_avalanche_mass_sum) {
    if (!h || !h->slide_ready) return fail(PBSM3D_ERR_INVALID, "call pbsm3d_slide_init first");
    CU(cudaSetDevice(h->device));
    TRY(fetch_chm(h, delta_avalanche_snowdepth, h->sl.dsd, 1, 0));
    TRY(fetch_chm(h, delta_avalanche_mass, h->sl.dmass, 1, 0));
    TRY(fetch_chm(h, delta_avalanche_snowdepth_sum, h->sl_sum_sd, 1, 0));
    TRY(fetch_chm(h, delta_avalanche_mass_sum, h->sl_sum_mass, 1, 0));
    return 0;
}
int pbsm3d_slide_set_state(pbsm3d_handle* h, const double* delta_avalanche_snowdepth, const double* delta_avalanche_mass,
                           const double* delta_avalanche_snowdepth_sum, const double* delta_avalanche_mass_sum) {
    if (!h || !h->slide_ready) return fail(PBSM3D_ERR_INVALID, "call pbsm3d_slide_init first");
    CU(cudaSetDevice(h->device));
    TRY(store_chm(h, h->sl.dsd, delta_avalanche_snowdepth));
    TRY(store_chm(h, h->sl.dmass, delta_avalanche_mass));
    TRY(store_chm(h, h->sl_sum_sd, delta_avalanche_snowdepth_sum));
    TRY(store_chm(h, h->sl_sum_mass, delta_avalanche_mass_sum));
    return 0;
}

int pbsm3d_time_kernel(pbsm3d_handle* h, int kernel, int reps, float* ms) {
    if (!h || !ms || reps < 1) return fail(PBSM3D_ERR_INVALID, "bad argument");
    if (!h->have_system) return fail(PBSM3D_ERR_INVALID, "run a step first");
    CU(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const double tol2 = h->cfg.tolerance * h->cfg.tolerance;
    // the sweep and residual kernels idle once the solve has converged: reopen it for the measurement
    CU(cudaMemsetAsync(&h->sc->susp_done, 0, sizeof(int), s));
    for (int pass = 0; pass < 2; ++pass) {  // one untimed launch, then `reps` timed ones
        int n = pass == 0 ? 1 : reps;
        if (pass == 1) CU(cudaEventRecord(h->ev[0], s));
        for (int k = 0; k < n; ++k) {
            switch (kernel) {
                case 0:
                    if (k == 0) TRY(l2_window(h, h->x, h->NS * sizeof(double)));
                    TRY(enqueue_sweeps(h, 1));
                    break;
                case 1: launch_residual(h, 0, tol2, 0); break;
                case 2: launch_assembly(h, h->last_forcing, h->last_dt); break;
                case 3:
                    LAUNCH(h, cg_spmv_kernel, red_grid(h->Tp), kRedThreads, h->dm, h->ddiag, h->doff, h->cg_p, h->cg_Ap, h->partial,
                           kRedBlocks, nullptr, h->red, tol2, 0);
                    break;
                case 4:
                    if (!h->cheb_ready) return fail(PBSM3D_ERR_INVALID, "Chebyshev is not set up on this handle");
                    // a_k = c_k = 0: the same traffic as a real iteration, and the converged iterate is only copied
                    LAUNCH(h, cheb_iter_kernel<0>, red_grid(h->Tp), kRedThreads, h->dm, h->offS, h->drhsS, h->ddiag,
                           h->h_sc->dep_buf ? h->qB : h->qA, h->h_sc->dep_buf ? h->qA : h->qB, 0.0, 0.0, k, h->partial,
                           kRedBlocks, nullptr, h->red, tol2, 0);
                    break;
                default: return fail(PBSM3D_ERR_INVALID, "unknown kernel id");
            }
        }
        if (pass == 1) CU(cudaEventRecord(h->ev[1], s));
    }
    TRY(l2_window(h, nullptr, 0));
    LAUNCH(h, flags_kernel, 1, 1, FLAGS_FORCE_SUSP_OK, h->sc, h->red, 0, tol2);
    CU(cudaStreamSynchronize(s));
    CU(cudaGetLastError());
    float t = 0;
    CU(cudaEventElapsedTime(&t, h->ev[0], h->ev[1]));
    *ms = t / reps;
    return 0;
}

}  // extern "C"

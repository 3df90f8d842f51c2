/*
 * pbsm3d.h — C-ABI of the B200-native PBSM3D hot path (libpbsm3d_b200.so).
 *
 * This is the boundary a CHM adaptor module binds.  It replaces, for one
 * MPI rank / one GPU, everything `PBSM3D::init` and `PBSM3D::run` do between reading the per-face
 * variable store and writing it back:
 *
 *   reference interface                                         replaced by
 *   ----------------------------------------------------------  -------------------------
 *   PBSM3D::PBSM3D(config_file)        PBSM3D.cpp:103-219       pbsm3d_config_defaults + pbsm3d_config
 *   PBSM3D::init(mesh&)                PBSM3D.cpp:221-398       pbsm3d_create
 *   NearestNeighborProblem ctor        LinearAlgebra.cpp:31-197 pbsm3d_create (static sparsity = mesh adjacency)
 *   PBSM3D::run(mesh&)                 PBSM3D.cpp:400-1748      pbsm3d_step / pbsm3d_step_device
 *   NearestNeighborProblem::Solve      LinearAlgebra.cpp:228-252  (inside pbsm3d_step)
 *   ghost_neighbors_communicate_variable  triangulation.cpp:1976-2079  (inside pbsm3d_step: peer memory over NVLink, NCCL fallback)
 *   getSolutionView                    LinearAlgebra.cpp:272-275  pbsm3d_get_solution
 *   writeSystemMatrixMarket (debug)    LinearAlgebra.cpp:277-287  pbsm3d_get_suspension_system / _deposition_system
 *   PBSM3D::checkpoint/load_checkpoint PBSM3D.cpp:1753-1773     pbsm3d_get_state / pbsm3d_set_state
 *   ~PBSM3D                                                     pbsm3d_destroy
 *   scale_wind_vert::run(mesh&)        scale_wind_vert.cpp:167-229  pbsm3d_scale_wind_vert   } the providers of two PBSM3D inputs,
 *   fetchr::run(face), every face      fetchr.cpp:54-119            pbsm3d_fetchr            } optionally fused into the step
 *                                                                   (pbsm3d_set_providers)
 *   snobal::run, drift_mass hook       snobal.cpp:363-387 -> sno::_adj_snow (sno.cpp:2527-2575)   pbsm3d_apply_drift     } the consumers of
 *   snow_slide::init / run / checkpoint  snow_slide.cpp:406-446 / :94-404 / :59-93            pbsm3d_slide_init / _run / _get_state / _set_state
 *   snobal::run, avalanche hook        snobal.cpp:389-408 -> sno::_adj_snow                       pbsm3d_apply_avalanche } PBSM3D's / snow_slide's output
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero
 * code on failure with a message available from pbsm3d_last_error() (the adaptor turns it into
 * CHM's module_error).  The caller owns all host buffers; the library owns all device memory.
 * A handle is bound to one CUDA device and is not re-entrant.  All floating point is fp64 (pbsm3d_config.fp32_sweep_streams
 * concerns the storage of coefficient copies the early sweeps read, not the arithmetic or the stopping rule).
 * There is no CPU fallback: without a CUDA device pbsm3d_create fails.
 */
#ifndef PBSM3D_B200_H
#define PBSM3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBSM3D_ABI_VERSION 8

enum {
    PBSM3D_OK = 0,
    PBSM3D_ERR_INVALID = 1,     /* bad argument / inconsistent mesh */
    PBSM3D_ERR_UNSUPPORTED = 2, /* optional reference path that is not implemented (never silently ignored) */
    PBSM3D_ERR_CUDA = 3,
    PBSM3D_ERR_NCCL = 4,
    PBSM3D_ERR_NOCONVERGE = 5   /* "Belos solver failed to converge", LinearAlgebra.cpp:236-243 */
};

enum {
    PBSM3D_DEP_AUTO = 0,       /* multicolour SOR (Chebyshev across ranks without peer memory); falls back to CG if it does not converge */
    PBSM3D_DEP_CG = 1,         /* Jacobi-preconditioned conjugate gradients */
    PBSM3D_DEP_CHEBYSHEV = 2,  /* Jacobi-preconditioned Chebyshev iteration (no global reductions) */
    PBSM3D_DEP_SOR = 3         /* multicolour SOR, Young's relaxation factor from the setup-time spectrum estimate; across ranks the
                                  sweep order is (colour, rank) and ghosts travel inside the colour passes (peer memory only) */
};

/* How ghost-face halos and the solvers' global reductions travel between the ranks of one NVSwitch box
 * (reference: MPI messages, triangulation.cpp:1976-2079).  PEER = direct stores into the partner's exported
 * staging buffer over NVLink (cudaIpc), chosen whenever every rank can map every other rank's arena; NCCL
 * send/recv + all-reduce otherwise, or when PBSM3D_HALO=nccl is set in the environment.  With PEER the halos of the
 * two dominant iterations (line sweeps, Chebyshev) are carried by the solver kernels themselves (stats.halo_fused);
 * PBSM3D_HALO=staged keeps them as separate push / wait-unpack launches (the path every other vector uses). */
enum {
    PBSM3D_HALO_NONE = 0,
    PBSM3D_HALO_NCCL = 1,
    PBSM3D_HALO_PEER = 2
};

enum {
    PBSM3D_SOLVER_AUTO = 0,     /* line relaxation, falling back to BiCGStab if it stagnates */
    PBSM3D_SOLVER_LINE = 1,     /* multicolour line Gauss-Seidel sweeps (exact vertical tridiagonal solve per face column) */
    PBSM3D_SOLVER_BICGSTAB = 2  /* right-preconditioned BiCGStab, column-tridiagonal preconditioner */
};

/* PBSM3D config keys, same names and defaults as the reference reads with cfg.get()
 * (PBSM3D.cpp:123-145 and :223-258).  Booleans are int 0/1. */
typedef struct pbsm3d_config {
    int nLayer;                   /* 10 */
    int do_fixed_settling;        /* false */
    double settling_velocity;     /* 0.5 */
    int do_sublimation;           /* true */
    int do_lateral_diff;          /* true */
    double smooth_coeff;          /* 820 */
    double min_sd_trans;          /* 0.1 */
    double cutoff;                /* 0.3 */
    double snow_diffusion_const;  /* 0.3 */
    int rouault_diffusion_coef;   /* false */
    int enable_veg;               /* true */
    int iterative_subl;           /* false — unsupported if true */
    int use_exp_fetch;            /* false */
    int use_tanh_fetch;           /* true */
    int use_PomLi_probability;    /* false; needs forcing.p_snow_hours */
    int z0_ustar_coupling;        /* false — unsupported if true */
    int use_subgrid_topo;         /* false — unsupported if true */
    int use_subgrid_topo_V2;      /* false — unsupported if true */
    int use_R94_lambda;           /* true */
    int debug_output;             /* false — unsupported if true */
    /* Solver controls.  The reference hard-codes these (LinearAlgebra.cpp:164-168). */
    double tolerance;             /* 1e-8: ||b-Ax||2/||b||2, x0 = 0 */
    int max_iterations;           /* 1000 */
    int solver;                   /* PBSM3D_SOLVER_AUTO (suspension system) */
    int deposition_solver;        /* PBSM3D_DEP_AUTO */
    int fp32_sweep_streams;       /* 1: line sweeps far from convergence stream fp32-rounded copies of their coefficients (fp64 x,
                                     fp64 arithmetic); the last sweeps and every residual check use the fp64 coefficients, so the
                                     stopping rule is unchanged.  0: fp64 streams throughout. */
} pbsm3d_config;

/* One rank's share of the mesh, flattened from CHM's triangulation in CHM's own face order.
 * Faces [0,n_local) are the owned faces `domain->face(i)` (ascending cell_global_id, one contiguous
 * global range per rank: triangulation.cpp:1482-1531); faces [n_local, n_local+n_ghost) are the
 * NEIGH ghosts sorted by cell_global_id (triangulation.cpp:1721-1770). */
typedef struct pbsm3d_mesh {
    int64_t n_global;             /* domain->size_global_faces() */
    int32_t n_local;              /* domain->size_faces() */
    int32_t n_ghost;
    const int64_t* global_id;     /* [n_local+n_ghost] cell_global_id */
    const int32_t* ghost_owner;   /* [n_ghost] owning rank (face->owner); may be NULL when n_ghost==0 */
    const int32_t* neigh;         /* [n_local][3] face->neighbor(j) as a local index; -1 = nullptr */
    const double* vertices;       /* [n_local+n_ghost][3 vertices][x,y,z] */
    const double* area;           /* [n_local] face->get_area() when the mesh has an "area" parameter, else NULL */
    const double* canopy_height;  /* [n_local] veg_attribute("CanopyHeight"), NULL = no vegetation info
                                     (then enable_veg turns off, PBSM3D.cpp:317-324) */
    const double* lai;            /* [n_local] veg_attribute("LAI") (use_R94_lambda), may be NULL otherwise */
    const double* stalk_number;   /* [n_local] or NULL -> 1   (PBSM3D.cpp:303-312) */
    const double* stalk_diameter; /* [n_local] or NULL -> 0.8 */
    const uint8_t* is_water;      /* [n_local] module_base::is_water(face), NULL = none */
    int32_t is_geographic;        /* domain->is_geographic(): vertex x/y are longitude/latitude in degrees.  core.cpp:809-821 then
                                     installs math::gis::distance = distance_latlong (haversine, coordinates.cpp:68-92), which the
                                     deposition matrix's dx uses (PBSM3D.cpp:1546).  pbsm3d_fetchr refuses such a mesh
                                     (PBSM3D_ERR_UNSUPPORTED: point_from_bearing_latlong returns (lat, lon) swapped). */
} pbsm3d_mesh;

/* Multi-GPU: one process per GPU.  NULL or n_ranks==1 means a single-rank run. */
typedef struct pbsm3d_comm {
    int32_t rank;
    int32_t n_ranks;
    const void* nccl_unique_id;   /* 128 bytes from pbsm3d_nccl_unique_id() on rank 0, broadcast by the host */
} pbsm3d_comm;

/* Per-face inputs of one timestep, the variables PBSM3D::run reads from the face store
 * (PBSM3D.cpp:436-449,468,670,880,927).  Each is [n_local]; -9999 / NaN mean "missing" as in CHM. */
typedef struct pbsm3d_forcing {
    const double* U_R;
    const double* U_2m_above_srf; /* may be NULL after pbsm3d_set_providers (derived on the device from U_R, snowdepthavg) */
    const double* snowdepthavg;
    const double* swe;
    const double* t;
    const double* rh;
    const double* vw_dir;
    const double* fetch;          /* may be NULL when neither fetch option is on (1000 m is used), or after pbsm3d_set_providers */
    const double* p_snow_hours;   /* hours since the last snowfall: read iff use_PomLi_probability (PBSM3D.cpp:853), else may be NULL */
} pbsm3d_forcing;

/* Per-face outputs PBSM3D::run writes back (provides(), PBSM3D.cpp:194-202).  Each [n_local];
 * any pointer may be NULL to skip that copy. */
typedef struct pbsm3d_outputs {
    double* Qsalt;
    double* Qsusp;
    double* Qsubl;
    double* Qsubl_mass;
    double* sum_subl;
    double* drift_mass;           /* unchanged from the previous step when no deposition solve ran */
    double* sum_drift;
    double* pbsm_more_than_avail; /* sticky 0/1 flag */
    double* blowingsnow_probability; /* use_PomLi_probability: written on the faces that saltate this step (PBSM3D.cpp:861), all
                                        others keep their previous value (-9999 until first written) */
} pbsm3d_outputs;

typedef struct pbsm3d_stats {
    int32_t suspension_present;
    int32_t deposition_present;
    int32_t suspension_iterations;   /* full sweeps (line) or matvec pairs (BiCGStab) until the stopping rule held */
    int32_t deposition_iterations;
    int32_t suspension_solver_used;  /* PBSM3D_SOLVER_LINE / _BICGSTAB */
    int32_t kernel_launches;         /* kernels of this library launched by the step */
    double suspension_residual;      /* achieved ||b-Ax||2/||b||2 */
    double deposition_residual;
    double suspension_rhs_max;
    double deposition_rhs_max;
    float ms_assembly;               /* CUDA-event times of the phases of this step */
    float ms_suspension_solve;
    float ms_flux_and_halo;
    float ms_deposition;
    float ms_total;
    float ms_line_sweeps;            /* CUDA-event time of the first `sweeps_timed` line sweeps of this step */
    int32_t sweeps_timed;            /* full sweeps (all colours) inside ms_line_sweeps */
    int32_t sweeps_timed_fp32;       /* of those, the leading ones that streamed fp32 coefficient copies ... */
    float ms_line_sweeps_fp32;       /* ... and their CUDA-event time (ms_line_sweeps - this = the fp64-stream sweeps) */
    int32_t n_colours;               /* colour classes of the internal face order */
    int32_t deposition_solver_used;  /* PBSM3D_DEP_CG / _CHEBYSHEV / _SOR */
    int32_t host_syncs;              /* stream synchronisations the step needed (1 when every prediction held) */
    int32_t halo_exchanges;          /* ghost-face halo exchanges the step enqueued (0 on a single rank) */
    int32_t halo_transport;          /* PBSM3D_HALO_NONE / _NCCL / _PEER */
    int32_t halo_fused;              /* of halo_exchanges: carried by the solver kernels themselves (direct stores into the
                                        partner's ghost buffer from the producing kernel; no pack/unpack launch) */
    int32_t residual_checks;         /* evaluations of ||b - A x|| of the suspension system in this step */
    int32_t sweeps_fp32_x;           /* of sweeps_timed_fp32: the leading ones that also kept the iterate x in fp32 STORAGE */
    int32_t persistent_kernels;      /* 1: each solve ran as one cooperative launch (single rank); then ms_line_sweeps is the
                                        duration of that launch = sweeps_timed sweeps + residual_checks checks */
    int32_t active_set;              /* 1: the persistent line solver skipped the columns outside its active set (columns whose
                                        right-hand side and whose neighbours' iterates are still exactly zero: their update is a
                                        no-op, so every iterate is bit-identical to the full sweep's).  Used on the steps where at
                                        most 40 % of the faces have a non-zero right-hand side; PBSM3D_ACTIVE_SET=0 / 1: never / always */
    int32_t faces_with_rhs;          /* local faces with a non-zero right-hand side in this step (the saltating faces) */
    int64_t column_updates_fp32_x;   /* persistent line solver, this rank: face-column updates executed in the sweeps of each */
    int64_t column_updates_fp32;     /*   storage phase (fp32 x + fp32 coefficients / fp32 coefficients / all fp64): the sum is */
    int64_t column_updates_fp64;     /*   sweeps_timed x local faces without the active set, less with it */
    int64_t columns_checked;         /* face columns evaluated by the residual checks (residual_checks x local faces without it) */
} pbsm3d_stats;

typedef struct pbsm3d_handle pbsm3d_handle;

int pbsm3d_abi_version(void);
const char* pbsm3d_last_error(void);
void pbsm3d_config_defaults(pbsm3d_config* cfg);

/* rank 0 calls this, the host broadcasts the 128 bytes, every rank passes them in pbsm3d_comm. */
int pbsm3d_nccl_unique_id(void* out_128_bytes);

/* Page-locked host memory for the forcing/output staging arrays: with pinned buffers pbsm3d_step overlaps the PCIe
 * transfers with the assembly and the solves (pageable buffers work too, without the overlap).  NULL on failure. */
void* pbsm3d_host_alloc(size_t bytes);
void pbsm3d_host_free(void* p);

int pbsm3d_create(const pbsm3d_config* cfg, const pbsm3d_mesh* mesh, int device, const pbsm3d_comm* comm,
                  pbsm3d_handle** out);
void pbsm3d_destroy(pbsm3d_handle* h);

/* One PBSM3D::run.  Host buffers; H2D of the forcing and D2H of the outputs are part of the call. */
int pbsm3d_step(pbsm3d_handle* h, double dt, const pbsm3d_forcing* forcing, const pbsm3d_outputs* out,
                pbsm3d_stats* stats);
/* Same with DEVICE pointers (forcing already resident, outputs left on the device). */
int pbsm3d_step_device(pbsm3d_handle* h, double dt, const pbsm3d_forcing* forcing_dev, const pbsm3d_outputs* out_dev,
                       pbsm3d_stats* stats);

/* Checkpoint state (PBSM3D.cpp:1753-1773 persists sum_drift; sum_subl, drift_mass and the sticky flag
 * are the other values that survive between steps).  Each [n_local]; NULL = skip. */
int pbsm3d_get_state(pbsm3d_handle* h, double* sum_drift, double* sum_subl, double* drift_mass, double* more_than_avail);
int pbsm3d_set_state(pbsm3d_handle* h, const double* sum_drift, const double* sum_subl, const double* drift_mass,
                     const double* more_than_avail);

/* ---- inspection (parity tests; mirrors the reference's commented-out MatrixMarket dump hooks) ---- */

/* Device-computed face geometry, each [3][n_local] or [n_local] (NULL = skip). */
int pbsm3d_get_geometry(pbsm3d_handle* h, double* nx, double* ny, double* edge_length, double* area, double* dx,
                        double* cx, double* cy, double* cz);
/* Internal layout (tests / diagnostics): number of colour classes, number of padded slots, and for every local
 * face its slot in the colour-major device order and its colour (each [n_local]; any pointer may be NULL). */
int pbsm3d_get_layout(pbsm3d_handle* h, int32_t* n_colours, int32_t* n_slots, int32_t* slot_of_face, int32_t* colour_of_face);
/* Suspended concentration of the last step, layout x[z*n_local + local_id] (LinearAlgebra.cpp:81). */
int pbsm3d_get_solution(pbsm3d_handle* h, double* x);
/* Last assembled suspension system in extruded-ELL form, each [nLayer][n_local] except lat [3][nLayer][n_local]
 * and rhs0/c_salt [n_local] (the RHS is non-zero only in layer 0); saltation is [n_local] 0/1. */
int pbsm3d_get_suspension_system(pbsm3d_handle* h, double* diag, double* lat, double* below, double* above,
                                 double* rhs0, double* u_z, double* csubl, double* c_salt, uint8_t* saltation);
/* Last deposition system: diag [n_local], off [3][n_local], rhs [n_local], solution q [n_local]. */
int pbsm3d_get_deposition_system(pbsm3d_handle* h, double* diag, double* off, double* rhs, double* q);

/* ---- providers of two PBSM3D inputs, on the device ----
 * scale_wind_vert (src/modules/scale_wind_vert.cpp:27-229): U_R [+ snowdepthavg] -> U_2m_above_srf; in domain mode every
 * face then takes the thin plate spline (src/interpolation/TPSpline.cpp:40-173) of its edge neighbours' values.
 * fetchr (src/modules/fetchr.cpp:27-119): vw_dir -> fetch, `steps` nearest-face-centre queries up-wind per face; the
 * search covers this rank's owned faces.  Vegetation comes from pbsm3d_mesh.canopy_height / .lai (NULL = the mesh has
 * no vegetation parameters, face->has_vegetation() false). */
typedef struct pbsm3d_wind_config {
    int32_t ignore_canopy;       /* scale_wind_vert "ignore_canopy", default false (scale_wind_vert.cpp:161) */
    int32_t point_mode;          /* 1: point_scale only (CHM's point mode / run(face)); 0: domain mode with the neighbour spline */
    int32_t fetch_steps;         /* fetchr "steps", default 10 (fetchr.cpp:34) */
    int32_t fetch_incl_veg;      /* "incl_veg", default true (:43) */
    double fetch_max_distance;   /* "max_distance", default 1000 m (:36) */
    double fetch_I;              /* "I", default 0.06 m/m (:41) */
} pbsm3d_wind_config;
void pbsm3d_wind_config_defaults(pbsm3d_wind_config* cfg);
/* Each array [n_local] in CHM face order; device_ptrs: 0 = host buffers, 1 = device buffers.  snowdepthavg may be NULL
 * (no module provides the optional input).  cfg NULL = defaults. */
int pbsm3d_scale_wind_vert(pbsm3d_handle* h, const pbsm3d_wind_config* cfg, const double* U_R, const double* snowdepthavg,
                           double* U_2m_above_srf, int device_ptrs);
int pbsm3d_fetchr(pbsm3d_handle* h, const pbsm3d_wind_config* cfg, const double* vw_dir, double* fetch, int device_ptrs);
/* Fuse the providers into the step: after this call pbsm3d_step / pbsm3d_step_device accept forcing->U_2m_above_srf == NULL
 * and (with exp/tanh fetch on) forcing->fetch == NULL and derive them on the device before the assembly, so the two arrays
 * never cross PCIe.  cfg NULL switches the fusion off again. */
int pbsm3d_set_providers(pbsm3d_handle* h, const pbsm3d_wind_config* cfg);

/* ---- the consumer of PBSM3D's output: snobal's snowpack mass adjustment on the device ----
 * snobal::run (src/modules/snobal.cpp:363-387) hands each face's drift_mass to sno::_adj_snow (third_party/snobal/sno.cpp:2527-2575,
 * with _adj_layers :2617-2696, _calc_layers :2366-2405, _layer_mass :1564-1580, _cold_content :2321-2329): erosion removes depth at
 * the pack's own density, deposition adds depth at `drift_density`; layers are re-partitioned, a pack that falls below the mass
 * threshold becomes liquid water.  The per-face snowpack state is the caller's (snobal's `sno` members of the same names), SoA,
 * each [n_local] in CHM face order; every field is read and written in place.  Bit-identical to the compiled sno.cpp. */
typedef struct pbsm3d_snowpack {
    double *z_s, *m_s, *rho;   /* total depth (m), specific mass (kg/m^2), density */
    int32_t* layer_count;      /* 0, 1 or 2 */
    double *z_s_0, *z_s_l, *m_s_0, *m_s_l;      /* surface / lower layer depth and mass */
    double *cc_s, *cc_s_0, *cc_s_l;             /* cold contents (J/m^2) */
    double *T_s, *T_s_0, *T_s_l;                /* temperatures (K) */
    double *h2o_total, *h2o_vol, *h2o, *h2o_max, *h2o_sat;
} pbsm3d_snowpack;
typedef struct pbsm3d_snobal_config {
    double drift_density;      /* "drift_density", 300 kg/m^3 (snobal.cpp:83) */
    double threshold;          /* tstep_info[SMALL_TSTEP].threshold, 0.2 kg/m^2 (snobal.cpp:190): minimum mass of a layer */
    double max_active_layer;   /* "max_active_layer" = sno::max_z_s_0, 0.1 m (snobal.cpp:101) */
} pbsm3d_snobal_config;
void pbsm3d_snobal_config_defaults(pbsm3d_snobal_config* cfg);
/* drift_mass [n_local]: NULL = the handle's own device-resident drift_mass of the last pbsm3d_step (it then never crosses PCIe);
 * -9999 / NaN count as 0 (module_base::is_nan).  swe_out / snowdepth_out [n_local] (each may be NULL) receive m_s and z_s, the two
 * face variables snobal hands back to PBSM3D's next step (snobal.cpp:468,491).  device_ptrs: 0 = every pointer (inside `pack`
 * too) is a host buffer, 1 = device buffers.  cfg NULL = defaults. */
int pbsm3d_apply_drift(pbsm3d_handle* h, const pbsm3d_snobal_config* cfg, const pbsm3d_snowpack* pack, const double* drift_mass,
                       double* swe_out, double* snowdepth_out, int device_ptrs);
/* The same adjustment for snow_slide's output (snobal.cpp:389-408): delta_avalanche_snowdepth is a VOLUME (m^3) and
 * delta_avalanche_mass a water-equivalent volume (m^3) per face; _adj_snow(volume / area, swe volume / area * 1000) with the
 * handle's face areas. */
int pbsm3d_apply_avalanche(pbsm3d_handle* h, const pbsm3d_snobal_config* cfg, const pbsm3d_snowpack* pack,
                           const double* delta_avalanche_snowdepth, const double* delta_avalanche_mass, double* swe_out,
                           double* snowdepth_out, int device_ptrs);

/* ---- snow_slide on the device (src/modules/snow_slide.cpp) ----
 * Gravitational redistribution: a face whose slope-normal snow depth exceeds maxDepth = max(avalache_mult * slopeDeg^avalache_pow,
 * CanopyHeight) * max(0.001, cos(slope)) sheds the excess to its lower neighbours, highest snow surface first.  The reference
 * sweeps the faces SEQUENTIALLY in that order (snow_slide.cpp:171-330); the device runs the same updates as a dependency wavefront
 * and reproduces the sequential result (pbsm3d_slide.cuh).  Across ranks: forward halo of the vertical depth, the reverse
 * ghost -> owner exchange of what was routed across a partition edge (triangulation.cpp:2081-2186), and outer iterations while any
 * rank received transport (<= 26), as in the reference.  Faces with EQUAL sort keys (undefined order in the reference's
 * tbb::parallel_sort) take their turn in CHM face order.  Vegetation height comes from pbsm3d_mesh.canopy_height. */
typedef struct pbsm3d_slide_config {
    double avalache_mult;        /* "avalache_mult", 3178.4 (snow_slide.cpp:409; the reference's spelling) */
    double avalache_pow;         /* "avalache_pow", -1.998 (:410) */
    int32_t use_vertical_snow;   /* "use_vertical_snow", true (:33); read by the reference's constructor and never used */
} pbsm3d_slide_config;
typedef struct pbsm3d_slide_stats {
    int32_t iterations;          /* outer iterations (1 on a single rank) */
    int32_t wavefront_rounds;    /* dependency rounds of the sweeps */
    int32_t faces_fired;         /* faces that shed snow */
    int32_t frontier_rounds;     /* rounds of the live-set expansion */
    int32_t live_faces;          /* faces that took part (start candidates and, transitively, their later-ordered neighbours) */
    float ms_device;             /* CUDA-event time of the run without the copies of the outputs */
} pbsm3d_slide_stats;
void pbsm3d_slide_config_defaults(pbsm3d_slide_config* cfg);
/* snow_slide::init: maxDepth per face; zeroes delta_avalanche_*_sum.  cfg NULL = defaults. */
int pbsm3d_slide_init(pbsm3d_handle* h, const pbsm3d_slide_config* cfg);
/* snow_slide::run.  Inputs [n_local] in CHM face order: snowdepthavg (m, slope-normal), snowdepthavg_vert (m, vertical), swe (mm).
 * Outputs [n_local], any may be NULL: delta_avalanche_snowdepth (m^3), delta_avalanche_mass (m^3 of water), their running sums,
 * maxDepth.  stats may be NULL.  Fails with PBSM3D_ERR_INVALID ("Snowslide did not conserve mass") like the reference's throw. */
int pbsm3d_slide_run(pbsm3d_handle* h, const double* snowdepthavg, const double* snowdepthavg_vert, const double* swe,
                     double* delta_avalanche_snowdepth, double* delta_avalanche_mass, double* delta_avalanche_snowdepth_sum,
                     double* delta_avalanche_mass_sum, double* maxDepth, pbsm3d_slide_stats* stats, int device_ptrs);
/* Inspection (parity tests): the two per-face constants of init, maxDepth and max(0.001, cos(slope)), each [n_local], NULL = skip.
 * Under MPI the reference's outer iterations end in a cascade of ever smaller transfers across partition edges; whether a transfer
 * of one ulp still counts decides a FINITE change down-slope (a receiver's vertical depth is recomputed with the donor's slope,
 * snow_slide.cpp:297), so a partitioned run is sensitive to the last bit of these constants (libm's pow / cos).  Tests feed the
 * device's own constants to the oracle to compare the sweeps themselves. */
int pbsm3d_slide_get_constants(pbsm3d_handle* h, double* maxDepth, double* cos_slope);
/* Checkpoint (snow_slide.cpp:59-93): the four arrays the reference persists.  Each [n_local]; NULL = skip. */
int pbsm3d_slide_get_state(pbsm3d_handle* h, double* delta_avalanche_snowdepth, double* delta_avalanche_mass,
                           double* delta_avalanche_snowdepth_sum, double* delta_avalanche_mass_sum);
int pbsm3d_slide_set_state(pbsm3d_handle* h, const double* delta_avalanche_snowdepth, const double* delta_avalanche_mass,
                           const double* delta_avalanche_snowdepth_sum, const double* delta_avalanche_mass_sum);

/* Stand-alone kernels for measurement (bench.py roofline line, ncu): run `reps` launches of the named kernel on
 * the last assembled system and return the mean CUDA-event time per launch in milliseconds.
 * kernel: 0 = one full line sweep (all colour passes), 1 = residual (SpMV), 2 = assembly, 3 = deposition CG SpMV,
 * 4 = deposition Chebyshev iteration. */
int pbsm3d_time_kernel(pbsm3d_handle* h, int kernel, int reps, float* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* PBSM3D_B200_H */

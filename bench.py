#!/usr/bin/env python
"""bench.py — PBSM3D element-layer solves/s on B200 (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm on the box's host cores (CPU)

A "step" is one complete PBSM3D::run (saltation + suspension assembly, suspension solve, flux integration, halo,
deposition solve, drift update) on a synthetic mesh:
  N = 1  : config c2 of BASELINE.md — 708×708 squares of 30 m split into 1 002 528 triangles, nLayer 10, fp64,
           PBSM3D options = the functional-test block (functional_tests/mesh_versioning/json_mesh.json:82-99)
  N > 1  : weak scaling — the same generator at ≈1.0 M triangles per GPU, partitioned by CHM's contiguous
           global-id rule; ghost-face halos and global reductions go through peer memory over NVLink (cudaIpc
           arenas; PBSM3D_HALO=nccl forces NCCL send/recv + all-reduce instead).
`value` = (triangles × layers over all ranks) / (CUDA-event time of the step, max over ranks), forcing resident in HBM.
`e2e`   = the same through pbsm3d_step() with pinned HOST buffers (H2D of 8 forcing arrays + D2H of 8 outputs inside
          the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pbsm3d_element_layer_solves_per_s"
UNIT = "element-layer solves/s"
NLAYER = 10
FUNCTEST = dict(nLayer=NLAYER, smooth_coeff=6500, do_fixed_settling=1, settling_velocity=0.5, use_R94_lambda=0)


def mesh_side(n_gpus: int) -> int:
    return int(round(708 * np.sqrt(n_gpus)))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (profiling recipe's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def sweep_bytes_per_row(L: int, fp32_streams: bool = False) -> float:
    """Algorithmic HBM bytes of one full line Gauss-Seidel sweep per unknown row:
    3 row-scaled lateral coefficients + scaled sub-diagonal + cp (5×8, or 5×4 for the fp32-rounded copies the sweeps far from
    convergence stream) + x read once + x written once (2×8) = 56 (36) B/row, plus per face 3 int32 neighbour slots + the
    scaled rhs (20 B) amortised over L layers."""
    return (36.0 if fp32_streams else 56.0) + 20.0 / L


def persistent_solve_bytes(st, T, L):
    """Algorithmic HBM bytes of ONE launch of gs_persistent_kernel (the whole suspension solve of a step): per row and sweep
    30 B while the iterate is stored in fp32 (20 B fp32 coefficient copies + x in/out 2x4 + 2), 38 B with fp64 x, 58 B on fp64
    coefficients; 58 B per residual check (3 latS + belowS + cp + den + x, + slots/rhs amortised); 12 B once for the
    fp32 -> fp64 conversion of the iterate; 4 B for zeroing the fp32 iterate."""
    n, n32, nx = st["sweeps_timed"], st["sweeps_timed_fp32"], st["sweeps_fp32_x"]
    per_row = nx * (28.0 + 20.0 / L) + (n32 - nx) * (36.0 + 20.0 / L) + (n - n32) * (56.0 + 20.0 / L) \
        + st["residual_checks"] * (56.0 + 20.0 / L) + (16.0 if nx > 0 else 0.0)
    return per_row * T * L


# ------------------------------------------------------------------------------------------ CPU reference arm
N_FORCING = 3  # distinct synthetic forcing fields cycled over the steps (iteration counts differ from step to step)


def forcing_set(cx, cy):
    from chm_b200 import synthetic
    return [synthetic.forcing(cx, cy, seed=7, step=k) for k in range(N_FORCING)]


def cpu_reference_run(side, n_warm, n_steps):
    """The CPU restatement of the reference algorithm (OpenMP assembly, GMRES(30) + ILUT re-factorised every step, tol 1e-8)
    on the same generator, forcing cycle and PBSM3D options as the GPU arm.  Returns (triangles, [seconds per step], iters)."""
    from chm_b200 import synthetic
    from oracle.cpu_ref import CpuReference
    from oracle.pbsm3d_oracle import Config
    mesh = synthetic.uniform_mesh(side, side)
    geo = mesh.geometry()
    Fs = forcing_set(geo.cx, geo.cy)
    ref = CpuReference(Config.functional_test(NLAYER), mesh, geo)
    times, iters = [], None
    for k in range(n_warm + n_steps):
        t0 = time.perf_counter()
        r = ref.step(Fs[k % N_FORCING], 3600.0)
        dt = time.perf_counter() - t0
        if k >= n_warm:
            times.append(dt)
            iters = (r["stats"]["susp_iters"], r["stats"]["dep_iters"])
    return mesh.n_local, times, iters


def pick_cpu_side(budget_s, n_total):
    """Full config c2 (708x708 squares) when n_total steps fit the time budget on this host, else the quarter-size sample."""
    t0 = time.perf_counter()
    cpu_reference_run(354, 0, 1)
    t_quarter = time.perf_counter() - t0  # includes mesh generation: an upper bound
    return 708 if 4.0 * t_quarter * n_total <= budget_s else 354


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle.cpu_ref import host_threads
    side = pick_cpu_side(150.0, args.warmup + args.steps)
    ntri, times, iters = cpu_reference_run(side, args.warmup, args.steps)
    ms = 1e3 * float(np.mean(times))
    value = ntri * NLAYER / (ms * 1e-3)
    sample = (f"{side}x{side} squares = {ntri} triangles x {NLAYER} layers "
              f"({'config c2 in full' if side == 708 else '1/4 of config c2'}), same forcing cycle")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "uniform synthetic mesh (BASELINE c2 generator), nLayer 10, functional-test PBSM3D block; "
                               "CPU restatement of the reference: OpenMP assembly, GMRES(30)+ILUT(3.0,1e-4) local to each thread block, tol 1e-8",
                   "sample": sample, "gmres_iterations": iters},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample():
    """cpu_baseline leg of the product line: ≈10–30 s of the CPU restatement on the host cores."""
    from oracle.cpu_ref import host_threads
    side = pick_cpu_side(30.0, 4)
    ntri, times, _ = cpu_reference_run(side, 1, 3)
    v = ntri * NLAYER / float(np.median(times))
    return {"value": v, "unit": UNIT, "cores": host_threads(), "kind": "port",
            "sample": f"{side}x{side} squares = {ntri} triangles x {NLAYER} layers "
                      f"({'config c2 in full' if side == 708 else '1/4 of config c2'}), median of 3 steps after 1 warm-up; "
                      "C++/OpenMP restatement (GMRES(30)+ILUT local per thread), not the CHM binary"}


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--side", type=int, default=0, help="override squares per side (debug)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2 (default, the driver's line): uniform 1 M triangles per GPU; c3 / c4: BASELINE's variable-resolution "
                         "5 M / 10 M-triangle Delaunay meshes (strong scaling over the GPUs given), recorded in profiles/")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from chm_b200 import build, capi, synthetic
    from chm_b200.mesh import partition_mesh

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    build.build()
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())

    side = args.side or mesh_side(world)
    if args.workload == "c2":
        gmesh = synthetic.uniform_mesh(side, side)
        wl = f"BASELINE c2 generator: {side}x{side} squares of 30 m"
    else:
        target = {"c3": 5_000_000, "c4": 10_000_000}[args.workload]
        gmesh = synthetic.variable_mesh(target)
        wl = f"BASELINE {args.workload} generator: variable-resolution Delaunay mesh (10:1 area range, seed 20250101), target {target}"
    G = gmesh.n_local
    mesh = partition_mesh(gmesh, rank, world) if world > 1 else gmesh
    del gmesh
    T = mesh.n_local
    geo = mesh.geometry()
    Fs = forcing_set(geo.cx[:T], geo.cy[:T])
    cfg = capi.default_config(**FUNCTEST)
    h = capi.Handle(cfg, mesh, device=local_rank, rank=rank, n_ranks=world, unique_id=uid)

    names = capi.DEFAULT_FORCING_NAMES
    dev_in = [{n: torch.from_numpy(F[n]).cuda() for n in names} for F in Fs]
    dev_out = {n: torch.empty(T, dtype=torch.float64, device="cuda") for n in capi.DEFAULT_OUTPUT_NAMES}
    pin_in = [{n: torch.from_numpy(F[n]).pin_memory() for n in names} for F in Fs]
    pin_out = {n: torch.empty(T, dtype=torch.float64).pin_memory() for n in capi.DEFAULT_OUTPUT_NAMES}
    dptr = lambda d: {n: t.data_ptr() for n, t in d.items()}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm
    for k in range(args.warmup):
        st = h.step_ptr(3600.0, dptr(dev_in[k % N_FORCING]), dptr(dev_out), device=True)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev_ms, launches, sweep_ms, sweeps, sweep32_ms, sweeps32 = 0.0, 0, 0.0, 0, 0.0, 0
    solve_bytes, solve_launches, sweeps_x32, checks = 0.0, 0, 0, 0
    phases = {"ms_assembly": 0.0, "ms_suspension_solve": 0.0, "ms_flux_and_halo": 0.0, "ms_deposition": 0.0}
    t0 = time.perf_counter()
    syncs, susp_its, dep_its = 0, [], []
    for k in range(args.steps):
        st = h.step_ptr(3600.0, dptr(dev_in[(args.warmup + k) % N_FORCING]), dptr(dev_out), device=True)
        syncs += st["host_syncs"]
        susp_its.append(st["suspension_iterations"])
        dep_its.append(st["deposition_iterations"])
        ev_ms += st["ms_total"]
        launches += st["kernel_launches"]
        sweep_ms += st["ms_line_sweeps"]
        sweeps += st["sweeps_timed"]
        sweep32_ms += st["ms_line_sweeps_fp32"]
        sweeps32 += st["sweeps_timed_fp32"]
        if st.get("persistent_kernels") and st["suspension_present"]:
            solve_bytes += persistent_solve_bytes(st, T, NLAYER)
            solve_launches += 1
            sweeps_x32 += st["sweeps_fp32_x"]
            checks += st["residual_checks"]
        for k in phases:
            phases[k] += st[k]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    ms_step = reduce_max(ev_ms / args.steps)
    wall_ms = reduce_max(wall_ms)

    # ---- the early-out path: a calm hour — no face saltates, so no suspension and no deposition solve
    calm_ms = None
    try:
        Fc = synthetic.forcing(geo.cx[:T], geo.cy[:T], seed=7, step=0, calm=True)
        dev_calm = {n: torch.from_numpy(Fc[n]).cuda() for n in names}
        for k in range(3):
            stc = h.step_ptr(3600.0, dptr(dev_calm), dptr(dev_out), device=True)
        calm_ms = reduce_max(float(stc["ms_total"])) if not stc["suspension_present"] else None
        h.step_ptr(3600.0, dptr(dev_in[0]), dptr(dev_out), device=True)  # back to a drifting state before the e2e leg
        del dev_calm
    except Exception:  # the extra figure must never cost the line
        calm_ms = None

    # ---- end-to-end arm: pinned host buffers through the reference-facing call
    for k in range(3):
        h.step_ptr(3600.0, dptr(pin_in[k % N_FORCING]), dptr(pin_out), device=False)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        h.step_ptr(3600.0, dptr(pin_in[k % N_FORCING]), dptr(pin_out), device=False)
        _ = float(pin_out["drift_mass"][0])  # the step's result is read on the host
    barrier()
    e2e_ms = reduce_max(1e3 * (time.perf_counter() - t0) / args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the two providers of PBSM3D inputs on the device, alone and fused into the e2e call
    providers = None
    if args.workload == "c2":
        import ctypes as C
        cast = lambda t: C.cast(C.c_void_p(t.data_ptr()), capi.c_double_p)
        d0, scratch = dev_in[0], torch.empty(T, dtype=torch.float64, device="cuda")
        t_sw = t_fe = 0.0
        for k in range(12):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            capi._check(h.lib, h.lib.pbsm3d_scale_wind_vert(h.h, None, cast(d0["U_R"]), cast(d0["snowdepthavg"]), cast(scratch), 1))
            t1 = time.perf_counter()
            capi._check(h.lib, h.lib.pbsm3d_fetchr(h.h, None, cast(d0["vw_dir"]), cast(scratch), 1))
            t2 = time.perf_counter()
            if k >= 2:
                t_sw += (t1 - t0) / 10
                t_fe += (t2 - t1) / 10
        h.set_providers(capi.default_wind_config())
        part = [{n: p for n, p in dptr(pi).items() if n not in ("U_2m_above_srf", "fetch")} for pi in pin_in]
        for k in range(3):
            h.step_ptr(3600.0, part[k % N_FORCING], dptr(pin_out), device=False)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            h.step_ptr(3600.0, part[k % N_FORCING], dptr(pin_out), device=False)
            _ = float(pin_out["drift_mass"][0])
        barrier()
        fused_ms = reduce_max(1e3 * (time.perf_counter() - t0) / args.steps)
        h.set_providers(None)
        providers = {"scale_wind_vert_ms": reduce_max(1e3 * t_sw), "fetchr_ms": reduce_max(1e3 * t_fe),
                     "e2e_ms_with_providers_fused": fused_ms, "h2d_bytes_per_step_fused": 6 * 8 * T * world,
                     "note": "synchronous device-pointer calls (host wall clock); fused: U_2m_above_srf and fetch are derived on the "
                             "device inside pbsm3d_step (these steps use the derived fields, so iteration counts differ slightly)"}

    total_rows = G * NLAYER
    value = total_rows / (ms_step * 1e-3)
    e2e = total_rows / (e2e_ms * 1e-3)
    # ---- roofline of the dominant kernel, timed live inside the timed steps with CUDA events
    peak, peak_src = measured_peak_gbs()
    persistent = bool(st.get("persistent_kernels"))
    # per-pass path: the dominant kernel is the sweep on fp32-rounded coefficient streams when the step used it (most sweeps do)
    use32 = sweeps32 > 0
    n_dom = sweeps32 if use32 else sweeps
    avg_sweep_ms = (sweep32_ms if use32 else sweep_ms) / max(n_dom, 1)
    row_bytes = sweep_bytes_per_row(NLAYER, use32)
    ach = row_bytes * T * NLAYER / (avg_sweep_ms * 1e-3) / 1e9 if n_dom else 0.0
    avg64_ms = (sweep_ms - sweep32_ms) / max(sweeps - sweeps32, 1)
    ach64 = sweep_bytes_per_row(NLAYER) * T * NLAYER / (avg64_ms * 1e-3) / 1e9 if sweeps > sweeps32 else None
    traffic = traffic32 = None
    tp = os.path.join(ROOT, "profiles", "sweep_dram_bytes_per_launch.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic, traffic32 = tj.get("dram_bytes_per_launch"), tj.get("dram_bytes_per_launch_fp32_streams")
        except Exception:
            traffic = traffic32 = None
    if persistent and solve_launches:
        ach_p = solve_bytes / (sweep_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm",
                    "kernel": "gs_persistent_kernel<10> (one cooperative launch = the whole suspension solve of a step: line "
                              "Gauss-Seidel sweeps in three storage phases + residual checks, grid barriers between colour passes)",
                    "achieved": ach_p, "peak": peak, "unit": "GB/s", "frac": ach_p / peak, "traffic": traffic, "peak_source": peak_src,
                    "bytes_per_launch": solve_bytes / solve_launches, "avg_launch_ms": sweep_ms / solve_launches,
                    "launches_timed": int(solve_launches), "share_of_step": sweep_ms / max(ev_ms, 1e-9),
                    "per_launch": {"sweeps": sweeps / solve_launches, "of_which_fp32_coefficients": sweeps32 / solve_launches,
                                   "of_which_fp32_x": sweeps_x32 / solve_launches, "residual_checks": checks / solve_launches},
                    "bytes_per_row": {"fp32_x_sweep": 28.0 + 20.0 / NLAYER, "fp32_coefficient_sweep": 36.0 + 20.0 / NLAYER,
                                      "fp64_sweep": 56.0 + 20.0 / NLAYER, "residual_check": 56.0 + 20.0 / NLAYER}}
    else:
        roofline = {"bound": "hbm",
                    "kernel": ("gs_sweep_kernel<10, float> (fp32-rounded coefficient streams, fp64 x and arithmetic; " if use32
                               else "gs_sweep_kernel<10, double> (") + "one full sweep = all colour passes)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic if not use32 else traffic32, "peak_source": peak_src,
                    "bytes_per_launch": row_bytes * T * NLAYER, "avg_launch_ms": avg_sweep_ms,
                    "launches_timed": int(n_dom), "share_of_step": (sweep32_ms if use32 else sweep_ms) / max(ev_ms, 1e-9),
                    "fp64_stream_sweeps": {"launches_timed": int(sweeps - sweeps32), "avg_launch_ms": avg64_ms, "achieved": ach64,
                                           "frac": (ach64 / peak) if ach64 else None,
                                           "bytes_per_launch": sweep_bytes_per_row(NLAYER) * T * NLAYER,
                                           "share_of_step": (sweep_ms - sweep32_ms) / max(ev_ms, 1e-9)}}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if args.workload == "c2" else "strong",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{wl} -> {G} triangles x nLayer {NLAYER} "
                                   f"({total_rows} unknowns), Morton order, functional-test PBSM3D block, tol 1e-8"
                                   + ("" if world == 1 else f", {world} ranks by CHM contiguous global-id partition"),
                       "triangles": G, "nLayer": NLAYER, "solver": "multicolour line Gauss-Seidel (auto)", "colours": st["n_colours"],
                       "forcing": f"{N_FORCING} distinct seeded fields cycled over the steps; every solve starts from x0 = 0",
                       "suspension_iterations": susp_its, "deposition_iterations": dep_its,
                       "deposition_solver": {1: "Jacobi-CG", 2: "Jacobi-Chebyshev (auto)", 3: "multicolour SOR, Young's omega (auto)"}.get(st["deposition_solver_used"], "?"),
                       "suspension_residual": st["suspension_residual"], "deposition_residual": st["deposition_residual"],
                       "host_syncs_per_step": syncs / args.steps,
                       "halo_transport": {0: "none (single rank)", 1: "nccl", 2: "peer memory (cudaIpc over NVLink)"}[st["halo_transport"]],
                       "halo_exchanges_per_step": st["halo_exchanges"],
                       "halo_exchanges_inside_solver_kernels": st["halo_fused"],
                       "l2_policy": "working set of one step (~1 GB of coefficient streams per rank) exceeds the 126 MB L2; no flush needed",
                       "phases_ms": {k: v / args.steps for k, v in phases.items()}, "wall_ms_per_step": wall_ms,
                       "providers": providers, "calm_step_ms": calm_ms},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": 8 * 8 * T * world,
                    "d2h_bytes_per_step": 8 * 8 * T * world},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline and args.workload == "c2":
            line["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

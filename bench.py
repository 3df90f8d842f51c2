#!/usr/bin/env python
"""bench.py — PBSM3D element-layer solves/s on B200 (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 24 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm on the box's host cores (CPU)

A "step" is one complete PBSM3D::run (saltation + suspension assembly, suspension solve, flux integration, halo,
deposition solve, drift update) on a synthetic mesh.  The headline line (`value`, `e2e`, `roofline`) is

  N = 1  : config c2 of BASELINE.md — 708×708 squares of 30 m split into 1 002 528 triangles, nLayer 10, fp64,
           PBSM3D options = the functional-test block (functional_tests/mesh_versioning/json_mesh.json:82-99)
  N > 1  : weak scaling — the same generator at ≈1.0 M triangles per GPU, partitioned by CHM's contiguous
           global-id rule; ghost-face halos and global reductions go through peer memory over NVLink (cudaIpc
           arenas; PBSM3D_HALO=nccl forces NCCL send/recv + all-reduce instead).

`value` = (triangles × layers over all ranks) / (CUDA-event time of the step, max over ranks), forcing resident in HBM.
`e2e`   = the same through pbsm3d_step() with pinned HOST buffers (H2D of 8 forcing arrays + D2H of 8 outputs inside
          the timed region).

The same run adds, inside `config`:
  variants.default_block : the code-default PBSM3D block (pow-based settling, smooth_coeff 820) on the same mesh
  strong_c4              : BASELINE c4 — the 10 M-triangle variable-resolution Delaunay mesh, STRONG scaling over the N GPUs
                           (CHM contiguous partition), with the efficiency against the 1-GPU time of the same mesh
and, for N > 1, `parity_check`: true residuals of both linear systems recomputed off the library (numpy SpMV over the exported
rows + all-reduce), the deposition conservation identity, and the outputs against a 1-rank run of the same global mesh.
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
DEFAULT_BLOCK = dict(nLayer=NLAYER)  # PBSM3D.cpp:223-258 code defaults
C4_TARGET = 10_000_000
CACHE_DIR = os.environ.get("PBSM3D_BENCH_CACHE", "/tmp/pbsm3d_bench_cache")


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


# ------------------------------------------------------------------------------------------ algorithmic bytes
def sweep_bytes_per_row(L: int, fp32_streams: bool = False, fp32_x: bool = False) -> float:
    """One full line Gauss-Seidel sweep per unknown row: 3 row-scaled lateral coefficients + scaled sub-diagonal + cp
    (5×8 B, or 5×4 for the fp32-rounded copies) + x read once + written once (2×8, or 2×4 while the iterate is stored in fp32),
    plus per face 3 int32 neighbour slots + the scaled rhs (20 B) amortised over L layers."""
    return (20.0 if fp32_streams else 40.0) + (8.0 if fp32_x else 16.0) + 20.0 / L


def persistent_solve_bytes(st, T, L):
    """Algorithmic HBM bytes of ONE launch of gs_persistent_kernel (the whole suspension solve of a step): the sweeps of its
    three storage phases, 58 B/row per residual check (3 latS + belowS + cp + den + x, + slots/rhs amortised), and 16 B/row
    once for zeroing the fp32 iterate and converting it to fp64."""
    n, n32, nx = st["sweeps_timed"], st["sweeps_timed_fp32"], st["sweeps_fp32_x"]
    # face-column updates per storage phase: with the solver's active set (pbsm3d_stats.active_set) only the executed ones count
    cols = (st.get("column_updates_fp32_x", 0), st.get("column_updates_fp32", 0), st.get("column_updates_fp64", 0), st.get("columns_checked", 0))
    if not st.get("active_set") or sum(cols) == 0:
        cols = (nx * T, (n32 - nx) * T, (n - n32) * T, st["residual_checks"] * T)
    per_col = cols[0] * sweep_bytes_per_row(L, True, True) + cols[1] * sweep_bytes_per_row(L, True) + cols[2] * sweep_bytes_per_row(L) \
        + cols[3] * (56.0 + 20.0 / L)
    return per_col * L + (16.0 * T * L if nx > 0 else 0.0)


def assembly_bytes_per_row(L: int) -> float:
    """reads ≈190 B/face, writes 8(6-2/L) matrix values + rhs + u_z + csubl ≈ 89 B per element-layer at L = 10."""
    return 190.0 / L + 8.0 * (6.0 - 2.0 / L) + 24.0


# ------------------------------------------------------------------------------------------ CPU reference arm
N_FORCING = 3  # distinct synthetic forcing fields cycled over the steps (iteration counts differ from step to step)


def forcing_set(cx, cy):
    from chm_b200 import synthetic
    return [synthetic.forcing(cx, cy, seed=7, step=k) for k in range(N_FORCING)]


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(side, n_warm, n_steps):
    """The CPU port of the reference algorithm (OpenMP assembly, GMRES(30) + ILUT re-factorised every step, tol 1e-8)
    on the same generator, forcing cycle and PBSM3D options as the GPU arm, on ALL host threads (torchrun exports
    OMP_NUM_THREADS=1, so the thread count is passed explicitly).  Returns (triangles, [seconds per step], iters, threads)."""
    from chm_b200 import synthetic
    from oracle.cpu_ref import CpuReference
    from oracle.pbsm3d_oracle import Config
    mesh = synthetic.uniform_mesh(side, side)
    geo = mesh.geometry()
    Fs = forcing_set(geo.cx, geo.cy)
    ref = CpuReference(Config.functional_test(NLAYER), mesh, geo, n_threads=host_cores())
    times, iters, nthr = [], None, host_cores()
    for k in range(n_warm + n_steps):
        t0 = time.perf_counter()
        r = ref.step(Fs[k % N_FORCING], 3600.0)
        dt = time.perf_counter() - t0
        nthr = r["stats"]["n_threads"]
        if k >= n_warm:
            times.append(dt)
            iters = (r["stats"]["susp_iters"], r["stats"]["dep_iters"])
    return mesh.n_local, times, iters, nthr


def pick_cpu_side(budget_s, n_total):
    """Full config c2 (708x708 squares) when n_total steps fit the time budget on this host, else the quarter-size sample."""
    t0 = time.perf_counter()
    cpu_reference_run(354, 0, 1)
    t_quarter = time.perf_counter() - t0  # includes mesh generation: an upper bound
    return 708 if 4.0 * t_quarter * n_total <= budget_s else 354


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host cores.  At every N it runs the SAME bounded sample — config c2
    (what one GPU of the weak-scaled job holds) on all host threads — because the metric is a throughput and the host does not
    grow with N; `config.sample` says so.  Rank 0 only."""
    if rank != 0:
        return
    side = pick_cpu_side(150.0, args.warmup + args.steps)
    ntri, times, iters, nthr = cpu_reference_run(side, args.warmup, args.steps)
    ms = 1e3 * float(np.mean(times))
    value = ntri * NLAYER / (ms * 1e-3)
    sample = (f"{side}x{side} squares = {ntri} triangles x {NLAYER} layers "
              f"({'config c2 in full' if side == 708 else '1/4 of config c2'}; the per-GPU share of the N-GPU weak-scaled job), "
              f"same forcing cycle, {nthr} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "uniform synthetic mesh (BASELINE c2 generator), nLayer 10, functional-test PBSM3D block; "
                               "CPU PORT of the reference (not the CHM binary): OpenMP assembly, GMRES(30)+ILUT(3.0,1e-4) local to "
                               "each thread block, tol 1e-8",
                   "sample": sample, "gmres_iterations": iters, "median_ms_per_step": 1e3 * float(np.median(times))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthr, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample():
    """cpu_baseline leg of the product line: ≈10–30 s of the CPU port on the host cores."""
    side = pick_cpu_side(30.0, 4)
    ntri, times, _, nthr = cpu_reference_run(side, 1, 3)
    v = ntri * NLAYER / float(np.median(times))
    return {"value": v, "unit": UNIT, "cores": nthr, "kind": "port",
            "sample": f"{side}x{side} squares = {ntri} triangles x {NLAYER} layers "
                      f"({'config c2 in full' if side == 708 else '1/4 of config c2'}), median of 3 steps after 1 warm-up; "
                      "C++/OpenMP port of the reference algorithm (GMRES(30)+ILUT local per thread), not the CHM binary"}


# ------------------------------------------------------------------------------------------ GPU arm
class Job:
    """Process-wide plumbing: rank, device, barriers, reductions."""

    def __init__(self, solo=False):
        import torch
        self.torch = torch
        self.rank = 0 if solo else int(os.environ.get("RANK", "0"))
        self.world = 1 if solo else int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.uid = None

    def init(self):
        torch = self.torch
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.uid = self.new_uid()

    def new_uid(self):
        """A fresh NCCL unique id for one more library communicator (rank 0 makes it, everyone gets it)."""
        torch = self.torch
        from chm_b200 import capi
        if self.world == 1:
            return None
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            buf.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        self.dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def reduce(self, x: float, op="max") -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return float(t.item())


def timed_steps(job, h, dev_in, dev_out, steps, warmup):
    """W untimed + K timed device-resident steps, bracketed by barrier + synchronize.  Returns per-step CUDA-event times (the
    library's own events on its compute stream), their mean/median as the max over ranks, and the aggregated stats."""
    dptr = lambda d: {n: t.data_ptr() for n, t in d.items()}
    nf = len(dev_in)
    for k in range(warmup):
        st = h.step_ptr(3600.0, dptr(dev_in[k % nf]), dptr(dev_out), device=True)
    job.barrier()
    acc = {"ms": [], "launches": 0, "sweep_ms": 0.0, "sweeps": 0, "sweep32_ms": 0.0, "sweeps32": 0, "sweeps_x32": 0, "checks": 0,
           "solve_bytes": 0.0, "solve_launches": 0, "syncs": 0, "susp_its": [], "dep_its": [],
           "phases": {"ms_assembly": 0.0, "ms_suspension_solve": 0.0, "ms_flux_and_halo": 0.0, "ms_deposition": 0.0}}
    t0 = time.perf_counter()
    for k in range(steps):
        st = h.step_ptr(3600.0, dptr(dev_in[(warmup + k) % nf]), dptr(dev_out), device=True)
        acc["ms"].append(st["ms_total"])
        acc["launches"] += st["kernel_launches"]
        acc["syncs"] += st["host_syncs"]
        acc["susp_its"].append(st["suspension_iterations"])
        acc["dep_its"].append(st["deposition_iterations"])
        acc["sweep_ms"] += st["ms_line_sweeps"]
        acc["sweeps"] += st["sweeps_timed"]
        acc["sweep32_ms"] += st["ms_line_sweeps_fp32"]
        acc["sweeps32"] += st["sweeps_timed_fp32"]
        if st.get("persistent_kernels") and st["suspension_present"]:
            acc["solve_bytes"] += persistent_solve_bytes(st, h.T, h.L)
            acc["solve_launches"] += 1
            acc["sweeps_x32"] += st["sweeps_fp32_x"]
            acc["checks"] += st["residual_checks"]
            acc["col_updates"] = acc.get("col_updates", 0) + st.get("column_updates_fp32_x", 0) + st.get("column_updates_fp32", 0) + st.get("column_updates_fp64", 0)
            acc["col_sweeps"] = acc.get("col_sweeps", 0) + st["sweeps_timed"] * h.T
        for p in acc["phases"]:
            acc["phases"][p] += st[p]
    job.barrier()
    acc["wall_ms"] = job.reduce(1e3 * (time.perf_counter() - t0) / steps)
    acc["ms_step"] = job.reduce(float(np.mean(acc["ms"])))
    acc["ms_median"] = job.reduce(float(np.median(acc["ms"])))
    acc["last"] = st
    return acc


def time_next_rows(h, T, local_rank):
    """SURVEY §8f ranks 3 and 4 on the bench handle's device (N = 1 only): snobal's drift_mass consumer on the handle's own
    device-resident drift_mass with a device-resident snowpack (no PCIe), and snow_slide on steep synthetic terrain of the same
    size.  Extra figures: failures are reported, never fatal."""
    import ctypes as C
    import torch
    from chm_b200 import capi, synthetic
    out = {}
    try:
        rng = np.random.default_rng(5)
        z_s = rng.uniform(0.0, 1.5, T)
        rho = rng.uniform(80.0, 500.0, T)
        two = z_s > 0.1
        host = {"z_s": z_s, "m_s": rho * z_s, "rho": rho, "layer_count": np.where(two, 2, 1).astype(np.int32),
                "z_s_0": np.where(two, 0.1, z_s), "z_s_l": np.where(two, z_s - 0.1, 0.0)}
        host["m_s_0"], host["m_s_l"] = rho * host["z_s_0"], rho * host["z_s_l"]
        for k in ("cc_s", "cc_s_0", "cc_s_l", "h2o_total", "h2o_vol", "h2o", "h2o_max", "h2o_sat"):
            host[k] = np.zeros(T)
        for k in ("T_s", "T_s_0", "T_s_l"):
            host[k] = np.full(T, 265.0)
        dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in host.items()}
        pk = capi.Snowpack()
        for k in capi.SNOWPACK_FIELDS:
            setattr(pk, k, C.cast(C.c_void_p(dev[k].data_ptr()), capi.c_int32_p if k == "layer_count" else capi.c_double_p))
        ts = []
        for k in range(12):
            t0 = time.perf_counter()
            capi._check(h.lib, h.lib.pbsm3d_apply_drift(h.h, None, C.byref(pk), None, None, None, 1))
            ts.append(time.perf_counter() - t0)
        ms = 1e3 * float(np.median(ts[2:]))
        nbytes = T * (19 * 8 * 2 - 8 + 8 + 4)  # 18 doubles + 1 int32 read and written, drift_mass + its slot index read
        out["snobal_apply_drift"] = {"ms": ms, "bytes": nbytes, "GB_per_s": nbytes / (ms * 1e-3) / 1e9,
                                     "what": "synchronous call, device-resident snowpack (19 SoA fields) and the handle's own drift_mass"}
    except Exception as e:
        out["snobal_apply_drift"] = {"error": repr(e)[:200]}
    try:
        side = int(round((T / 2) ** 0.5))
        m = synthetic.with_elevation(synthetic.uniform_mesh(side, side))
        geo = m.geometry()
        rng = np.random.default_rng(11)
        f = 0.5 + 0.25 * (np.sin(geo.cx / 2100.0) * np.cos(geo.cy / 1700.0) + np.sin((geo.cx + geo.cy) / 3900.0))
        sd = np.abs(1.0 * np.clip(f, 0.02, None) * (1.0 + 0.05 * rng.standard_normal(m.n_local)))
        hs = capi.Handle(capi.default_config(nLayer=2), m, device=local_rank)
        hs.slide_init()
        # the vertical depth is the slope-normal one over cos(slope), from the vertices
        v = m.vertex[m.elem]
        nrm = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
        cosf = np.maximum(0.001, np.abs(nrm[:, 2]) / np.linalg.norm(nrm, axis=1))
        res = [hs.slide_run(sd, sd / cosf, sd * 300.0)[1] for _ in range(5)]
        z = np.zeros(m.n_local)
        calm = [hs.slide_run(z, z, z)[1]["ms_device"] for _ in range(4)]
        out["snow_slide"] = {"triangles": int(m.n_local), "terrain": "synthetic.alpine_terrain (slopes to 68 degrees), 1 m snow cover",
                             "faces_fired": int(res[-1]["faces_fired"]), "wavefront_rounds": int(res[-1]["wavefront_rounds"]),
                             "ms_device": float(np.median([r["ms_device"] for r in res[1:]])), "calm_ms_device": float(np.median(calm[1:]))}
        hs.close()
    except Exception as e:
        out["snow_slide"] = {"error": repr(e)[:200]}
    return out


def cached_mesh(job, target):
    """The variable-resolution Delaunay mesh of BASELINE c3/c4 (seed 20250101).  Rank 0 generates it once per box and caches it
    under CACHE_DIR (the driver runs N = 1, 2, 4, 8 back to back on one box); everyone loads the cache."""
    from chm_b200 import synthetic
    from chm_b200.mesh import TriMesh
    path = os.path.join(CACHE_DIR, f"variable_{target}.npz")
    t0 = time.perf_counter()
    made = False
    if job.rank == 0 and not os.path.exists(path):
        os.makedirs(CACHE_DIR, exist_ok=True)
        m = synthetic.variable_mesh(target)
        tmp = path + f".tmp{os.getpid()}.npz"
        np.savez(tmp, vertex=m.vertex, elem=m.elem, neigh=m.neigh)
        os.replace(tmp, path)
        made = True
        del m
    job.barrier()
    d = np.load(path)
    mesh = TriMesh(d["vertex"], d["elem"], d["neigh"], {})
    return mesh, {"generated_this_run": made, "seconds": time.perf_counter() - t0}


def run_strong_c4(job, steps):
    """BASELINE c4: ≈10 M variable-resolution triangles × 10 layers, STRONG scaling over the job's GPUs, CHM contiguous
    partition (triangulation.cpp:1482-1531).  Efficiency = t(1 GPU) / (N · t(N GPUs)) with t(1) measured on this box (cached
    by the N = 1 run; measured on rank 0 while the others wait when the cache is empty)."""
    import torch
    from chm_b200 import capi
    from chm_b200.mesh import partition_mesh
    t_begin = time.perf_counter()
    gmesh, info = cached_mesh(job, C4_TARGET)
    G = gmesh.n_local
    t1_path = os.path.join(CACHE_DIR, f"c4_t1_{G}.json")

    def one(jb, mesh, rank, world, uid, k, w):
        T = mesh.n_local
        geo = mesh.geometry()
        Fs = forcing_set(geo.cx[:T], geo.cy[:T])
        del geo
        h = capi.Handle(capi.default_config(**FUNCTEST), mesh, device=job.local_rank, rank=rank, n_ranks=world, unique_id=uid)
        dev_in = [{n: torch.from_numpy(F[n]).cuda() for n in capi.DEFAULT_FORCING_NAMES} for F in Fs]
        dev_out = {n: torch.empty(T, dtype=torch.float64, device="cuda") for n in capi.DEFAULT_OUTPUT_NAMES}
        acc = timed_steps(jb, h, dev_in, dev_out, k, w)
        st = acc["last"]
        res = {"ms_per_step": acc["ms_step"], "median_ms_per_step": acc["ms_median"], "suspension_iterations": acc["susp_its"],
               "deposition_iterations": acc["dep_its"], "colours": st["n_colours"], "suspension_residual": st["suspension_residual"],
               "deposition_residual": st["deposition_residual"], "phases_ms": {p: v / k for p, v in acc["phases"].items()},
               "launches_per_step": acc["launches"] / k, "local_triangles": T}
        h.close()
        del dev_in, dev_out
        torch.cuda.empty_cache()
        return res

    k, w = max(3, min(steps, 6)), 3
    out = {"workload": f"BASELINE c4 generator: variable-resolution Delaunay mesh (10:1 area range, seed 20250101), target {C4_TARGET} "
                       f"-> {G} triangles x nLayer {NLAYER}, Morton order, functional-test PBSM3D block, tol 1e-8",
           "triangles": G, "n_gpus": job.world, "scaling": "strong", "steps": k, "warmup": w, "mesh": info}
    t1 = None
    if os.path.exists(t1_path):
        try:
            t1 = json.load(open(t1_path))
        except Exception:
            t1 = None
    if job.world == 1:
        res = one(job, gmesh, 0, 1, None, k, w)
        out.update(res)
        out["value"] = G * NLAYER / (res["ms_per_step"] * 1e-3)
        out["efficiency_vs_1gpu"] = 1.0
        try:
            json.dump({"ms_per_step": res["ms_per_step"], "suspension_iterations": res["suspension_iterations"]}, open(t1_path, "w"))
        except Exception:
            pass
    else:
        if t1 is None:  # no 1-GPU time from an earlier run on this box: rank 0 measures it now, the others wait
            if job.rank == 0:
                r1 = one(Job(solo=True), gmesh, 0, 1, None, 3, 3)
                try:
                    json.dump({"ms_per_step": r1["ms_per_step"], "suspension_iterations": r1["suspension_iterations"]}, open(t1_path, "w"))
                except Exception:
                    pass
            job.barrier()
            t1 = json.load(open(t1_path)) if os.path.exists(t1_path) else None
        uid = job.new_uid()
        mesh = partition_mesh(gmesh, job.rank, job.world)
        del gmesh
        res = one(job, mesh, job.rank, job.world, uid, k, w)
        out.update(res)
        out["value"] = G * NLAYER / (res["ms_per_step"] * 1e-3)
        out["one_gpu"] = t1
        out["efficiency_vs_1gpu"] = (t1["ms_per_step"] / (job.world * res["ms_per_step"])) if t1 else None
    out["seconds_total"] = time.perf_counter() - t_begin
    return out


def parity_check(job, gmesh, mesh, h, Fg, cfg_kw, nl):
    """N > 1: the partitioned solve against quantities recomputed OFF the library.
    (1) true relative residuals of both linear systems: numpy SpMV over the rows each rank exports
        (pbsm3d_get_suspension_system / _deposition_system) with the all-gathered solution, one all-reduce;
    (2) the deposition conservation identity  sum_i area_i q_i = sum_i rhs_i  (the off-diagonals of a row and its neighbours'
        cancel pairwise);
    (3) Qsusp / Qsalt / drift_mass / the suspended concentration against a 1-rank run of the same GLOBAL mesh on rank 0's GPU."""
    import torch
    from chm_b200 import capi
    dist, world, rank = job.dist, job.world, job.rank
    T, G = mesh.n_local, gmesh.n_local
    sizes = np.asarray(mesh.local_sizes, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    s0 = int(starts[rank])
    outs, st = h.step(3600.0, {k: v[s0:s0 + T] for k, v in Fg.items()})

    def gather(a):  # [.., T] per rank -> [.., G] on every rank (owned ranges are contiguous in the global numbering)
        a = np.ascontiguousarray(a, dtype=np.float64)
        lead = a.shape[:-1]
        tmax = int(sizes.max())
        pad = np.zeros(lead + (tmax,))
        pad[..., :T] = a
        t = torch.from_numpy(pad).cuda()
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        return np.concatenate([allt[r].cpu().numpy()[..., :int(sizes[r])] for r in range(world)], axis=-1)

    gid = mesh.global_id  # [T + ghosts]
    nb = mesh.neigh  # local ids, -1 none, >= T ghost
    has = nb >= 0
    nbg = np.where(has, gid[np.maximum(nb, 0)], 0)  # global id of each neighbour
    # ---- (1a) suspension system
    s = h.suspension_system()
    xg = gather(h.solution())  # [L, G]
    x_own = xg[:, s0:s0 + T]
    r = -s["diag"] * x_own
    r[0] += s["rhs0"]
    for j in range(3):
        r -= s["lat"][j] * np.where(has[:, j][None], xg[:, nbg[:, j]], 0.0)
    r[1:] -= s["below"][1:] * x_own[:-1]
    r[:-1] -= s["above"][:-1] * x_own[1:]
    rr = job.reduce(float((r * r).sum()), "sum")
    bb = job.reduce(float((s["rhs0"] ** 2).sum()), "sum")
    del s, r
    # ---- (1b) + (2) deposition system
    d = h.deposition_system()
    qg = gather(d["q"])
    rd = d["rhs"] - d["diag"] * d["q"]
    for j in range(3):
        rd -= d["off"][j] * np.where(has[:, j], qg[nbg[:, j]], 0.0)
    rrd = job.reduce(float((rd * rd).sum()), "sum")
    bbd = job.reduce(float((d["rhs"] ** 2).sum()), "sum")
    area = mesh.geometry().area
    lhs = job.reduce(float((area * d["q"]).sum()), "sum")
    rhs = job.reduce(float(d["rhs"].sum()), "sum")
    rhs_abs = job.reduce(float(np.abs(d["rhs"]).sum()), "sum")
    # ---- (3) the same global mesh on ONE rank
    got = {k: gather(outs[k]) for k in ("Qsusp", "Qsalt", "drift_mass")}
    res = {"mesh": f"{G} triangles over {world} ranks, forcing field 0", "suspension_true_residual": float(np.sqrt(rr / bb)) if bb > 0 else None,
           "suspension_residual_reported": st["suspension_residual"],
           "deposition_true_residual": float(np.sqrt(rrd / bbd)) if bbd > 0 else None,
           "deposition_residual_reported": st["deposition_residual"],
           "deposition_conservation_rel": abs(lhs - rhs) / max(rhs_abs, 1e-300),
           "how": "numpy SpMV over the exported rows with the all-gathered solution + one all-reduce; 1-rank run of the global mesh on rank 0"}
    if rank == 0:
        h1 = capi.Handle(capi.default_config(**cfg_kw), gmesh, device=job.local_rank)
        o1, st1 = h1.step(3600.0, Fg)
        x1 = h1.solution()
        rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
        res["vs_single_rank_rel_l2"] = {k: rel(got[k], o1[k]) for k in got}
        res["vs_single_rank_rel_l2"]["suspended_concentration"] = rel(xg, x1)
        res["iterations_single_rank"] = [st1["suspension_iterations"], st1["deposition_iterations"]]
        res["iterations_partitioned"] = [st["suspension_iterations"], st["deposition_iterations"]]
        h1.close()
        worst = max(res["vs_single_rank_rel_l2"].values())
        res["ok"] = bool(res["suspension_true_residual"] <= 1.05e-8 and res["deposition_true_residual"] <= 1.05e-8
                         and res["deposition_conservation_rel"] <= 1e-6 and worst <= 1e-6)
        res["bars"] = "residuals <= 1e-8 (the reference's stopping rule), conservation <= 1e-6, outputs vs 1 rank <= 1e-6 rel. L2 (north_star)"
    job.barrier()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24, help="BASELINE.md §3: 24 steps")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the strong-scaling block on the 10 M-triangle mesh")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--side", type=int, default=0, help="override squares per side (debug)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5", "c5shard"],
                    help="c2 (default, the driver's line): uniform 1 M triangles per GPU; c3 / c4 / c5: BASELINE's variable-resolution "
                         "5 M / 10 M / 50 M-triangle Delaunay meshes as the HEADLINE workload (strong scaling over the GPUs given; c5 "
                         "with nLayer 20), recorded in profiles/")
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
    from chm_b200 import build, capi, synthetic
    from chm_b200.mesh import partition_mesh

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    if local_rank == 0:
        build.build()
    job = Job()
    job.init()
    job.barrier()  # the other ranks load the library only after local rank 0 has (re)built it
    barrier, reduce_max = job.barrier, job.reduce

    nl = NLAYER
    cfg_kw = dict(FUNCTEST)
    side = args.side or mesh_side(world)
    if args.workload == "c2":
        gmesh = synthetic.uniform_mesh(side, side)
        wl = f"BASELINE c2 generator: {side}x{side} squares of 30 m"
    else:
        # c5shard: one GPU's share of config c5 (50 M triangles x nLayer 20 over 8 GPUs = 6.25 M x 20 = 125 M unknowns) run on ONE GPU
        target = {"c3": 5_000_000, "c4": C4_TARGET, "c5": 50_000_000, "c5shard": 6_250_000}[args.workload]
        if args.workload in ("c5", "c5shard"):
            nl = 20
            cfg_kw["nLayer"] = 20
        gmesh, _ = cached_mesh(job, target)
        wl = f"BASELINE {args.workload} generator: variable-resolution Delaunay mesh (10:1 area range, seed 20250101), target {target}"
    G = gmesh.n_local
    mesh = partition_mesh(gmesh, rank, world) if world > 1 else gmesh
    keep_global = world > 1 and args.workload == "c2" and not args.no_parity
    if not keep_global:
        del gmesh
    T = mesh.n_local
    geo = mesh.geometry()
    Fs = forcing_set(geo.cx[:T], geo.cy[:T])
    cfg = capi.default_config(**cfg_kw)
    h = capi.Handle(cfg, mesh, device=local_rank, rank=rank, n_ranks=world, unique_id=job.uid)

    names = capi.DEFAULT_FORCING_NAMES
    dev_in = [{n: torch.from_numpy(F[n]).cuda() for n in names} for F in Fs]
    dev_out = {n: torch.empty(T, dtype=torch.float64, device="cuda") for n in capi.DEFAULT_OUTPUT_NAMES}
    pin_in = [{n: torch.from_numpy(F[n]).pin_memory() for n in names} for F in Fs]
    pin_out = {n: torch.empty(T, dtype=torch.float64).pin_memory() for n in capi.DEFAULT_OUTPUT_NAMES}
    dptr = lambda d: {n: t.data_ptr() for n, t in d.items()}

    # ---- device-resident arm (the headline `value`)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    acc = timed_steps(job, h, dev_in, dev_out, args.steps, args.warmup)
    st = acc["last"]
    ms_step, wall_ms = acc["ms_step"], acc["wall_ms"]

    # ---- the assembly alone (CUDA events, mean of 20 launches of face_prelude_kernel + assemble_kernel on the last forcing)
    asm_ms = None
    try:
        asm_ms = reduce_max(h.time_kernel(2, 20))
    except Exception:
        asm_ms = None

    # ---- the early-out path: a calm hour — no face saltates, so no suspension and no deposition solve
    calm_ms = calm_launches = None
    try:
        Fc = synthetic.forcing(geo.cx[:T], geo.cy[:T], seed=7, step=0, calm=True)
        dev_calm = {n: torch.from_numpy(Fc[n]).cuda() for n in names}
        for k in range(3):
            stc = h.step_ptr(3600.0, dptr(dev_calm), dptr(dev_out), device=True)
        if not stc["suspension_present"]:
            calm_ms, calm_launches = reduce_max(float(stc["ms_total"])), stc["kernel_launches"]
        h.step_ptr(3600.0, dptr(dev_in[0]), dptr(dev_out), device=True)  # back to a drifting state before the e2e leg
        del dev_calm
    except Exception:  # the extra figure must never cost the line
        calm_ms = None

    # ---- the consumers on the far side of the path (N = 1: extra figures, not the headline)
    next_rows = time_next_rows(h, T, local_rank) if world == 1 and args.workload == "c2" and not args.no_variants else None

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
        l_sw, l_fe = [], []
        for k in range(14):
            barrier()
            t0 = time.perf_counter()
            capi._check(h.lib, h.lib.pbsm3d_scale_wind_vert(h.h, None, cast(d0["U_R"]), cast(d0["snowdepthavg"]), cast(scratch), 1))
            t1 = time.perf_counter()
            capi._check(h.lib, h.lib.pbsm3d_fetchr(h.h, None, cast(d0["vw_dir"]), cast(scratch), 1))
            t2 = time.perf_counter()
            if k >= 2:
                l_sw.append(t1 - t0)
                l_fe.append(t2 - t1)
        # median over the calls (a rank that leaves the host barrier late makes its partners' halo wait look like kernel time)
        t_sw, t_fe = float(np.median(l_sw)), float(np.median(l_fe))
        sw_max = reduce_max(1e3 * float(np.max(l_sw)))
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
        providers = {"scale_wind_vert_ms": reduce_max(1e3 * t_sw), "fetchr_ms": reduce_max(1e3 * t_fe), "scale_wind_vert_ms_worst_call": sw_max,
                     "e2e_ms_with_providers_fused": fused_ms, "h2d_bytes_per_step_fused": 6 * 8 * T * world,
                     "note": "synchronous device-pointer calls (host wall clock, ranks aligned by a barrier before each call); fused: "
                             "U_2m_above_srf and fetch are derived on the device inside pbsm3d_step (these steps use the derived fields, "
                             "so iteration counts differ slightly)"}

    total_rows = G * nl
    value = total_rows / (ms_step * 1e-3)
    e2e = total_rows / (e2e_ms * 1e-3)
    # ---- roofline of the dominant kernel, timed live inside the timed steps with CUDA events
    peak, peak_src = measured_peak_gbs()
    persistent = bool(st.get("persistent_kernels"))
    sweeps, sweeps32, sweep_ms, sweep32_ms = acc["sweeps"], acc["sweeps32"], acc["sweep_ms"], acc["sweep32_ms"]
    ev_ms = float(np.sum(acc["ms"]))
    traffic = traffic32 = traffic_src = None
    tp = os.path.join(ROOT, "profiles", "sweep_dram_bytes_per_launch.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic, traffic32 = tj.get("dram_bytes_per_launch"), tj.get("dram_bytes_per_launch_fp32_streams")
            traffic_src = tj.get("source")
            if persistent:
                traffic = tj.get("dram_bytes_per_launch_persistent")
                traffic_src = (tj.get("persistent") or {}).get("source", traffic_src)
        except Exception:
            traffic = traffic32 = None
    if not (T == 1002528 and nl == 10):  # the ncu captures are of config c2 on one GPU; no figure for other shapes
        traffic = traffic32 = None
    if persistent and acc["solve_launches"]:
        nlch = acc["solve_launches"]
        ach_p = acc["solve_bytes"] / (sweep_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm",
                    "kernel": f"gs_persistent_kernel<{nl}> (one cooperative launch = the whole suspension solve of a step: line "
                              "Gauss-Seidel sweeps in three storage phases + residual checks, grid barriers between colour passes)",
                    "achieved": ach_p, "peak": peak, "unit": "GB/s", "frac": ach_p / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "traffic_frac_of_peak": (traffic / (sweep_ms / nlch * 1e-3) / 1e9 / peak) if traffic else None,
                    "peak_source": peak_src, "bytes_per_launch": acc["solve_bytes"] / nlch, "avg_launch_ms": sweep_ms / nlch,
                    "launches_timed": int(nlch), "share_of_step": sweep_ms / max(ev_ms, 1e-9),
                    "active_set": {"on": bool(st.get("active_set")),
                                   "column_updates_executed": (acc.get("col_updates", 0) / max(acc.get("col_sweeps", 0), 1)) if st.get("active_set") else 1.0,
                                   "note": "share of sweeps x faces column updates the solver executed: columns whose right-hand side "
                                           "and whose neighbours' iterates are still exactly zero are skipped (a no-op update, iterates "
                                           "bit-identical); bytes_per_launch counts the executed updates only"},
                    "per_launch": {"sweeps": sweeps / nlch, "of_which_fp32_coefficients": sweeps32 / nlch,
                                   "of_which_fp32_x": acc["sweeps_x32"] / nlch, "residual_checks": acc["checks"] / nlch},
                    "bytes_per_row": {"fp32_x_sweep": sweep_bytes_per_row(nl, True, True),
                                      "fp32_coefficient_sweep": sweep_bytes_per_row(nl, True),
                                      "fp64_sweep": sweep_bytes_per_row(nl), "residual_check": 56.0 + 20.0 / nl}}
    else:
        use32 = sweeps32 > 0
        n_dom = sweeps32 if use32 else sweeps
        avg_sweep_ms = (sweep32_ms if use32 else sweep_ms) / max(n_dom, 1)
        row_bytes = sweep_bytes_per_row(nl, use32)
        ach = row_bytes * T * nl / (avg_sweep_ms * 1e-3) / 1e9 if n_dom else 0.0
        avg64_ms = (sweep_ms - sweep32_ms) / max(sweeps - sweeps32, 1)
        ach64 = sweep_bytes_per_row(nl) * T * nl / (avg64_ms * 1e-3) / 1e9 if sweeps > sweeps32 else None
        roofline = {"bound": "hbm",
                    "kernel": (f"gs_sweep(_halo)_kernel<{nl}, float> (fp32-rounded coefficient streams, fp64 x and arithmetic; " if use32
                               else f"gs_sweep(_halo)_kernel<{nl}, double> (") + "one full sweep = all colour passes)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic if not use32 else traffic32, "traffic_source": traffic_src, "peak_source": peak_src,
                    "bytes_per_launch": row_bytes * T * nl, "avg_launch_ms": avg_sweep_ms,
                    "launches_timed": int(n_dom), "share_of_step": (sweep32_ms if use32 else sweep_ms) / max(ev_ms, 1e-9),
                    "fp64_stream_sweeps": {"launches_timed": int(sweeps - sweeps32), "avg_launch_ms": avg64_ms, "achieved": ach64,
                                           "frac": (ach64 / peak) if ach64 else None,
                                           "bytes_per_launch": sweep_bytes_per_row(nl) * T * nl,
                                           "share_of_step": (sweep_ms - sweep32_ms) / max(ev_ms, 1e-9)}}
    if asm_ms:
        ab = assembly_bytes_per_row(nl) * T * nl
        roofline["assembly"] = {"kernels": "face_prelude_kernel + assemble_kernel, stand-alone (mean of 20 launch pairs, CUDA events)",
                                "ms": asm_ms, "algorithmic_bytes": ab, "achieved": ab / (asm_ms * 1e-3) / 1e9,
                                "frac": ab / (asm_ms * 1e-3) / 1e9 / peak,
                                "note": "fp64-issue bound, not HBM bound (profiles/r2c_summary.md)"}

    # ---- N > 1: the partitioned solve checked off the library
    parity = None
    if keep_global:
        try:
            gg = gmesh.geometry()
            Fg = synthetic.forcing(gg.cx, gg.cy, seed=7, step=0)
            del gg
            parity = parity_check(job, gmesh, mesh, h, Fg, cfg_kw, nl)
        except Exception as e:  # never lose the line to the extra check
            parity = {"error": repr(e)}
        del gmesh

    # ---- second variant of BASELINE.md §3: the code-default PBSM3D block on the same mesh
    variants = None
    if not args.no_variants and args.workload == "c2":
        try:
            hv = capi.Handle(capi.default_config(**DEFAULT_BLOCK), mesh, device=local_rank, rank=rank, n_ranks=world,
                             unique_id=job.new_uid())
            kv = max(3, min(args.steps, 6))
            av = timed_steps(job, hv, dev_in, dev_out, kv, 3)
            sv = av["last"]
            variants = {"default_block": {
                "what": "PBSM3D.cpp:223-258 code defaults (do_fixed_settling false: omega = 1.1e7 r^1.8; smooth_coeff 820; use_R94_lambda true)",
                "ms_per_step": av["ms_step"], "median_ms_per_step": av["ms_median"], "value": total_rows / (av["ms_step"] * 1e-3),
                "sweeps": av["susp_its"], "deposition_iterations": av["dep_its"],
                "solver_used": {1: "multicolour line Gauss-Seidel", 2: "BiCGStab + column-tridiagonal preconditioner"}.get(
                    sv["suspension_solver_used"], "?"),
                "phases_ms": {p: v / kv for p, v in av["phases"].items()}, "steps": kv, "launches_per_step": av["launches"] / kv,
                "suspension_residual": sv["suspension_residual"]}}
            hv.close()
            del hv
        except Exception as e:
            variants = {"default_block": {"error": repr(e)}}
        # ---- saltation on wind-exposed patches only (most hours of a real winter): the line solver's active set against the
        #      plain sweep on the same fields (PBSM3D_ACTIVE_SET is read when a handle is created)
        if world == 1:
            try:
                Fp = [synthetic.patchy_forcing(geo.cx[:T], geo.cy[:T], seed=7, step=k) for k in range(N_FORCING)]
                dev_p = [{n: torch.from_numpy(F[n]).cuda() for n in names} for F in Fp]
                pv = {"what": "synthetic.patchy_forcing: the c2 fields with the reference-height wind cut to 3 m/s outside "
                              "wind-exposed patches; same mesh and PBSM3D block", "steps": 6}
                saved_env = os.environ.get("PBSM3D_ACTIVE_SET")
                for key, env in (("active_set", None), ("plain_sweep", "0")):
                    if env is None:
                        os.environ.pop("PBSM3D_ACTIVE_SET", None)
                    else:
                        os.environ["PBSM3D_ACTIVE_SET"] = env
                    hp = capi.Handle(cfg, mesh, device=local_rank, rank=rank, n_ranks=world, unique_id=job.new_uid())
                    ap_ = timed_steps(job, hp, dev_p, dev_out, 6, 3)
                    sp = ap_["last"]
                    nup = sp.get("column_updates_fp32_x", 0) + sp.get("column_updates_fp32", 0) + sp.get("column_updates_fp64", 0)
                    pv[key] = {"ms_per_step": ap_["ms_step"], "ms_suspension_solve": ap_["phases"]["ms_suspension_solve"] / 6,
                               "sweeps": ap_["susp_its"][:3], "active_set_used": bool(sp.get("active_set")),
                               "faces_with_rhs_share": sp.get("faces_with_rhs", 0) / T,
                               "column_updates_executed": nup / max(sp["sweeps_timed"] * T, 1)}
                    hp.close()
                    del hp
                if saved_env is None:
                    os.environ.pop("PBSM3D_ACTIVE_SET", None)
                else:
                    os.environ["PBSM3D_ACTIVE_SET"] = saved_env
                variants["patchy_saltation"] = pv
                del dev_p
            except Exception as e:
                variants["patchy_saltation"] = {"error": repr(e)}
    h.close()
    del dev_in, dev_out, pin_in, pin_out
    torch.cuda.empty_cache()

    # ---- north_star's scaling config: 10 M triangles, strong scaling
    strong = None
    if not args.no_c4 and args.workload == "c2" and not args.side:
        try:
            strong = run_strong_c4(job, args.steps)
        except Exception as e:
            strong = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if args.workload == "c2" else "strong",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{wl} -> {G} triangles x nLayer {nl} "
                                   f"({total_rows} unknowns), Morton order, functional-test PBSM3D block, tol 1e-8"
                                   + ("" if world == 1 else f", {world} ranks by CHM contiguous global-id partition"),
                       "triangles": G, "nLayer": nl, "solver": "multicolour line Gauss-Seidel (auto)", "colours": st["n_colours"],
                       "forcing": f"{N_FORCING} distinct seeded fields cycled over the steps; every solve starts from x0 = 0",
                       "median_ms_per_step": acc["ms_median"],
                       "suspension_iterations": acc["susp_its"], "deposition_iterations": acc["dep_its"],
                       "deposition_solver": {1: "Jacobi-CG", 2: "Jacobi-Chebyshev (auto)", 3: "multicolour SOR, Young's omega (auto)"}.get(
                           st["deposition_solver_used"], "?"),
                       "suspension_residual": st["suspension_residual"], "deposition_residual": st["deposition_residual"],
                       "host_syncs_per_step": acc["syncs"] / args.steps,
                       "persistent_solver_kernels": persistent,
                       "halo_transport": {0: "none (single rank)", 1: "nccl", 2: "peer memory (cudaIpc over NVLink)"}[st["halo_transport"]],
                       "halo_exchanges_per_step": st["halo_exchanges"],
                       "halo_exchanges_inside_solver_kernels": st["halo_fused"],
                       "l2_policy": "working set of one step (~1 GB of coefficient streams per rank) exceeds the 126 MB L2; no flush needed",
                       "phases_ms": {k: v / args.steps for k, v in acc["phases"].items()}, "wall_ms_per_step": wall_ms,
                       "providers": providers, "next_rows": next_rows, "calm_step_ms": calm_ms, "calm_step_launches": calm_launches,
                       "variants": variants, "strong_c4": strong,
                       "cpu_baseline_is": "a C++/OpenMP PORT of the reference algorithm, not the CHM binary (DESIGN.md §6)"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": 8 * 8 * T * world,
                    "d2h_bytes_per_step": 8 * 8 * T * world},
            "gpu_launches": int(acc["launches"]),
            "roofline": roofline,
            "clocks": clocks,
        }
        if parity is not None:
            line["parity_check"] = parity
        if world == 1 and not args.no_cpu_baseline and args.workload == "c2":
            line["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(line), flush=True)
    if world > 1:
        job.dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Stand-alone kernel timings on config c2 (or --side N) through pbsm3d_time_kernel: CUDA-event mean per launch.

    python tools/time_kernels.py [--side 708] [--block functest|default] [--reps 20]

Prints one JSON line: assembly / sweep / residual times with their algorithmic GB/s.  Environment knobs of the library
(PBSM3D_ASSEMBLY=column, PBSM3D_ASM_MINB=2|3|4, ...) are read at handle creation, so run it once per setting."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=708)
    ap.add_argument("--block", default="functest", choices=["functest", "default"])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--nlayer", type=int, default=10)
    a = ap.parse_args()
    from chm_b200 import build, capi, synthetic
    build.build()
    mesh = synthetic.uniform_mesh(a.side, a.side)
    geo = mesh.geometry()
    F = synthetic.forcing(geo.cx, geo.cy, seed=7, step=0)
    kw = dict(nLayer=a.nlayer)
    if a.block == "functest":
        kw.update(smooth_coeff=6500, do_fixed_settling=1, settling_velocity=0.5, use_R94_lambda=0)
    h = capi.Handle(capi.default_config(**kw), mesh)
    outs, st = h.step(3600.0, F)
    T, L = mesh.n_local, a.nlayer
    rows = T * L
    out = {"triangles": T, "nLayer": L, "block": a.block, "env": {k: v for k, v in os.environ.items() if k.startswith("PBSM3D_")},
           "step": {k: st[k] for k in ("ms_total", "ms_assembly", "ms_suspension_solve", "ms_deposition", "suspension_iterations",
                                       "deposition_iterations", "kernel_launches")},
           "saltating_fraction": float(np.mean(outs["Qsalt"] > 0))}
    for name, kid, bpr in (("assembly", 2, 84.0 + 190.0 / L), ("sweep_fp64", 0, 56.0 + 20.0 / L), ("residual", 1, 56.0 + 20.0 / L)):
        ms = h.time_kernel(kid, a.reps)
        out[name] = {"ms": ms, "alg_GBps": bpr * rows / (ms * 1e-3) / 1e9, "bytes_per_row": bpr}
    print(json.dumps(out), flush=True)
    h.close()


if __name__ == "__main__":
    main()

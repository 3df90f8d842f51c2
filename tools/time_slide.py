#!/usr/bin/env python
"""snow_slide on one B200 vs the reference's own snow_slide.cpp on the host (oracle/_ref/libchmref.so, single thread: its sweep is
sequential; its sort is tbb::parallel_sort in CHM, std::sort in this build).  Prints one JSON line.
    python tools/time_slide.py [side=708] [cpu_side=354] [deep=1.0]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chm_b200 import capi, synthetic  # noqa: E402
from oracle import chm_ref, slide_oracle as so  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 708
cpu_side = int(sys.argv[2]) if len(sys.argv) > 2 else 354
deep = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0


def case(n):
    m = synthetic.with_elevation(synthetic.uniform_mesh(n, n))
    geo = m.geometry()
    slope = so.face_slope(m.face_vertices().reshape(-1, 3, 3))
    return m, geo, slope, so.synthetic_snow(geo.cx, geo.cy, slope, seed=11, deep=deep)


out = {"what": "snow_slide::run, steep synthetic terrain (synthetic.alpine_terrain), uniform 30 m mesh", "deep_m": deep}
m, geo, slope, (sd, sdv, swe) = case(side)
h = capi.Handle(capi.default_config(nLayer=2), m)
h.slide_init()
runs = []
for k in range(6):
    t = time.perf_counter()
    o, st = h.slide_run(sd, sdv, swe)
    runs.append((time.perf_counter() - t, st))
st = runs[-1][1]
z = np.zeros(m.n_local)
h.slide_init()
calm = [h.slide_run(z, z, z)[1]["ms_device"] for _ in range(4)]
out["gpu"] = {"triangles": int(m.n_local), "candidates": int(np.count_nonzero(sd > so.max_depth(slope, None))), "faces_fired": st["faces_fired"],
              "wavefront_rounds": st["wavefront_rounds"], "frontier_rounds": st["frontier_rounds"], "live_faces": st["live_faces"], "ms_device": float(np.median([r[1]["ms_device"] for r in runs[1:]])),
              "ms_host_call_with_copies": float(np.median([r[0] for r in runs[1:]]) * 1e3), "calm_ms_device": float(np.median(calm[1:])),
              "moved_m3_water": float(np.abs(o["delta_avalanche_mass"]).sum() / 2)}
h.close()
if chm_ref.available() and cpu_side > 0:
    m2, geo2, slope2, (sd2, sdv2, swe2) = case(cpu_side)
    t = time.perf_counter()
    ref = so.ReferenceSlide(m2.vertex, m2.elem, m2.neigh, None, None)
    t_setup = time.perf_counter() - t
    ts = []
    for k in range(3):
        t = time.perf_counter()
        r = ref.run(sd2, sdv2, swe2)
        ts.append(time.perf_counter() - t)
    out["cpu_reference"] = {"kind": "reference (snow_slide.cpp compiled unmodified; stand-in face store)", "cores": 1, "triangles": int(m2.n_local),
                            "ms_per_run": float(np.median(ts) * 1e3), "setup_s": t_setup,
                            "ms_per_run_scaled_to_gpu_size": float(np.median(ts) * 1e3 * m.n_local / m2.n_local)}
    t = time.perf_counter()
    stt = so.SlideState(m2.face_vertices().reshape(-1, 3, 3), m2.neigh, geo2.area)
    oo = so.run_single(stt, sd2, sdv2, swe2)
    out["numpy_oracle_ms"] = (time.perf_counter() - t) * 1e3
    sc = np.abs(r["delta_avalanche_mass"]).max()
    out["oracle_vs_reference"] = float(np.max(np.abs(oo["delta_avalanche_mass"] - r["delta_avalanche_mass"])) / sc)
print(json.dumps(out))

"""Golden vectors for snow_slide — OUTPUTS OF THE REFERENCE ITSELF: src/modules/snow_slide.cpp compiled
unmodified into oracle/_ref/libchmref.so (oracle/refbuild/Makefile; oneTBB's concurrent_vector / parallel_sort stand-ins in
stubs/tbb/) and driven like CHM drives a module: ctor(config) -> init(mesh) -> run(mesh), twice (the *_sum variables accumulate).
Run HERE, where /root/reference exists:

    python tests/golden/make_golden_slide.py     ->  tests/golden/golden_slide.npz

Cases: granger1m (985 faces; avalache_mult lowered to 900 so that its 35-degree slopes release) and slope (2618 faces) with the
default parameters under a 14 m snow cover and with (500, -1.7) under 5 m; slope with a vegetation height on a third of the faces
(maxDepth = CanopyHeight there); and, for the partition-edge code path, rank 1 of 3 of the slope mesh with its ghost neighbours
attached (USE_MPI is not defined in this build, so the exchanges are compiled out: what is pinned is the rank-local sweep and the
ghost accumulators it leaves behind).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from chm_b200.mesh import partition_mesh  # noqa: E402
from conftest import load_mesh  # noqa: E402
from oracle import chm_ref, slide_oracle as so  # noqa: E402

CASES = [("granger_m900", "granger1m", {"avalache_mult": 900.0}, 6.0, False),
         ("slope_default", "slope", {}, 14.0, False),
         ("slope_custom", "slope", {"avalache_mult": 500.0, "avalache_pow": -1.7}, 5.0, False),
         ("slope_veg", "slope", {}, 9.0, True)]


def main():
    assert chm_ref.build() and chm_ref.available()
    out = {}
    for tag, name, cfg, deep, veg in CASES:
        m = load_mesh(name)
        geo = m.geometry()
        slope = so.face_slope(m.face_vertices().reshape(-1, 3, 3))
        sd, sdv, swe = so.synthetic_snow(geo.cx, geo.cy, slope, seed=3, deep=deep)
        params = None
        if veg:
            rng = np.random.default_rng(17)
            canopy = np.where(rng.random(m.n_local) < 0.33, rng.uniform(0.5, 6.0, m.n_local), 0.0)
            params = {"CanopyHeight": canopy}
            out[f"{tag}_canopy"] = canopy
        ref = so.ReferenceSlide(m.vertex, m.elem, m.neigh, params, cfg)
        out[f"{tag}_cfg"] = np.array([cfg.get("avalache_mult", 3178.4), cfg.get("avalache_pow", -1.998)])
        out[f"{tag}_sd"], out[f"{tag}_sdv"], out[f"{tag}_swe"] = sd, sdv, swe
        for run in (1, 2):
            r = ref.run(sd, sdv, swe)
            for k in so.ReferenceSlide.VARS:
                out[f"{tag}_run{run}_{k}"] = r[k]
        out[f"{tag}_checkpoint"] = ref.checkpoint()
        print(tag, "faces that changed:", int(np.count_nonzero(out[f"{tag}_run1_delta_avalanche_mass"])))
    # rank-local sweep with ghost neighbours
    m = load_mesh("slope")
    geo = m.geometry()
    slope = so.face_slope(m.face_vertices().reshape(-1, 3, 3))
    sd, sdv, swe = so.synthetic_snow(geo.cx, geo.cy, slope, seed=3, deep=14.0)
    p = partition_mesh(m, 1, 3)
    T = p.n_local
    V = p.face_vertices().reshape(-1, 3, 3)
    gids = p.global_id
    ref = so.ReferenceSlide(p.vertex, p.elem[:T], p.neigh, None, None, ghosts=dict(vertices=V[T:], area=geo.area[gids[T:]]))
    r = ref.run(sd[gids[:T]], sdv[gids[:T]], swe[gids[:T]], sdv[gids[T:]])
    for k in so.ReferenceSlide.VARS + so.ReferenceSlide.GHOST_VARS:
        out[f"rank1of3_{k}"] = r[k]
    print("rank1of3 ghosts that received:", int(np.count_nonzero(r["ghost_ss_delta_avalanche_swe"])))
    out["depends"] = np.array(ref.depends())
    out["provides"] = np.array(ref.provides())
    path = os.path.join(ROOT, "tests", "golden", "golden_slide.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

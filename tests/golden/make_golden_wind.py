"""Golden vectors for the two providers of PBSM3D inputs — OUTPUTS OF THE REFERENCE ITSELF:
src/modules/scale_wind_vert.cpp and src/modules/fetchr.cpp compiled unmodified by oracle/refbuild/Makefile into
oracle/_ref/libchmref.so (thin plate spline and kd-tree restated in the stand-in headers, see there), driven on the
reference's bundled meshes.  Run HERE, where /root/reference exists:

    python tests/golden/make_golden_wind.py     ->  tests/golden/golden_wind.npz

Cases (each on granger1m = 985 faces and slope = 2618 faces, real terrain):
  veg    random CanopyHeight (40 % of faces, 0.2-12 m) + LAI, snow depth with -9999 holes and one 60 m "avalanche" face
  bare   no vegetation parameters, no snowdepthavg provider (the optional input absent)
  ignore veg parameters present but ignore_canopy / incl_veg = false, fetchr with steps 7, max_distance 650, I 0.03
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from chm_b200 import synthetic  # noqa: E402
from conftest import load_mesh  # noqa: E402
from oracle import chm_ref  # noqa: E402


def inputs(mesh, seed):
    geo = mesh.geometry()
    T = mesh.n_local
    F = synthetic.forcing(geo.cx, geo.cy, seed=seed)
    rng = np.random.default_rng(seed)
    canopy = np.where(rng.random(T) < 0.4, rng.uniform(0.2, 12.0, T), 0.0)
    lai = rng.uniform(0.3, 3.0, T)
    sd = F["snowdepthavg"].copy()
    sd[::7] = -9999.0
    sd[5] = 60.0
    vw = np.mod(F["vw_dir"] + rng.uniform(-180, 180, T), 360.0)  # every azimuth
    return F["U_R"], sd, vw, canopy, lai


def main():
    chm_ref.build()
    out = {}
    for name, seed in (("granger1m", 21), ("slope", 22)):
        mesh = load_mesh(name)
        U_R, sd, vw, canopy, lai = inputs(mesh, seed)
        for k, v in (("U_R", U_R), ("sd", sd), ("vw_dir", vw), ("canopy", canopy), ("lai", lai)):
            out[f"{name}_{k}"] = v
        veg = chm_ref.ReferencePBSM3D(mesh.vertex, mesh.elem, mesh.neigh, {"CanopyHeight": canopy, "LAI": lai}, {"nLayer": 5})
        bare = chm_ref.ReferencePBSM3D(mesh.vertex, mesh.elem, mesh.neigh, None, {"nLayer": 5})
        out[f"{name}_veg_u2"] = veg.scale_wind_vert(U_R, sd)
        out[f"{name}_veg_u2_point"] = veg.scale_wind_vert(U_R, sd, point_only=True)
        out[f"{name}_veg_fetch"] = veg.fetchr(vw)
        out[f"{name}_bare_u2"] = bare.scale_wind_vert(U_R, None)
        out[f"{name}_bare_u2_point"] = bare.scale_wind_vert(U_R, None, point_only=True)
        out[f"{name}_bare_fetch"] = bare.fetchr(vw)
        out[f"{name}_ignore_u2"] = veg.scale_wind_vert(U_R, sd, cfg={"ignore_canopy": True})
        out[f"{name}_ignore_fetch"] = veg.fetchr(vw, cfg={"incl_veg": False, "steps": 7, "max_distance": 650.0, "I": 0.03})
        for c in ("veg", "bare", "ignore"):
            f = out[f"{name}_{c}_fetch"]
            print(name, c, "fetch values", dict(zip(*np.unique(f, return_counts=True))))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_wind.npz"), **out)


if __name__ == "__main__":
    main()

"""Generate the committed fixtures under tests/golden/ (run HERE, where /root/reference exists).

1. Mesh fixtures: the reference's bundled meshes, re-encoded as npz (vertex/elem/neigh/params in CHM face
   order after the cell_global_id permutation) because /root/reference does not exist on the GPU box.
     granger1m.npz  <- test_data/meshes/granger1m.mesh + .param           (985 triangles)
     slope.npz      <- functional_tests/mesh_versioning/slope.mesh + .param (2618 triangles)
     slope_metis.npz<- functional_tests/mesh_versioning/slope.metis.mesh   (METIS permutation + 31 local_size)
2. Oracle golden vectors (the reference pins nothing for PBSM3D — SURVEY §8c): outputs of
   oracle/pbsm3d_oracle.py with the direct solver for 24 steps on granger1m (nLayer=5, code defaults) and
   3 steps on slope (nLayer=10, functional-test block).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from chm_b200 import synthetic  # noqa: E402
from chm_b200.mesh import read_chm_mesh  # noqa: E402
from oracle.pbsm3d_oracle import Config, PBSM3DOracle  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def save_mesh(name, mesh):
    d = {"vertex": mesh.vertex, "elem": mesh.elem, "neigh": mesh.neigh}
    for k, v in mesh.params.items():
        d["param_" + k] = v
    if mesh.local_sizes is not None:
        d["local_sizes"] = mesh.local_sizes
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def run_golden(name, mesh, cfg, nsteps, keep_c_steps=(0,)):
    geo = mesh.geometry()
    o = PBSM3DOracle(cfg, mesh.neigh, geo, mesh.global_id, mesh.n_global, mesh.params)
    out = {}
    for k in range(nsteps):
        calm = (k % 8 == 5)  # a calm hour now and then: early-out path + stale drift_mass
        F = synthetic.forcing(geo.cx, geo.cy, seed=7, step=k, calm=calm)
        r = o.step(F, 3600.0, solver="direct")
        for v in ("Qsusp", "Qsalt", "Qsubl", "drift_mass", "sum_drift", "sum_subl"):
            out[f"{v}_{k}"] = r[v]
        out[f"present_{k}"] = np.array([r["suspension_present"], r["deposition_present"]], dtype=np.int8)
        if k in keep_c_steps:
            out[f"c_{k}"] = r["c"]
            a = r["asm"]
            out[f"diag_{k}"], out[f"lat_{k}"], out[f"below_{k}"], out[f"above_{k}"] = a.diag, a.lat, a.below, a.above
            out[f"rhs_{k}"] = a.rhs[0]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


if __name__ == "__main__":
    g = read_chm_mesh(f"{REF}/test_data/meshes/granger1m.mesh", [f"{REF}/test_data/meshes/granger1m.param"])
    save_mesh("granger1m", g)
    s = read_chm_mesh(f"{REF}/functional_tests/mesh_versioning/slope.mesh", [f"{REF}/functional_tests/mesh_versioning/slope.param"])
    save_mesh("slope", s)
    sm = read_chm_mesh(f"{REF}/functional_tests/mesh_versioning/slope.metis.mesh", [])
    save_mesh("slope_metis", sm)
    run_golden("golden_granger1m_L5_default", g, Config(nLayer=5), 24)
    run_golden("golden_slope_L10_functest", s, Config.functional_test(10), 3)
    print("ok")

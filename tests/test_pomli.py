"""use_PomLi_probability (PBSM3D.cpp:848-866): Pomeroy & Li's upscaled probability of blowing snow scales the saltation
concentration and is provided as `blowingsnow_probability`.  Vectors: tests/golden/golden_pomli.npz, outputs of the reference's
own PBSM3D.cpp (oracle/_ref) on granger1m with 20 % part-buried shrubs, stalk and R94 vegetation, 3 steps (the middle one calm:
the face variable keeps its previous value)."""
import os

import numpy as np
import pytest

from chm_b200 import capi, synthetic
from chm_b200.mesh import TriMesh
from conftest import GOLDEN, load_mesh, rel_l2
from oracle.pbsm3d_oracle import Config, PBSM3DOracle

CASES = {"pomli_stalks": False, "pomli_R94": True}


def setup(name):
    g = np.load(os.path.join(GOLDEN, "golden_pomli.npz"))
    base = load_mesh("granger1m")
    mesh = TriMesh(base.vertex, base.elem, base.neigh, dict(base.params, **synthetic.shrub_params(base.n_local, canopy=0.45)))
    geo = mesh.geometry()
    forc = []
    for k in range(3):
        F = synthetic.forcing(geo.cx, geo.cy, seed=3, step=k, calm=(k == 1))
        F["p_snow_hours"] = g[f"{name}/p_snow_hours_{k}"]
        forc.append(F)
    return g, mesh, geo, forc


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_the_reference(name):
    g, mesh, geo, forc = setup(name)
    o = PBSM3DOracle(Config(nLayer=5, use_PomLi_probability=True, use_R94_lambda=CASES[name]), mesh.neigh, geo, mesh.global_id,
                     mesh.n_global, mesh.params)
    for k, F in enumerate(forc):
        r = o.step(F, 3600.0)
        ref_p = g[f"{name}/blowingsnow_probability_{k}"]
        assert np.array_equal(r["blowingsnow_probability"] == -9999.0, ref_p == -9999.0)
        assert np.max(np.abs(r["blowingsnow_probability"] - ref_p)) <= 1e-14
        for v in ("Qsalt", "Qsusp", "drift_mass", "sum_drift"):
            assert rel_l2(r[v], g[f"{name}/{v}_{k}"]) <= 1e-11, (v, k)
    # not a no-op: the probability moves Qsalt, and the two vegetation models differ
    o2 = PBSM3DOracle(Config(nLayer=5, use_R94_lambda=CASES[name]), mesh.neigh, geo, mesh.global_id, mesh.n_global, mesh.params)
    assert rel_l2(o2.step(forc[0], 3600.0)["Qsalt"], g[f"{name}/Qsalt_0"]) > 0.05
    assert rel_l2(g["pomli_stalks/Qsalt_0"], g["pomli_R94/Qsalt_0"]) > 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_matches_the_reference(name):
    g, mesh, geo, forc = setup(name)
    h = capi.Handle(capi.default_config(nLayer=5, use_PomLi_probability=1, use_R94_lambda=int(CASES[name])), mesh)
    for k, F in enumerate(forc):
        outs, st = h.step(3600.0, F)
        ref_p = g[f"{name}/blowingsnow_probability_{k}"]
        assert np.array_equal(outs["blowingsnow_probability"] == -9999.0, ref_p == -9999.0)  # unset / kept on non-saltating faces
        assert np.max(np.abs(outs["blowingsnow_probability"] - ref_p)) <= 1e-13
        assert [st["suspension_present"], st["deposition_present"]] == list(g[f"{name}/present_{k}"])
        for v in ("Qsalt", "Qsusp", "Qsubl", "drift_mass", "sum_drift"):
            assert rel_l2(outs[v], g[f"{name}/{v}_{k}"]) <= 1e-6, (v, k)
        if k == 0:
            s = h.suspension_system()
            assert np.max(np.abs(s["rhs0"] - g[f"{name}/rhs_0"]) / np.maximum(np.abs(g[f"{name}/rhs_0"]), 1e-300)) <= 1e-12
    h.close()


@pytest.mark.gpu
def test_p_snow_hours_is_required_only_with_the_option():
    mesh = load_mesh("granger1m")
    geo = mesh.geometry()
    F = synthetic.forcing(geo.cx, geo.cy)
    h = capi.Handle(capi.default_config(nLayer=5, use_PomLi_probability=1), mesh)
    with pytest.raises(capi.Pbsm3dError, match="forcing array missing"):
        h.step(3600.0, F)
    h.close()
    h = capi.Handle(capi.default_config(nLayer=5), mesh)
    outs, _ = h.step(3600.0, F)
    assert "blowingsnow_probability" not in outs
    h.close()

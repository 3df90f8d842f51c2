"""Two independent restatements of the reference must agree: oracle/pbsm3d_oracle.py (numpy + exact direct solve)
and oracle/pbsm3d_ref.cpp (C++/OpenMP, GMRES(30) + rank-local ILUT — also the CPU baseline of bench.py)."""
import numpy as np
import pytest

from chm_b200 import synthetic
from chm_b200.mesh import TriMesh
from conftest import max_rel, rel_l2
from oracle.cpu_ref import CpuReference, host_threads
from oracle.pbsm3d_oracle import Config, PBSM3DOracle


@pytest.mark.parametrize("cfg", [Config(nLayer=5), Config.functional_test(10),
                                 Config(nLayer=4, rouault_diffusion_coef=True, do_lateral_diff=False, use_exp_fetch=True, use_tanh_fetch=False)],
                         ids=["default_L5", "functest_L10", "rouault_expfetch"])
def test_cpp_reference_matches_numpy_oracle(slope, cfg):
    geo = slope.geometry()
    o = PBSM3DOracle(cfg, slope.neigh, geo, slope.global_id, slope.n_global, slope.params)
    c = CpuReference(cfg, slope, geo, dump_system=True)
    for k in range(3):
        F = synthetic.forcing(geo.cx, geo.cy, step=k, calm=(k == 1), fetch_const=None)
        r = o.step(F, 3600.0)
        rc = c.step(F, 3600.0, tolerance=1e-12)
        s, a = rc["system"], r["asm"]
        for name, ref in (("diag", a.diag), ("below", a.below), ("above", a.above), ("u_z", a.u_z), ("csubl", a.csubl)):
            assert max_rel(s[name], ref) <= 1e-12, name
        assert np.max(np.abs(s["lat"] - a.lat) / np.abs(a.diag)[None]) <= 1e-12
        assert max_rel(s["rhs0"], a.rhs[0]) <= 1e-12 and max_rel(s["c_salt"], a.c_salt) <= 1e-12
        assert np.array_equal(s["saltation"].astype(bool), a.saltation)
        assert bool(rc["stats"]["susp_present"]) == bool(r["suspension_present"])
        for v in ("c", "Qsusp", "Qsalt", "Qsubl", "drift_mass", "sum_drift", "sum_subl"):
            assert rel_l2(rc[v], r[v]) <= 1e-8, (v, k)


def test_cpp_reference_at_reference_tolerance_and_veg(slope):
    params = dict(slope.params, **synthetic.shrub_params(slope.n_local))
    mesh = TriMesh(slope.vertex, slope.elem, slope.neigh, params)
    geo = mesh.geometry()
    water = np.arange(mesh.n_local) % 13 == 0
    cfg = Config.functional_test(10)
    F = synthetic.forcing(geo.cx, geo.cy, seed=9)
    r = PBSM3DOracle(cfg, mesh.neigh, geo, mesh.global_id, mesh.n_global, params, water).step(F, 3600.0)
    for nth in (1, 3):
        rc = CpuReference(cfg, mesh, geo, is_water=water, n_threads=nth).step(F, 3600.0)
        st = rc["stats"]
        assert st["n_threads"] == nth and 0 < st["susp_iters"] <= 200 and st["susp_resid"] <= 1e-8
        assert rel_l2(rc["c"], r["c"]) <= 1e-6 and rel_l2(rc["drift_mass"], r["drift_mass"]) <= 1e-6
    assert host_threads() >= 1

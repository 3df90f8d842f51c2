"""Pins the oracle (and, with -m gpu, the CUDA path) to OUTPUTS OF THE REFERENCE ITSELF.

tests/golden/golden_*.npz are produced by the reference's own src/modules/PBSM3D.cpp, Atmosphere.cpp and coordinates.cpp,
compiled unmodified by oracle/refbuild/Makefile (stand-in headers replace the libraries this image lacks) and driven by
tests/golden/make_golden.py.  Three layers:
  1. CPU: oracle/pbsm3d_oracle.py against the committed reference vectors, every supported config variant.
  2. CPU, when oracle/_ref/libchmref.so is present (it travels to the GPU box; /root/reference does not): the live
     library reproduces the committed vectors, and agrees with the oracle on fresh random cases.
  3. GPU: the CUDA path through the C-ABI against the same reference vectors.
Tolerances: assembled coefficients <= 1e-12 relative (numpy/libm vs glibc pow/exp/log rounding), exact-solve outputs
<= 1e-11; CUDA at the reference's solver tolerance 1e-8 <= 1e-6 relative L2 (north_star), <= 1e-8 at tolerance 1e-11.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

from chm_b200 import synthetic
from chm_b200.mesh import TriMesh
from conftest import GOLDEN, load_mesh, rel_l2
from make_golden import cfg_dict, variant_cases  # the generator's own case table
from oracle import chm_ref
from oracle.pbsm3d_oracle import (Config, PBSM3DOracle, bearing_to_cartesian, log_scale_wind,
                                  saturated_vapour_pressure)

OUT = ("Qsusp", "Qsalt", "Qsubl", "Qsubl_mass", "drift_mass", "sum_drift", "sum_subl")
N = 985  # granger1m
CASES = variant_cases(N)
needs_ref = pytest.mark.skipif(not chm_ref.available(), reason="oracle/_ref/libchmref.so not built (needs /root/reference)")


def is_water_of(extra, table):
    if "landcover" not in extra:
        return None
    return np.array([bool(table.get(f"landcover.{int(c)}.is_water", False)) for c in extra["landcover"]])


def oracle_case(granger, name):
    cfg, extra, table, tweak = CASES[name]
    params = dict(granger.params, **extra)
    mesh = TriMesh(granger.vertex, granger.elem, granger.neigh, params)
    o = PBSM3DOracle(cfg, mesh.neigh, mesh.geometry(), mesh.global_id, mesh.n_global, params, is_water_of(extra, table))
    return mesh, o, tweak


def check_assembly_arrays(get, g, tag, tol):
    """get(key) -> array of the implementation under test; g = golden npz."""
    diag = g[f"{tag}diag_0"]
    assert np.max(np.abs(get("diag") - diag) / np.abs(diag)) <= tol
    assert np.max(np.abs(get("lat") - g[f"{tag}lat_0"]) / np.abs(diag)[None]) <= tol
    rhs = g[f"{tag}rhs_0"]
    assert np.max(np.abs(get("rhs0") - rhs)) <= tol * max(np.abs(rhs).max(), 1e-300)


# ----------------------------------------------------------------------------------------------- 1. oracle vs reference vectors
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_vectors(granger, name):
    g = np.load(os.path.join(GOLDEN, "golden_variants.npz"))
    mesh, o, tweak = oracle_case(granger, name)
    geo = mesh.geometry()
    tag = name + "/"
    for k in range(2):
        F = tweak(synthetic.forcing(geo.cx, geo.cy, seed=3, step=k))
        r = o.step(F, 3600.0)
        assert [int(r["suspension_present"]), int(r["deposition_present"])] == g[f"{tag}present_{k}"].tolist()
        for v in OUT:
            assert rel_l2(r[v], g[f"{tag}{v}_{k}"]) <= 1e-11, (v, k)
        # the reference only ever writes 1 into pbsm_more_than_avail (PBSM3D.cpp:1727); elsewhere the store default stays
        assert np.array_equal(r["pbsm_more_than_avail"] == 1, g[f"{tag}pbsm_more_than_avail_{k}"] == 1)
        if k == 0:
            a = r["asm"]
            check_assembly_arrays({"diag": a.diag, "lat": a.lat, "rhs0": a.rhs[0]}.__getitem__, g, tag, 1e-13)
            assert rel_l2(r["c"], g[f"{tag}c_0"]) <= 1e-11
            assert rel_l2(r["dep"][2], g[f"{tag}dep_rhs_0"]) <= 1e-11


@pytest.mark.parametrize("golden,meshname,cfg", [
    ("golden_granger1m_L5_default", "granger1m", Config(nLayer=5)),
    ("golden_slope_L10_functest", "slope", Config.functional_test(10)),
])
def test_oracle_systems_match_reference_entry_by_entry(golden, meshname, cfg):
    """Every coefficient of both linear systems of step 0, in the reference's own numbering."""
    mesh = load_mesh(meshname)
    g = np.load(os.path.join(GOLDEN, golden + ".npz"))
    geo = mesh.geometry()
    o = PBSM3DOracle(cfg, mesh.neigh, geo, mesh.global_id, mesh.n_global, mesh.params)
    r = o.step(synthetic.forcing(geo.cx, geo.cy, seed=7, step=0), 3600.0)
    a = r["asm"]
    scale = np.abs(g["diag_0"])
    for k, mine in (("diag", a.diag), ("below", a.below), ("above", a.above)):
        assert np.max(np.abs(mine - g[k + "_0"]) / scale) <= 1e-13, k
    assert np.max(np.abs(a.lat - g["lat_0"]) / scale[None]) <= 1e-13
    assert rel_l2(a.rhs[0], g["rhs_0"]) <= 1e-13
    d, off, drhs = r["dep"]
    assert np.max(np.abs(d - g["dep_diag_0"]) / g["dep_diag_0"]) <= 1e-14
    assert np.max(np.abs(off - g["dep_off_0"])) <= 1e-14 * g["dep_diag_0"].max()
    assert rel_l2(drhs, g["dep_rhs_0"]) <= 1e-11 and rel_l2(r["q_dep"], g["q_dep_0"]) <= 1e-11
    # module contract (PBSM3D.cpp:103-219)
    assert g["depends"].tolist() == ["U_2m_above_srf", "vw_dir", "swe", "t", "rh", "U_R", "fetch"]
    assert g["provides"].tolist() == ["pbsm_more_than_avail", "global_cell_id", "blowingsnow_probability", "Qsubl",
                                      "Qsubl_mass", "sum_subl", "drift_mass", "Qsusp", "Qsalt", "sum_drift"]


def test_helpers_match_reference_vectors():
    """Atmosphere::log_scale_wind / saturatedVapourPressure (Atmosphere.cpp:32-38,62-80) and
    math::gis::bearing_to_cartesian (coordinates.cpp:112-131) as compiled from the reference."""
    g = np.load(os.path.join(GOLDEN, "golden_helpers.npz"))
    assert np.max(np.abs(log_scale_wind(g["u"], 50.0, g["zout"], g["sd"], g["z0"]) / g["log_scale_wind"] - 1)) <= 4e-16
    assert np.max(np.abs(saturated_vapour_pressure(g["tk"]) / g["es"] - 1)) <= 4e-16
    bx, by = bearing_to_cartesian(g["bearing"])
    assert np.max(np.abs(bx - g["bx"])) <= 2e-16 and np.max(np.abs(by - g["by"])) <= 2e-16


def test_module_twin_declares_the_reference_contract():
    from chm_b200.module import PBSM3D
    g = np.load(os.path.join(GOLDEN, "golden_granger1m_L5_default.npz"))
    m = PBSM3D({"nLayer": 5})
    assert m._depends == g["depends"].tolist() and m._provides == g["provides"].tolist()


# ----------------------------------------------------------------------------------------------- 2. live reference library
@needs_ref
def test_live_reference_reproduces_committed_vectors(granger):
    g = np.load(os.path.join(GOLDEN, "golden_granger1m_L5_default.npz"))
    geo = granger.geometry()
    ref = chm_ref.ReferencePBSM3D(granger.vertex, granger.elem, granger.neigh, granger.params, cfg_dict(Config(nLayer=5)))
    for k in range(7):  # includes the calm hour k = 5
        r = ref.step(synthetic.forcing(geo.cx, geo.cy, seed=7, step=k, calm=(k % 8 == 5)), 3600.0)
        for v in OUT:
            assert rel_l2(r[v], g[f"{v}_{k}"]) <= 1e-13, (v, k)  # the direct solve may differ by a rounding
    ref.close()


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_live_reference_equals_oracle_on_random_cases(seed):
    """Fresh meshes, forcing and configs the fixtures have not seen."""
    rng = np.random.default_rng(100 + seed)
    mesh = synthetic.variable_mesh(1500 + 400 * seed, seed=seed) if seed % 2 else synthetic.uniform_mesh(17 + seed, 23)
    n = mesh.n_local
    params = synthetic.shrub_params(n, frac=0.3, seed=seed, canopy=float(rng.uniform(0.5, 1.5)))
    cfg = Config(nLayer=int(rng.integers(2, 9)), do_fixed_settling=bool(seed & 1), settling_velocity=float(rng.uniform(0.2, 0.8)),
                 do_sublimation=bool(seed & 2) or True, do_lateral_diff=not bool(seed & 2), smooth_coeff=float(int(rng.uniform(500, 7000))),
                 min_sd_trans=float(rng.uniform(0.05, 0.4)), cutoff=float(rng.uniform(0.1, 0.6)),
                 snow_diffusion_const=float(rng.uniform(0.2, 1.0)), rouault_diffusion_coef=(seed == 3),
                 use_R94_lambda=bool(seed % 2), use_exp_fetch=(seed == 2), use_tanh_fetch=(seed != 2))
    geo = mesh.geometry()
    o = PBSM3DOracle(cfg, mesh.neigh, geo, mesh.global_id, mesh.n_global, params)
    ref = chm_ref.ReferencePBSM3D(mesh.vertex, mesh.elem, mesh.neigh, params, cfg_dict(cfg))
    for k in range(3):
        F = synthetic.forcing(geo.cx, geo.cy, seed=50 + seed, step=k, calm=(k == 1 and seed == 0), fetch_const=None)
        a, b = o.step(F, 3600.0), ref.step(F, 3600.0)
        A = o.suspension_csr(a["asm"])
        assert abs(A - b["susp"][0]).max() <= 1e-13 * abs(b["susp"][0]).max()
        assert rel_l2(a["asm"].rhs.reshape(-1), b["susp"][1]) <= 1e-13
        for v in OUT + ("c", "q_dep"):
            assert rel_l2(a[v], b[v]) <= 1e-11, (v, k)
    # checkpoint()/load_checkpoint() of the reference (PBSM3D.cpp:1753-1773) carry exactly sum_drift
    assert np.array_equal(ref.checkpoint(), ref.get_var("sum_drift"))
    ref.load_checkpoint(np.arange(n, dtype=float))
    assert np.array_equal(ref.get_var("sum_drift"), np.arange(n, dtype=float))
    ref.close()


@needs_ref
def test_reference_config_parsing_quirks(granger):
    """cfg.get("smooth_coeff", 820) is an int read (PBSM3D.cpp:247): a fractional value falls back to the default 820."""
    geo = granger.geometry()
    F = synthetic.forcing(geo.cx, geo.cy)
    d = cfg_dict(Config(nLayer=5))
    d["smooth_coeff"] = 6500.5
    frac = chm_ref.ReferencePBSM3D(granger.vertex, granger.elem, granger.neigh, granger.params, d).step(F, 3600.0)
    d["smooth_coeff"] = 820
    dflt = chm_ref.ReferencePBSM3D(granger.vertex, granger.elem, granger.neigh, granger.params, d).step(F, 3600.0)
    assert np.array_equal(frac["drift_mass"], dflt["drift_mass"])
    with pytest.raises(RuntimeError, match="Cannot specify both"):
        chm_ref.ReferencePBSM3D(granger.vertex, granger.elem, granger.neigh, granger.params,
                                dict(cfg_dict(Config(nLayer=5)), use_exp_fetch=True, use_tanh_fetch=True))


# ----------------------------------------------------------------------------------------------- 3. CUDA vs reference vectors
def _kw(cfg: Config):
    d = cfg_dict(cfg)
    drop = ("iterative_subl", "use_PomLi_probability", "z0_ustar_coupling", "use_subgrid_topo", "use_subgrid_topo_V2", "debug_output")
    return {k: (int(v) if isinstance(v, bool) else v) for k, v in d.items() if k not in drop}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("tol,bar", [(1e-8, 1e-6), (1e-11, 1e-8)], ids=["ref_tol", "tight"])
def test_cuda_matches_reference_vectors(granger, name, tol, bar):
    from chm_b200 import capi
    g = np.load(os.path.join(GOLDEN, "golden_variants.npz"))
    cfg, extra, table, tweak = CASES[name]
    mesh = TriMesh(granger.vertex, granger.elem, granger.neigh, dict(granger.params, **extra))
    geo = mesh.geometry()
    h = capi.Handle(capi.default_config(tolerance=tol, **_kw(cfg)), mesh, is_water=is_water_of(extra, table))
    tag = name + "/"
    for k in range(2):
        F = tweak(synthetic.forcing(geo.cx, geo.cy, seed=3, step=k))
        outs, st = h.step(3600.0, F)
        assert [st["suspension_present"], st["deposition_present"]] == g[f"{tag}present_{k}"].tolist()
        for v in ("Qsusp", "Qsalt", "Qsubl", "drift_mass", "sum_drift", "sum_subl"):
            assert rel_l2(outs[v], g[f"{tag}{v}_{k}"]) <= bar, (v, k)
        assert np.array_equal(outs["pbsm_more_than_avail"] == 1, g[f"{tag}pbsm_more_than_avail_{k}"] == 1)
        if k == 0:
            s = h.suspension_system()
            if name != "missing_values":  # swe == 0 turns the availability test into a test on rounding noise
                check_assembly_arrays(s.__getitem__, g, tag, 1e-12)
            assert rel_l2(h.solution(), g[f"{tag}c_0"]) <= bar
    h.close()

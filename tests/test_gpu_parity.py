"""Parity of the CUDA path against the oracle, through the C-ABI (run with -m gpu on the B200 box).

Tolerances (BASELINE.md §5 / north_star): face ordering, neighbour indexing and geometry bit-exact; assembled
coefficients ≤ 1e-12 relative; suspended concentration, fluxes and drift ≤ 1e-6 relative L2 (fp64), against the
oracle's exact sparse direct solve.
"""
import os

import numpy as np
import pytest

from chm_b200 import capi, synthetic
from chm_b200.mesh import TriMesh
from chm_b200.module import Domain, PBSM3D, module_error
from conftest import GOLDEN, functest_kw, load_mesh, max_rel, rel_l2
from oracle.pbsm3d_oracle import Config, PBSM3DOracle

pytestmark = pytest.mark.gpu

ASM_TOL = 1e-12
L2_TOL = 1e-6


def oracle_for(mesh, cfg, is_water=None):
    return PBSM3DOracle(cfg, mesh.neigh, mesh.geometry(), mesh.global_id, mesh.n_global, mesh.params, is_water)


def check_assembly(h, asm):
    s = h.suspension_system()
    for k, ref in (("diag", asm.diag), ("below", asm.below), ("above", asm.above), ("u_z", asm.u_z), ("csubl", asm.csubl)):
        assert max_rel(s[k], ref) <= ASM_TOL, k
    # lateral coefficients: some are sums of a large and a tiny term; compare against the row scale
    scale = np.abs(asm.diag)[None]
    assert np.max(np.abs(s["lat"] - asm.lat) / scale) <= ASM_TOL
    assert max_rel(s["rhs0"], asm.rhs[0]) <= ASM_TOL
    assert max_rel(s["c_salt"], asm.c_salt) <= ASM_TOL
    assert np.array_equal(s["saltation"].astype(bool), asm.saltation)


@pytest.mark.parametrize("meshname", ["granger1m", "slope", "slope_metis"])
def test_geometry_and_indexing_bit_exact(meshname):
    mesh = load_mesh(meshname)
    h = capi.Handle(capi.default_config(nLayer=5), mesh)
    g, ref = h.geometry(), mesh.geometry()
    for k in ("nx", "ny", "elen", "area", "dx"):
        assert np.array_equal(g[k], getattr(ref, k)), k
    for k in ("cx", "cy", "cz"):
        assert np.array_equal(g[k], getattr(ref, k)[:mesh.n_local]), k
    h.close()


def test_geometry_without_area_parameter_bit_exact():
    mesh = synthetic.variable_mesh(6000)
    h = capi.Handle(capi.default_config(nLayer=4), mesh)
    g, ref = h.geometry(), mesh.geometry()
    for k in ("nx", "ny", "elen", "area", "dx"):
        assert np.array_equal(g[k], getattr(ref, k)), k
    h.close()


CASES = [
    ("default_L5", Config(nLayer=5), dict(nLayer=5)),
    ("functest_L10", Config.functional_test(10), functest_kw(10)),
    ("generic_L7", Config(nLayer=7), dict(nLayer=7)),  # runtime-L sweep kernel
    ("L2_min", Config(nLayer=2), dict(nLayer=2)),
    ("L20", Config.functional_test(20), functest_kw(20)),
    ("no_subl_no_latdiff", Config(nLayer=5, do_sublimation=False, do_lateral_diff=False),
     dict(nLayer=5, do_sublimation=0, do_lateral_diff=0)),
    ("rouault", Config(nLayer=5, rouault_diffusion_coef=True), dict(nLayer=5, rouault_diffusion_coef=1)),
    ("exp_fetch", Config(nLayer=5, use_exp_fetch=True, use_tanh_fetch=False), dict(nLayer=5, use_exp_fetch=1, use_tanh_fetch=0)),
    ("no_fetch", Config(nLayer=5, use_tanh_fetch=False), dict(nLayer=5, use_tanh_fetch=0)),
]


@pytest.mark.parametrize("name,ocfg,kw", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("solver", [capi.SOLVER_LINE, capi.SOLVER_BICGSTAB], ids=["line", "bicgstab"])
def test_step_matches_oracle(granger, name, ocfg, kw, solver):
    geo = granger.geometry()
    F = synthetic.forcing(geo.cx, geo.cy, seed=3, fetch_const=None if "fetch" in name else 1000.0)
    r = oracle_for(granger, ocfg).step(F, 3600.0, solver="direct")
    h = capi.Handle(capi.default_config(solver=solver, tolerance=1e-11, **kw), granger)
    outs, st = h.step(3600.0, F)
    check_assembly(h, r["asm"])
    assert st["suspension_present"] == 1 and st["deposition_present"] == 1
    assert st["suspension_solver_used"] == solver
    assert rel_l2(h.solution(), r["c"]) <= 1e-9
    for v in ("Qsusp", "Qsalt", "Qsubl", "drift_mass", "sum_drift", "sum_subl"):
        assert rel_l2(outs[v], r[v]) <= 1e-8, v
    d = h.deposition_system()
    assert max_rel(d["diag"], r["dep"][0]) <= ASM_TOL and max_rel(d["off"], r["dep"][1]) <= ASM_TOL
    assert np.max(np.abs(d["rhs"] - r["dep"][2])) <= 1e-7 * np.abs(r["dep"][2]).max()  # carries the solve error of Qsusp
    h.close()


def test_reference_tolerance_meets_the_parity_bar(slope):
    """At the reference's own stopping rule (1e-8) the GPU solution is within 1e-6 of the exact solve."""
    geo = slope.geometry()
    F = synthetic.forcing(geo.cx, geo.cy)
    r = oracle_for(slope, Config.functional_test(10)).step(F, 3600.0)
    for solver in (capi.SOLVER_AUTO, capi.SOLVER_BICGSTAB):
        h = capi.Handle(capi.default_config(solver=solver, **functest_kw(10)), slope)
        outs, st = h.step(3600.0, F)
        assert st["suspension_residual"] <= 1e-8 and st["deposition_residual"] <= 1e-8
        assert rel_l2(h.solution(), r["c"]) <= L2_TOL
        for v in ("Qsusp", "Qsalt", "drift_mass"):
            assert rel_l2(outs[v], r[v]) <= L2_TOL, v
        h.close()


@pytest.mark.parametrize("golden,meshname,cfgd,nsteps", [
    ("golden_granger1m_L5_default", "granger1m", {"nLayer": "5"}, 24),
    ("golden_slope_L10_functest", "slope", {"nLayer": "10", "smooth_coeff": "6500", "do_fixed_settling": "true",
                                            "settling_velocity": "0.5", "use_R94_lambda": "false"}, 3),
])
def test_golden_sequence_through_module_interface(golden, meshname, cfgd, nsteps):
    """Config c1 shape: 24 hourly steps on the bundled mesh, driven like CHM drives a module
    (ctor(config) → init(mesh) → run(mesh) per step), compared with the committed golden vectors."""
    mesh = load_mesh(meshname)
    g = np.load(os.path.join(GOLDEN, golden + ".npz"))
    mod = PBSM3D(cfgd)
    dom = Domain(mesh, dt=3600.0)
    mod.init(dom)
    geo = mesh.geometry()
    for k in range(nsteps):
        F = synthetic.forcing(geo.cx, geo.cy, seed=7, step=k, calm=(k % 8 == 5))
        for name, val in F.items():
            dom[name] = val
        st = mod.run(dom)
        assert [st["suspension_present"], st["deposition_present"]] == g[f"present_{k}"].tolist()
        for v in ("Qsusp", "Qsalt", "Qsubl", "drift_mass", "sum_drift", "sum_subl"):
            assert rel_l2(dom[v], g[f"{v}_{k}"]) <= L2_TOL, (v, k)
    assert rel_l2(mod.handle.solution(), np.zeros_like(g["c_0"])) >= 0  # solution view is readable
    chk = mod.checkpoint(dom)
    assert np.array_equal(chk["PBSM3D:sum_drift"], dom["sum_drift"])
    mod.close()


def test_calm_step_keeps_stale_drift_mass(granger):
    geo = granger.geometry()
    h = capi.Handle(capi.default_config(nLayer=5), granger)
    o0, s0 = h.step(3600.0, synthetic.forcing(geo.cx, geo.cy, step=0))
    o1, s1 = h.step(3600.0, synthetic.forcing(geo.cx, geo.cy, step=1, calm=True))
    assert s1["suspension_present"] == 0 and s1["deposition_present"] == 0 and s1["suspension_iterations"] == 0
    assert np.array_equal(o1["drift_mass"], o0["drift_mass"]) and np.array_equal(o1["sum_drift"], o0["sum_drift"])
    assert not o1["Qsusp"].any() and not o1["Qsalt"].any() and not h.solution().any()
    fresh = capi.Handle(capi.default_config(nLayer=5), granger)
    o, s = fresh.step(3600.0, synthetic.forcing(geo.cx, geo.cy, step=1, calm=True))
    assert (o["drift_mass"] == -9999.0).all()  # never written: the face-store default
    h.close()
    fresh.close()


@pytest.mark.parametrize("r94", [True, False], ids=["R94_LAI", "stalks"])
def test_vegetation_and_water(slope, r94):
    params = dict(slope.params, **synthetic.shrub_params(slope.n_local))
    mesh = TriMesh(slope.vertex, slope.elem, slope.neigh, params)
    water = (np.arange(mesh.n_local) % 17 == 0)
    ocfg = Config(nLayer=5, use_R94_lambda=r94)
    geo = mesh.geometry()
    F = synthetic.forcing(geo.cx, geo.cy, seed=5)
    r = oracle_for(mesh, ocfg, water).step(F, 3600.0)
    h = capi.Handle(capi.default_config(nLayer=5, use_R94_lambda=int(r94), tolerance=1e-11), mesh, is_water=water)
    outs, st = h.step(3600.0, F)
    check_assembly(h, r["asm"])
    assert not r["asm"].saltation[water].any()
    assert (r["asm"].u_z == 0.01).any()  # some layers sit inside the canopy
    for v in ("Qsusp", "Qsalt", "drift_mass"):
        assert rel_l2(outs[v], r[v]) <= 1e-8, v
    h.close()


def test_missing_values_follow_chm_sentinels(granger):
    geo = granger.geometry()
    F = synthetic.forcing(geo.cx, geo.cy)
    F["snowdepthavg"] = np.where(np.arange(granger.n_local) % 3 == 0, -9999.0, F["snowdepthavg"])
    F["swe"] = np.where(np.arange(granger.n_local) % 5 == 0, np.nan, F["swe"])
    r = oracle_for(granger, Config(nLayer=5)).step(F, 3600.0)
    h = capi.Handle(capi.default_config(nLayer=5, tolerance=1e-11), granger)
    outs, st = h.step(3600.0, F)
    s = h.suspension_system()
    # swe == 0 makes the reference's availability test a test on rounding noise: compare the
    # faces where it cannot fire, and the flag everywhere
    assert np.array_equal(s["saltation"].astype(bool), r["asm"].saltation)
    ok = ~np.isnan(F["swe"])
    assert max_rel(s["c_salt"][ok], r["asm"].c_salt[ok]) <= ASM_TOL
    assert max_rel(s["diag"], r["asm"].diag) <= ASM_TOL
    h.close()


def test_state_roundtrip_and_checkpoint(granger):
    geo = granger.geometry()
    F = synthetic.forcing(geo.cx, geo.cy)
    a = capi.Handle(capi.default_config(nLayer=5), granger)
    a.step(3600.0, F)
    st = a.get_state()
    b = capi.Handle(capi.default_config(nLayer=5), granger)
    b.set_state(**st)
    oa, _ = a.step(3600.0, F)
    ob, _ = b.step(3600.0, F)
    # the restored handle schedules its sweeps without history, so the two runs stop at different (both converged)
    # iterates: equal to solver tolerance, and exactly equal in the state that was carried over
    assert rel_l2(oa["sum_drift"], ob["sum_drift"]) <= L2_TOL and rel_l2(oa["sum_subl"], ob["sum_subl"]) <= L2_TOL
    assert np.array_equal(b.get_state()["pbsm_more_than_avail"], a.get_state()["pbsm_more_than_avail"])
    a.close()
    b.close()


def test_error_paths_on_device(granger):
    bad = TriMesh(granger.vertex, granger.elem, granger.neigh.copy(), granger.params)
    bad.neigh[3, 1] = 5000
    with pytest.raises(capi.Pbsm3dError, match="out of bound neighbors"):
        capi.Handle(capi.default_config(nLayer=5), bad)
    with pytest.raises(capi.Pbsm3dError, match="nLayer"):
        capi.Handle(capi.default_config(nLayer=0), granger)
    h = capi.Handle(capi.default_config(nLayer=5, max_iterations=3, solver=capi.SOLVER_LINE), granger)
    geo = granger.geometry()
    with pytest.raises(capi.Pbsm3dError) as e:
        h.step(3600.0, synthetic.forcing(geo.cx, geo.cy))
    assert e.value.code == 5  # "failed to converge" (LinearAlgebra.cpp:236-243)
    mod = PBSM3D({"nLayer": 5, "max_iterations": 3, "solver": 1})
    dom = Domain(granger)
    mod.init(dom)
    for k, v in synthetic.forcing(geo.cx, geo.cy).items():
        dom[k] = v
    with pytest.raises(module_error, match="converge"):
        mod.run(dom)
    h.close()
    mod.close()


def test_fetch_is_required_while_a_fetch_option_is_on(granger):
    """The reference declares depends("fetch") when use_exp_fetch or use_tanh_fetch (the default) is set (PBSM3D.cpp:181-184) and
    would fail module linking without a provider; here a step without the array is refused instead of running on fetch = 1000 m."""
    geo = granger.geometry()
    F = {k: v for k, v in synthetic.forcing(geo.cx, geo.cy).items() if k != "fetch"}
    h = capi.Handle(capi.default_config(nLayer=5), granger)
    with pytest.raises(capi.Pbsm3dError, match="forcing array missing"):
        h.step(3600.0, F)
    h.close()


@pytest.mark.parametrize("which", ["granger1m", "slope_metis", "uniform", "variable"])
def test_colour_major_layout_is_a_proper_colouring(which):
    """The device order is a permutation of CHM's faces into colour classes in which no two edge-neighbours share
    a colour (what makes the in-place line Gauss-Seidel pass race-free); structured meshes need 2 colours."""
    mesh = {"uniform": lambda: synthetic.uniform_mesh(40, 30), "variable": lambda: synthetic.variable_mesh(5000)}.get(
        which, lambda: load_mesh(which))()
    h = capi.Handle(capi.default_config(nLayer=5), mesh)
    nc, ns, slot, colour = h.layout()
    T = mesh.n_local
    assert 1 <= nc <= 4 and ns >= T and ns % 32 == 0
    assert len(np.unique(slot)) == T and slot.min() >= 0 and slot.max() < ns
    for j in range(3):
        nb = mesh.neigh[:, j]
        has = (nb >= 0) & (nb < T)
        assert (colour[has] != colour[nb[has]]).all()
    for c in range(nc):  # CHM order is kept inside a class, classes are contiguous slot ranges
        s = slot[colour == c]
        assert (np.diff(s) == 1).all()
    if which == "uniform":
        assert nc == 2
    h.close()


def test_iteration_prediction_never_changes_results(slope):
    """Sweep/CG counts are predicted from the previous step (the solve itself always starts from x0 = 0): a handle
    that has seen other forcing must return the same fields as a fresh one."""
    geo = slope.geometry()
    kw = functest_kw(10)
    warm = capi.Handle(capi.default_config(**kw), slope)
    seq = [synthetic.forcing(geo.cx, geo.cy, seed=7, step=k, calm=(k == 2)) for k in range(5)]
    for F in seq:
        ow, sw = warm.step(3600.0, F)
        fresh = capi.Handle(capi.default_config(**kw), slope)
        of, sf = fresh.step(3600.0, F)
        assert sw["suspension_present"] == sf["suspension_present"] and sw["deposition_present"] == sf["deposition_present"]
        if sf["suspension_present"]:
            assert sw["suspension_residual"] <= 1e-8 and sf["suspension_residual"] <= 1e-8
        for v in ("Qsusp", "Qsalt", "Qsubl"):
            assert rel_l2(ow[v], of[v]) <= 1e-7, v
        if sf["deposition_present"]:
            assert rel_l2(ow["drift_mass"], of["drift_mass"]) <= 1e-6
        fresh.close()
    warm.close()


@pytest.mark.parametrize("meshname", ["slope", "variable"])
@pytest.mark.parametrize("dep", [capi.DEP_CG, capi.DEP_CHEBYSHEV, capi.DEP_SOR, capi.DEP_AUTO], ids=["cg", "chebyshev", "sor", "auto"])
def test_deposition_solvers_match_direct_solve(meshname, dep):
    """The device solvers of the (SPD) deposition system against the oracle's sparse direct solve; the Chebyshev iteration
    and the multicolour SOR use a spectrum estimate made once per mesh and must report the true residual."""
    mesh = load_mesh("slope") if meshname == "slope" else synthetic.variable_mesh(8000)
    geo = mesh.geometry()
    F = synthetic.forcing(geo.cx, geo.cy, seed=9)
    r = oracle_for(mesh, Config.functional_test(6)).step(F, 3600.0)
    h = capi.Handle(capi.default_config(deposition_solver=dep, tolerance=1e-11, **functest_kw(6)), mesh)
    for rep in range(3):  # first step: unknown iteration count; later steps: predicted schedule
        outs, st = h.step(3600.0, F)
        assert st["deposition_present"] == 1 and st["deposition_residual"] <= 1e-11
        assert st["deposition_solver_used"] == {capi.DEP_CG: capi.DEP_CG, capi.DEP_CHEBYSHEV: capi.DEP_CHEBYSHEV,
                                                capi.DEP_SOR: capi.DEP_SOR, capi.DEP_AUTO: capi.DEP_SOR}[dep]  # AUTO on one rank: SOR
        d = h.deposition_system()
        A = oracle_for(mesh, Config.functional_test(6)).deposition_csr(d["diag"], d["off"])
        res = np.linalg.norm(d["rhs"] - A @ d["q"]) / np.linalg.norm(d["rhs"])
        assert res <= 2e-11, (rep, res)  # the reported residual is the true one
        assert rel_l2(d["q"], r["q_dep"]) <= 1e-8
        assert rel_l2(outs["drift_mass"], r["drift_mass"]) <= 1e-8
    assert st["host_syncs"] == 1  # steady state: every prediction held, one synchronisation per step
    if dep == capi.DEP_SOR:  # about twice as fast as Chebyshev-accelerated Jacobi, sweep for iteration
        hc = capi.Handle(capi.default_config(deposition_solver=capi.DEP_CHEBYSHEV, tolerance=1e-11, **functest_kw(6)), mesh)
        _, sc = hc.step(3600.0, F)
        assert st["deposition_iterations"] <= 0.75 * sc["deposition_iterations"], (st["deposition_iterations"], sc["deposition_iterations"])
        hc.close()
    h.close()


def test_fp32_sweep_streams_do_not_move_the_answer(slope):
    """Sweeps far from convergence may stream fp32-rounded copies of their coefficients (fp64 x, fp64 arithmetic); the last
    sweeps and every residual check use the fp64 coefficients, so the result meets the same stopping rule and agrees with the
    all-fp64 schedule and the direct solve to the solver tolerance."""
    geo = slope.geometry()
    F = synthetic.forcing(geo.cx, geo.cy, seed=4)
    r = oracle_for(slope, Config.functional_test(10)).step(F, 3600.0)
    res = {}
    for flag in (1, 0):
        h = capi.Handle(capi.default_config(tolerance=1e-10, fp32_sweep_streams=flag, **functest_kw(10)), slope)
        for rep in range(3):  # the fp32 phase needs a sweep-count prediction: it starts with the second step
            outs, st = h.step(3600.0, F)
        res[flag] = (h.solution(), outs, st)
        assert st["suspension_residual"] <= 1e-10 and st["host_syncs"] == 1
        assert rel_l2(h.solution(), r["c"]) <= 1e-8
        h.close()
    assert res[1][2]["sweeps_timed_fp32"] > 0 and res[0][2]["sweeps_timed_fp32"] == 0
    assert abs(res[1][2]["suspension_iterations"] - res[0][2]["suspension_iterations"]) <= 2
    assert rel_l2(res[1][0], res[0][0]) <= 1e-9
    for v in ("Qsusp", "drift_mass"):
        assert rel_l2(res[1][1][v], res[0][1][v]) <= 1e-8


def _patchy(cx, cy, seed=7):
    """Bench forcing with saltation confined to wind-exposed patches (elsewhere 3 m/s: no saltation, zero right-hand side)."""
    return synthetic.patchy_forcing(cx, cy, seed=seed, scale=1500.0, threshold=0.4)


@pytest.mark.parametrize("L", [10, 12], ids=["L10", "L12-generic"])
@pytest.mark.parametrize("meshname", ["uniform", "variable"])
def test_active_set_never_changes_an_iterate(meshname, L, monkeypatch):
    """The persistent line solver skips columns whose right-hand side and whose neighbours' iterates are still exactly zero
    (tests/models/active_set_model.py): that update is a no-op, so the solution, every output and the sweep count must be
    IDENTICAL to the full sweep's (PBSM3D_ACTIVE_SET=0), on every step of a handle (fp64-only first step, fp32 phases later),
    while fewer column updates are executed."""
    mesh = synthetic.uniform_mesh(120, 120) if meshname == "uniform" else synthetic.variable_mesh(30000)
    geo = mesh.geometry()
    forc = [_patchy(geo.cx, geo.cy, seed=3), _patchy(geo.cx, geo.cy, seed=4), synthetic.forcing(geo.cx, geo.cy, seed=7)]
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("PBSM3D_ACTIVE_SET", flag)
        h = capi.Handle(capi.default_config(**functest_kw(L)), mesh)
        res[flag] = []
        for F in forc + forc[:1]:
            outs, st = h.step(3600.0, F)
            res[flag].append((h.solution().copy(), {k: v.copy() for k, v in outs.items()}, st))
        h.close()
    T = mesh.n_local
    saved = []
    for (x1, o1, s1), (x0, o0, s0) in zip(res["1"], res["0"]):
        assert s1["persistent_kernels"] == 1 and s1["active_set"] == 1 and s0["active_set"] == 0
        assert s1["suspension_present"] == 1
        assert s1["suspension_iterations"] == s0["suspension_iterations"] and s1["suspension_residual"] == s0["suspension_residual"]
        assert np.array_equal(x1, x0)
        for v in o1:
            assert np.array_equal(o1[v], o0[v]), v
        n1 = s1["column_updates_fp32_x"] + s1["column_updates_fp32"] + s1["column_updates_fp64"]
        n0 = s0["column_updates_fp32_x"] + s0["column_updates_fp32"] + s0["column_updates_fp64"]
        assert n0 == s0["sweeps_timed"] * T
        assert s0["residual_checks"] * T <= s0["columns_checked"] <= s0["residual_checks"] * (T + 256)  # padding slots included
        assert 0 < n1 <= n0 and s1["columns_checked"] <= s0["columns_checked"]
        saved.append(1.0 - n1 / n0)
    assert max(saved[:2]) > 0.10, saved  # the patchy fields leave a good part of the domain untouched


def test_active_set_switches_on_by_itself_when_saltation_is_patchy(monkeypatch):
    """Default (PBSM3D_ACTIVE_SET unset): the active set is used on the steps where at most 40 % of the faces have a non-zero
    right-hand side, and the plain sweep otherwise; either way the results equal the other mode's bit for bit."""
    monkeypatch.delenv("PBSM3D_ACTIVE_SET", raising=False)
    mesh = synthetic.uniform_mesh(120, 120)
    geo = mesh.geometry()
    T = mesh.n_local
    h = capi.Handle(capi.default_config(**functest_kw(10)), mesh)
    monkeypatch.setenv("PBSM3D_ACTIVE_SET", "0")
    h0 = capi.Handle(capi.default_config(**functest_kw(10)), mesh)
    seen = set()
    for F in (_patchy(geo.cx, geo.cy, seed=4), synthetic.forcing(geo.cx, geo.cy, seed=7), _patchy(geo.cx, geo.cy, seed=3)):
        outs, st = h.step(3600.0, F)
        o0, s0 = h0.step(3600.0, F)
        assert st["faces_with_rhs"] == s0["faces_with_rhs"] == int(np.count_nonzero(h.suspension_system()["rhs0"]))
        assert st["active_set"] == (1 if st["faces_with_rhs"] <= 0.4 * T else 0) and s0["active_set"] == 0
        seen.add(st["active_set"])
        assert np.array_equal(h.solution(), h0.solution())
        for v in outs:
            assert np.array_equal(outs[v], o0[v]), v
    assert seen == {0, 1}
    h.close()
    h0.close()


@pytest.mark.parametrize("L", [15, 20, 12])
def test_all_layer_count_specialisations(slope, L):
    """nLayer 15 and 20 have compile-time sweep/residual kernels of their own (config c5 uses 20), 12 takes the generic path; each on
    fp64 and fp32-rounded sweep streams (the latter start with the second step of a handle)."""
    geo = slope.geometry()
    F = synthetic.forcing(geo.cx, geo.cy, seed=11)
    r = oracle_for(slope, Config.functional_test(L)).step(F, 3600.0)
    h = capi.Handle(capi.default_config(tolerance=1e-10, **functest_kw(L)), slope)
    for rep in range(3):
        outs, st = h.step(3600.0, F)
        assert st["suspension_residual"] <= 1e-10
        assert rel_l2(h.solution(), r["c"]) <= 1e-8, (L, rep)
        for v in ("Qsusp", "Qsalt", "Qsubl", "drift_mass"):
            assert rel_l2(outs[v], r[v]) <= 1e-7, (L, rep, v)
    assert st["sweeps_timed_fp32"] > 0
    h.close()


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 3), (5, 4)])
def test_tiny_meshes(nx, ny):
    """2, 12 and 40 faces: fewer faces than a warp, every face on the hull, a spectrum estimate with next to no Krylov space."""
    mesh = synthetic.uniform_mesh(nx, ny)
    geo = mesh.geometry()
    F = synthetic.forcing(geo.cx, geo.cy, seed=2)
    F["U_R"] = np.full(mesh.n_local, 14.0)
    F["U_2m_above_srf"] = np.full(mesh.n_local, 9.0)
    r = oracle_for(mesh, Config.functional_test(5)).step(F, 3600.0)
    h = capi.Handle(capi.default_config(tolerance=1e-10, **functest_kw(5)), mesh)
    for rep in range(2):
        outs, st = h.step(3600.0, F)
        assert st["suspension_present"] == int(r["suspension_present"]) and st["deposition_present"] == int(r["deposition_present"])
        assert rel_l2(h.solution(), r["c"]) <= 1e-8
        for v in ("Qsusp", "Qsalt", "drift_mass"):
            assert rel_l2(outs[v], r[v]) <= 1e-7, v
    assert np.array_equal(h.fetchr(F["vw_dir"]), __import__("oracle.wind_oracle", fromlist=["fetchr"]).fetchr(F["vw_dir"], geo.cx, geo.cy, geo.cz, None))
    h.close()

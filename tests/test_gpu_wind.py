"""-m gpu: the device providers of U_2m_above_srf and fetch through the C-ABI against

* tests/golden/golden_wind.npz — outputs of the reference's scale_wind_vert.cpp / fetchr.cpp compiled unmodified;
* the numpy oracle (oracle/wind_oracle.py) on larger synthetic meshes, incl. the nearest-centre search against a kd-tree;
* the fused step: PBSM3D with the two inputs derived on the device = PBSM3D fed the oracle's arrays.

Tolerances: point_scale 1e-13 (a handful of log/exp), spline 1e-10 relative (its basis -(log x + c + E1(x)) cancels to
~x, so E1's last bits are amplified), fetch bit-exact (it is a multiple of the step, 0 or max_distance).
"""
import os

import numpy as np
import pytest

from chm_b200 import capi, module, synthetic
from chm_b200.mesh import TriMesh
from conftest import GOLDEN, functest_kw, load_mesh, rel_l2
from oracle import wind_oracle as wo

pytestmark = pytest.mark.gpu


def with_params(mesh, **params):
    return TriMesh(mesh.vertex, mesh.elem, mesh.neigh, dict(mesh.params, **params))


@pytest.mark.parametrize("name", ["granger1m", "slope"])
def test_reference_golden_vectors(name):
    g = np.load(os.path.join(GOLDEN, "golden_wind.npz"))
    base = load_mesh(name)
    U_R, sd, vw, canopy, lai = (g[f"{name}_{k}"] for k in ("U_R", "sd", "vw_dir", "canopy", "lai"))
    veg = capi.Handle(capi.default_config(nLayer=5), with_params(base, CanopyHeight=canopy, LAI=lai))
    bare_mesh = TriMesh(base.vertex, base.elem, base.neigh, {k: v for k, v in base.params.items() if k not in ("CanopyHeight", "LAI")})
    bare = capi.Handle(capi.default_config(nLayer=5), bare_mesh)
    mrel = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b)))
    assert mrel(veg.scale_wind_vert(U_R, sd, capi.default_wind_config(point_mode=1)), g[f"{name}_veg_u2_point"]) <= 1e-13
    assert mrel(veg.scale_wind_vert(U_R, sd), g[f"{name}_veg_u2"]) <= 1e-10
    assert mrel(bare.scale_wind_vert(U_R, None, capi.default_wind_config(point_mode=1)), g[f"{name}_bare_u2_point"]) <= 1e-13
    assert mrel(bare.scale_wind_vert(U_R, None), g[f"{name}_bare_u2"]) <= 1e-10
    assert mrel(veg.scale_wind_vert(U_R, sd, capi.default_wind_config(ignore_canopy=1)), g[f"{name}_ignore_u2"]) <= 1e-10
    assert np.array_equal(veg.fetchr(vw), g[f"{name}_veg_fetch"])
    assert np.array_equal(bare.fetchr(vw), g[f"{name}_bare_fetch"])
    wc = capi.default_wind_config(fetch_steps=7, fetch_max_distance=650.0, fetch_I=0.03, fetch_incl_veg=0)
    assert np.array_equal(veg.fetchr(vw, wc), g[f"{name}_ignore_fetch"])
    veg.close()
    bare.close()


def test_lai_missing_is_an_error_not_a_default():
    base = load_mesh("granger1m")
    m = TriMesh(base.vertex, base.elem, base.neigh, {"CanopyHeight": np.full(base.n_local, 3.0)})
    h = capi.Handle(capi.default_config(nLayer=5, use_R94_lambda=0), m)
    with pytest.raises(capi.Pbsm3dError, match="LAI"):
        h.scale_wind_vert(np.full(base.n_local, 8.0))
    assert np.all(h.scale_wind_vert(np.full(base.n_local, 8.0), cfg=capi.default_wind_config(ignore_canopy=1)) > 0.1)
    h.close()


def test_against_oracle_on_synthetic_meshes():
    rng = np.random.default_rng(12)
    for mesh in (synthetic.variable_mesh(60000, seed=4), synthetic.uniform_mesh(150, 150)):
        T = mesh.n_local
        canopy = np.where(rng.random(T) < 0.3, rng.uniform(0.05, 8.0, T), 0.0)
        lai = rng.uniform(0.2, 3.5, T)
        vz = mesh.vertex.copy()
        vz[:, 2] += 60.0 * np.sin(vz[:, 0] / 400.0) * np.cos(vz[:, 1] / 300.0)  # slopes around fetchr's I = 0.06
        m = TriMesh(vz, mesh.elem, mesh.neigh, {"CanopyHeight": canopy, "LAI": lai})
        geo = m.geometry()
        F = synthetic.forcing(geo.cx, geo.cy, seed=3)
        vw = rng.uniform(0.0, 360.0, T)
        h = capi.Handle(capi.default_config(nLayer=5), m)
        u_pt = wo.point_scale(F["U_R"], F["snowdepthavg"], canopy, lai)
        assert np.max(np.abs(h.scale_wind_vert(F["U_R"], F["snowdepthavg"], capi.default_wind_config(point_mode=1)) - u_pt) / u_pt) <= 1e-13
        # the spline on a sample of faces (the numpy oracle loops in Python), incl. hull faces with 1-2 neighbours
        u = h.scale_wind_vert(F["U_R"], F["snowdepthavg"])
        hull = np.where((m.neigh < 0).any(axis=1))[0][:200]
        sample = np.concatenate([hull, rng.choice(T, 2000, replace=False)])
        for i in sample:
            pts = np.array([(geo.cx[n], geo.cy[n], u_pt[n]) for n in m.neigh[i] if n >= 0])
            ref = max(0.1, wo.thin_plate_spline(pts, (geo.cx[i], geo.cy[i])))
            assert abs(u[i] - ref) <= 1e-10 * ref, i
        f = h.fetchr(vw)
        fo = wo.fetchr(vw, geo.cx, geo.cy, geo.cz, canopy)
        assert np.array_equal(f, fo)
        assert len(np.unique(f)) >= 6
        h.close()


def test_fused_providers_feed_the_step():
    """pbsm3d_set_providers: the step derives U_2m_above_srf and fetch on the device; same results as feeding it the arrays."""
    mesh = synthetic.uniform_mesh(90, 90)
    geo = mesh.geometry()
    F = synthetic.forcing(geo.cx, geo.cy, fetch_const=None)
    h1 = capi.Handle(capi.default_config(**functest_kw(8)), mesh)
    F1 = dict(F)
    F1["U_2m_above_srf"] = h1.scale_wind_vert(F["U_R"], F["snowdepthavg"])
    F1["fetch"] = h1.fetchr(F["vw_dir"])
    o1, s1 = h1.step(3600.0, F1)
    h2 = capi.Handle(capi.default_config(**functest_kw(8)), mesh)
    with pytest.raises(capi.Pbsm3dError):
        h2.step(3600.0, {k: v for k, v in F.items() if k != "U_2m_above_srf"})  # not switched on: the input is required
    h2.set_providers(capi.default_wind_config())
    o2, s2 = h2.step(3600.0, {k: v for k, v in F.items() if k not in ("U_2m_above_srf", "fetch")})
    assert s2["suspension_present"] and s2["deposition_present"]
    for k in ("Qsalt", "Qsusp", "Qsubl", "drift_mass", "sum_drift"):
        assert np.array_equal(o1[k], o2[k]), k
    assert np.array_equal(h1.solution(), h2.solution())
    h1.close()
    h2.close()


def test_module_mirrors():
    mesh = with_params(load_mesh("granger1m"), CanopyHeight=np.full(985, 0.4), LAI=np.full(985, 1.0))
    geo = mesh.geometry()
    dom = module.Domain(mesh)
    F = synthetic.forcing(geo.cx, geo.cy)
    for k in ("U_R", "snowdepthavg", "vw_dir", "swe", "t", "rh"):
        dom[k] = F[k]
    pb = module.PBSM3D({"nLayer": 5})
    sw, fe = module.scale_wind_vert({"ignore_canopy": "false"}), module.fetchr({"steps": 10})
    assert sw.get_depends() == ["U_R"] and sw.get_optionals() == ["snowdepthavg"] and sw.get_provides() == ["U_2m_above_srf"]
    assert fe.get_depends() == ["vw_dir"] and fe.get_provides() == ["fetch"]
    pb.init(dom); sw.init(dom); fe.init(dom)
    sw.run(dom, pb)
    fe.run(dom, pb)
    u, _ = wo.scale_wind_vert(F["U_R"], mesh.neigh, geo.cx, geo.cy, F["snowdepthavg"], mesh.params["CanopyHeight"], mesh.params["LAI"])
    assert np.max(np.abs(dom["U_2m_above_srf"] - u) / u) <= 1e-10
    assert np.array_equal(dom["fetch"], wo.fetchr(F["vw_dir"], geo.cx, geo.cy, geo.cz, mesh.params["CanopyHeight"]))
    st = pb.run(dom)
    assert st["suspension_present"] in (0, 1)
    pb.close()
    with pytest.raises(module.module_error):
        module.fetchr({"stepz": 3})

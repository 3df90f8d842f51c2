"""CPU: oracle/wind_oracle.py (scale_wind_vert, thin plate spline, fetchr) against the reference.

* the thin plate spline on the reference's OWN known-answer tests (src/tests/test_interpolation.cpp:47-170);
* point_scale / domain-mode scale_wind_vert / fetchr on tests/golden/golden_wind.npz, which holds outputs of the reference's
  scale_wind_vert.cpp and fetchr.cpp compiled unmodified (tests/golden/make_golden_wind.py);
* where oracle/_ref/libchmref.so is present, the live library on a fresh random mesh.
"""
import os

import numpy as np
import pytest

from chm_b200 import synthetic
from conftest import GOLDEN, load_mesh
from oracle import chm_ref, wind_oracle as wo


def test_tps_reference_known_answers():
    # TEST_F(InterpTest, spline2) / interpolation_class_static_size2: ASSERT_DOUBLE_EQ(result, 22.30217013945628) = 4 ULP
    s = np.array([(-1276639.4142831599, 1408220.6433826166, 22.241299818717572),
                  (-1276628.96002623, 1408213.5776356135, 22.423794697169313),
                  (-1276628.8896492834, 1408225.6645281466, 22.301020204404736)])
    r = wo.thin_plate_spline(s, (-1276633.6294519969, 1408220.6575855566))
    assert abs(r - 22.30217013945628) <= 4 * np.spacing(22.30217013945628)
    # TEST_F(InterpTest, spline) and its three variants: ASSERT_LT(fabs(s(xy,query) - 15.795), 1)
    s5 = np.array([(69., 76., 20.820), (59., 64., 10.910), (75., 52., 10.380), (86., 73., 14.600), (88., 53., 10.560)])
    assert abs(wo.thin_plate_spline(s5, (69., 67.)) - 15.795) < 1


@pytest.mark.parametrize("name", ["granger1m", "slope"])
def test_oracle_reproduces_reference_outputs(name):
    g = np.load(os.path.join(GOLDEN, "golden_wind.npz"))
    mesh = load_mesh(name)
    geo = mesh.geometry()
    U_R, sd, vw, canopy, lai = (g[f"{name}_{k}"] for k in ("U_R", "sd", "vw_dir", "canopy", "lai"))
    cases = {"veg": dict(snowdepthavg=sd, canopy_height=canopy, lai=lai), "bare": dict(),
             "ignore": dict(snowdepthavg=sd, canopy_height=canopy, lai=lai, ignore_canopy=True)}
    for c, kw in cases.items():
        u2, u_pt = wo.scale_wind_vert(U_R, mesh.neigh, geo.cx, geo.cy, **kw)
        assert np.max(np.abs(u2 - g[f"{name}_{c}_u2"]) / g[f"{name}_{c}_u2"]) <= 1e-12, c
        if c != "ignore":
            assert np.max(np.abs(u_pt - g[f"{name}_{c}_u2_point"]) / g[f"{name}_{c}_u2_point"]) <= 1e-14, c
    # fetch is a multiple of the step, 0 or max_distance: exact
    assert np.array_equal(wo.fetchr(vw, geo.cx, geo.cy, geo.cz, canopy), g[f"{name}_veg_fetch"])
    assert np.array_equal(wo.fetchr(vw, geo.cx, geo.cy, geo.cz, None), g[f"{name}_bare_fetch"])
    assert np.array_equal(wo.fetchr(vw, geo.cx, geo.cy, geo.cz, canopy, steps=7, max_distance=650.0, I=0.03, incl_veg=False),
                          g[f"{name}_ignore_fetch"])
    # the cases are not degenerate
    assert len(np.unique(g[f"{name}_bare_fetch"])) >= 8 and (g[f"{name}_veg_fetch"] == 0).any()


def test_point_scale_branches():
    """Every branch of scale_wind_vert.cpp:48-136 on hand-picked faces (values from the formulas in Atmosphere.cpp:32-46)."""
    U = np.full(6, 10.0)
    sd = np.array([0.5, 0.5, 0.5, 3.0, 49.0, -9999.0])
    can = np.array([0.0, 10.0, 4.0, 2.0, 1.0, 0.0])
    lai = np.full(6, 2.0)
    u = wo.point_scale(U, sd, can, lai)
    ls = wo.log_scale_wind
    assert np.isclose(u[0], ls(10.0, 50.0, 2.5, 0.5))                                    # no canopy
    utop = ls(10.0, 50.0, 10.0, 0.5)                                                     # 2 m above snow is below the canopy bottom
    assert np.isclose(u[1], max(0.1, ls(wo.exp_scale_wind(utop, 10.0, 5.0, 2.0), 5.0, 2.5, 0.5)))
    utop = ls(10.0, 50.0, 4.0, 0.5)                                                      # between canopy bottom and top
    assert np.isclose(u[2], wo.exp_scale_wind(utop, 4.0, 2.5, 2.0))
    assert np.isclose(u[3], ls(10.0, 50.0, 5.0, 3.0))                                    # snow above the canopy
    assert u[4] == 10.0                                                                  # 2 m above snow >= 50 m: U_R
    assert np.isclose(u[5], ls(10.0, 50.0, 2.0, 0.0))                                    # missing snow depth -> 0
    assert (wo.point_scale(np.full(3, 1e-3)) == 0.1).all()                                # floor


@pytest.mark.skipif(not chm_ref.available(), reason="oracle/_ref/libchmref.so not built (needs /root/reference)")
def test_live_reference_on_a_fresh_mesh():
    m = synthetic.variable_mesh(1500, seed=99)
    geo = m.geometry()
    T = m.n_local
    rng = np.random.default_rng(5)
    F = synthetic.forcing(geo.cx, geo.cy, seed=5)
    canopy = np.where(rng.random(T) < 0.5, rng.uniform(0.05, 9.0, T), 0.0)
    lai = rng.uniform(0.2, 4.0, T)
    vz = m.vertex.copy()
    vz[:, 2] += 40.0 * np.sin(vz[:, 0] / 150.0) * np.cos(vz[:, 1] / 110.0)  # slopes on both sides of I = 0.06
    m = type(m)(vz, m.elem, m.neigh, m.params)
    geo = m.geometry()
    ref = chm_ref.ReferencePBSM3D(m.vertex, m.elem, m.neigh, {"CanopyHeight": canopy, "LAI": lai}, {"nLayer": 5})
    u_ref = ref.scale_wind_vert(F["U_R"], F["snowdepthavg"])
    u, _ = wo.scale_wind_vert(F["U_R"], m.neigh, geo.cx, geo.cy, F["snowdepthavg"], canopy, lai)
    assert np.max(np.abs(u - u_ref) / u_ref) <= 1e-12
    vw = rng.uniform(0, 360, T)
    assert np.array_equal(wo.fetchr(vw, geo.cx, geo.cy, geo.cz, canopy), ref.fetchr(vw))
    # the C++ restatement of the spline the compiled module calls agrees with the numpy one on the reference's KAT
    s = np.array([(-1276639.4142831599, 1408220.6433826166, 22.241299818717572),
                  (-1276628.96002623, 1408213.5776356135, 22.423794697169313),
                  (-1276628.8896492834, 1408225.6645281466, 22.301020204404736)])
    assert abs(chm_ref.tpspline(s, (-1276633.6294519969, 1408220.6575855566)) - 22.30217013945628) <= 4 * np.spacing(22.3)
    ref.close()

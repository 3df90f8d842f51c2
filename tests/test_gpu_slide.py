"""-m gpu: snow_slide on the device through the C-ABI (pbsm3d_slide_init / pbsm3d_slide_run) against

* tests/golden/golden_slide.npz: outputs of the reference's own snow_slide.cpp (compiled unmodified, make_golden_slide.py);
* the sequential oracle (oracle/slide_oracle.py) on steep synthetic terrain with tens of thousands of faces: the device runs the
  reference's sequential sweep as a dependency wavefront, so the results must agree to rounding (1e-10 of the largest entry: one
  `pow` in maxDepth, amplified by depth / excess), not merely "conserve mass".
"""
import os

import numpy as np
import pytest

from chm_b200 import capi, synthetic
from chm_b200.mesh import TriMesh
from conftest import GOLDEN, load_mesh
from oracle import slide_oracle as so

pytestmark = pytest.mark.gpu
TOL = 1e-10
CASES = {"granger_m900": "granger1m", "slope_default": "slope", "slope_custom": "slope", "slope_veg": "slope"}


def close(a, b, tol=TOL):
    scale = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / scale <= tol


@pytest.mark.parametrize("tag", list(CASES))
def test_reference_golden_vectors(tag):
    g = np.load(os.path.join(GOLDEN, "golden_slide.npz"))
    m = load_mesh(CASES[tag])
    if f"{tag}_canopy" in g.files:
        m = TriMesh(m.vertex, m.elem, m.neigh, dict(m.params, CanopyHeight=g[f"{tag}_canopy"]))
    else:
        m = TriMesh(m.vertex, m.elem, m.neigh, {k: v for k, v in m.params.items() if k not in ("CanopyHeight", "LAI")})
    h = capi.Handle(capi.default_config(nLayer=2, use_R94_lambda=0), m)
    mult, power = g[f"{tag}_cfg"]
    h.slide_init(avalache_mult=mult, avalache_pow=power)
    for run in (1, 2):
        o, st = h.slide_run(g[f"{tag}_sd"], g[f"{tag}_sdv"], g[f"{tag}_swe"])
        for k in capi.SLIDE_OUTPUTS:
            assert close(o[k], g[f"{tag}_run{run}_{k}"]), (run, k)
        assert st["iterations"] == 1 and st["faces_fired"] > 10
    chk = h.slide_get_state()
    for i, k in enumerate(capi.SLIDE_OUTPUTS[:4]):
        assert close(chk[k], g[f"{tag}_checkpoint"][i]), k
    # load_checkpoint into a fresh handle: the sums continue from the restored values
    h2 = capi.Handle(capi.default_config(nLayer=2, use_R94_lambda=0), m)
    h2.slide_init(avalache_mult=mult, avalache_pow=power)
    h2.slide_set_state(**chk)
    o3a, _ = h.slide_run(g[f"{tag}_sd"], g[f"{tag}_sdv"], g[f"{tag}_swe"])
    o3b, _ = h2.slide_run(g[f"{tag}_sd"], g[f"{tag}_sdv"], g[f"{tag}_swe"])
    for k in capi.SLIDE_OUTPUTS:
        assert np.array_equal(o3a[k], o3b[k]), k
    h.close()
    h2.close()


@pytest.mark.parametrize("n,deep,order", [(100, 3.0, "morton"), (160, 2.0, "morton"), (90, 5.0, "native")])
def test_sequential_sweep_reproduced_on_steep_terrain(n, deep, order):
    m = synthetic.with_elevation(synthetic.uniform_mesh(n, n, order=order))
    geo = m.geometry()
    st = so.SlideState(m.face_vertices().reshape(-1, 3, 3), m.neigh, geo.area)
    sd, sdv, swe = so.synthetic_snow(geo.cx, geo.cy, st.slope, seed=n, deep=deep)
    want = so.run_single(st, sd, sdv, swe)
    h = capi.Handle(capi.default_config(nLayer=2), m)
    h.slide_init()
    got, stats = h.slide_run(sd, sdv, swe)
    assert np.count_nonzero(want["delta_avalanche_mass"]) > 1000
    for k in capi.SLIDE_OUTPUTS:
        assert close(got[k], want[k]), k
    # the same faces changed, not just the same totals
    assert np.array_equal(got["delta_avalanche_mass"] != 0, want["delta_avalanche_mass"] != 0)
    assert stats["wavefront_rounds"] < m.n_local // 10      # a wavefront, not a serial walk
    # run it again: deterministic to the bit, and the sums doubled
    again, _ = h.slide_run(sd, sdv, swe)
    assert np.array_equal(again["delta_avalanche_mass"], got["delta_avalanche_mass"])
    assert close(again["delta_avalanche_mass_sum"], 2 * want["delta_avalanche_mass"])
    h.close()


def test_variable_resolution_mesh_and_no_snow():
    m = synthetic.with_elevation(synthetic.variable_mesh(30000, seed=4))
    geo = m.geometry()
    st = so.SlideState(m.face_vertices().reshape(-1, 3, 3), m.neigh, geo.area)
    sd, sdv, swe = so.synthetic_snow(geo.cx, geo.cy, st.slope, seed=2, deep=3.0)
    want = so.run_single(st, sd, sdv, swe)
    h = capi.Handle(capi.default_config(nLayer=2), m)
    h.slide_init()
    got, stats = h.slide_run(sd, sdv, swe)
    for k in capi.SLIDE_OUTPUTS:
        assert close(got[k], want[k]), k
    # nothing above its holding depth: no face takes part, no wavefront round, outputs exactly zero
    z = np.zeros(m.n_local)
    h.slide_init()
    calm, cs = h.slide_run(z, z, z)
    assert cs["wavefront_rounds"] == 0 and cs["faces_fired"] == 0
    assert not calm["delta_avalanche_mass"].any() and not calm["delta_avalanche_snowdepth_sum"].any()
    h.close()


def test_slide_feeds_the_snowpack(tmp_path):
    """snow_slide -> snobal's avalanche hook (snobal.cpp:389-408) without leaving the library's conventions."""
    from oracle import snobal_oracle as sno
    m = synthetic.with_elevation(synthetic.uniform_mesh(50, 50))
    geo = m.geometry()
    st = so.SlideState(m.face_vertices().reshape(-1, 3, 3), m.neigh, geo.area)
    pack = sno.synthetic_state(m.n_local, seed=8)
    sd = pack["z_s"]
    sdv = sd / np.maximum(0.001, np.cos(st.slope))
    h = capi.Handle(capi.default_config(nLayer=2), m)
    h.slide_init(avalache_mult=600.0)
    o, stats = h.slide_run(sd, sdv, pack["m_s"])
    assert stats["faces_fired"] > 50
    got = h.apply_avalanche(dict(pack, layer_count=pack["layer_count"].astype(np.int32)), o["delta_avalanche_snowdepth"], o["delta_avalanche_mass"])
    want = sno.apply_avalanche(pack, o["delta_avalanche_snowdepth"], o["delta_avalanche_mass"], h.geometry()["area"])
    for k in sno.FIELDS:
        assert np.array_equal(np.asarray(got[k], dtype=np.float64), want[k], equal_nan=True), k
    h.close()

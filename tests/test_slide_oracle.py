"""CPU: the restatement of snow_slide (oracle/slide_oracle.py) against the reference's own snow_slide.cpp — the committed vectors
(tests/golden/golden_slide.npz, made by tests/golden/make_golden_slide.py from the compiled reference) and, where
oracle/_ref/libchmref.so is present, the live library.  Tolerance 1e-10 relative to the largest entry: maxDepth carries one
`pow` (numpy vs glibc differ in the last bit) and the released depth `snowdepth - maxDepth` amplifies it by depth / excess."""
import os

import numpy as np
import pytest

from chm_b200 import synthetic
from chm_b200.mesh import partition_mesh
from conftest import GOLDEN, load_mesh
from oracle import chm_ref, slide_oracle as so

TOL = 1e-10
CASES = {"granger_m900": "granger1m", "slope_default": "slope", "slope_custom": "slope", "slope_veg": "slope"}


def close(a, b, tol=TOL):
    scale = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / scale <= tol


def state_for(g, tag):
    m = load_mesh(CASES[tag])
    mult, power = g[f"{tag}_cfg"]
    canopy = g[f"{tag}_canopy"] if f"{tag}_canopy" in g.files else None
    st = so.SlideState(m.face_vertices().reshape(-1, 3, 3), m.neigh, m.geometry().area, canopy=canopy,
                       cfg=dict(avalache_mult=mult, avalache_pow=power))
    return m, st


@pytest.mark.parametrize("tag", list(CASES))
def test_oracle_reproduces_reference_vectors(tag):
    g = np.load(os.path.join(GOLDEN, "golden_slide.npz"))
    m, st = state_for(g, tag)
    for run in (1, 2):
        o = so.run_single(st, g[f"{tag}_sd"], g[f"{tag}_sdv"], g[f"{tag}_swe"])
        for k in so.ReferenceSlide.VARS:
            assert close(o[k], g[f"{tag}_run{run}_{k}"]), (run, k)
    assert np.count_nonzero(g[f"{tag}_run1_delta_avalanche_mass"]) > 20   # the fixture does move snow
    chk = g[f"{tag}_checkpoint"]                                          # snow_slide.cpp:59-76 persists these four
    for i, k in enumerate(so.ReferenceSlide.VARS[:4]):
        assert np.array_equal(chk[i], g[f"{tag}_run2_{k}"])


def test_contract_lists():
    g = np.load(os.path.join(GOLDEN, "golden_slide.npz"))
    assert list(g["depends"]) == ["snowdepthavg", "swe"]                  # snowdepthavg_vert is read without being declared (:121)
    assert {"delta_avalanche_mass", "delta_avalanche_snowdepth", "delta_avalanche_mass_sum", "delta_avalanche_snowdepth_sum",
            "maxDepth"} <= set(g["provides"])


def test_rank_local_sweep_with_ghosts():
    g = np.load(os.path.join(GOLDEN, "golden_slide.npz"))
    m = load_mesh("slope")
    geo = m.geometry()
    p = partition_mesh(m, 1, 3)
    T, gid = p.n_local, p.global_id
    V = p.face_vertices().reshape(-1, 3, 3)
    st = so.SlideState(V[:T], p.neigh, geo.area[gid[:T]], ghost_vertices=V[T:], ghost_area=geo.area[gid[T:]])
    sd, sdv, swe = (g[f"slope_default_{k}"] for k in ("sd", "sdv", "swe"))
    a, b, c = sd[gid[:T]].copy(), sdv[gid[:T]].copy(), swe[gid[:T]] / 1000.0
    d, e = np.zeros(T), np.zeros(T)
    acc = st.sweep(a, b, c, d, e, sdv[gid[T:]])
    z = np.zeros(T)
    st.absorb(a, b, c, d, e, z, z, z, z)
    assert close(d, g["rank1of3_delta_avalanche_snowdepth"]) and close(e, g["rank1of3_delta_avalanche_mass"])
    for k, name in enumerate(so.ReferenceSlide.GHOST_VARS):
        assert close(acc[k], g[f"rank1of3_{name}"]), name
    assert np.count_nonzero(acc[3]) > 5


def test_properties_on_steep_terrain():
    m = synthetic.with_elevation(synthetic.uniform_mesh(60, 60))
    geo = m.geometry()
    st = so.SlideState(m.face_vertices().reshape(-1, 3, 3), m.neigh, geo.area)
    assert np.degrees(st.slope.max()) > 45
    sd, sdv, swe = so.synthetic_snow(geo.cx, geo.cy, st.slope, seed=5, deep=3.0)
    o = so.run_single(st, sd, sdv, swe)
    moved = o["delta_avalanche_mass"]
    assert np.count_nonzero(moved) > 500
    # water volume is conserved up to what leaves through the domain edge
    boundary = (m.neigh < 0).any(axis=1)
    assert moved.sum() <= 1e-9 * np.abs(moved).sum()
    interior_only = so.run_single(so.SlideState(m.face_vertices().reshape(-1, 3, 3), m.neigh, geo.area), np.where(boundary, 0.0, sd),
                                  np.where(boundary, 0.0, sdv), np.where(boundary, 0.0, swe))
    # swe_copy + what moved is the input
    assert np.allclose(o["swe_copy"] * geo.area - swe / 1000.0 * geo.area, moved, rtol=0, atol=1e-9 * np.abs(moved).max())
    # a face that fired ends at its holding depth unless a later face routed snow back onto it
    fired = (o["delta_avalanche_snowdepth"] < 0)
    assert np.all(o["snowdepthavg_copy"][fired] >= st.maxDepth[fired] * (1 - 1e-12))
    assert interior_only["iterations"] == 1


@pytest.mark.skipif(not chm_ref.available(), reason="oracle/_ref/libchmref.so not built (needs /root/reference)")
def test_live_reference_on_steep_synthetic_terrain():
    m = synthetic.with_elevation(synthetic.uniform_mesh(40, 40))
    geo = m.geometry()
    st = so.SlideState(m.face_vertices().reshape(-1, 3, 3), m.neigh, geo.area)
    sd, sdv, swe = so.synthetic_snow(geo.cx, geo.cy, st.slope, seed=9, deep=3.0)
    ref = so.ReferenceSlide(m.vertex, m.elem, m.neigh, None, None)
    assert np.max(np.abs(ref.slope() - st.slope)) <= 1e-15
    o, r = so.run_single(st, sd, sdv, swe), ref.run(sd, sdv, swe)
    assert np.count_nonzero(r["delta_avalanche_mass"]) > 300
    for k in so.ReferenceSlide.VARS:
        assert close(o[k], r[k]), k


@pytest.mark.parametrize("n,deep", [(40, 3.0), (50, 1.2)])
def test_device_schedule_model_equals_the_sequential_sweep(n, deep):
    """The device runs the reference's sequential sweep as a dependency wavefront (live-set expansion, then rounds in which a face
    takes its turn once every earlier face within two edges has).  tests/models/slide_wavefront_model.py is that schedule in
    numpy: it must give the sequential sweep's result bit for bit."""
    from models import slide_wavefront_model as wm
    m = synthetic.with_elevation(synthetic.uniform_mesh(n, n))
    geo = m.geometry()
    st = so.SlideState(m.face_vertices().reshape(-1, 3, 3), m.neigh, geo.area)
    sd, sdv, swe = so.synthetic_snow(geo.cx, geo.cy, st.slope, seed=n, deep=deep)
    o = so.run_single(st, sd, sdv, swe)
    dsd, dmass, rounds, live, fired = wm.run(st, sd, sdv, swe)
    assert np.array_equal(dsd, o["delta_avalanche_snowdepth"]) and np.array_equal(dmass, o["delta_avalanche_mass"])
    assert fired > 500 and rounds < m.n_local // 8

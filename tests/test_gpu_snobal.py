"""-m gpu: snobal's drift_mass / avalanche consumer on the device through the C-ABI
(pbsm3d_apply_drift / pbsm3d_apply_avalanche) — BIT-EXACT against

* tests/golden/golden_snobal.npz: outputs of the reference's own sno.cpp (compiled unmodified, tests/golden/make_golden_snobal.py);
* the numpy oracle (oracle/snobal_oracle.py) on a 200 k-face synthetic mesh, incl. the fused path: PBSM3D step -> the handle's
  device-resident drift_mass -> snowpack, no host round trip of drift_mass.
"""
import os

import numpy as np
import pytest

from chm_b200 import capi, synthetic
from chm_b200.mesh import TriMesh
from conftest import GOLDEN, functest_kw
from oracle import snobal_oracle as so

pytestmark = pytest.mark.gpu


def gpu_state(st):
    return dict(st, layer_count=np.asarray(st["layer_count"]).astype(np.int32))


def assert_same(got, want):
    for k in so.FIELDS:
        assert np.array_equal(np.asarray(got[k], dtype=np.float64), want[k], equal_nan=True), k


@pytest.fixture(scope="module")
def handle2048():
    base = synthetic.uniform_mesh(32, 32)  # 2048 triangles
    assert base.n_local == 2048
    # the fixture's face areas as the mesh's "area" parameter (face->get_area() returns it when present, triangulation.hpp:1830-1856)
    g = np.load(os.path.join(GOLDEN, "golden_snobal.npz"))
    mesh = TriMesh(base.vertex, base.elem, base.neigh, {"area": g["area"]})
    h = capi.Handle(capi.default_config(nLayer=5), mesh)
    yield h, mesh
    h.close()


@pytest.mark.parametrize("case", ["default", "custom"])
def test_reference_golden_vectors(handle2048, case):
    h, mesh = handle2048
    g = np.load(os.path.join(GOLDEN, "golden_snobal.npz"))
    st = {k: g[f"in_{k}"] for k in so.FIELDS}
    dd, th, mz = g[f"{case}_cfg"]
    cfg = capi.SnobalConfig(dd, th, mz)
    r = h.apply_drift(gpu_state(st), g["drift_mass"], cfg)
    assert_same(r, {k: g[f"{case}_drift_{k}"] for k in so.FIELDS})
    assert np.array_equal(r["swe"], r["m_s"]) and np.array_equal(r["snowdepthavg"], r["z_s"])
    r2 = h.apply_drift({k: r[k] for k in so.FIELDS}, g["drift_mass"], cfg)
    assert_same(r2, {k: g[f"{case}_drift2_{k}"] for k in so.FIELDS})
    assert np.array_equal(h.geometry()["area"], g["area"])
    a = h.apply_avalanche(gpu_state(st), g["delta_avalanche_snowdepth"], g["delta_avalanche_mass"], cfg)
    assert_same(a, {k: g[f"{case}_aval_{k}"] for k in so.FIELDS})


def test_large_mesh_and_fused_drift_mass():
    mesh = synthetic.uniform_mesh(320, 320)  # 204 800 triangles
    T = mesh.n_local
    h = capi.Handle(capi.default_config(**functest_kw(10)), mesh)
    geo = mesh.geometry()
    F = synthetic.forcing(geo.cx, geo.cy, seed=3)
    outs, stats = h.step(3600.0, F)
    assert stats["deposition_present"] == 1
    st = so.synthetic_state(T, seed=41)
    want = so.apply_drift(st, outs["drift_mass"])
    assert_same(h.apply_drift(gpu_state(st), outs["drift_mass"]), want)   # drift_mass handed in by the caller
    assert_same(h.apply_drift(gpu_state(st), None), want)                 # the handle's own, never leaving the device
    d = so.synthetic_drift(st, seed=42)
    assert_same(h.apply_drift(gpu_state(st), d), so.apply_drift(st, d))
    h.close()


def test_missing_state_array_is_refused(handle2048):
    h, _ = handle2048
    pk = capi.Snowpack()
    rc = h.lib.pbsm3d_apply_drift(h.h, None, pk, None, None, None, 0)
    assert rc == 1 and b"required" in h.lib.pbsm3d_last_error()


def test_coupled_loop_pbsm3d_into_the_snowpack(granger):
    """Config c1's shape (bundled granger1m mesh, nLayer 5, 24 hourly steps) with the loop closed on the device: PBSM3D's
    drift_mass goes into each face's snowpack (pbsm3d_apply_drift on the handle's own device-resident drift_mass) and the pack's
    swe / depth are the next hour's inputs.  Checked hour by hour: the pack against the snobal oracle fed the device's drift_mass
    (bit-exact), PBSM3D's outputs against the PBSM3D oracle fed the same inputs (1e-6)."""
    from oracle.pbsm3d_oracle import Config, PBSM3DOracle
    mesh = granger
    geo = mesh.geometry()
    T = mesh.n_local
    h = capi.Handle(capi.default_config(nLayer=5), mesh)
    o = PBSM3DOracle(Config(nLayer=5), mesh.neigh, geo, mesh.global_id, mesh.n_global, mesh.params)
    pack = so.synthetic_state(T, seed=12)
    moved = 0
    for hour in range(24):
        F = synthetic.forcing(geo.cx, geo.cy, seed=7, step=hour, calm=(hour % 6 == 5))
        F["swe"] = pack["m_s"].copy()                 # snobal.cpp:468,491: what snobal hands to the next PBSM3D step
        F["snowdepthavg"] = pack["z_s"].copy()
        outs, st = h.step(3600.0, F)
        r = o.step(F, 3600.0)
        for v in ("Qsusp", "Qsalt", "drift_mass", "sum_drift"):
            scale = np.linalg.norm(r[v])
            assert np.linalg.norm(outs[v] - r[v]) <= 1e-6 * scale + 1e-300, (v, hour)
        want = so.apply_drift(pack, outs["drift_mass"])
        got = h.apply_drift(gpu_state(pack), None)    # the handle's own drift_mass, never leaving the device
        assert_same(got, want)
        assert np.array_equal(got["swe"], want["m_s"]) and np.array_equal(got["snowdepthavg"], want["z_s"])
        moved += int(np.count_nonzero(want["m_s"] != pack["m_s"]))
        pack = want
    assert moved > 24 * 50 and np.all(pack["m_s"] >= 0)
    h.close()

"""CPU: the numpy restatement of snobal's drift_mass / avalanche consumer (oracle/snobal_oracle.py) against the reference's own
sno.cpp — the committed vectors (tests/golden/golden_snobal.npz, made by tests/golden/make_golden_snobal.py from the compiled
reference) and, where oracle/_ref/libsnoref.so is present, the live library on fresh random packs.  Bit-exact: every operation is
one IEEE add / multiply / divide in the reference's order."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import snobal_oracle as so


def state_of(g, prefix):
    return {k: g[f"{prefix}_{k}"] for k in so.FIELDS}


def assert_same(a, b):
    for k in so.FIELDS:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("case", ["default", "custom"])
def test_oracle_reproduces_reference_vectors(case):
    g = np.load(os.path.join(GOLDEN, "golden_snobal.npz"))
    st = state_of(g, "in")
    dd, th, mz = g[f"{case}_cfg"]
    r = so.apply_drift(st, g["drift_mass"], dd, th, mz)
    assert_same(r, state_of(g, f"{case}_drift"))
    assert_same(so.apply_drift(r, g["drift_mass"], dd, th, mz), state_of(g, f"{case}_drift2"))
    assert_same(so.apply_avalanche(st, g["delta_avalanche_snowdepth"], g["delta_avalanche_mass"], g["area"], th, mz),
                state_of(g, f"{case}_aval"))


def test_vectors_reach_every_branch():
    g = np.load(os.path.join(GOLDEN, "golden_snobal.npz"))
    before, after = g["in_layer_count"].astype(int), g["default_drift_layer_count"].astype(int)
    seen = {(int(a), int(b)) for a, b in zip(before, after)}
    assert {(1, 0), (1, 1), (1, 2), (2, 0), (2, 1), (2, 2), (0, 0), (0, 1), (0, 2)} <= seen  # sno.cpp:2632-2640 + bare ground
    assert np.any(g["default_drift_rho"] == so.MAX_SNOW_DENSITY) or np.any(g["default_aval_rho"] == so.MAX_SNOW_DENSITY)
    assert np.any(g["default_drift_h2o_total"] > g["in_h2o_total"])      # a sub-threshold pack turned into water
    assert np.any(np.isnan(g["drift_mass"])) and np.any(g["drift_mass"] == -9999.0) and np.any(g["drift_mass"] == 0.0)


def test_properties():
    st = so.synthetic_state(5000, seed=21)
    same = so.apply_drift(st, np.zeros(5000))
    for k in ("z_s", "m_s", "layer_count", "T_s", "cc_s", "h2o_total"):
        assert np.array_equal(same[k], st[k]), k                          # drift_mass = 0 leaves a consistent pack alone
    d = so.synthetic_drift(st, seed=22)
    r = so.apply_drift(st, d)
    assert np.all(r["m_s"] >= 0) and np.all(r["z_s"] >= 0) and np.all(r["rho"] <= so.MAX_SNOW_DENSITY)
    two = r["layer_count"] == 2
    assert np.allclose(r["z_s_0"][two] + r["z_s_l"][two], r["z_s"][two], rtol=1e-14)
    assert np.all(r["z_s"][r["layer_count"] == 0] == 0)
    # mass either stays snow or becomes water: m_s + h2o_total changes by the applied mass wherever nothing was clipped at 0
    dm = np.where(np.isnan(d) | (d == -9999.0), 0.0, d)
    kept = st["m_s"] + dm >= 0
    tot0, tot1 = st["m_s"] + st["h2o_total"], r["m_s"] + r["h2o_total"]
    assert np.allclose((tot1 - tot0)[kept], dm[kept], rtol=0, atol=1e-9)


@pytest.mark.skipif(not so.reference_available(), reason="oracle/_ref/libsnoref.so not built (needs /root/reference)")
def test_live_reference_agrees_on_fresh_packs():
    for seed in (31, 32):
        st = so.synthetic_state(20000, seed=seed)
        d = so.synthetic_drift(st, seed=seed + 100)
        assert_same(so.apply_drift(st, d, 250.0, 0.2, 0.1), so.reference_apply_drift(st, d, 250.0, 0.2, 0.1))
        area = np.random.default_rng(seed).uniform(30.0, 8000.0, 20000)
        dv, dm = so.synthetic_avalanche(st, area, seed=seed + 200)
        assert_same(so.apply_avalanche(st, dv, dm, area), so.reference_apply_avalanche(st, dv, dm, area))

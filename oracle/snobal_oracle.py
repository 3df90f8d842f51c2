"""CPU restatement of the step on the far side of PBSM3D: snobal applies `drift_mass` to each face's snowpack
(SURVEY §8f rank 3).  TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/).

    src/modules/snobal.cpp:363-385        mass = is_nan(drift_mass) ? 0 : drift_mass; erosion removes depth at the pack's density,
                                          deposition adds depth at `drift_density`; sbal->_adj_snow(mass / density, mass)
    third_party/snobal/sno.cpp:2527-2575  _adj_snow      :2617-2696  _adj_layers     :2366-2405  _calc_layers
                              :1564-1580  _layer_mass    :2321-2329  _cold_content   :523-533    heat_stor
    third_party/snobal/snomacros.h        FREEZE 273.16, MAX_SNOW_DENSITY 750, MIN_SNOW_TEMP -75, CP_ICE(t)

Pinned to the reference's own code: oracle/_ref/libsnoref.so is sno.cpp compiled unmodified (oracle/refbuild/Makefile);
`reference_apply_drift` drives it, tests/golden/golden_snobal_drift.npz holds its outputs (tests/golden/make_golden_snobal.py).
The restatement is a per-face scalar loop written with numpy masks; every operation is a single IEEE add/multiply/divide in the
order the reference writes it, so agreement is bit-exact.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

FIELDS = ("z_s", "m_s", "rho", "layer_count", "z_s_0", "z_s_l", "m_s_0", "m_s_l", "cc_s", "cc_s_0", "cc_s_l", "T_s", "T_s_0", "T_s_l",
          "h2o_total", "h2o_vol", "h2o", "h2o_max", "h2o_sat")
FREEZE = 2.7316e2
MAX_SNOW_DENSITY = 750.0
MIN_SNOW_TEMP = -75.0
DEFAULTS = dict(drift_density=300.0, threshold=0.2, max_z_s_0=0.1)  # snobal.cpp:83, :131, :101

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libsnoref.so")
REF = os.environ.get("CHM_REFERENCE", "/root/reference")


def cp_ice(t):
    return ((0.024928 + (0.00176 * t)) * 4.186798188) / (1 * 0.001)  # CAL_TO_J(...) / G_TO_KG(1)


def cold_content(temp, mass):
    return np.where(temp < FREEZE, cp_ice(temp) * mass * (temp - FREEZE), 0.0)


def apply_drift(state: Dict[str, np.ndarray], drift_mass: np.ndarray, drift_density: float = 300.0, threshold: float = 0.2,
                max_z_s_0: float = 0.1) -> Dict[str, np.ndarray]:
    """snobal.cpp:363-387.  Returns the new state (a dict of arrays, FIELDS); the input is not modified."""
    mass = np.asarray(drift_mass, dtype=np.float64)
    mass = np.where((np.abs(mass - -9999.0) < 1e-5) | np.isnan(mass), 0.0, mass)
    with np.errstate(divide="ignore", invalid="ignore"):
        dz = mass / np.where(mass < 0.0, np.asarray(state["rho"], dtype=np.float64), drift_density)
    return adj_snow(state, dz, mass, threshold, max_z_s_0)


def apply_avalanche(state: Dict[str, np.ndarray], delta_avalanche_snowdepth, delta_avalanche_mass, area, threshold: float = 0.2,
                    max_z_s_0: float = 0.1) -> Dict[str, np.ndarray]:
    """snobal.cpp:389-408: snow_slide's per-face volume / swe-volume deltas become a depth and a mass change."""
    area = np.asarray(area, dtype=np.float64)
    d_depth = np.asarray(delta_avalanche_snowdepth, dtype=np.float64) / area
    d_mass = np.asarray(delta_avalanche_mass, dtype=np.float64) / area * 1000
    return adj_snow(state, d_depth, d_mass, threshold, max_z_s_0)


def adj_snow(state: Dict[str, np.ndarray], dz: np.ndarray, mass: np.ndarray, threshold: float = 0.2, max_z_s_0: float = 0.1):
    """sno::_adj_snow(delta_z_s, delta_m_s) (sno.cpp:2527-2575) for every face."""
    s = {k: np.array(state[k], dtype=np.float64, copy=True) for k in FIELDS}
    with np.errstate(divide="ignore", invalid="ignore"):
        # ---- _adj_snow
        z_s = s["z_s"] + dz
        m_s = s["m_s"] + mass
        cc_s, cc_s_0, cc_s_l, m_s_0 = s["cc_s"], s["cc_s_0"], s["cc_s_l"], s["m_s_0"]
        neg = (m_s < 0) | (z_s < 0)
        m_s = np.where(neg, 0.0, m_s)
        cc_s = np.where(neg, 0.0, cc_s)
        m_s_0 = np.where(neg, 0.0, m_s_0)
        cc_s_0 = np.where(neg, 0.0, cc_s_0)
        rho = np.where(z_s != 0.0, m_s / z_s, 0.0)
        clip = rho > MAX_SNOW_DENSITY
        rho = np.where(clip, MAX_SNOW_DENSITY, rho)
        z_s = np.where(clip, m_s / rho, z_s)
    adj_layers = clip | (dz != 0.0)
    # ---- _calc_layers (only where _adj_layers runs)
    prev = s["layer_count"].astype(np.int64)
    none = m_s <= threshold
    one = ~none & (z_s < max_z_s_0)
    two = ~none & ~one
    z0n = np.where(none, 0.0, np.where(one, z_s, max_z_s_0))
    zln = np.where(two, z_s - max_z_s_0, 0.0)
    thin = two & (zln * rho < threshold)
    lc = np.where(none, 0, np.where(one | thin, 1, 2))
    z0n = np.where(thin, z_s, z0n)
    zln = np.where(thin, 0.0, zln)
    zsn = np.where(none, 0.0, z_s)
    layer_count = np.where(adj_layers, lc, prev)
    z_s = np.where(adj_layers, zsn, z_s)
    z_s_0 = np.where(adj_layers, z0n, s["z_s_0"])
    z_s_l = np.where(adj_layers, zln, s["z_s_l"])
    # ---- _adj_layers, branch layer_count == 0
    gone = adj_layers & (layer_count == 0)
    h2o_total = np.where(gone & (m_s > 0.0), s["h2o_total"] + m_s, s["h2o_total"])
    rho = np.where(gone, 0.0, rho)
    m_s = np.where(gone, 0.0, m_s)
    cc_s = np.where(gone, 0.0, cc_s)
    m_s_0 = np.where(gone, 0.0, m_s_0)
    cc_s_0 = np.where(gone, 0.0, cc_s_0)
    T_s = np.where(gone, MIN_SNOW_TEMP + FREEZE, s["T_s"])
    T_s_0 = np.where(gone, MIN_SNOW_TEMP + FREEZE, s["T_s_0"])
    gone2 = gone & (prev == 2)
    m_s_l = np.where(gone2, 0.0, s["m_s_l"])
    cc_s_l = np.where(gone2, 0.0, cc_s_l)
    T_s_l = np.where(gone2, MIN_SNOW_TEMP + FREEZE, s["T_s_l"])
    h2o_vol, h2o, h2o_max, h2o_sat = (np.where(gone, 0.0, s[k]) for k in ("h2o_vol", "h2o", "h2o_max", "h2o_sat"))
    # ---- _layer_mass: in _adj_layers (layer_count > 0) and on a pure mass change
    lm = ~gone
    m_s_0 = np.where(lm, np.where(layer_count == 0, 0.0, rho * z_s_0), m_s_0)
    m_s_l = np.where(lm, np.where(layer_count == 2, rho * z_s_l, 0.0), m_s_l)
    grow = adj_layers & ~gone & (prev == 1) & (layer_count == 2)
    T_s_l = np.where(grow, T_s, T_s_l)
    cc_s_l = np.where(grow, cold_content(T_s_l, m_s_l), cc_s_l)
    shrink = adj_layers & ~gone & (prev == 2) & (layer_count == 1)
    T_s_l = np.where(shrink, MIN_SNOW_TEMP + FREEZE, T_s_l)
    cc_s_l = np.where(shrink, 0.0, cc_s_l)
    out = dict(z_s=z_s, m_s=m_s, rho=rho, layer_count=layer_count.astype(np.float64), z_s_0=z_s_0, z_s_l=z_s_l, m_s_0=m_s_0, m_s_l=m_s_l,
               cc_s=cc_s, cc_s_0=cc_s_0, cc_s_l=cc_s_l, T_s=T_s, T_s_0=T_s_0, T_s_l=T_s_l, h2o_total=h2o_total, h2o_vol=h2o_vol,
               h2o=h2o, h2o_max=h2o_max, h2o_sat=h2o_sat)
    return out


def synthetic_state(n: int, seed: int = 5) -> Dict[str, np.ndarray]:
    """Consistent two-layer snowpacks plus the edge cases the reference's branches test: no snow, one thin layer, a pack at the
    layer threshold, density at the clip, a lower layer too light to exist."""
    rng = np.random.default_rng(seed)
    z_s = rng.uniform(0.0, 1.5, n)
    rho = rng.uniform(80.0, 500.0, n)
    kind = rng.integers(0, 8, n)
    z_s = np.where(kind == 0, 0.0, z_s)               # bare ground
    z_s = np.where(kind == 1, rng.uniform(0.001, 0.09, n), z_s)  # one thin layer
    z_s = np.where(kind == 2, 0.1 + rng.uniform(0, 1e-3, n), z_s)  # just above the active-layer depth
    rho = np.where(kind == 3, 749.0, rho)             # next to MAX_SNOW_DENSITY
    m_s = rho * z_s
    none = m_s <= 0.2
    layer = np.where(none, 0, np.where((z_s < 0.1) | ((z_s - 0.1) * rho < 0.2), 1, 2))
    z_s = np.where(none, 0.0, z_s)
    m_s = np.where(none, 0.0, m_s)
    rho = np.where(none, 0.0, rho)
    z_s_0 = np.where(layer == 2, 0.1, z_s)
    z_s_l = np.where(layer == 2, z_s - 0.1, 0.0)
    T_s_0 = np.where(none, MIN_SNOW_TEMP + FREEZE, FREEZE - rng.uniform(0.0, 25.0, n))
    T_s_l = np.where(layer == 2, FREEZE - rng.uniform(0.0, 10.0, n), MIN_SNOW_TEMP + FREEZE)
    T_s = np.where(none, MIN_SNOW_TEMP + FREEZE, np.where(layer == 2, 0.5 * (T_s_0 + T_s_l), T_s_0))
    m_s_0, m_s_l = rho * z_s_0, rho * z_s_l
    h2o_sat = np.where(none, 0.0, rng.uniform(0, 1, n))
    h2o_max = np.where(none, 0.0, 0.01 * z_s * 1000.0)
    st = dict(z_s=z_s, m_s=m_s, rho=rho, layer_count=layer.astype(np.float64), z_s_0=z_s_0, z_s_l=z_s_l, m_s_0=m_s_0, m_s_l=m_s_l,
              cc_s=cold_content(T_s, m_s), cc_s_0=cold_content(T_s_0, m_s_0), cc_s_l=cold_content(T_s_l, m_s_l), T_s=T_s, T_s_0=T_s_0,
              T_s_l=T_s_l, h2o_total=rng.uniform(0, 3, n), h2o_vol=np.where(none, 0.0, rng.uniform(0, 0.01, n)),
              h2o=h2o_sat * h2o_max, h2o_max=h2o_max, h2o_sat=h2o_sat)
    return st


def synthetic_drift(state: Dict[str, np.ndarray], seed: int = 6) -> np.ndarray:
    """drift_mass values that reach every branch: deposition, erosion within / beyond the pack, exact zero, -9999 and NaN."""
    rng = np.random.default_rng(seed)
    n = state["m_s"].shape[0]
    d = rng.normal(0.0, 12.0, n)
    pick = rng.integers(0, 10, n)
    d = np.where(pick == 0, 0.0, d)
    d = np.where(pick == 1, -9999.0, d)
    d = np.where(pick == 2, np.nan, d)
    d = np.where(pick == 3, -state["m_s"] * rng.uniform(0.9, 1.2, n), d)  # erode (almost / more than) everything
    d = np.where(pick == 4, rng.uniform(100.0, 600.0, n), d)               # heavy deposition: density clip / second layer
    return d


# ------------------------------------------------------------------ the reference's own sno.cpp (oracle/_ref/libsnoref.so)
def build_reference(force: bool = False) -> Optional[str]:
    if not os.path.isdir(os.path.join(REF, "third_party", "snobal")):
        return LIB if os.path.exists(LIB) else None
    args = ["make", "-C", os.path.join(HERE, "refbuild"), f"REF={REF}", "../_ref/libsnoref.so"] + (["-B"] if force else [])
    res = subprocess.run(args, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building oracle/_ref/libsnoref.so failed:\n" + res.stdout + res.stderr)
    return LIB


def reference_available() -> bool:
    return os.path.exists(LIB)


def reference_apply_drift(state, drift_mass, drift_density=300.0, threshold=0.2, max_z_s_0=0.1):
    L = C.CDLL(LIB)
    assert L.chmref_sno_fields() == len(FIELDS)
    n = len(drift_mass)
    buf = np.ascontiguousarray(np.stack([np.asarray(state[k], dtype=np.float64) for k in FIELDS]))
    dm = np.ascontiguousarray(drift_mass, dtype=np.float64)
    L.chmref_sno_apply_drift.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double]
    L.chmref_sno_apply_drift(n, buf.ctypes.data, dm.ctypes.data, float(drift_density), float(threshold), float(max_z_s_0))
    return {k: buf[i].copy() for i, k in enumerate(FIELDS)}


def reference_apply_avalanche(state, delta_avalanche_snowdepth, delta_avalanche_mass, area, threshold=0.2, max_z_s_0=0.1):
    L = C.CDLL(LIB)
    n = len(area)
    buf = np.ascontiguousarray(np.stack([np.asarray(state[k], dtype=np.float64) for k in FIELDS]))
    a, b, ar = (np.ascontiguousarray(v, dtype=np.float64) for v in (delta_avalanche_snowdepth, delta_avalanche_mass, area))
    L.chmref_sno_apply_avalanche.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_double, C.c_double]
    L.chmref_sno_apply_avalanche(n, buf.ctypes.data, a.ctypes.data, b.ctypes.data, ar.ctypes.data, float(threshold), float(max_z_s_0))
    return {k: buf[i].copy() for i, k in enumerate(FIELDS)}


def synthetic_avalanche(state: Dict[str, np.ndarray], area: np.ndarray, seed: int = 8):
    """(delta_avalanche_snowdepth, delta_avalanche_mass): donors lose part or all of their pack, receivers gain; most faces 0."""
    rng = np.random.default_rng(seed)
    n = area.shape[0]
    pick = rng.integers(0, 6, n)
    frac = np.where(pick == 0, -rng.uniform(0.1, 1.0, n), np.where(pick == 1, rng.uniform(0.1, 3.0, n), 0.0))
    frac = np.where(pick == 2, -1.0, frac)
    dvol = frac * state["z_s"] * area
    dswe = frac * state["m_s"] / 1000.0 * area
    recv_empty = (pick == 3)
    dvol = np.where(recv_empty, rng.uniform(0.0, 0.8, n) * area, dvol)
    dswe = np.where(recv_empty, dvol * rng.uniform(0.1, 0.9, n), dswe)   # densities 100..900 kg/m^3: some clip at 750
    return dvol, dswe

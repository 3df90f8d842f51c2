"""Golden vectors for the consumer of PBSM3D's drift_mass — OUTPUTS OF THE REFERENCE ITSELF:
third_party/snobal/sno.cpp compiled unmodified by oracle/refbuild/Makefile into oracle/_ref/libsnoref.so and driven through the
module glue of src/modules/snobal.cpp:363-408 (restated in oracle/refbuild/sno_harness.cpp).  Run HERE, where /root/reference exists:

    python tests/golden/make_golden_snobal.py     ->  tests/golden/golden_snobal.npz

Cases, 2048 faces each (oracle/snobal_oracle.py:synthetic_*): packs of 0 / 1 / 2 layers incl. bare ground, a pack at the
active-layer depth, density next to the 750 kg/m^3 clip; drift_mass with deposition, erosion within / beyond the pack, exact 0,
-9999 and NaN; non-default (drift_density, threshold, max_active_layer); avalanche volume deltas (donors emptied, receivers on bare
ground, deposits denser than the clip).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import snobal_oracle as so  # noqa: E402

N = 2048


def main():
    assert so.build_reference() and so.reference_available()
    out = {}
    st = so.synthetic_state(N, seed=5)
    drift = so.synthetic_drift(st, seed=6)
    area = np.random.default_rng(9).uniform(40.0, 6000.0, N)
    dvol, dswe = so.synthetic_avalanche(st, area, seed=8)
    for k in so.FIELDS:
        out["in_" + k] = st[k]
    out.update(drift_mass=drift, area=area, delta_avalanche_snowdepth=dvol, delta_avalanche_mass=dswe)
    cases = {"default": dict(so.DEFAULTS), "custom": dict(drift_density=180.0, threshold=1.5, max_z_s_0=0.25)}
    for name, kw in cases.items():
        r = so.reference_apply_drift(st, drift, **kw)
        r2 = so.reference_apply_drift(r, drift, **kw)  # a second hour with the same drift: packs that changed layer count move again
        a = so.reference_apply_avalanche(st, dvol, dswe, area, kw["threshold"], kw["max_z_s_0"])
        for k in so.FIELDS:
            out[f"{name}_drift_{k}"] = r[k]
            out[f"{name}_drift2_{k}"] = r2[k]
            out[f"{name}_aval_{k}"] = a[k]
        out[f"{name}_cfg"] = np.array([kw["drift_density"], kw["threshold"], kw["max_z_s_0"]])
    path = os.path.join(ROOT, "tests", "golden", "golden_snobal.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

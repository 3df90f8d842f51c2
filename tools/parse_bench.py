#!/usr/bin/env python
"""Print the essentials of a bench.py JSON line (last line starting with '{' of the given file)."""
import json
import sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
c = d["config"]
print("N", d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "median", round(c["median_ms_per_step"], 3), {k: round(v, 3) for k, v in c["phases_ms"].items()},
      "its", c["suspension_iterations"][:3], c["deposition_iterations"][:3], "launches/step", d["gpu_launches"] / d["steps"],
      "calm", c["calm_step_ms"], "e2e", round(d["e2e"]["ms_per_step"], 3), "persistent", c.get("persistent_solver_kernels"))
r = d["roofline"]
print("roofline", round(r["frac"], 3), r["kernel"][:40], "share", round(r["share_of_step"], 3), "asm", r.get("assembly", {}).get("ms"))
if d.get("parity_check"):
    print("parity", json.dumps(d["parity_check"])[:900])
v = (c.get("variants") or {}).get("default_block")
if v:
    print("default_block", {k: v.get(k) for k in ("ms_per_step", "sweeps", "solver_used", "launches_per_step", "error")})
s = c.get("strong_c4")
if s:
    print("strong_c4", {k: s.get(k) for k in ("ms_per_step", "efficiency_vs_1gpu", "one_gpu", "phases_ms", "suspension_iterations", "launches_per_step", "seconds_total", "error")})
print("providers", c.get("providers") and {k: c["providers"][k] for k in ("scale_wind_vert_ms", "fetchr_ms", "e2e_ms_with_providers_fused")})
if d.get("cpu_baseline"):
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])

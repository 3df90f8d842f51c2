"""CPU: the parts of bench.py's contract that do not need a GPU — the reference arm prints one JSON line with the agreed keys
(it times the CPU restatement of the reference on this host), and the roofline helper's byte counts match DESIGN.md §3."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sweep_bytes_per_row():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.sweep_bytes_per_row(10) == 58.0            # 5 fp64 streams + x read + x write + (3 slots + rhs)/L
    assert bench.sweep_bytes_per_row(10, True) == 38.0      # float4 {latS0..2, belowS} + cp32 instead of the five fp64 streams
    assert bench.mesh_side(1) == 708 and bench.mesh_side(4) == 1416


@pytest.mark.timeout(600)
def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "pbsm3d_element_layer_solves_per_s"
    for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["value"] > 0 and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]

"""chm_b200/csrc/pbsm3d_math.cuh (flog / fexp / fpow, the elementary functions of the assembly kernels) against libm.

The header is `__host__ __device__`: this test compiles tests/fastmath_host.cu for the HOST with nvcc and runs it on the CPU,
2·10^6 samples over the argument ranges of the path.  The device build of the same functions differs only in the reciprocal /
square-root seeds (MUFU + Newton instead of the host's divide) and is covered by the 1e-12 coefficient parity tests (-m gpu)."""
import json
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_fast_math_matches_libm(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "fastmath_host")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-x", "cu", os.path.join(HERE, "fastmath_host.cu"), "-o", exe],
                   check=True, capture_output=True)
    err = json.loads(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    assert err["flog"] <= 4e-16 and err["fexp"] <= 6e-16, err
    assert err["fpow"] <= 4e-15, err  # |y ln x| up to 28 here; <= 20 on the path
    assert err["fcbrt"] <= 1e-15, err

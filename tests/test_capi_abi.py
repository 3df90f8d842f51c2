"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/pbsm3d.h declares, reports the
reference's config defaults, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from chm_b200 import build, capi
from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    build.build()
    return capi.load_library()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pbsm3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pbsm3d_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pbsm3d.h but not exported"
    assert set(syms) == set(capi.SYMBOLS), "ctypes table and header disagree"
    assert lib.pbsm3d_abi_version() == capi.ABI_VERSION == 8


def test_config_defaults_are_the_reference_defaults(lib):
    c = capi.default_config()
    # PBSM3D.cpp:123-145, 223-258; LinearAlgebra.cpp:166-168
    expect = dict(nLayer=10, do_fixed_settling=0, settling_velocity=0.5, do_sublimation=1, do_lateral_diff=1,
                  smooth_coeff=820.0, min_sd_trans=0.1, cutoff=0.3, snow_diffusion_const=0.3, rouault_diffusion_coef=0,
                  enable_veg=1, iterative_subl=0, use_exp_fetch=0, use_tanh_fetch=1, use_PomLi_probability=0,
                  z0_ustar_coupling=0, use_subgrid_topo=0, use_subgrid_topo_V2=0, use_R94_lambda=1, debug_output=0,
                  tolerance=1e-8, max_iterations=1000, solver=0, deposition_solver=0, fp32_sweep_streams=1)
    for k, v in expect.items():
        assert getattr(c, k) == v, k
    with pytest.raises(KeyError):
        capi.default_config(not_a_key=1)


def test_struct_layouts_match_header(lib, tmp_path):
    """The ctypes mirrors against the C compiler's view of include/pbsm3d.h: sizes and the offset of every struct's last field."""
    import subprocess
    probes = {"pbsm3d_forcing": (capi.Forcing, "fetch"), "pbsm3d_outputs": (capi.Outputs, "pbsm_more_than_avail"),
              "pbsm3d_comm": (capi.Comm, None), "pbsm3d_stats": (capi.Stats, "persistent_kernels"), "pbsm3d_mesh": (capi.Mesh, None),
              "pbsm3d_config": (capi.Config, "fp32_sweep_streams"), "pbsm3d_wind_config": (capi.WindConfig, "fetch_I")}
    lines = []
    for cname, (_, last) in probes.items():
        lines.append(f'printf("{cname} %zu %zu\\n", sizeof({cname}), {"offsetof(" + cname + ", " + last + ")" if last else "(size_t)0"});')
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pbsm3d.h"\nint main(void){' + "".join(lines) + "return 0;}\n")
    exe = tmp_path / "probe"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.splitlines():
        cname, size, off = line.split()
        ct, last = probes[cname]
        assert C.sizeof(ct) == int(size), cname
        if last:
            assert getattr(ct, last).offset == int(off), (cname, last)
    for name, _ in capi.Stats._fields_ + capi.Config._fields_:  # every mirrored field is a member of the C struct
        assert name in open(os.path.join(inc, "pbsm3d.h")).read(), name

def test_create_fails_loudly_without_cuda(lib, granger):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this is the no-GPU behaviour")
    with pytest.raises(capi.Pbsm3dError) as e:
        capi.Handle(capi.default_config(nLayer=5), granger)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


def test_argument_errors_do_not_need_a_gpu(lib, granger):
    # the reference's own config errors come back as error codes + message (→ module_error in the adaptor)
    with pytest.raises(capi.Pbsm3dError, match="Cannot specify both exp_fetch and tanh_fetch"):
        capi.Handle(capi.default_config(use_exp_fetch=1, use_tanh_fetch=1), granger)
    with pytest.raises(capi.Pbsm3dError, match="settling velocity must be positive"):
        capi.Handle(capi.default_config(settling_velocity=-1.0), granger)
    for k in ("iterative_subl", "z0_ustar_coupling", "use_subgrid_topo", "debug_output"):
        with pytest.raises(capi.Pbsm3dError) as e:
            capi.Handle(capi.default_config(**{k: 1}), granger)
        assert e.value.code == 2
    assert lib.pbsm3d_create(None, None, 0, None, None) == 1
    assert b"null" in lib.pbsm3d_last_error()


def test_missing_library_is_an_import_error(tmp_path):
    with pytest.raises(ImportError, match="no CPU fallback"):
        capi.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under chm_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "chm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src, os.path.join(dirpath, f)

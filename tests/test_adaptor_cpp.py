"""The C++ CHM adaptor (chm_b200/host/PBSM3D_gpu.cpp) — the file a CHM maintainer drops into src/modules/ — compiled
here against the stand-in CHM types and driven like CHM's core drives a module: ctor(config) → init(mesh) → run(mesh)."""
import os
import subprocess

import numpy as np
import pytest

from chm_b200 import build, synthetic
from conftest import load_mesh, rel_l2
from oracle.pbsm3d_oracle import Config, PBSM3DOracle

OUT = ["Qsalt", "Qsusp", "Qsubl", "Qsubl_mass", "sum_subl", "drift_mass", "sum_drift", "pbsm_more_than_avail"]


def test_adaptor_compiles_against_the_c_abi():
    build.build()
    exe = build.build_adaptor()
    assert os.path.exists(exe)
    # it registers the same dependency lists as PBSM3D::PBSM3D (PBSM3D.cpp:105-202)
    src = open(os.path.join(build.HOST, "PBSM3D_gpu.cpp")).read()
    for v in ("U_2m_above_srf", "vw_dir", "swe", "t", "rh", "U_R"):
        assert f'depends("{v}")' in src
    for v in OUT + ["global_cell_id", "blowingsnow_probability"]:
        assert f'provides("{v}")' in src


def write_case(d, mesh, cfg_lines, forcings):
    np.ascontiguousarray(mesh.vertex, dtype=np.float64).tofile(d / "vertex.bin")
    np.ascontiguousarray(mesh.elem, dtype=np.int32).tofile(d / "elem.bin")
    np.ascontiguousarray(mesh.neigh, dtype=np.int32).tofile(d / "neigh.bin")
    if "area" in mesh.params:
        np.ascontiguousarray(mesh.params["area"], dtype=np.float64).tofile(d / "area.bin")
    (d / "config.txt").write_text("\n".join(cfg_lines) + "\n")
    for k, F in enumerate(forcings):
        for n, a in F.items():
            np.ascontiguousarray(a, dtype=np.float64).tofile(d / f"forcing_{k}_{n}.bin")


@pytest.mark.gpu
def test_adaptor_runs_like_a_chm_module_and_matches_the_oracle(tmp_path):
    mesh = load_mesh("granger1m")
    geo = mesh.geometry()
    forc = [synthetic.forcing(geo.cx, geo.cy, seed=7, step=k, calm=(k == 1)) for k in range(3)]
    write_case(tmp_path, mesh, ["nLayer 5"], forc)
    exe = build.build_adaptor()
    res = subprocess.run([exe, str(tmp_path), "3"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    o = PBSM3DOracle(Config(nLayer=5), mesh.neigh, geo, mesh.global_id, mesh.n_global, mesh.params)
    for k, F in enumerate(forc):
        r = o.step(F, 3600.0)
        for v in ("Qsalt", "Qsusp", "Qsubl", "sum_subl"):
            got = np.fromfile(tmp_path / f"out_{k}_{v}.bin")
            assert rel_l2(got, r[v]) <= 1e-6, (v, k)
        got = np.fromfile(tmp_path / f"out_{k}_drift_mass.bin")
        # a face variable keeps its previous value on a step without a deposition solve (the calm step)
        assert rel_l2(got, r["drift_mass"]) <= 1e-6, k
        assert rel_l2(np.fromfile(tmp_path / f"out_{k}_sum_drift.bin"), r["sum_drift"]) <= 1e-6
    assert rel_l2(np.fromfile(tmp_path / "checkpoint_sum_drift.bin"), r["sum_drift"]) <= 1e-6


@pytest.mark.gpu
def test_adaptor_turns_library_errors_into_module_error(tmp_path):
    mesh = load_mesh("granger1m")
    geo = mesh.geometry()
    write_case(tmp_path, mesh, ["nLayer 5", "iterative_subl true"], [synthetic.forcing(geo.cx, geo.cy)])
    res = subprocess.run([build.build_adaptor(), str(tmp_path), "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 1 and "module_error" in res.stderr and "iterative_subl" in res.stderr

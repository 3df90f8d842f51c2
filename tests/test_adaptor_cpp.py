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
    for v in ("U_2m_above_srf", "vw_dir", "swe", "t", "rh", "U_R", "fetch"):
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


@pytest.mark.gpu
def test_adaptor_with_fused_providers(tmp_path):
    """"fuse_providers": the adaptor stops depending on U_2m_above_srf / fetch; the library derives them on the device
    (scale_wind_vert + fetchr kernels).  Same outputs as the plain adaptor fed the oracle's provider outputs."""
    from oracle import wind_oracle as wo
    mesh = load_mesh("granger1m")
    geo = mesh.geometry()
    F = synthetic.forcing(geo.cx, geo.cy, seed=7, step=0)
    u2, _ = wo.scale_wind_vert(F["U_R"], mesh.neigh, geo.cx, geo.cy, F["snowdepthavg"])
    fetch = wo.fetchr(F["vw_dir"], geo.cx, geo.cy, geo.cz, None)
    exe = build.build_adaptor()
    plain, fused = tmp_path / "plain", tmp_path / "fused"
    plain.mkdir(); fused.mkdir()
    write_case(plain, mesh, ["nLayer 5"], [dict(F, U_2m_above_srf=u2, fetch=fetch)])
    garbage = dict(F, U_2m_above_srf=np.full(mesh.n_local, -9999.0), fetch=np.full(mesh.n_local, -9999.0))  # must not be read
    write_case(fused, mesh, ["nLayer 5", "fuse_providers true"], [garbage])
    for d in (plain, fused):
        res = subprocess.run([exe, str(d), "1"], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
    for v in ("Qsalt", "Qsusp", "Qsubl", "drift_mass"):
        a, b = np.fromfile(plain / f"out_0_{v}.bin"), np.fromfile(fused / f"out_0_{v}.bin")
        assert np.abs(b).max() > 0 and rel_l2(b, a) <= 1e-9, v


def test_snow_slide_adaptor_declares_the_reference_contract():
    """chm_b200/host/snow_slide_gpu.cpp: the depends / provides lists of snow_slide.cpp:30-39 (read from the compiled reference)."""
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "golden_slide.npz"))
    src = open(os.path.join(build.HOST, "snow_slide_gpu.cpp")).read()
    for v in g["depends"]:
        assert f'depends("{v}")' in src
    for v in g["provides"]:
        if not v.startswith("ghost_ss_"):       # the reference's MPI scratch variables: the exchange lives in the library
            assert f'provides("{v}")' in src
    for key in ("avalache_mult", "avalache_pow", "use_vertical_snow"):
        assert f'cfg.get("{key}"' in src


@pytest.mark.gpu
def test_snow_slide_adaptor_matches_the_reference_vectors(tmp_path):
    """snow_slide_gpu driven like a CHM module (ctor -> init -> run -> run -> checkpoint) on granger1m, after PBSM3D_gpu has
    flattened the mesh: the variables it provides = the reference's own snow_slide.cpp (golden_slide.npz, granger_m900)."""
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "golden_slide.npz"))
    mesh = load_mesh("granger1m")
    geo = mesh.geometry()
    write_case(tmp_path, mesh, ["nLayer 5"], [synthetic.forcing(geo.cx, geo.cy, seed=7, step=0)])
    (tmp_path / "slide_config.txt").write_text("avalache_mult 900\n")
    for n, k in (("snowdepthavg", "sd"), ("snowdepthavg_vert", "sdv"), ("swe", "swe")):
        np.ascontiguousarray(g[f"granger_m900_{k}"]).tofile(tmp_path / f"slide_{n}.bin")
    res = subprocess.run([build.build_adaptor(), str(tmp_path), "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "snow_slide run 1" in res.stdout
    for run in (0, 1):
        for v in ("delta_avalanche_snowdepth", "delta_avalanche_mass", "delta_avalanche_snowdepth_sum", "delta_avalanche_mass_sum", "maxDepth"):
            got, want = np.fromfile(tmp_path / f"slide_out_{run}_{v}.bin"), g[f"granger_m900_run{run + 1}_{v}"]
            assert np.max(np.abs(got - want)) <= 1e-10 * np.max(np.abs(want)), (run, v)
    assert np.array_equal(np.fromfile(tmp_path / "slide_checkpoint_mass_sum.bin"), np.fromfile(tmp_path / "slide_out_1_delta_avalanche_mass_sum.bin"))

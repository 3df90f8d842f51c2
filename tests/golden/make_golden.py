"""Generate the committed fixtures under tests/golden/ (run HERE, where /root/reference exists).

1. Mesh fixtures: the reference's bundled meshes, re-encoded as npz (vertex/elem/neigh/params in CHM face
   order after the cell_global_id permutation) because /root/reference does not exist on the GPU box.
     granger1m.npz  <- test_data/meshes/granger1m.mesh + .param           (985 triangles)
     slope.npz      <- functional_tests/mesh_versioning/slope.mesh + .param (2618 triangles)
     slope_metis.npz<- functional_tests/mesh_versioning/slope.metis.mesh   (METIS permutation + 31 local_size)
2. Golden vectors — OUTPUTS OF THE REFERENCE ITSELF: the reference's own src/modules/PBSM3D.cpp (+ Atmosphere.cpp,
   coordinates.cpp), compiled unmodified by oracle/refbuild/Makefile into oracle/_ref/libchmref.so and driven through
   oracle/chm_ref.py (init → run per step, exactly as CHM drives a module).  The linear solves inside are sparse direct
   solves (the contract of NearestNeighborProblem::Solve), so the vectors are the exact solution of the systems the
   reference assembles.
     golden_granger1m_L5_default.npz   24 hourly steps, nLayer 5, code defaults (config c1's shape), calm hours at k%8==5
     golden_slope_L10_functest.npz     3 steps, nLayer 10, functional-test PBSM3D block
     golden_variants.npz               every supported config key / vegetation / water / missing-value variant, 2 steps
     golden_pomli.npz                  use_PomLi_probability with stalk and R94 vegetation, 3 steps (blowingsnow_probability too)
     golden_helpers.npz                the reference's scalar helpers on a sample grid

    python tests/golden/make_golden.py
"""
import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from chm_b200 import synthetic  # noqa: E402
from chm_b200.mesh import TriMesh, read_chm_mesh  # noqa: E402
from oracle import chm_ref  # noqa: E402
from oracle.pbsm3d_oracle import Config  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
OUTPUTS = ("Qsusp", "Qsalt", "Qsubl", "Qsubl_mass", "drift_mass", "sum_drift", "sum_subl", "pbsm_more_than_avail")


def save_mesh(name, mesh):
    d = {"vertex": mesh.vertex, "elem": mesh.elem, "neigh": mesh.neigh}
    for k, v in mesh.params.items():
        d["param_" + k] = v
    if mesh.local_sizes is not None:
        d["local_sizes"] = mesh.local_sizes
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def cfg_dict(cfg: Config):
    d = dataclasses.asdict(cfg)
    d["smooth_coeff"] = int(d["smooth_coeff"])  # PBSM3D.cpp:247 reads it with an int default
    return d


def ell_view(A, neigh, T, L):
    """The reference's CSR (rows/cols layer*G+id) as the extruded-ELL arrays the tests compare: diag, lat, below, above."""
    A = A.tocsr()
    idx = np.arange(T)
    diag = np.empty((L, T)); lat = np.zeros((3, L, T)); below = np.zeros((L, T)); above = np.zeros((L, T))
    for z in range(L):
        r = z * T + idx
        diag[z] = np.asarray(A[r, r]).ravel()
        for f in range(3):
            has = neigh[:, f] >= 0
            lat[f, z, has] = np.asarray(A[r[has], z * T + neigh[has, f]]).ravel()
        if z > 0:
            below[z] = np.asarray(A[r, r - T]).ravel()
        if z < L - 1:
            above[z] = np.asarray(A[r, r + T]).ravel()
    return diag, lat, below, above


def record(out, tag, k, r, mesh, L, keep_system, compact=False):
    T = mesh.n_local
    for v in OUTPUTS:
        out[f"{tag}{v}_{k}"] = r[v]
    A, b = r["susp"]
    Ad, bd = r["dep"]
    susp = np.abs(b).max(initial=0.0) > 1e-12  # PBSM3D.cpp:1424-1427
    dep = bool(susp and np.abs(bd).max(initial=0.0) > 1e-12)  # :1661-1664
    out[f"{tag}present_{k}"] = np.array([susp, dep], dtype=np.int8)
    if keep_system:
        out[f"{tag}c_{k}"] = r["c"]
        out[f"{tag}diag_{k}"], out[f"{tag}lat_{k}"], out[f"{tag}below_{k}"], out[f"{tag}above_{k}"] = ell_view(A, mesh.neigh, T, L)
        out[f"{tag}rhs_{k}"] = b[:T]
        assert not b[T:].any()
        if compact:  # variants: the vertical couplings and the static deposition matrix are pinned by the main sequences
            del out[f"{tag}below_{k}"], out[f"{tag}above_{k}"]
            out[f"{tag}dep_rhs_{k}"] = bd
            return
        out[f"{tag}dep_diag_{k}"] = Ad.diagonal()
        off = np.zeros((3, T))
        for j in range(3):
            has = mesh.neigh[:, j] >= 0
            off[j, has] = np.asarray(Ad.tocsr()[np.arange(T)[has], mesh.neigh[has, j]]).ravel()
        out[f"{tag}dep_off_{k}"] = off
        out[f"{tag}dep_rhs_{k}"] = bd
        out[f"{tag}q_dep_{k}"] = r["q_dep"]


def run_golden(name, mesh, cfg, nsteps, keep_c_steps=(0,)):
    geo = mesh.geometry()
    ref = chm_ref.ReferencePBSM3D(mesh.vertex, mesh.elem, mesh.neigh, mesh.params, cfg_dict(cfg))
    out = {}
    for k in range(nsteps):
        calm = (k % 8 == 5)  # a calm hour now and then: early-out path + stale drift_mass
        F = synthetic.forcing(geo.cx, geo.cy, seed=7, step=k, calm=calm)
        r = ref.step(F, 3600.0)
        record(out, "", k, r, mesh, int(cfg.nLayer), k in keep_c_steps)
    out["depends"] = np.array(ref.depends())
    out["provides"] = np.array(ref.provides())
    ref.close()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def variant_cases(n):
    """name -> (Config, params added to granger1m, landcover table, forcing tweak)."""
    shrub = synthetic.shrub_params(n)
    lc = {"landcover": np.where(np.arange(n) % 17 == 0, 1.0, 2.0)}
    table = {"landcover.1.is_water": True, "landcover.2.is_water": False}
    ident = lambda F: F
    def missing(F):
        F = dict(F)
        F["snowdepthavg"] = np.where(np.arange(n) % 3 == 0, -9999.0, F["snowdepthavg"])
        # swe is missing only where the snow depth is missing too (no saltation there).  On a saltating face swe == 0
        # makes the availability test `mass < 0 && |mass| > swe` (PBSM3D.cpp:910) a test on the rounding noise of a sum
        # that is analytically zero — its outcome depends on the last bit of libm's sin/cos, not on the algorithm.
        F["swe"] = np.where(np.arange(n) % 6 == 0, np.nan, F["swe"])
        return F
    def varfetch(F):
        return dict(F, fetch=np.linspace(0.0, 1000.0, n))
    return {
        "default_L5": (Config(nLayer=5), {}, {}, ident),
        "functest_L10": (Config.functional_test(10), {}, {}, ident),
        "generic_L7": (Config(nLayer=7), {}, {}, ident),
        "L2_min": (Config(nLayer=2), {}, {}, ident),
        "no_subl_no_latdiff": (Config(nLayer=5, do_sublimation=False, do_lateral_diff=False), {}, {}, ident),
        "rouault": (Config(nLayer=5, rouault_diffusion_coef=True), {}, {}, ident),
        "exp_fetch": (Config(nLayer=5, use_exp_fetch=True, use_tanh_fetch=False), {}, {}, varfetch),
        "tanh_fetch_var": (Config(nLayer=5), {}, {}, varfetch),
        "no_fetch": (Config(nLayer=5, use_tanh_fetch=False), {}, {}, ident),
        "fixed_settling_0p3": (Config(nLayer=5, do_fixed_settling=True, settling_velocity=0.3, snow_diffusion_const=0.5,
                                      smooth_coeff=2000.0, min_sd_trans=0.3, cutoff=0.5), {}, {}, ident),
        "veg_R94": (Config(nLayer=5, use_R94_lambda=True), shrub, {}, ident),
        "veg_stalks": (Config(nLayer=5, use_R94_lambda=False), shrub, {}, ident),
        "veg_stalk_defaults": (Config(nLayer=5, use_R94_lambda=False), {"CanopyHeight": shrub["CanopyHeight"]}, {}, ident),
        "veg_disabled": (Config(nLayer=5, enable_veg=False), shrub, {}, ident),
        "water": (Config(nLayer=5, use_R94_lambda=False), dict(shrub, **lc), table, ident),
        "missing_values": (Config(nLayer=5), {}, {}, missing),
    }


def run_variants(mesh):
    geo = mesh.geometry()
    out = {}
    for name, (cfg, extra, table, tweak) in variant_cases(mesh.n_local).items():
        params = dict(mesh.params, **extra)
        ref = chm_ref.ReferencePBSM3D(mesh.vertex, mesh.elem, mesh.neigh, params, cfg_dict(cfg), table)
        for k in range(2):
            F = tweak(synthetic.forcing(geo.cx, geo.cy, seed=3, step=k))
            r = ref.step(F, 3600.0)
            record(out, name + "/", k, r, mesh, int(cfg.nLayer), k == 0, compact=True)
        ref.close()
    np.savez_compressed(os.path.join(OUT, "golden_variants.npz"), **out)


def run_pomli(mesh):
    """use_PomLi_probability (PBSM3D.cpp:848-866): stalk vegetation (z0v > 0 on part-buried shrubs) and R94 (N = dv = 0),
    p_snow_hours from a seeded field; 3 steps so that blowingsnow_probability shows its keep-previous-value behaviour."""
    geo = mesh.geometry()
    n = mesh.n_local
    shrub = synthetic.shrub_params(n, canopy=0.45)  # 0.2-1.5 m of snow: buried, and exposed by less than `cutoff`
    out = {}
    for name, r94 in (("pomli_stalks", False), ("pomli_R94", True)):
        cfg = Config(nLayer=5, use_PomLi_probability=True, use_R94_lambda=r94)
        ref = chm_ref.ReferencePBSM3D(mesh.vertex, mesh.elem, mesh.neigh, dict(mesh.params, **shrub), cfg_dict(cfg))
        for k in range(3):
            F = synthetic.forcing(geo.cx, geo.cy, seed=3, step=k, calm=(k == 1))
            F["p_snow_hours"] = np.random.default_rng(40 + k).uniform(0.5, 240.0, n)
            r = ref.step(F, 3600.0)
            record(out, name + "/", k, r, mesh, 5, k == 0, compact=True)
            out[f"{name}/blowingsnow_probability_{k}"] = ref.get_var("blowingsnow_probability")
            out[f"{name}/p_snow_hours_{k}"] = F["p_snow_hours"]
        ref.close()
    np.savez_compressed(os.path.join(OUT, "golden_pomli.npz"), **out)


def run_helpers():
    rng = np.random.default_rng(5)
    u = rng.uniform(0.5, 25, 200); zout = rng.uniform(0.3, 49, 200); sd = rng.uniform(0, 0.2, 200)
    z0 = np.where(rng.random(200) < 0.5, 0.01, rng.uniform(0.001, 0.05, 200))
    tk = rng.uniform(230, 285, 200)
    bearing = np.concatenate([rng.uniform(0, 360, 190), [0, 90, 180, 270, 360, 89.999, 90.001, 45, 225, 315]])
    xy = np.array([chm_ref.bearing_to_cartesian(b) for b in bearing])
    np.savez_compressed(os.path.join(OUT, "golden_helpers.npz"), u=u, zout=zout, sd=sd, z0=z0, tk=tk, bearing=bearing,
                        log_scale_wind=np.array([chm_ref.log_scale_wind(a, 50.0, b, c, d) for a, b, c, d in zip(u, zout, sd, z0)]),
                        es=np.array([chm_ref.saturated_vapour_pressure(t) for t in tk]), bx=xy[:, 0], by=xy[:, 1])


if __name__ == "__main__":
    chm_ref.build()
    if "--pomli-only" in sys.argv:  # added after the other vectors were committed; those are not regenerated
        d = np.load(os.path.join(OUT, "granger1m.npz"))
        run_pomli(TriMesh(d["vertex"], d["elem"], d["neigh"], {k[6:]: d[k] for k in d.files if k.startswith("param_")}))
        print("ok")
        sys.exit(0)
    g = read_chm_mesh(f"{REF}/test_data/meshes/granger1m.mesh", [f"{REF}/test_data/meshes/granger1m.param"])
    save_mesh("granger1m", g)
    s = read_chm_mesh(f"{REF}/functional_tests/mesh_versioning/slope.mesh", [f"{REF}/functional_tests/mesh_versioning/slope.param"])
    save_mesh("slope", s)
    sm = read_chm_mesh(f"{REF}/functional_tests/mesh_versioning/slope.metis.mesh", [])
    save_mesh("slope_metis", sm)
    run_golden("golden_granger1m_L5_default", g, Config(nLayer=5), 24)
    run_golden("golden_slope_L10_functest", s, Config.functional_test(10), 3)
    run_variants(g)
    run_pomli(g)
    run_helpers()
    print("ok")

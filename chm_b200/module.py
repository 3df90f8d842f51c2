"""Host-side mirror of CHM's plugin contract for the PBSM3D path (Python twin of host/PBSM3D_gpu.hpp).

The reference registers ``class PBSM3D : public module_base`` (src/modules/PBSM3D.hpp:306-423,
module_base.hpp:534-540) and the core calls ``PBSM3D(config_file)``, ``init(mesh&)``, ``run(mesh&)``,
``checkpoint``/``load_checkpoint``.  Inputs and outputs live in each face's variable store
(``(*face)["name"_s]``, default -9999).  This file keeps those names, argument meanings and error
behaviour so parity tests read like tests of the reference module; all numerics happen in
libpbsm3d_b200.so through the C-ABI (chm_b200/capi.py).  There is no CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from . import capi
from .mesh import TriMesh

MISSING = -9999.0  # variablestorage default (src/timeseries/variablestorage.hpp)


class module_error(RuntimeError):
    """CHM's module_error (src/exception.hpp:133-134)."""


class Domain:
    """What ``mesh& domain`` gives a domain-parallel module: the local faces and their variable store.

    ``domain[name]`` is the SoA view of ``(*domain->face(i))[name]`` for i in [0, size_faces()).
    """

    def __init__(self, mesh: TriMesh, dt: float = 3600.0, is_water: Optional[np.ndarray] = None):
        self.mesh = mesh
        self.dt = float(dt)  # global_param->dt()
        self.is_water = is_water
        self._vars: Dict[str, np.ndarray] = {}

    def size_faces(self) -> int:
        return self.mesh.n_local

    def size_global_faces(self) -> int:
        return self.mesh.n_global

    def init_face_data(self, names):
        """core → triangulation::init_face_data (triangulation.cpp:2536-2560): every variable starts missing."""
        for n in names:
            self._vars.setdefault(n, np.full(self.mesh.n_local, MISSING))

    def __getitem__(self, name: str) -> np.ndarray:
        if name not in self._vars:
            raise module_error(f"Variable {name} does not exist.")  # SAFE_CHECKS behaviour of the hash store
        return self._vars[name]

    def __setitem__(self, name: str, value):
        self._vars[name] = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64), (self.mesh.n_local,))).copy()

    def has(self, name: str) -> bool:
        return name in self._vars


class PBSM3D:
    """Drop-in for the reference module, backed by the B200 library.

    Same config keys and defaults as PBSM3D.cpp:123-145,223-258; same depends/provides lists as
    PBSM3D.cpp:103-219 (debug_output and the sub-grid options are refused, not ignored).
    """

    name = "PBSM3D"
    parallel = "domain"  # parallel::domain

    CONFIG_KEYS = {
        "nLayer", "do_fixed_settling", "settling_velocity", "do_sublimation", "do_lateral_diff", "smooth_coeff",
        "min_sd_trans", "cutoff", "snow_diffusion_const", "rouault_diffusion_coef", "enable_veg", "iterative_subl",
        "use_exp_fetch", "use_tanh_fetch", "use_PomLi_probability", "z0_ustar_coupling", "use_subgrid_topo",
        "use_subgrid_topo_V2", "use_R94_lambda", "debug_output",
        # solver controls (hard-coded in LinearAlgebra.cpp:164-168; exposed here)
        "tolerance", "max_iterations", "solver", "deposition_solver",
        # keys CHM's config block carries that this path does not read
        "N", "dv",
    }

    def __init__(self, cfg: Optional[dict] = None, device: int = 0, rank: int = 0, n_ranks: int = 1,
                 unique_id: Optional[bytes] = None):
        cfg = dict(cfg or {})
        unknown = set(cfg) - self.CONFIG_KEYS
        if unknown:
            raise module_error(f"PBSM3D: unknown config key(s) {sorted(unknown)}")
        self.cfg = {k: self._coerce(v) for k, v in cfg.items() if k not in ("N", "dv")}
        self._depends: List[str] = []
        self._provides: List[str] = []
        self.device, self.rank, self.n_ranks, self.unique_id = device, rank, n_ranks, unique_id
        get = lambda k, d: self.cfg.get(k, d)
        # --- PBSM3D::PBSM3D (PBSM3D.cpp:103-219)
        for v in ("U_2m_above_srf", "vw_dir", "swe", "t", "rh", "U_R"):
            self.depends(v)
        self.provides("pbsm_more_than_avail")
        self.provides("global_cell_id")
        use_exp, use_tanh = bool(get("use_exp_fetch", False)), bool(get("use_tanh_fetch", True))
        if use_exp and use_tanh:
            raise module_error("PBSM3d: Cannot specify both exp_fetch and tanh_fetch")
        self.depends("fetch" if (use_exp or use_tanh) else "p_snow_hours")
        self.provides("blowingsnow_probability")
        for v in ("Qsubl", "Qsubl_mass", "sum_subl", "drift_mass", "Qsusp", "Qsalt", "sum_drift"):
            self.provides(v)
        self.handle: Optional[capi.Handle] = None

    @staticmethod
    def _coerce(v):
        if isinstance(v, str):  # CHM configs carry everything as strings ("true", "10", "0.3")
            s = v.strip().lower()
            if s in ("true", "false"):
                return s == "true"
            try:
                return int(s)
            except ValueError:
                return float(s)
        return v

    def depends(self, name: str):
        self._depends.append(name)

    def provides(self, name: str):
        self._provides.append(name)

    def get_depends(self) -> List[str]:
        return list(self._depends)

    def get_provides(self) -> List[str]:
        return list(self._provides)

    # ------------------------------------------------------------------ init(mesh&)
    def init(self, domain: Domain):
        """PBSM3D::init (PBSM3D.cpp:221-398): flatten the mesh to device SoA, build both linear systems' structure."""
        domain.init_face_data(self._provides)
        cc = capi.default_config()
        for k, v in self.cfg.items():
            if k == "smooth_coeff":
                # the reference reads it as an INT (cfg.get("smooth_coeff", 820), PBSM3D.cpp:235): ptree's stream translator yields the
                # default when the text does not parse completely as an int ("820.7" -> 820), as PBSM3D_gpu.cpp does
                v = int(v) if float(v).is_integer() else 820
            setattr(cc, k, type(getattr(cc, k))(v))
        try:
            self.handle = capi.Handle(cc, domain.mesh, device=self.device, rank=self.rank, n_ranks=self.n_ranks,
                                      unique_id=self.unique_id, is_water=domain.is_water)
        except capi.Pbsm3dError as e:
            raise module_error(str(e)) from e
        domain["sum_drift"] = 0.0  # (*face)["sum_drift"_s]=0, PBSM3D.cpp:391

    # ------------------------------------------------------------------ run(mesh&)
    def run(self, domain: Domain) -> dict:
        """PBSM3D::run (PBSM3D.cpp:400-1748): gather the face store → one library call → scatter."""
        if self.handle is None:
            raise module_error("PBSM3D::run called before init")
        F = {"U_R": domain["U_R"], "U_2m_above_srf": domain["U_2m_above_srf"], "swe": domain["swe"], "t": domain["t"],
             "rh": domain["rh"], "vw_dir": domain["vw_dir"]}
        # snowdepthavg is read without being declared (PBSM3D.cpp:446); missing → -9999 → treated as 0
        F["snowdepthavg"] = domain["snowdepthavg"] if domain.has("snowdepthavg") else np.full(domain.size_faces(), MISSING)
        if "fetch" in self._depends:
            F["fetch"] = domain["fetch"]
        if self.cfg.get("use_PomLi_probability", False):
            F["p_snow_hours"] = domain["p_snow_hours"]  # read whether or not it was declared (PBSM3D.cpp:137-141 vs :853)
        try:
            outs, stats = self.handle.step(domain.dt, F)
        except capi.Pbsm3dError as e:
            raise module_error(str(e)) from e
        for k, v in outs.items():
            if k == "drift_mass" and not stats["deposition_present"]:
                continue  # the reference leaves the face variable untouched on such steps
            if k == "pbsm_more_than_avail":
                cur = domain["pbsm_more_than_avail"]
                domain[k] = np.where(v > 0, 1.0, cur)  # only ever set to 1 (PBSM3D.cpp:1727)
                continue
            if k == "sum_drift" and not stats["deposition_present"]:
                continue
            domain[k] = v
        return stats

    # ------------------------------------------------------------------ checkpoint (PBSM3D.cpp:1753-1773)
    def checkpoint(self, domain: Domain) -> Dict[str, np.ndarray]:
        return {"PBSM3D:sum_drift": domain["sum_drift"].copy()}

    def load_checkpoint(self, domain: Domain, chk: Dict[str, np.ndarray]):
        domain["sum_drift"] = chk["PBSM3D:sum_drift"]
        self.handle.set_state(sum_drift=chk["PBSM3D:sum_drift"])

    def close(self):
        if self.handle is not None:
            self.handle.close()
            self.handle = None


class scale_wind_vert:
    """Mirror of the reference module (src/modules/scale_wind_vert.cpp:27-229) on the device kernels of a PBSM3D handle.

    depends U_R, optional snowdepthavg, provides U_2m_above_srf; config key ``ignore_canopy`` (default false).  In CHM
    the module is data-parallel in point mode and domain-parallel otherwise (:140-142); ``point_mode`` selects which.
    The mesh substrate (geometry, neighbours, vegetation, halo plan) is the PBSM3D handle's, so it takes the module."""

    name = "scale_wind_vert"

    def __init__(self, cfg: Optional[dict] = None, point_mode: bool = False):
        cfg = dict(cfg or {})
        unknown = set(cfg) - {"ignore_canopy"}
        if unknown:
            raise module_error(f"scale_wind_vert: unknown config key(s) {sorted(unknown)}")
        self.wind_cfg_kw = dict(ignore_canopy=int(bool(PBSM3D._coerce(cfg.get("ignore_canopy", False)))), point_mode=int(point_mode))
        self._depends, self._optional, self._provides = ["U_R"], ["snowdepthavg"], ["U_2m_above_srf"]
        self.parallel = "data" if point_mode else "domain"

    def get_depends(self):
        return list(self._depends)

    def get_optionals(self):
        return list(self._optional)

    def get_provides(self):
        return list(self._provides)

    def init(self, domain: Domain):
        domain.init_face_data(self._provides)

    def run(self, domain: Domain, pbsm: PBSM3D):
        if pbsm.handle is None:
            raise module_error("scale_wind_vert::run needs an initialised PBSM3D handle (it owns the device mesh)")
        sd = domain["snowdepthavg"] if domain.has("snowdepthavg") else None  # has_optional("snowdepthavg")
        try:
            domain["U_2m_above_srf"] = pbsm.handle.scale_wind_vert(domain["U_R"], sd, capi.default_wind_config(**self.wind_cfg_kw))
        except capi.Pbsm3dError as e:
            raise module_error(str(e)) from e


class fetchr:
    """Mirror of the reference module (src/modules/fetchr.cpp:27-119): depends vw_dir, provides fetch; config keys
    ``steps`` (10), ``max_distance`` (1000), ``I`` (0.06), ``incl_veg`` (true)."""

    name = "fetchr"
    parallel = "data"

    def __init__(self, cfg: Optional[dict] = None):
        cfg = dict(cfg or {})
        unknown = set(cfg) - {"steps", "max_distance", "I", "incl_veg"}
        if unknown:
            raise module_error(f"fetchr: unknown config key(s) {sorted(unknown)}")
        c = {k: PBSM3D._coerce(v) for k, v in cfg.items()}
        self.wind_cfg_kw = dict(fetch_steps=int(c.get("steps", 10)), fetch_max_distance=float(c.get("max_distance", 1000.0)),
                                fetch_I=float(c.get("I", 0.06)), fetch_incl_veg=int(bool(c.get("incl_veg", True))))
        self._depends, self._provides = ["vw_dir"], ["fetch"]

    def get_depends(self):
        return list(self._depends)

    def get_provides(self):
        return list(self._provides)

    def init(self, domain: Domain):
        domain.init_face_data(self._provides)

    def run(self, domain: Domain, pbsm: PBSM3D):
        if pbsm.handle is None:
            raise module_error("fetchr::run needs an initialised PBSM3D handle (it owns the device mesh)")
        try:
            domain["fetch"] = pbsm.handle.fetchr(domain["vw_dir"], capi.default_wind_config(**self.wind_cfg_kw))
        except capi.Pbsm3dError as e:
            raise module_error(str(e)) from e


class snow_slide:
    """Mirror of the reference module (src/modules/snow_slide.cpp:27-446) on the device kernels of a PBSM3D handle.

    depends snowdepthavg, swe (and reads snowdepthavg_vert without declaring it, :121); provides the ten variables of :35-50
    (the ghost_ss_* ones are the reference's MPI scratch space: declared for the contract, never filled here — the exchange
    happens inside pbsm3d_slide_run); config keys ``use_vertical_snow`` (true; read and unused by the reference),
    ``avalache_mult`` (3178.4), ``avalache_pow`` (-1.998).  parallel::domain."""

    name = "snow_slide"
    parallel = "domain"
    OUTPUTS = ("delta_avalanche_snowdepth", "delta_avalanche_mass", "delta_avalanche_snowdepth_sum", "delta_avalanche_mass_sum", "maxDepth")

    def __init__(self, cfg: Optional[dict] = None):
        cfg = dict(cfg or {})
        unknown = set(cfg) - {"use_vertical_snow", "avalache_mult", "avalache_pow"}
        if unknown:
            raise module_error(f"snow_slide: unknown config key(s) {sorted(unknown)}")
        c = {k: PBSM3D._coerce(v) for k, v in cfg.items()}
        self.slide_cfg_kw = dict(avalache_mult=float(c.get("avalache_mult", 3178.4)), avalache_pow=float(c.get("avalache_pow", -1.998)),
                                 use_vertical_snow=int(bool(c.get("use_vertical_snow", True))))
        self._depends = ["snowdepthavg", "swe"]
        self._provides = ["delta_avalanche_mass", "delta_avalanche_snowdepth", "delta_avalanche_mass_sum", "delta_avalanche_snowdepth_sum",
                          "maxDepth", "ghost_ss_snowdepthavg_vert_copy", "ghost_ss_snowdepthavg_to_xfer", "ghost_ss_swe_to_xfer",
                          "ghost_ss_delta_avalanche_snowdepth", "ghost_ss_delta_avalanche_swe"]
        self.stats = None

    def get_depends(self):
        return list(self._depends)

    def get_provides(self):
        return list(self._provides)

    def init(self, domain: Domain, pbsm: PBSM3D):
        if pbsm.handle is None:
            raise module_error("snow_slide::init needs an initialised PBSM3D handle (it owns the device mesh)")
        domain.init_face_data(self._provides)
        try:
            pbsm.handle.slide_init(**self.slide_cfg_kw)
        except capi.Pbsm3dError as e:
            raise module_error(str(e)) from e
        domain["delta_avalanche_snowdepth_sum"] = 0.0   # snow_slide.cpp:442-443
        domain["delta_avalanche_mass_sum"] = 0.0
        self._first = True

    def run(self, domain: Domain, pbsm: PBSM3D):
        if pbsm.handle is None or not hasattr(self, "_first"):
            raise module_error("snow_slide::run before init")
        try:
            out, self.stats = pbsm.handle.slide_run(domain["snowdepthavg"], domain["snowdepthavg_vert"], domain["swe"])
        except capi.Pbsm3dError as e:
            raise module_error(str(e)) from e
        for n in self.OUTPUTS:
            domain[n] = out[n]

    def checkpoint(self, domain: Domain, pbsm: PBSM3D) -> Dict[str, np.ndarray]:
        """snow_slide.cpp:59-76."""
        s = pbsm.handle.slide_get_state()
        return {"snow_slide:" + k: v for k, v in s.items()}

    def load_checkpoint(self, domain: Domain, pbsm: PBSM3D, chk: Dict[str, np.ndarray]):
        """snow_slide.cpp:78-92."""
        pbsm.handle.slide_set_state(**{k.split(":", 1)[1]: v for k, v in chk.items()})
        domain["delta_avalanche_snowdepth_sum"] = chk["snow_slide:delta_avalanche_snowdepth_sum"]
        domain["delta_avalanche_mass_sum"] = chk["snow_slide:delta_avalanche_mass_sum"]

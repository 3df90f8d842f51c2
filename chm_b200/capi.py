"""ctypes binding of include/pbsm3d.h (libpbsm3d_b200.so).  Plumbing only: no arithmetic happens here.

This is the Python twin of the cgo/JNI-style stub INTEGRATION.md shows for CHM's C++ adaptor: it declares
exactly the symbols the header declares and fails loudly when the CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

from .mesh import TriMesh

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpbsm3d_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)

SOLVER_AUTO, SOLVER_LINE, SOLVER_BICGSTAB = 0, 1, 2
DEP_AUTO, DEP_CG, DEP_CHEBYSHEV, DEP_SOR = 0, 1, 2, 3
HALO_NONE, HALO_NCCL, HALO_PEER = 0, 1, 2
ABI_VERSION = 8  # include/pbsm3d.h PBSM3D_ABI_VERSION
ERR_NAMES = {1: "INVALID", 2: "UNSUPPORTED", 3: "CUDA", 4: "NCCL", 5: "NOCONVERGE"}


class Config(C.Structure):
    _fields_ = [
        ("nLayer", C.c_int), ("do_fixed_settling", C.c_int), ("settling_velocity", C.c_double),
        ("do_sublimation", C.c_int), ("do_lateral_diff", C.c_int), ("smooth_coeff", C.c_double),
        ("min_sd_trans", C.c_double), ("cutoff", C.c_double), ("snow_diffusion_const", C.c_double),
        ("rouault_diffusion_coef", C.c_int), ("enable_veg", C.c_int), ("iterative_subl", C.c_int),
        ("use_exp_fetch", C.c_int), ("use_tanh_fetch", C.c_int), ("use_PomLi_probability", C.c_int),
        ("z0_ustar_coupling", C.c_int), ("use_subgrid_topo", C.c_int), ("use_subgrid_topo_V2", C.c_int),
        ("use_R94_lambda", C.c_int), ("debug_output", C.c_int),
        ("tolerance", C.c_double), ("max_iterations", C.c_int), ("solver", C.c_int), ("deposition_solver", C.c_int),
        ("fp32_sweep_streams", C.c_int),
    ]


class Mesh(C.Structure):
    _fields_ = [
        ("n_global", C.c_int64), ("n_local", C.c_int32), ("n_ghost", C.c_int32),
        ("global_id", c_int64_p), ("ghost_owner", c_int32_p), ("neigh", c_int32_p), ("vertices", c_double_p),
        ("area", c_double_p), ("canopy_height", c_double_p), ("lai", c_double_p), ("stalk_number", c_double_p),
        ("stalk_diameter", c_double_p), ("is_water", c_uint8_p), ("is_geographic", C.c_int32),
    ]


class Comm(C.Structure):
    _fields_ = [("rank", C.c_int32), ("n_ranks", C.c_int32), ("nccl_unique_id", C.c_void_p)]


class Forcing(C.Structure):
    _fields_ = [(n, c_double_p) for n in ("U_R", "U_2m_above_srf", "snowdepthavg", "swe", "t", "rh", "vw_dir", "fetch", "p_snow_hours")]


class Outputs(C.Structure):
    _fields_ = [(n, c_double_p) for n in ("Qsalt", "Qsusp", "Qsubl", "Qsubl_mass", "sum_subl", "drift_mass", "sum_drift",
                                          "pbsm_more_than_avail", "blowingsnow_probability")]


class Stats(C.Structure):
    _fields_ = [
        ("suspension_present", C.c_int32), ("deposition_present", C.c_int32), ("suspension_iterations", C.c_int32),
        ("deposition_iterations", C.c_int32), ("suspension_solver_used", C.c_int32), ("kernel_launches", C.c_int32),
        ("suspension_residual", C.c_double), ("deposition_residual", C.c_double), ("suspension_rhs_max", C.c_double),
        ("deposition_rhs_max", C.c_double), ("ms_assembly", C.c_float), ("ms_suspension_solve", C.c_float),
        ("ms_flux_and_halo", C.c_float), ("ms_deposition", C.c_float), ("ms_total", C.c_float), ("ms_line_sweeps", C.c_float),
        ("sweeps_timed", C.c_int32), ("sweeps_timed_fp32", C.c_int32), ("ms_line_sweeps_fp32", C.c_float), ("n_colours", C.c_int32), ("deposition_solver_used", C.c_int32), ("host_syncs", C.c_int32),
        ("halo_exchanges", C.c_int32), ("halo_transport", C.c_int32), ("halo_fused", C.c_int32),
        ("residual_checks", C.c_int32), ("sweeps_fp32_x", C.c_int32), ("persistent_kernels", C.c_int32),
        ("active_set", C.c_int32), ("faces_with_rhs", C.c_int32), ("column_updates_fp32_x", C.c_int64), ("column_updates_fp32", C.c_int64),
        ("column_updates_fp64", C.c_int64), ("columns_checked", C.c_int64),
    ]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class WindConfig(C.Structure):
    """pbsm3d_wind_config: scale_wind_vert / fetchr config keys (scale_wind_vert.cpp:161, fetchr.cpp:34-43)."""
    _fields_ = [("ignore_canopy", C.c_int32), ("point_mode", C.c_int32), ("fetch_steps", C.c_int32), ("fetch_incl_veg", C.c_int32),
                ("fetch_max_distance", C.c_double), ("fetch_I", C.c_double)]


SNOWPACK_FIELDS = ("z_s", "m_s", "rho", "layer_count", "z_s_0", "z_s_l", "m_s_0", "m_s_l", "cc_s", "cc_s_0", "cc_s_l", "T_s", "T_s_0", "T_s_l",
                   "h2o_total", "h2o_vol", "h2o", "h2o_max", "h2o_sat")


class Snowpack(C.Structure):
    """pbsm3d_snowpack: SoA view of snobal's per-face `sno` members (third_party/snobal/sno.h)."""
    _fields_ = [(n, c_int32_p if n == "layer_count" else c_double_p) for n in SNOWPACK_FIELDS]


class SnobalConfig(C.Structure):
    """pbsm3d_snobal_config (snobal.cpp:83,101,190)."""
    _fields_ = [("drift_density", C.c_double), ("threshold", C.c_double), ("max_active_layer", C.c_double)]


class SlideConfig(C.Structure):
    """pbsm3d_slide_config (snow_slide.cpp:33,409-410)."""
    _fields_ = [("avalache_mult", C.c_double), ("avalache_pow", C.c_double), ("use_vertical_snow", C.c_int32)]


class SlideStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("wavefront_rounds", C.c_int32), ("faces_fired", C.c_int32), ("frontier_rounds", C.c_int32),
                ("live_faces", C.c_int32), ("ms_device", C.c_float)]


SLIDE_OUTPUTS = ("delta_avalanche_snowdepth", "delta_avalanche_mass", "delta_avalanche_snowdepth_sum", "delta_avalanche_mass_sum", "maxDepth")
FORCING_NAMES = [n for n, _ in Forcing._fields_]
OUTPUT_NAMES = [n for n, _ in Outputs._fields_]
# what the default path reads / writes (p_snow_hours and blowingsnow_probability belong to use_PomLi_probability)
DEFAULT_FORCING_NAMES = [n for n in FORCING_NAMES if n != "p_snow_hours"]
DEFAULT_OUTPUT_NAMES = [n for n in OUTPUT_NAMES if n != "blowingsnow_probability"]

# every symbol include/pbsm3d.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pbsm3d_abi_version": (C.c_int, []),
    "pbsm3d_last_error": (C.c_char_p, []),
    "pbsm3d_config_defaults": (None, [C.POINTER(Config)]),
    "pbsm3d_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "pbsm3d_host_alloc": (C.c_void_p, [C.c_size_t]),
    "pbsm3d_host_free": (None, [C.c_void_p]),
    "pbsm3d_create": (C.c_int, [C.POINTER(Config), C.POINTER(Mesh), C.c_int, C.POINTER(Comm), C.POINTER(C.c_void_p)]),
    "pbsm3d_destroy": (None, [C.c_void_p]),
    "pbsm3d_step": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(Forcing), C.POINTER(Outputs), C.POINTER(Stats)]),
    "pbsm3d_step_device": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(Forcing), C.POINTER(Outputs), C.POINTER(Stats)]),
    "pbsm3d_get_state": (C.c_int, [C.c_void_p] + [c_double_p] * 4),
    "pbsm3d_set_state": (C.c_int, [C.c_void_p] + [c_double_p] * 4),
    "pbsm3d_get_geometry": (C.c_int, [C.c_void_p] + [c_double_p] * 8),
    "pbsm3d_get_layout": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
    "pbsm3d_get_solution": (C.c_int, [C.c_void_p, c_double_p]),
    "pbsm3d_get_suspension_system": (C.c_int, [C.c_void_p] + [c_double_p] * 8 + [c_uint8_p]),
    "pbsm3d_get_deposition_system": (C.c_int, [C.c_void_p] + [c_double_p] * 4),
    "pbsm3d_time_kernel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "pbsm3d_wind_config_defaults": (None, [C.POINTER(WindConfig)]),
    "pbsm3d_scale_wind_vert": (C.c_int, [C.c_void_p, C.POINTER(WindConfig), c_double_p, c_double_p, c_double_p, C.c_int]),
    "pbsm3d_fetchr": (C.c_int, [C.c_void_p, C.POINTER(WindConfig), c_double_p, c_double_p, C.c_int]),
    "pbsm3d_set_providers": (C.c_int, [C.c_void_p, C.POINTER(WindConfig)]),
    "pbsm3d_slide_config_defaults": (None, [C.POINTER(SlideConfig)]),
    "pbsm3d_slide_init": (C.c_int, [C.c_void_p, C.POINTER(SlideConfig)]),
    "pbsm3d_slide_run": (C.c_int, [C.c_void_p] + [c_double_p] * 8 + [C.POINTER(SlideStats), C.c_int]),
    "pbsm3d_slide_get_constants": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "pbsm3d_slide_get_state": (C.c_int, [C.c_void_p] + [c_double_p] * 4),
    "pbsm3d_slide_set_state": (C.c_int, [C.c_void_p] + [c_double_p] * 4),
    "pbsm3d_snobal_config_defaults": (None, [C.POINTER(SnobalConfig)]),
    "pbsm3d_apply_drift": (C.c_int, [C.c_void_p, C.POINTER(SnobalConfig), C.POINTER(Snowpack), c_double_p, c_double_p, c_double_p, C.c_int]),
    "pbsm3d_apply_avalanche": (C.c_int, [C.c_void_p, C.POINTER(SnobalConfig), C.POINTER(Snowpack), c_double_p, c_double_p, c_double_p,
                                         c_double_p, C.c_int]),
}

_lib = None


class Pbsm3dError(RuntimeError):
    """What the CHM adaptor raises as module_error."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[{ERR_NAMES.get(code, code)}] {msg}")
        self.code = code


def load_library(path: Optional[str] = None):
    """dlopen the CUDA library.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(f"{p} not found: build it with `python -m chm_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.pbsm3d_abi_version() != ABI_VERSION:
        raise ImportError("pbsm3d ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _check(lib, rc: int):
    if rc != 0:
        raise Pbsm3dError(rc, lib.pbsm3d_last_error().decode())


def default_config(**overrides) -> Config:
    lib = load_library()
    cfg = Config()
    lib.pbsm3d_config_defaults(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise KeyError(f"unknown PBSM3D config key {k}")
        setattr(cfg, k, type(getattr(cfg, k))(v))
    return cfg


def default_wind_config(**overrides) -> WindConfig:
    lib = load_library()
    cfg = WindConfig()
    lib.pbsm3d_wind_config_defaults(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise KeyError(f"unknown wind config key {k}")
        setattr(cfg, k, type(getattr(cfg, k))(v))
    return cfg


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    _check(lib, lib.pbsm3d_nccl_unique_id(buf))
    return buf.raw


class Handle:
    """Owns one pbsm3d_handle (one rank, one GPU)."""

    def __init__(self, cfg: Config, mesh: TriMesh, device: int = 0, rank: int = 0, n_ranks: int = 1,
                 unique_id: Optional[bytes] = None, is_water: Optional[np.ndarray] = None):
        self.lib = load_library()
        self.T = mesh.n_local
        self.L = int(cfg.nLayer)
        self.pomli = bool(cfg.use_PomLi_probability)
        self._keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a

        m = Mesh()
        m.n_global = mesh.n_global
        m.n_local = mesh.n_local
        m.n_ghost = mesh.n_ghost
        m.global_id = arr(mesh.global_id, np.int64).ctypes.data_as(c_int64_p)
        m.ghost_owner = arr(mesh.ghost_owner, np.int32).ctypes.data_as(c_int32_p) if mesh.n_ghost else None
        m.neigh = arr(mesh.neigh, np.int32).ctypes.data_as(c_int32_p)
        m.vertices = _dp(arr(mesh.face_vertices(), np.float64))
        p = mesh.params
        m.area = _dp(arr(p["area"], np.float64)) if "area" in p else None
        m.canopy_height = _dp(arr(p["CanopyHeight"], np.float64)) if "CanopyHeight" in p else None
        m.lai = _dp(arr(p["LAI"], np.float64)) if "LAI" in p else None
        m.stalk_number = _dp(arr(p["stalk_number"], np.float64)) if "stalk_number" in p else None
        m.stalk_diameter = _dp(arr(p["stalk_diameter"], np.float64)) if "stalk_diameter" in p else None
        m.is_water = arr(is_water, np.uint8).ctypes.data_as(c_uint8_p) if is_water is not None else None
        m.is_geographic = 1 if getattr(mesh, "is_geographic", False) else 0
        comm = None
        if n_ranks > 1:
            self._uid = C.create_string_buffer(unique_id, 128)
            comm = Comm(rank, n_ranks, C.cast(self._uid, C.c_void_p))
        h = C.c_void_p()
        _check(self.lib, self.lib.pbsm3d_create(C.byref(cfg), C.byref(m), device, C.byref(comm) if comm else None, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.pbsm3d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ stepping
    def step(self, dt: float, forcing: Dict[str, np.ndarray], want=None):
        """Host-buffer entry point (what the CHM adaptor calls).  Returns (outputs dict, stats dict)."""
        if want is None:
            want = OUTPUT_NAMES if self.pomli else DEFAULT_OUTPUT_NAMES
        f = Forcing()
        keep = []
        for n in FORCING_NAMES:
            if n in forcing and forcing[n] is not None:
                a = np.ascontiguousarray(forcing[n], dtype=np.float64)
                if a.shape != (self.T,):
                    raise ValueError(f"forcing {n} must be [{self.T}]")
                keep.append(a)
                setattr(f, n, _dp(a))
        o = Outputs()
        outs = {n: np.empty(self.T) for n in want}
        for n, a in outs.items():
            setattr(o, n, _dp(a))
        st = Stats()
        _check(self.lib, self.lib.pbsm3d_step(self.h, dt, C.byref(f), C.byref(o), C.byref(st)))
        return outs, st.asdict()

    def step_ptr(self, dt: float, forcing_ptrs: Dict[str, int], output_ptrs: Dict[str, int], device: bool):
        """Raw-pointer entry (pinned host buffers or device buffers owned by the caller, e.g. torch tensors)."""
        f = Forcing()
        for n, p in forcing_ptrs.items():
            setattr(f, n, C.cast(C.c_void_p(p), c_double_p))
        o = Outputs()
        for n, p in output_ptrs.items():
            setattr(o, n, C.cast(C.c_void_p(p), c_double_p))
        st = Stats()
        fn = self.lib.pbsm3d_step_device if device else self.lib.pbsm3d_step
        _check(self.lib, fn(self.h, dt, C.byref(f), C.byref(o), C.byref(st)))
        return st.asdict()

    # ------------------------------------------------------------------ providers of PBSM3D inputs
    def scale_wind_vert(self, U_R, snowdepthavg=None, cfg: Optional[WindConfig] = None) -> np.ndarray:
        """scale_wind_vert::run on the device: U_R [+ snowdepthavg] -> U_2m_above_srf (host arrays)."""
        u = np.ascontiguousarray(U_R, dtype=np.float64)
        sd = None if snowdepthavg is None else np.ascontiguousarray(snowdepthavg, dtype=np.float64)
        out = np.empty(self.T)
        _check(self.lib, self.lib.pbsm3d_scale_wind_vert(self.h, C.byref(cfg) if cfg is not None else None, _dp(u), _dp(sd), _dp(out), 0))
        return out

    def fetchr(self, vw_dir, cfg: Optional[WindConfig] = None) -> np.ndarray:
        """fetchr::run for every face on the device: vw_dir -> fetch (host arrays)."""
        v = np.ascontiguousarray(vw_dir, dtype=np.float64)
        out = np.empty(self.T)
        _check(self.lib, self.lib.pbsm3d_fetchr(self.h, C.byref(cfg) if cfg is not None else None, _dp(v), _dp(out), 0))
        return out

    def set_providers(self, cfg: Optional[WindConfig]):
        """Fuse the providers into step(): U_2m_above_srf / fetch left out of the forcing are derived on the device."""
        _check(self.lib, self.lib.pbsm3d_set_providers(self.h, C.byref(cfg) if cfg is not None else None))

    # ------------------------------------------------------------------ the consumer of drift_mass (snobal's _adj_snow)
    def _snowpack(self, state: Dict[str, np.ndarray]):
        arrs = {}
        pk = Snowpack()
        for n in SNOWPACK_FIELDS:
            a = np.array(state[n], dtype=np.int32 if n == "layer_count" else np.float64, copy=True)
            if a.shape != (self.T,):
                raise ValueError(f"snowpack field {n} must be [{self.T}]")
            arrs[n] = a
            setattr(pk, n, a.ctypes.data_as(c_int32_p if n == "layer_count" else c_double_p))
        return pk, arrs

    def apply_drift(self, state: Dict[str, np.ndarray], drift_mass=None, cfg: Optional[SnobalConfig] = None):
        """snobal.cpp:363-387 for every face on the device.  drift_mass None = the handle's own (last step's, device-resident).
        Returns the new state (the input is not modified) with two extra keys `swe`, `snowdepthavg`."""
        pk, arrs = self._snowpack(state)
        dm = None if drift_mass is None else np.ascontiguousarray(drift_mass, dtype=np.float64)
        swe, sd = np.empty(self.T), np.empty(self.T)
        _check(self.lib, self.lib.pbsm3d_apply_drift(self.h, C.byref(cfg) if cfg is not None else None, C.byref(pk), _dp(dm), _dp(swe),
                                                     _dp(sd), 0))
        arrs.update(swe=swe, snowdepthavg=sd)
        return arrs

    def apply_avalanche(self, state: Dict[str, np.ndarray], delta_snowdepth, delta_mass, cfg: Optional[SnobalConfig] = None):
        """snobal.cpp:389-408 for every face on the device (snow_slide's volumes / the handle's face areas)."""
        pk, arrs = self._snowpack(state)
        dv = np.ascontiguousarray(delta_snowdepth, dtype=np.float64)
        dm = np.ascontiguousarray(delta_mass, dtype=np.float64)
        swe, sd = np.empty(self.T), np.empty(self.T)
        _check(self.lib, self.lib.pbsm3d_apply_avalanche(self.h, C.byref(cfg) if cfg is not None else None, C.byref(pk), _dp(dv), _dp(dm),
                                                         _dp(swe), _dp(sd), 0))
        arrs.update(swe=swe, snowdepthavg=sd)
        return arrs

    # ------------------------------------------------------------------ snow_slide
    def slide_init(self, **cfg_overrides):
        """snow_slide::init (maxDepth per face; the running sums start at 0)."""
        cfg = SlideConfig()
        self.lib.pbsm3d_slide_config_defaults(C.byref(cfg))
        for k, v in cfg_overrides.items():
            if not hasattr(cfg, k):
                raise KeyError(f"unknown snow_slide config key {k}")
            setattr(cfg, k, type(getattr(cfg, k))(v))
        _check(self.lib, self.lib.pbsm3d_slide_init(self.h, C.byref(cfg)))

    def slide_run(self, snowdepthavg, snowdepthavg_vert, swe):
        """snow_slide::run on host arrays [T].  Returns (outputs dict, stats dict)."""
        ins = [np.ascontiguousarray(a, dtype=np.float64) for a in (snowdepthavg, snowdepthavg_vert, swe)]
        for a in ins:
            if a.shape != (self.T,):
                raise ValueError(f"snow_slide inputs must be [{self.T}]")
        outs = {n: np.empty(self.T) for n in SLIDE_OUTPUTS}
        st = SlideStats()
        _check(self.lib, self.lib.pbsm3d_slide_run(self.h, *[_dp(a) for a in ins], *[_dp(outs[n]) for n in SLIDE_OUTPUTS], C.byref(st), 0))
        return outs, {n: getattr(st, n) for n, _ in st._fields_}

    def slide_constants(self):
        """(maxDepth, max(0.001, cos(slope))) per face as the device computed them in slide_init."""
        a, b = np.empty(self.T), np.empty(self.T)
        _check(self.lib, self.lib.pbsm3d_slide_get_constants(self.h, _dp(a), _dp(b)))
        return a, b

    def slide_get_state(self):
        s = {n: np.empty(self.T) for n in SLIDE_OUTPUTS[:4]}
        _check(self.lib, self.lib.pbsm3d_slide_get_state(self.h, *[_dp(s[n]) for n in SLIDE_OUTPUTS[:4]]))
        return s

    def slide_set_state(self, **arrays):
        arrs = [None if arrays.get(n) is None else np.ascontiguousarray(arrays[n], dtype=np.float64) for n in SLIDE_OUTPUTS[:4]]
        _check(self.lib, self.lib.pbsm3d_slide_set_state(self.h, *[_dp(a) for a in arrs]))

    # ------------------------------------------------------------------ inspection
    def geometry(self):
        T = self.T
        g = {k: np.empty((3, T)) for k in ("nx", "ny", "elen", "dx")}
        g.update({k: np.empty(T) for k in ("area", "cx", "cy", "cz")})
        _check(self.lib, self.lib.pbsm3d_get_geometry(self.h, _dp(g["nx"]), _dp(g["ny"]), _dp(g["elen"]), _dp(g["area"]),
                                                      _dp(g["dx"]), _dp(g["cx"]), _dp(g["cy"]), _dp(g["cz"])))
        return g

    def layout(self):
        """(n_colours, n_slots, slot_of_face[T], colour_of_face[T]) of the colour-major device order."""
        nc, ns = C.c_int32(), C.c_int32()
        slot = np.empty(self.T, dtype=np.int32)
        colour = np.empty(self.T, dtype=np.int32)
        _check(self.lib, self.lib.pbsm3d_get_layout(self.h, C.byref(nc), C.byref(ns), slot.ctypes.data_as(c_int32_p),
                                                    colour.ctypes.data_as(c_int32_p)))
        return int(nc.value), int(ns.value), slot, colour

    def solution(self) -> np.ndarray:
        x = np.empty((self.L, self.T))
        _check(self.lib, self.lib.pbsm3d_get_solution(self.h, _dp(x)))
        return x

    def suspension_system(self):
        L, T = self.L, self.T
        s = {k: np.empty((L, T)) for k in ("diag", "below", "above", "u_z", "csubl")}
        s["lat"] = np.empty((3, L, T))
        s["rhs0"] = np.empty(T)
        s["c_salt"] = np.empty(T)
        s["saltation"] = np.empty(T, dtype=np.uint8)
        _check(self.lib, self.lib.pbsm3d_get_suspension_system(
            self.h, _dp(s["diag"]), _dp(s["lat"]), _dp(s["below"]), _dp(s["above"]), _dp(s["rhs0"]), _dp(s["u_z"]),
            _dp(s["csubl"]), _dp(s["c_salt"]), s["saltation"].ctypes.data_as(c_uint8_p)))
        return s

    def deposition_system(self):
        T = self.T
        d = {"diag": np.empty(T), "off": np.empty((3, T)), "rhs": np.empty(T), "q": np.empty(T)}
        _check(self.lib, self.lib.pbsm3d_get_deposition_system(self.h, _dp(d["diag"]), _dp(d["off"]), _dp(d["rhs"]), _dp(d["q"])))
        return d

    def get_state(self):
        T = self.T
        s = {k: np.empty(T) for k in ("sum_drift", "sum_subl", "drift_mass", "pbsm_more_than_avail")}
        _check(self.lib, self.lib.pbsm3d_get_state(self.h, _dp(s["sum_drift"]), _dp(s["sum_subl"]), _dp(s["drift_mass"]),
                                                   _dp(s["pbsm_more_than_avail"])))
        return s

    def set_state(self, sum_drift=None, sum_subl=None, drift_mass=None, pbsm_more_than_avail=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64)
                for a in (sum_drift, sum_subl, drift_mass, pbsm_more_than_avail)]
        _check(self.lib, self.lib.pbsm3d_set_state(self.h, *[_dp(a) for a in arrs]))

    def time_kernel(self, kernel: int, reps: int = 20) -> float:
        ms = C.c_float()
        _check(self.lib, self.lib.pbsm3d_time_kernel(self.h, kernel, reps, C.byref(ms)))
        return float(ms.value)

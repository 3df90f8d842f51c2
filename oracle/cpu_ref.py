"""ctypes wrapper of oracle/pbsm3d_ref.cpp (C++/OpenMP restatement: the second oracle and the CPU baseline).
TEST INFRASTRUCTURE ONLY — see the header of pbsm3d_ref.cpp."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpbsm3d_ref.so")
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
up = C.POINTER(C.c_ubyte)


class RefConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nLayer", "do_fixed_settling", "do_sublimation", "do_lateral_diff", "rouault",
                                       "enable_veg", "use_exp_fetch", "use_tanh_fetch", "use_R94_lambda", "n_threads")] + \
               [(n, C.c_double) for n in ("settling_velocity", "smooth_coeff", "min_sd_trans", "cutoff", "snow_diffusion_const",
                                          "tolerance", "ilut_drop", "ilut_fill")] + \
               [("gmres_restart", C.c_int), ("max_iterations", C.c_int)]


class RefMesh(C.Structure):
    _fields_ = [("T", C.c_int), ("neigh", ip), ("nx", dp), ("ny", dp), ("elen", dp), ("dx", dp), ("area", dp), ("zc", dp),
                ("canopy", dp), ("lai", dp), ("stalk_n", dp), ("stalk_dv", dp), ("water", up)]


class RefForcing(C.Structure):
    _fields_ = [(n, dp) for n in ("U_R", "u2", "sd", "swe", "t", "rh", "vw_dir", "fetch")]


class RefOut(C.Structure):
    _fields_ = [(n, dp) for n in ("c", "Qsusp", "Qsalt", "Qsubl", "sum_drift", "sum_subl", "drift_mass", "more_avail",
                                  "diag", "lat", "below", "above", "rhs0", "u_z", "csubl", "c_salt")] + \
               [("salt", up)] + \
               [(n, C.c_int) for n in ("susp_present", "dep_present", "susp_iters", "dep_iters", "n_threads", "pad")] + \
               [(n, C.c_double) for n in ("susp_resid", "dep_resid", "s_assembly", "s_factor", "s_solve", "s_deposition", "s_total")]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "pbsm3d_ref.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libpbsm3d_ref.so"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.pbsm3d_ref_step.restype = C.c_int
        _lib.pbsm3d_ref_step.argtypes = [C.POINTER(RefConfig), C.POINTER(RefMesh), C.POINTER(RefForcing), C.c_double,
                                         C.POINTER(RefOut)]
        _lib.pbsm3d_ref_threads.restype = C.c_int
    return _lib


def host_threads() -> int:
    return int(lib().pbsm3d_ref_threads())


class CpuReference:
    """One global mesh, the reference algorithm on the host cores (GMRES(30) + rank-local ILUT)."""

    def __init__(self, ocfg, mesh, geo, is_water=None, n_threads: int = 0, dump_system: bool = False):
        self.L, self.T = int(ocfg.nLayer), mesh.n_local
        T, L = self.T, self.L
        self._k = []
        params = mesh.params
        has_veg = "CanopyHeight" in params
        veg = bool(ocfg.enable_veg and has_veg)
        c = RefConfig()
        c.nLayer, c.do_fixed_settling, c.do_sublimation = L, int(ocfg.do_fixed_settling), int(ocfg.do_sublimation)
        c.do_lateral_diff, c.rouault, c.enable_veg = int(ocfg.do_lateral_diff), int(ocfg.rouault_diffusion_coef), int(veg)
        c.use_exp_fetch, c.use_tanh_fetch, c.use_R94_lambda = int(ocfg.use_exp_fetch), int(ocfg.use_tanh_fetch), int(ocfg.use_R94_lambda)
        c.n_threads = n_threads
        c.settling_velocity, c.smooth_coeff, c.min_sd_trans = ocfg.settling_velocity, ocfg.smooth_coeff, ocfg.min_sd_trans
        c.cutoff, c.snow_diffusion_const = ocfg.cutoff, ocfg.snow_diffusion_const
        c.tolerance, c.ilut_drop, c.ilut_fill = 1e-8, 1e-4, 3.0  # LinearAlgebra.cpp:168,184-185
        c.gmres_restart, c.max_iterations = 30, 1000  # LinearAlgebra.cpp:166-167
        self.cfg = c
        m = RefMesh()
        m.T = T
        m.neigh = self._a(mesh.neigh, np.int32).ctypes.data_as(ip)
        for k in ("nx", "ny", "elen", "dx", "area"):
            setattr(m, k, self._d(getattr(geo, k)))
        m.zc = self._d(geo.cz[:T])
        if veg:
            m.canopy = self._d(params["CanopyHeight"])
            if ocfg.use_R94_lambda:
                m.lai = self._d(params["LAI"])
            else:
                if "stalk_number" in params:
                    m.stalk_n = self._d(params["stalk_number"])
                if "stalk_diameter" in params:
                    m.stalk_dv = self._d(params["stalk_diameter"])
        if is_water is not None:
            m.water = self._a(is_water, np.uint8).ctypes.data_as(up)
        self.mesh = m
        self.state = {k: np.zeros(T) for k in ("sum_drift", "sum_subl", "more_avail")}
        self.state["drift_mass"] = np.full(T, -9999.0)
        self.dump = dump_system

    def _a(self, a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        self._k.append(a)
        return a

    def _d(self, a):
        return self._a(a, np.float64).ctypes.data_as(dp)

    def step(self, F: Dict[str, np.ndarray], dt: float, tolerance: Optional[float] = None):
        T, L = self.T, self.L
        if tolerance is not None:
            self.cfg.tolerance = tolerance
        f = RefForcing()
        keep = []
        for cn, pn in (("U_R", "U_R"), ("u2", "U_2m_above_srf"), ("sd", "snowdepthavg"), ("swe", "swe"), ("t", "t"),
                       ("rh", "rh"), ("vw_dir", "vw_dir"), ("fetch", "fetch")):
            if pn in F:
                a = np.ascontiguousarray(F[pn], dtype=np.float64)
                keep.append(a)
                setattr(f, cn, a.ctypes.data_as(dp))
        o = RefOut()
        res = {"c": np.zeros((L, T)), "Qsusp": np.zeros(T), "Qsalt": np.zeros(T), "Qsubl": np.zeros(T)}
        for k, a in res.items():
            setattr(o, k, a.ctypes.data_as(dp))
        for k, a in self.state.items():
            setattr(o, k, a.ctypes.data_as(dp))
        if self.dump:
            sysd = {k: np.zeros((L, T)) for k in ("diag", "below", "above", "u_z", "csubl")}
            sysd["lat"] = np.zeros((3, L, T))
            sysd["rhs0"] = np.zeros(T)
            sysd["c_salt"] = np.zeros(T)
            for k, a in sysd.items():
                setattr(o, k, a.ctypes.data_as(dp))
            salt = np.zeros(T, dtype=np.uint8)
            o.salt = salt.ctypes.data_as(up)
            sysd["saltation"] = salt
            res["system"] = sysd
        rc = lib().pbsm3d_ref_step(C.byref(self.cfg), C.byref(self.mesh), C.byref(f), dt, C.byref(o))
        if rc:
            raise RuntimeError("pbsm3d_ref_step failed")
        res.update({k: v.copy() for k, v in self.state.items()})
        res["stats"] = {n: getattr(o, n) for n in ("susp_present", "dep_present", "susp_iters", "dep_iters", "n_threads",
                                                   "susp_resid", "dep_resid", "s_assembly", "s_factor", "s_solve",
                                                   "s_deposition", "s_total")}
        return res

"""ctypes wrapper of oracle/_ref/libchmref.so — the reference's own PBSM3D.cpp / Atmosphere.cpp / coordinates.cpp,
compiled unmodified from /root/reference against the stand-in headers in oracle/refbuild/stubs (see the Makefile there).

TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/): it pins oracle/pbsm3d_oracle.py and generates the golden
vectors under tests/golden/ref_*.npz.  /root/reference exists only in the build container; the built .so is git-ignored
but travels to the GPU box with the gpurun snapshot.  `available()` says whether the library is there.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libchmref.so")
REF = os.environ.get("CHM_REFERENCE", "/root/reference")

_lib = None
_SOLVE_T = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                       C.POINTER(C.c_double), C.POINTER(C.c_double))


def build(force: bool = False) -> Optional[str]:
    """Compile the reference sources where they lie (needs /root/reference); returns the .so path or None."""
    if not os.path.isdir(os.path.join(REF, "src", "modules")):
        return LIB if os.path.exists(LIB) else None
    args = ["make", "-C", os.path.join(HERE, "refbuild"), f"REF={REF}"] + (["-B"] if force else [])
    res = subprocess.run(args, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building oracle/_ref/libchmref.so failed:\n" + res.stdout + res.stderr)
    return LIB


def available() -> bool:
    return os.path.exists(LIB)


@_SOLVE_T
def _direct_solve(n, rowptr, col, val, rhs, x):
    """NearestNeighborProblem::Solve's contract (A x = b) by a sparse direct factorisation."""
    try:
        rp = np.ctypeslib.as_array(rowptr, shape=(n + 1,))
        nnz = int(rp[n])
        A = sp.csr_matrix((np.ctypeslib.as_array(val, shape=(nnz,)).copy(), np.ctypeslib.as_array(col, shape=(nnz,)).copy(),
                           rp.copy()), shape=(n, n))
        b = np.ctypeslib.as_array(rhs, shape=(n,))
        np.ctypeslib.as_array(x, shape=(n,))[:] = spla.splu(A.tocsc()).solve(b)
        return 0
    except Exception:  # pragma: no cover
        return 1


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libchmref.so is missing (build it in the container that has /root/reference)")
        L = C.CDLL(LIB)
        L.chmref_create.restype = C.c_void_p
        L.chmref_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_char_p),
                                    C.c_void_p, C.c_char_p, C.c_char_p]
        L.chmref_last_error.restype = C.c_char_p
        for f in ("chmref_set_var", "chmref_get_var"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.chmref_run.argtypes = [C.c_void_p, C.c_double]
        L.chmref_destroy.argtypes = [C.c_void_p]
        L.chmref_system_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.chmref_system.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        for f in ("chmref_n_depends", "chmref_n_provides"):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("chmref_depend", "chmref_provide"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
            getattr(L, f).restype = C.c_char_p
        L.chmref_checkpoint.argtypes = [C.c_void_p, C.c_void_p]
        L.chmref_load_checkpoint.argtypes = [C.c_void_p, C.c_void_p]
        L.chmref_log_scale_wind.restype = C.c_double
        L.chmref_log_scale_wind.argtypes = [C.c_double] * 5
        L.chmref_saturatedVapourPressure.restype = C.c_double
        L.chmref_saturatedVapourPressure.argtypes = [C.c_double]
        L.chmref_bearing_to_cartesian.argtypes = [C.c_double, C.c_void_p]
        L.chmref_distance_UTM.restype = C.c_double
        L.chmref_distance_UTM.argtypes = [C.c_void_p, C.c_void_p]
        L.chmref_run_scale_wind_vert.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.chmref_run_fetchr.argtypes = [C.c_void_p, C.c_char_p]
        L.chmref_tpspline.restype = C.c_double
        L.chmref_tpspline.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.chmref_set_solver(_direct_solve)
        _lib = L
    return _lib


def _cfg_text(cfg: Dict) -> bytes:
    out = []
    for k, v in cfg.items():
        if isinstance(v, (bool, np.bool_)):
            v = "true" if v else "false"
        elif isinstance(v, (int, np.integer)):
            v = str(int(v))
        else:
            v = repr(float(v))
        out.append(f"{k}={v}")
    return "\n".join(out).encode()


class ReferencePBSM3D:
    """The reference module on one (unpartitioned) mesh.

    vertex [V,3], elem [T,3], neigh [T,3] as in CHM's .mesh files; params: name -> [T] (NaN = absent on that face);
    cfg: PBSM3D config keys; landcover: {"landcover.<id>.<key>": value} global parameter table."""

    OUTPUTS = ("Qsalt", "Qsusp", "Qsubl", "Qsubl_mass", "sum_subl", "drift_mass", "sum_drift", "pbsm_more_than_avail")

    def __init__(self, vertex, elem, neigh, params: Optional[Dict[str, np.ndarray]], cfg: Dict, landcover: Optional[Dict] = None):
        L = lib()
        vertex = np.asarray(vertex, dtype=np.float64)
        elem = np.asarray(elem, dtype=np.int64)
        self.T = T = elem.shape[0]
        self.L = int(cfg.get("nLayer", 10))
        v = vertex[elem]  # [T,3,3]
        vx, vy, vz = (np.ascontiguousarray(v[:, :, k]) for k in range(3))
        ng = np.ascontiguousarray(np.asarray(neigh, dtype=np.int32))
        params = params or {}
        names = list(params)
        pv = np.ascontiguousarray(np.stack([np.asarray(params[n], dtype=np.float64) for n in names])) if names else np.zeros((0, T))
        arr = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
        self.h = L.chmref_create(T, vx.ctypes.data, vy.ctypes.data, vz.ctypes.data, ng.ctypes.data, len(names), arr,
                                 pv.ctypes.data, _cfg_text(cfg), _cfg_text(landcover or {}))
        if not self.h:
            raise RuntimeError("reference PBSM3D init failed: " + L.chmref_last_error().decode())

    def close(self):
        if self.h:
            lib().chmref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def depends(self):
        L = lib()
        return [L.chmref_depend(self.h, i).decode() for i in range(L.chmref_n_depends(self.h))]

    def provides(self):
        L = lib()
        return [L.chmref_provide(self.h, i).decode() for i in range(L.chmref_n_provides(self.h))]

    def set_var(self, name, vals):
        a = np.ascontiguousarray(np.asarray(vals, dtype=np.float64))
        assert a.shape == (self.T,)
        lib().chmref_set_var(self.h, name.encode(), a.ctypes.data)

    def get_var(self, name):
        out = np.empty(self.T)
        lib().chmref_get_var(self.h, name.encode(), out.ctypes.data)
        return out

    def system(self, which: int):
        """(A csr in the reference numbering layer*G+global_id, rhs, solution) of the last run(); 0 suspension, 1 deposition."""
        L = lib()
        n, nnz = C.c_int(), C.c_int()
        L.chmref_system_size(self.h, which, C.byref(n), C.byref(nnz))
        rp = np.empty(n.value + 1, dtype=np.int32)
        col = np.empty(nnz.value, dtype=np.int64)
        val = np.empty(nnz.value)
        rhs = np.empty(n.value)
        sol = np.empty(n.value)
        L.chmref_system(self.h, which, rp.ctypes.data, col.ctypes.data, val.ctypes.data, rhs.ctypes.data, sol.ctypes.data)
        A = sp.csr_matrix((val, col, rp), shape=(n.value, n.value))
        A.sum_duplicates()
        return A, rhs, sol

    def step(self, F: Dict[str, np.ndarray], dt: float) -> Dict[str, np.ndarray]:
        """PBSM3D::run(mesh&) on the forcing F (the variables the module depends on + snowdepthavg)."""
        for k, v in F.items():
            self.set_var(k, v)
        if lib().chmref_run(self.h, float(dt)) != 0:
            raise RuntimeError("reference PBSM3D::run threw: " + lib().chmref_last_error().decode())
        out = {k: self.get_var(k) for k in self.OUTPUTS}
        A, b, x = self.system(0)
        out["c"] = x.reshape(self.L, self.T)
        out["susp"] = (A, b)
        Ad, bd, q = self.system(1)
        out["dep"] = (Ad, bd)
        out["q_dep"] = q
        return out

    def scale_wind_vert(self, U_R, snowdepthavg=None, cfg: Optional[Dict] = None, point_only: bool = False):
        """The reference's scale_wind_vert (scale_wind_vert.cpp, compiled unmodified) on this mesh: domain mode
        (point_scale + neighbour thin plate spline) or, with point_only, run(face) per face.  Returns U_2m_above_srf."""
        self.set_var("U_R", U_R)
        if snowdepthavg is not None:
            self.set_var("snowdepthavg", snowdepthavg)
        if lib().chmref_run_scale_wind_vert(self.h, _cfg_text(cfg or {}), int(snowdepthavg is not None), int(point_only)) != 0:
            raise RuntimeError("reference scale_wind_vert threw: " + lib().chmref_last_error().decode())
        return self.get_var("U_2m_above_srf")

    def fetchr(self, vw_dir, cfg: Optional[Dict] = None):
        """The reference's fetchr (fetchr.cpp, compiled unmodified), run(face) for every face.  Returns fetch."""
        self.set_var("vw_dir", vw_dir)
        if lib().chmref_run_fetchr(self.h, _cfg_text(cfg or {})) != 0:
            raise RuntimeError("reference fetchr threw: " + lib().chmref_last_error().decode())
        return self.get_var("fetch")

    def checkpoint(self):
        out = np.empty(self.T)
        lib().chmref_checkpoint(self.h, out.ctypes.data)
        return out

    def load_checkpoint(self, sum_drift):
        a = np.ascontiguousarray(np.asarray(sum_drift, dtype=np.float64))
        lib().chmref_load_checkpoint(self.h, a.ctypes.data)


def tpspline(sample_xyv, query_xy):
    """stubs/interpolation.hpp (the C++ restatement of TPSpline.cpp the compiled scale_wind_vert.cpp calls)."""
    a = np.ascontiguousarray(np.asarray(sample_xyv, dtype=np.float64))
    q = np.ascontiguousarray(np.asarray(query_xy, dtype=np.float64))
    return lib().chmref_tpspline(a.shape[0], a.ctypes.data, q.ctypes.data)


def log_scale_wind(u, Z_in, Z_out, sd, z0=0.01):
    return lib().chmref_log_scale_wind(u, Z_in, Z_out, sd, z0)


def saturated_vapour_pressure(t_kelvin):
    return lib().chmref_saturatedVapourPressure(t_kelvin)


def bearing_to_cartesian(bearing):
    xy = np.empty(2)
    lib().chmref_bearing_to_cartesian(float(bearing), xy.ctypes.data)
    return xy[0], xy[1]


def distance_utm(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return lib().chmref_distance_UTM(a.ctypes.data, b.ctypes.data)

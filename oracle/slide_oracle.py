"""CPU restatement of CHM's snow_slide module: gravitational redistribution of snow that exceeds a slope-
dependent holding depth, highest surface first.  TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/).

    src/modules/snow_slide.cpp:406-446   init: maxDepth = max(mult * slopeDeg^pow, CanopyHeight) * max(0.001, cos(slope))
                              :94-404    run: per outer iteration — copies of snowdepthavg / snowdepthavg_vert / swe (first iteration
                                         only), faces sorted by centre elevation + vertical snow depth (descending), ONE SEQUENTIAL
                                         SWEEP in that order: a face whose (slope-normal) depth exceeds maxDepth sheds the excess to
                                         its lower neighbours, weights = elevation differences; a missing neighbour takes its share
                                         out of the domain; a ghost neighbour's share goes into ghost accumulators that travel back
                                         to the owner (ghost_to_neighbors_communicate_variable, triangulation.cpp:2081-2186), and
                                         another outer iteration follows while any rank received such a share (<= 26 iterations)
    src/mesh/triangulation.hpp:1501-1523, 1549-1574   face slope = acos(nz of the unit normal)

Behaviours kept as written: the receiver's vertical depth is recomputed with the DONOR's slope (:297); in outer iterations after
the first the running per-run deltas are added to the *_sum variables again (:356-357); faces with equal sort keys have no defined
order (tbb::parallel_sort) — here ties fall back to the face index, and fixtures keep neighbouring keys distinct.

Pinned to the reference's own code: oracle/_ref/libchmref.so contains snow_slide.cpp compiled unmodified (oracle/refbuild);
`ReferenceSlide` drives it; tests/golden/golden_slide.npz holds its outputs (tests/golden/make_golden_slide.py).  The multi-rank
composition (reverse exchange + outer iterations) is restated here from the source and is NOT pinned by a reference run (no MPI in
this image); what is pinned about it is the rank-local sweep with ghost neighbours, which the compiled reference runs as is.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import numpy as np

DEFAULTS = dict(avalache_mult=3178.4, avalache_pow=-1.998)  # snow_slide.cpp:409-410 (the reference's spelling)


def face_slope(vertices: np.ndarray) -> np.ndarray:
    """vertices [T,3,3] -> slope [T] (rad).  CGAL::unit_normal(v0,v1,v2) then acos(norm_dot(n, ez))."""
    a = vertices[:, 1] - vertices[:, 0]
    b = vertices[:, 2] - vertices[:, 0]
    nx = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    ny = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
    nz = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    ln = np.sqrt(nx * nx + ny * ny + nz * nz)
    nx, ny, nz = nx / ln, ny / ln, nz / ln
    dot = nx * 0.0 + ny * 0.0 + nz * 1.0
    na = np.sqrt(nx * nx + ny * ny + nz * nz)
    return np.arccos(dot / (na * 1.0))


def max_depth(slope: np.ndarray, canopy: Optional[np.ndarray], mult: float = 3178.4, power: float = -1.998) -> np.ndarray:
    zc = np.zeros_like(slope) if canopy is None else np.asarray(canopy, dtype=np.float64)
    slope_deg = np.maximum(10.0, slope * 180 / math.pi)
    return np.maximum(mult * np.power(slope_deg, power), zc) * np.maximum(0.001, np.cos(slope))


class SlideState:
    """Per-face arrays of one rank: T owned faces (+ nG ghost neighbours)."""

    def __init__(self, vertices, neigh, area, canopy=None, ghost_vertices=None, ghost_area=None, cfg: Optional[Dict] = None):
        cfg = dict(DEFAULTS, **(cfg or {}))
        self.T = T = neigh.shape[0]
        self.neigh = np.asarray(neigh, dtype=np.int64)          # -1 none, >= T ghost (index T+g)
        self.cz = (vertices[:, 0, 2] + vertices[:, 1, 2] + vertices[:, 2, 2]) / 3
        self.area = np.asarray(area, dtype=np.float64)
        self.slope = face_slope(vertices)
        self.cosf = np.maximum(0.001, np.cos(self.slope))
        self.maxDepth = max_depth(self.slope, canopy, cfg["avalache_mult"], cfg["avalache_pow"])
        self.nG = 0 if ghost_vertices is None else ghost_vertices.shape[0]
        if self.nG:
            self.g_cz = (ghost_vertices[:, 0, 2] + ghost_vertices[:, 1, 2] + ghost_vertices[:, 2, 2]) / 3
            self.g_area = np.asarray(ghost_area, dtype=np.float64)
        self.sum_sd = np.zeros(T)     # delta_avalanche_snowdepth_sum
        self.sum_mass = np.zeros(T)   # delta_avalanche_mass_sum

    # ---- one outer iteration's sequential sweep on this rank (snow_slide.cpp:171-330)
    def sweep(self, sd, sdv, swe, dsd, dmass, g_sdv):
        T = self.T
        g_sd_x = np.zeros(self.nG); g_swe_x = np.zeros(self.nG); g_dsd = np.zeros(self.nG); g_dswe = np.zeros(self.nG)
        key = self.cz + sdv
        order = np.lexsort((np.arange(T), -key))   # descending key, ties by index
        nb = self.neigh
        maxD, cz, area, cosf = self.maxDepth, self.cz, self.area, self.cosf
        cand = sd > maxD
        # a face can only fire if it exceeds maxDepth at its turn; only faces that are candidates or neighbours of firing faces can
        for f in order:
            if not sd[f] > maxD[f]:
                continue
            snow, snow_v, w_e = sd[f], sdv[f], swe[f]
            del_depth = snow - maxD[f]
            del_swe = w_e * (1 - maxD[f] / snow)
            orig_mass = del_swe * area[f]
            z_s = cz[f] + snow_v
            w = [0.0, 0.0, 0.0]
            w_dem = 0.0
            for j in range(3):
                n = nb[f, j]
                if n < 0:
                    w[j] = max(0.0, z_s - cz[f])
                elif n >= T:
                    w[j] = max(0.0, z_s - (self.g_cz[n - T] + g_sdv[n - T]))
                else:
                    w[j] = max(0.0, z_s - (cz[n] + sdv[n]))
                w_dem += w[j]
            if w_dem == 0:
                continue
            w = [x / w_dem for x in w]
            out_mass = 0.0
            for j in range(3):
                n = nb[f, j]
                if n < 0:
                    out_mass += del_swe * area[f] * w[j]
                    continue
                n_area = self.g_area[n - T] if n >= T else area[n]
                d_sd = del_depth * (area[f] / n_area) * w[j]
                d_swe = del_swe * (area[f] / n_area) * w[j]
                d_sd_m3 = del_depth * area[f] * w[j]
                d_swe_m3 = del_swe * area[f] * w[j]
                if n >= T:
                    g = n - T
                    g_sd_x[g] += d_sd; g_swe_x[g] += d_swe; g_dsd[g] += d_sd_m3; g_dswe[g] += d_swe_m3
                else:
                    sd[n] += d_sd
                    swe[n] += d_swe
                    sdv[n] = sd[n] / cosf[f]          # the DONOR's slope (snow_slide.cpp:297)
                    dsd[n] += d_sd_m3
                    dmass[n] += d_swe_m3
                out_mass += del_swe * area[f] * w[j]
            sd[f] = maxD[f]
            sdv[f] = sd[f] / cosf[f]
            swe[f] = w_e * maxD[f] / snow
            dsd[f] -= del_depth * area[f]
            dmass[f] -= del_swe * area[f]
            if abs(orig_mass - out_mass) > 0.0001:
                raise RuntimeError("Snowslide did not conserve mass")
        return g_sd_x, g_swe_x, g_dsd, g_dswe

    # ---- after the (reverse) exchange: snow_slide.cpp:343-358
    def absorb(self, sd, sdv, swe, dsd, dmass, x_sd, x_swe, x_dsd, x_dswe):
        sd += x_sd
        sdv += x_sd / self.cosf
        swe += x_swe
        dsd += x_dsd
        dmass += x_dswe
        self.sum_sd += dsd
        self.sum_mass += dmass
        return int(np.count_nonzero(x_dsd > 0))


def run_single(state: SlideState, snowdepthavg, snowdepthavg_vert, swe_mm) -> Dict[str, np.ndarray]:
    """snow_slide::run on one rank without ghosts: exactly one outer iteration."""
    sd = np.array(snowdepthavg, dtype=np.float64, copy=True)
    sdv = np.array(snowdepthavg_vert, dtype=np.float64, copy=True)
    swe = np.asarray(swe_mm, dtype=np.float64) / 1000.0
    dsd = np.zeros(state.T); dmass = np.zeros(state.T)
    state.sweep(sd, sdv, swe, dsd, dmass, np.zeros(0))
    z = np.zeros(state.T)
    state.absorb(sd, sdv, swe, dsd, dmass, z, z, z, z)
    return dict(delta_avalanche_snowdepth=dsd, delta_avalanche_mass=dmass, delta_avalanche_snowdepth_sum=state.sum_sd.copy(),
                delta_avalanche_mass_sum=state.sum_mass.copy(), maxDepth=state.maxDepth.copy(), snowdepthavg_copy=sd,
                snowdepthavg_vert_copy=sdv, swe_copy=swe, iterations=1)


def run_partitioned(states: List[SlideState], ghost_owner: List[np.ndarray], ghost_owner_local: List[np.ndarray], sd_in, sdv_in, swe_in):
    """All ranks of a partitioned mesh, emulated in one process (snow_slide.cpp:94-404 under USE_MPI).
    ghost_owner[r][g] / ghost_owner_local[r][g]: owning rank and the owner's local index of rank r's ghost g.
    sd_in / sdv_in / swe_in: per-rank input arrays.  Returns per-rank output dicts."""
    P = len(states)
    sd = [np.array(a, dtype=np.float64, copy=True) for a in sd_in]
    sdv = [np.array(a, dtype=np.float64, copy=True) for a in sdv_in]
    swe = [np.asarray(a, dtype=np.float64) / 1000.0 for a in swe_in]
    dsd = [np.zeros(s.T) for s in states]
    dmass = [np.zeros(s.T) for s in states]
    iterations = 0
    while True:
        # forward halo of snowdepthavg_vert_copy (owner -> ghost)
        g_sdv = [sdv[0][:0]] * P
        for r in range(P):
            g_sdv[r] = np.array([sdv[ghost_owner[r][g]][ghost_owner_local[r][g]] for g in range(states[r].nG)])
        acc = [states[r].sweep(sd[r], sdv[r], swe[r], dsd[r], dmass[r], g_sdv[r]) for r in range(P)]
        # reverse exchange: the owner's variable is SET to what the partner holds on its ghost copy; partners in ascending rank
        # order, so when two ranks hold the same face as a ghost the higher rank's value stays (triangulation.cpp:2172-2182)
        recv = [[np.zeros(s.T) for _ in range(4)] for s in states]
        for q in range(P):                      # sender (ghost holder), ascending = the order the owner unpacks its partners
            for g in range(states[q].nG):
                o, l = ghost_owner[q][g], ghost_owner_local[q][g]
                for k in range(4):
                    recv[o][k][l] = acc[q][k][g]
        moved = 0
        for r in range(P):
            moved += states[r].absorb(sd[r], sdv[r], swe[r], dsd[r], dmass[r], *recv[r])
        iterations += 1
        done = moved == 0
        if not done and iterations > 25:
            done = True
        if done:
            break
    return [dict(delta_avalanche_snowdepth=dsd[r], delta_avalanche_mass=dmass[r], delta_avalanche_snowdepth_sum=states[r].sum_sd.copy(),
                 delta_avalanche_mass_sum=states[r].sum_mass.copy(), maxDepth=states[r].maxDepth.copy(), snowdepthavg_copy=sd[r],
                 snowdepthavg_vert_copy=sdv[r], swe_copy=swe[r], iterations=iterations) for r in range(P)]


def synthetic_snow(cx, cy, slope, seed: int = 3, deep: float = 4.0):
    """(snowdepthavg, snowdepthavg_vert, swe[mm]): a smooth snow cover, deep enough on the steep faces for slides to start."""
    rng = np.random.default_rng(seed)
    k = rng.uniform(0.5, 2.0, 4) / 2500.0
    ph = rng.uniform(0, 2 * np.pi, 4)
    f = 0.5 + 0.25 * (np.sin(k[0] * cx + ph[0]) * np.cos(k[1] * cy + ph[1]) + np.sin(k[2] * (cx + cy) + ph[2]) * np.cos(k[3] * (cx - cy) + ph[3]))
    sd = deep * np.clip(f, 0.02, None) * (1.0 + 0.05 * rng.standard_normal(cx.shape[0]))
    sd = np.where(rng.random(cx.shape[0]) < 0.05, 0.0, np.abs(sd))
    rho = rng.uniform(150.0, 420.0, cx.shape[0])
    return sd, sd / np.maximum(0.001, np.cos(slope)), sd * rho   # swe in mm = kg/m^2


# ------------------------------------------------------------------ the reference's own snow_slide.cpp (oracle/_ref/libchmref.so)
class ReferenceSlide:
    """snow_slide.cpp compiled unmodified, on one rank's view of a mesh: owned faces + optional ghost neighbours."""

    VARS = ("delta_avalanche_snowdepth", "delta_avalanche_mass", "delta_avalanche_snowdepth_sum", "delta_avalanche_mass_sum", "maxDepth")
    GHOST_VARS = ("ghost_ss_snowdepthavg_to_xfer", "ghost_ss_swe_to_xfer", "ghost_ss_delta_avalanche_snowdepth", "ghost_ss_delta_avalanche_swe")

    def __init__(self, vertex, elem, neigh, params=None, cfg: Optional[Dict] = None, ghosts: Optional[Dict] = None):
        from . import chm_ref
        self._cr = chm_ref
        L = chm_ref.lib()
        L.chmref_add_ghosts.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3
        for f in ("chmref_set_ghost_var", "chmref_get_ghost_var"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.chmref_slide_init.argtypes = [C.c_void_p, C.c_char_p]
        L.chmref_slide_run.argtypes = [C.c_void_p]
        L.chmref_slide_checkpoint.argtypes = [C.c_void_p, C.c_void_p]
        L.chmref_face_slope.argtypes = [C.c_void_p, C.c_int]
        L.chmref_face_slope.restype = C.c_double
        for f in ("chmref_slide_n_depends", "chmref_slide_n_provides"):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("chmref_slide_depend", "chmref_slide_provide"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
            getattr(L, f).restype = C.c_char_p
        neigh = np.asarray(neigh, dtype=np.int64)
        T = neigh.shape[0]
        nb = np.where(neigh >= T, -1, neigh)
        self.base = chm_ref.ReferencePBSM3D(vertex, elem, nb, params, {"nLayer": 2, "use_R94_lambda": False})  # PBSM3D only hosts the mesh here
        self.T = T
        self.nG = 0
        if ghosts is not None:
            gv = np.ascontiguousarray(ghosts["vertices"], dtype=np.float64)   # [nG,3,3]
            self.nG = gv.shape[0]
            vx, vy, vz = (np.ascontiguousarray(gv[:, :, k]) for k in range(3))
            ga = np.ascontiguousarray(ghosts["area"], dtype=np.float64)
            ff, jj = np.nonzero(neigh >= T)
            af, ae = np.ascontiguousarray(ff, dtype=np.int32), np.ascontiguousarray(jj, dtype=np.int32)
            ag = np.ascontiguousarray(neigh[ff, jj] - T, dtype=np.int32)
            L.chmref_add_ghosts(self.base.h, self.nG, vx.ctypes.data, vy.ctypes.data, vz.ctypes.data, ga.ctypes.data, len(af),
                                af.ctypes.data, ae.ctypes.data, ag.ctypes.data)
        if L.chmref_slide_init(self.base.h, chm_ref._cfg_text(cfg or {})) != 0:
            raise RuntimeError("reference snow_slide init failed: " + L.chmref_last_error().decode())

    def depends(self):
        L = self._cr.lib()
        return [L.chmref_slide_depend(self.base.h, i).decode() for i in range(L.chmref_slide_n_depends(self.base.h))]

    def provides(self):
        L = self._cr.lib()
        return [L.chmref_slide_provide(self.base.h, i).decode() for i in range(L.chmref_slide_n_provides(self.base.h))]

    def slope(self):
        L = self._cr.lib()
        return np.array([L.chmref_face_slope(self.base.h, i) for i in range(self.T)])

    def run(self, snowdepthavg, snowdepthavg_vert, swe_mm, ghost_sdv=None):
        L = self._cr.lib()
        self.base.set_var("snowdepthavg", snowdepthavg)
        self.base.set_var("snowdepthavg_vert", snowdepthavg_vert)
        self.base.set_var("swe", swe_mm)
        if self.nG:
            a = np.ascontiguousarray(ghost_sdv, dtype=np.float64)
            L.chmref_set_ghost_var(self.base.h, b"ghost_ss_snowdepthavg_vert_copy", a.ctypes.data)
        if L.chmref_slide_run(self.base.h) != 0:
            raise RuntimeError("reference snow_slide::run threw: " + L.chmref_last_error().decode())
        out = {k: self.base.get_var(k) for k in self.VARS}
        for k in self.GHOST_VARS:
            g = np.empty(self.nG)
            if self.nG:
                L.chmref_get_ghost_var(self.base.h, k.encode(), g.ctypes.data)
            out[k] = g
        return out

    def checkpoint(self):
        out = np.empty((4, self.T))
        self._cr.lib().chmref_slide_checkpoint(self.base.h, out.ctypes.data)
        return out

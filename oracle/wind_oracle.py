"""CPU oracle (numpy, fp64) for the two per-face providers of PBSM3D inputs.
TEST INFRASTRUCTURE ONLY: only tests/ may import this file; the product (chm_b200/) never does.

* ``scale_wind_vert``  U_R (50 m) -> U_2m_above_srf          src/modules/scale_wind_vert.cpp:48-229
* ``thin_plate_spline``  the interpolant its domain mode uses  src/interpolation/TPSpline.cpp:40-173
* ``fetchr``  upwind fetch (Lapen & Martz 1993)               src/modules/fetchr.cpp:54-119

PARITY STATUS.  ``thin_plate_spline`` is pinned on the reference's own known-answer tests
(src/tests/test_interpolation.cpp:47-170: the 3-point case to ASSERT_DOUBLE_EQ, the 5-point case to |.-15.795| < 1;
tests/test_wind_oracle.py).  ``scale_wind_vert::point_scale`` and ``fetchr::run`` are pinned on the reference's own
translation units compiled unmodified against stand-in headers (oracle/refbuild -> oracle/_ref/libchmref.so; golden
vectors tests/golden/golden_wind.npz, generator tests/golden/make_golden_wind.py).  Restated from published
definitions because the library is absent from this image: GSL ``gsl_sf_expint_E1`` (scipy.special.exp1 here), Eigen
``FullPivLU`` (numpy's LU here), CGAL's kd-tree nearest-neighbour query (scipy cKDTree / brute force here).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
from scipy.special import exp1

Z_U_R = 50.0  # Atmosphere.h:31
Z0_SNOW = 0.01  # Snow.h:31


def is_nan(x):
    """module_base::is_nan (module_base.hpp:471-479): -9999 sentinel or NaN."""
    x = np.asarray(x, dtype=np.float64)
    return (np.abs(x - -9999.0) < 1e-5) | np.isnan(x)


def log_scale_wind(u, z_in, z_out, snowdepth, z0=Z0_SNOW):
    """Atmosphere::log_scale_wind (Atmosphere.cpp:32-38)."""
    return u * np.log((z_out - (snowdepth + z0)) / z0) / np.log((z_in - (snowdepth + z0)) / z0)


def exp_scale_wind(u, z_in, z_out, alpha):
    """Atmosphere::exp_scale_wind (Atmosphere.cpp:41-46)."""
    return u * np.exp(alpha * (z_out / z_in - 1.0))


def point_scale(U_R, snowdepthavg=None, canopy_height=None, lai=None, ignore_canopy=False):
    """scale_wind_vert::point_scale (scale_wind_vert.cpp:48-136), vectorised over faces.

    canopy_height None = the mesh has no vegetation parameters (face->has_vegetation() false everywhere)."""
    U_R = np.asarray(U_R, dtype=np.float64)
    n = U_R.shape[0]
    sd = np.zeros(n) if snowdepthavg is None else np.where(is_nan(snowdepthavg), 0.0, np.asarray(snowdepthavg, dtype=np.float64))
    ztop = np.zeros(n)
    if not ignore_canopy and canopy_height is not None:
        ztop = np.asarray(canopy_height, dtype=np.float64)
    zbot = ztop / 2.0
    z2 = sd + 2.0
    out = np.empty(n)
    with np.errstate(all="ignore"):
        plain = log_scale_wind(U_R, Z_U_R, z2, sd)
        in_canopy = (not ignore_canopy) & (ztop > 0.0) & (z2 < ztop)
        u = plain.copy()
        if in_canopy.any():
            if lai is None:
                raise ValueError("Parameter LAI does not exist.")  # veg_attribute throws (triangulation.hpp:1685-1688)
            alpha = np.asarray(lai, dtype=np.float64)
            safe_top = np.where(ztop > 0, ztop, 1.0)
            u_top = log_scale_wind(U_R, Z_U_R, safe_top, sd)
            u_bot = exp_scale_wind(u_top, safe_top, zbot, alpha)
            below = log_scale_wind(u_bot, np.where(zbot > 0, zbot, 1.0), z2, sd)
            within = exp_scale_wind(u_top, safe_top, z2, alpha)
            canopy_u = np.where(z2 < zbot, below, within)
            canopy_u = np.where(sd < ztop, canopy_u, plain)
            u = np.where(in_canopy, canopy_u, plain)
        u = np.maximum(0.1, u)  # std::max(0.1, NaN) keeps 0.1
        u = np.where(np.isnan(u), 0.1, u)
        out = np.where(z2 >= Z_U_R, U_R, u)
    return out


# ------------------------------------------------------------------------------------------ thin plate spline
TPS_C = 0.577215  # "euler constant" as the reference truncates it (TPSpline.cpp:197)
TPS_WEIGHT = 0.01  # TPSpline.cpp:198


def tps_basis(d):
    """Rd = -(log(x) + c + E1(x)),  x = (d*weight/2)^2   (TPSpline.cpp:83-94, TPSBasis.hpp)."""
    x = (d * TPS_WEIGHT / 2.0) * (d * TPS_WEIGHT / 2.0)
    return -(np.log(x) + TPS_C + exp1(x))


def thin_plate_spline(sample_xyz, query_xy) -> float:
    """thin_plate_spline::operator() (TPSpline.cpp:40-173) for one query.  sample_xyz: (n,3) rows (x, y, value)."""
    s = np.asarray(sample_xyz, dtype=np.float64)
    n = s.shape[0]
    size = n + 1
    A = np.zeros((size, size))
    for i in range(n):
        for j in range(i, n):
            xd, yd = s[i, 0] - s[j, 0], s[i, 1] - s[j, 1]
            if xd == 0.0 and yd == 0.0:
                continue
            Rd = tps_basis(np.sqrt(xd * xd + yd * yd))
            A[i, j + 1] = Rd
            A[j, i + 1] = Rd
    A[:, 0] = 1.0
    A[size - 1, :] = 1.0
    A[size - 1, 0] = 0.0
    b = np.zeros(size)
    b[:n] = s[:, 2]
    x = np.linalg.solve(A, b)
    z0 = x[0]
    for i in range(1, size):
        xd, yd = s[i - 1, 0] - query_xy[0], s[i - 1, 1] - query_xy[1]
        z0 = z0 + x[i] * tps_basis(np.sqrt(xd * xd + yd * yd))
    return float(z0)


def scale_wind_vert(U_R, neigh, cx, cy, snowdepthavg=None, canopy_height=None, lai=None, ignore_canopy=False,
                    ghost_u2: Optional[np.ndarray] = None):
    """scale_wind_vert::run(mesh&) (scale_wind_vert.cpp:167-229): point_scale, halo, then every face takes the thin
    plate spline of its (<= 3) neighbours' values at its own centre, floored at 0.1.

    neigh [T,3] with -1 none and >= T ghost (T+g); cx, cy [T+nG]; ghost_u2 [nG] = the partners' point_scale values."""
    T = neigh.shape[0]
    u = point_scale(U_R, snowdepthavg, canopy_height, lai, ignore_canopy)
    uall = u if ghost_u2 is None else np.concatenate([u, ghost_u2])
    out = np.empty(T)
    for i in range(T):
        pts = [(cx[n], cy[n], uall[n]) for n in neigh[i] if n >= 0]
        if pts:
            out[i] = max(0.1, thin_plate_spline(np.array(pts), (cx[i], cy[i])))
        else:
            out[i] = max(0.1, u[i])
    return out, u


# ------------------------------------------------------------------------------------------ fetchr
def fetchr(vw_dir, cx, cy, cz, canopy_height=None, steps=10, max_distance=1000.0, I=0.06, incl_veg=True,
           search_cx=None, search_cy=None, search_cz=None, search_canopy=None):
    """fetchr::run (fetchr.cpp:54-119) for every face.  The nearest-centroid search (face::find_closest_face,
    triangulation.hpp:1543-1546 -> triangulation.cpp:170-186, UTM point_from_bearing coordinates.cpp:60-71) runs over
    the `search_*` faces (default: the faces themselves).  canopy_height None = no vegetation parameters."""
    from scipy.spatial import cKDTree

    vw_dir = np.asarray(vw_dir, dtype=np.float64)
    T = vw_dir.shape[0]
    scx = cx if search_cx is None else search_cx
    scy = cy if search_cy is None else search_cy
    scz = cz if search_cz is None else search_cz
    scan = canopy_height if search_canopy is None else search_canopy
    tree = cKDTree(np.stack([scx, scy], axis=1))
    size_of_step = max_distance / steps
    fetch = np.full(T, max_distance)
    done = np.zeros(T, dtype=bool)
    has_veg = canopy_height is not None
    if incl_veg and has_veg:
        tall = np.asarray(canopy_height) > 1
        fetch[tall] = 0.0
        done |= tall
    b = vw_dir * (np.pi / 180.0)
    z0_2, nn, h = 0.001, 1.0 / 0.8, 5.0
    for j in range(1, steps + 1):
        distance = j * size_of_step
        qx = cx[:T] + distance * np.sin(b)
        qy = cy[:T] + distance * np.cos(b)
        _, f = tree.query(np.stack([qx, qy], axis=1))
        ztop = np.zeros(T)
        if incl_veg and scan is not None:
            ztop = np.asarray(scan, dtype=np.float64)[f]
        z_test = scz[f] + ztop
        z_core = cz[:T] + distance * I
        z0_1 = 0.12 * ztop
        with np.errstate(all="ignore"):
            x_sss = np.power((33.33333333 * h - 25.0 * z0_2) / (np.log(z0_1 / z0_2) * z0_2), nn) * z0_2
        hit = (z_test >= z_core) | (incl_veg & (distance < x_sss))
        hit &= ~done
        fetch[hit] = distance
        done |= hit
    return fetch

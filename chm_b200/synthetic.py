"""Synthetic meshes and forcing of the BASELINE.json shapes.

* ``uniform_mesh``   – nx×ny squares of side h, split on alternating diagonals, Morton-ordered
                       (708×708×30 m → 1 002 528 triangles = config c2).
* ``variable_mesh``  – Delaunay triangulation of points whose density follows a smooth 10:1 area
                       field (configs c3–c5), Morton-ordered.
* ``forcing``        – spatially smooth per-face forcing (U_R, U_2m_above_srf, snowdepthavg, swe, t, rh,
                       vw_dir, fetch) from a seeded generator.

Morton order of face centroids stands in for the METIS/RCM permutation CHM meshes are shipped with
(docs/meshgen.rst:18-31; METIS itself lives in the external mesher tool).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .mesh import TriMesh, reorder_faces


def build_neighbours(elem: np.ndarray) -> np.ndarray:
    """neigh[i,j] = face sharing the edge opposite vertex j of face i, -1 on the hull (CHM convention)."""
    T = elem.shape[0]
    e = elem.astype(np.int64)
    a = np.stack([e[:, 1], e[:, 2], e[:, 0]], axis=1).reshape(-1)  # edge j: vertices (j+1)%3,(j+2)%3
    b = np.stack([e[:, 2], e[:, 0], e[:, 1]], axis=1).reshape(-1)
    nv = int(e.max()) + 1
    key = np.minimum(a, b) * nv + np.maximum(a, b)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    same = ks[1:] == ks[:-1]
    neigh = np.full(3 * T, -1, dtype=np.int64)
    i0, i1 = order[:-1][same], order[1:][same]
    neigh[i0] = i1 // 3
    neigh[i1] = i0 // 3
    return neigh.reshape(T, 3).astype(np.int32)


def _part1by1(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64) & np.uint64(0xFFFFFFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
    v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
    v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
    v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
    return v


def morton_permutation(cx: np.ndarray, cy: np.ndarray, bits: int = 20) -> np.ndarray:
    """perm[k] = old face index that becomes global id k (the ``mesh.cell_global_id`` array)."""
    sx = (cx - cx.min()) / max(cx.max() - cx.min(), 1e-300)
    sy = (cy - cy.min()) / max(cy.max() - cy.min(), 1e-300)
    q = (1 << bits) - 1
    code = _part1by1((sx * q).astype(np.uint64)) | (_part1by1((sy * q).astype(np.uint64)) << np.uint64(1))
    return np.argsort(code, kind="stable").astype(np.int64)


def _terrain(x, y):
    return 1500.0 + 200.0 * np.sin(x / 3000.0) * np.cos(y / 2000.0)


def _finish(vertex, elem, order: str) -> TriMesh:
    # make every face CCW
    p = vertex[elem]
    cross = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    cw = cross < 0
    elem = elem.copy()
    elem[cw, 1], elem[cw, 2] = elem[cw, 2].copy(), elem[cw, 1].copy()
    neigh = build_neighbours(elem)
    mesh = TriMesh(vertex, elem, neigh)
    if order == "morton":
        c = vertex[elem].mean(axis=1)
        mesh = reorder_faces(mesh, morton_permutation(c[:, 0], c[:, 1]))
    return mesh


def uniform_mesh(nx: int = 708, ny: int = 708, h: float = 30.0, order: str = "morton",
                 x0: float = 488000.0, y0: float = 6710000.0) -> TriMesh:
    """nx×ny squares → 2·nx·ny triangles, diagonals alternating with (i+j) parity, UTM-like coords."""
    xs = x0 + h * np.arange(nx + 1)
    ys = y0 + h * np.arange(ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    vertex = np.stack([X.ravel(), Y.ravel(), _terrain(X.ravel() - x0, Y.ravel() - y0)], axis=1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    i, j = i.ravel(), j.ravel()
    v00 = j * (nx + 1) + i
    v10, v01, v11 = v00 + 1, v00 + nx + 1, v00 + nx + 2
    even = ((i + j) % 2) == 0
    t0 = np.where(even[:, None], np.stack([v00, v10, v11], 1), np.stack([v00, v10, v01], 1))
    t1 = np.where(even[:, None], np.stack([v00, v11, v01], 1), np.stack([v10, v11, v01], 1))
    elem = np.empty((2 * nx * ny, 3), dtype=np.int32)
    elem[0::2], elem[1::2] = t0, t1
    return _finish(vertex, elem, order)


def alpine_terrain(x, y):
    """Steep synthetic relief for snow_slide: slopes up to ~55 degrees, ridges ~3-6 km apart plus 1 km-scale gullies."""
    return (2000.0 + 900.0 * np.sin(x / 900.0) * np.cos(y / 700.0) + 150.0 * np.sin(x / 170.0 + 1.0) * np.sin(y / 210.0)
            + 40.0 * np.sin((x + 0.7 * y) / 61.0))


def with_elevation(mesh: TriMesh, fn=alpine_terrain, x0: float = 488000.0, y0: float = 6710000.0) -> TriMesh:
    """The same triangulation with vertex elevations fn(x - x0, y - y0) (global, unpartitioned meshes)."""
    v = mesh.vertex.copy()
    v[:, 2] = fn(v[:, 0] - x0, v[:, 1] - y0)
    return TriMesh(v, mesh.elem, mesh.neigh, dict(mesh.params), local_sizes=mesh.local_sizes)


def variable_mesh(n_tri_target: int, seed: int = 20250101, area_ratio: float = 10.0,
                  order: str = "morton", x0: float = 488000.0, y0: float = 6710000.0) -> TriMesh:
    """Variable-resolution Delaunay mesh, ≈n_tri_target triangles, areas spanning ≈area_ratio:1.

    Points are a jittered lattice in a warped coordinate whose Jacobian follows a smooth log-uniform
    area field (rejection-free, so the count is predictable); ``scipy.spatial.Delaunay`` then gives the
    irregular adjacency (its ``simplices`` are re-paired here so the neighbour convention is CHM's).
    """
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    n_pts = max(16, n_tri_target // 2)
    n = int(np.ceil(np.sqrt(n_pts)))
    u, v = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(n) + 0.5) / n, indexing="xy")
    u = (u + rng.uniform(-0.35, 0.35, u.shape) / n).ravel()
    v = (v + rng.uniform(-0.35, 0.35, v.shape) / n).ravel()
    # smooth 1-D stretchings; linear density varies by sqrt(area_ratio) along each axis ⇒ area by ratio
    a = np.sqrt(area_ratio)
    k = (a - 1.0) / (a + 1.0)

    def warp(s, cycles):
        return s + k * np.sin(2 * np.pi * cycles * s) / (2 * np.pi * cycles)

    mean_area = 450.0  # m², same scale as the uniform 30 m mesh
    side = np.sqrt(mean_area * n_tri_target)
    x = x0 + side * warp(u, 2.0)
    y = y0 + side * warp(v, 3.0)
    tri = Delaunay(np.stack([x, y], axis=1))
    elem = tri.simplices.astype(np.int32)
    # drop hull slivers (min angle < 10°), which only occur on the convex hull of a jittered lattice
    p = np.stack([x, y], axis=1)[elem]
    e0 = np.linalg.norm(p[:, 1] - p[:, 0], axis=1)
    e1 = np.linalg.norm(p[:, 2] - p[:, 1], axis=1)
    e2 = np.linalg.norm(p[:, 0] - p[:, 2], axis=1)
    area2 = np.abs((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0]))
    # sin(min angle) = 2A / (product of the two longest edges)
    es = np.sort(np.stack([e0, e1, e2], 1), axis=1)
    sin_min = area2 / (es[:, 1] * es[:, 2])
    elem = elem[sin_min > np.sin(np.deg2rad(10.0))]
    vertex = np.stack([x, y, _terrain(x - x0, y - y0)], axis=1)
    return _finish(vertex, elem, order)


def _smooth_field(cx, cy, rng, n_modes=6, scale=4000.0):
    """Zero-mean, roughly unit-amplitude smooth field from a few random plane waves."""
    f = np.zeros_like(cx)
    for _ in range(n_modes):
        kx, ky = rng.normal(0, 1.0 / scale, 2) * 2 * np.pi
        f += np.sin(kx * cx + ky * cy + rng.uniform(0, 2 * np.pi))
    return f / np.sqrt(n_modes / 2.0)


def forcing(cx: np.ndarray, cy: np.ndarray, seed: int = 7, step: int = 0, calm: bool = False,
            fetch_const: Optional[float] = 1000.0) -> Dict[str, np.ndarray]:
    """Per-face forcing for one timestep at centroids (cx, cy) (SURVEY.md §8d 'forcing per face').

    Fields are functions of position only (plus ``step``), so a partitioned run sees exactly the
    values the global run sees.
    """
    rng = np.random.default_rng(seed + 1000 * step)
    x = cx - 488000.0
    y = cy - 6710000.0
    g = lambda: _smooth_field(x, y, rng)
    U_R = np.clip(11.0 * np.exp(0.35 * g()), 4.0, 22.0)
    if calm:
        U_R = np.full_like(U_R, 1.5)
    sd = np.clip(0.85 + 0.45 * g(), 0.2, 1.5)
    swe = np.clip(265.0 + 130.0 * g(), 80.0, 450.0)
    t = np.clip(-13.5 + 8.0 * g(), -25.0, -2.0)
    rh = np.clip(75.0 + 14.0 * g(), 55.0, 95.0)
    vw_dir = 270.0 + np.clip(12.0 * g(), -25.0, 25.0)
    # scale_wind_vert: U_2m_above_srf = log_scale_wind(U_R, Z_U_R=50, 2+sd, sd)  (Atmosphere.cpp:32-38)
    z0 = 0.01
    u2 = U_R * np.log((2.0 + sd - (sd + z0)) / z0) / np.log((50.0 - (sd + z0)) / z0)
    u2 = np.maximum(0.1, u2)
    if fetch_const is None:
        fetch = np.clip(500.0 + 500.0 * g(), 0.0, 1000.0)
    else:
        fetch = np.full_like(U_R, fetch_const)
    return {"U_R": U_R, "U_2m_above_srf": u2, "snowdepthavg": sd, "swe": swe, "t": t, "rh": rh,
            "vw_dir": vw_dir, "fetch": fetch}


def patchy_forcing(cx: np.ndarray, cy: np.ndarray, seed: int = 7, step: int = 0, scale: float = 2500.0,
                   threshold: float = 0.5) -> Dict[str, np.ndarray]:
    """`forcing` with saltation confined to wind-exposed patches, as on most hours of a real winter: where a smooth field
    of length scale `scale` is below `threshold` the reference-height wind drops to 3 m/s (no saltation, zero right-hand
    side of the suspension system); about a fifth to a third of the faces keep the stormy wind."""
    f = forcing(cx, cy, seed=seed, step=step)
    s = _smooth_field(cx - 488000.0, cy - 6710000.0, np.random.default_rng(5 + seed + 1000 * step), scale=scale)
    f["U_R"] = np.where(s > threshold, f["U_R"], 3.0)
    sd, z0 = f["snowdepthavg"], 0.01
    f["U_2m_above_srf"] = np.maximum(0.1, f["U_R"] * np.log((2.0 + sd - (sd + z0)) / z0) / np.log((50.0 - (sd + z0)) / z0))
    return f


def shrub_params(n: int, frac: float = 0.2, seed: int = 11, canopy: float = 1.0) -> Dict[str, np.ndarray]:
    """Variant with `frac` of faces carrying shrubs (CanopyHeight `canopy`, N 1, dv 0.8), others bare.
    With snow depths of 0.2–1.5 m a 1 m canopy gives all three regimes: buried, partly exposed (lambda > 0)
    and exposed by more than `cutoff` (saltation inhibited)."""
    rng = np.random.default_rng(seed)
    shrub = rng.random(n) < frac
    return {"CanopyHeight": np.where(shrub, canopy, 0.0), "stalk_number": np.ones(n), "stalk_diameter": np.full(n, 0.8),
            "LAI": np.where(shrub, 1.0, 0.0)}

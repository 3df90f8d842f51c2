"""Triangle-mesh substrate for the PBSM3D hot path: CHM's face order, flattened to SoA.

What CHM keeps as a CGAL ``Triangulation_data_structure_2`` of heap ``face`` objects
(reference: src/mesh/triangulation.hpp:1173, triangulation.cpp:206-511) is held here as
plain arrays in CHM's own face order (ascending ``cell_global_id``):

* ``vertex[nv,3]``  – x, y, z
* ``elem[T,3]``     – vertex ids of each face (CCW)
* ``neigh[T,3]``    – ``neigh[i,j]`` is the face sharing the edge opposite vertex ``j`` of face ``i``
                      (-1 = domain boundary), exactly ``face->neighbor(j)``
* ``params``        – per-face parameters (``area``, ``CanopyHeight``, ``LAI`` ...)

Only host-side bookkeeping lives here (file reading, the ``cell_global_id`` permutation, CHM's
contiguous-range partition rule and ghost lists).  Edge normals / lengths / centroids used by the
product are computed on the device by ``pbsm3d_create`` (csrc/pbsm3d_setup.cu); the numpy versions in
this file exist so that tests can pin that kernel bit-for-bit (they are the geometry the CPU checker uses).
"""
from __future__ import annotations

import json
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

GHOST_NONE, GHOST_NEIGH, GHOST_DIST = 0, 1, 2  # triangulation.hpp GHOST_TYPE


def _strip_json_comments(text: str) -> str:
    """CHM configs/meshes are JSON with // and /* */ comments (src/utility/jsonstrip.cpp)."""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"(^|[^:\"])//[^\n]*", r"\1", text)


@dataclass
class TriMesh:
    """A (possibly rank-local) triangle mesh in CHM face order.

    For a rank-local mesh faces ``[0, n_local)`` are owned (ascending global id) and faces
    ``[n_local, n_local+n_ghost)`` are NEIGH ghosts sorted by global id, which makes each owner's
    ghosts contiguous (triangulation.cpp:1755-1762).  ``neigh`` holds local indices into that
    combined list; only owned faces have neighbour rows.
    """

    vertex: np.ndarray  # [nv,3] f64
    elem: np.ndarray  # [T(+ghost),3] i32
    neigh: np.ndarray  # [T,3] i32, local ids, -1 none
    params: Dict[str, np.ndarray] = field(default_factory=dict)  # each [T]
    n_global: int = 0
    global_id: Optional[np.ndarray] = None  # [T+ghost] i64
    n_ghost: int = 0
    ghost_owner: Optional[np.ndarray] = None  # [ghost] i32
    rank: int = 0
    n_ranks: int = 1
    local_sizes: Optional[np.ndarray] = None  # [P] faces owned by each rank (CHM mesh.local_size)
    is_geographic: bool = False

    def __post_init__(self):
        self.vertex = np.ascontiguousarray(self.vertex, dtype=np.float64)
        self.elem = np.ascontiguousarray(self.elem, dtype=np.int32)
        self.neigh = np.ascontiguousarray(self.neigh, dtype=np.int32)
        if self.n_global == 0:
            self.n_global = self.n_local
        if self.global_id is None:
            self.global_id = np.arange(self.n_local + self.n_ghost, dtype=np.int64)
        if self.ghost_owner is None:
            self.ghost_owner = np.zeros(self.n_ghost, dtype=np.int32)

    # ------------------------------------------------------------------ sizes
    @property
    def n_local(self) -> int:
        return int(self.neigh.shape[0])

    @property
    def n_faces_total(self) -> int:
        return int(self.elem.shape[0])

    # --------------------------------------------------------------- geometry
    def face_vertices(self) -> np.ndarray:
        """[T+ghost, 3 vertices, 3 coords] – what the C-ABI takes (no vertex indexing on device)."""
        return np.ascontiguousarray(self.vertex[self.elem])

    def geometry(self) -> "FaceGeometry":
        return face_geometry(self.face_vertices(), self.neigh, self.params.get("area"))


@dataclass
class FaceGeometry:
    nx: np.ndarray  # [3,T] outward unit edge normal x (edge j is shared with neigh[:,j])
    ny: np.ndarray  # [3,T]
    elen: np.ndarray  # [3,T] edge length
    area: np.ndarray  # [T]
    cx: np.ndarray  # [T+ghost] centroid
    cy: np.ndarray
    cz: np.ndarray
    dx: np.ndarray  # [3,T] centroid distance to neighbour j (2.0 where none, PBSM3D.cpp:1534)


def face_geometry(fv: np.ndarray, neigh: np.ndarray, area_param: Optional[np.ndarray] = None) -> FaceGeometry:
    """Plain-fp64 restatement of the face geometry CHM evaluates per call.

    reference: triangulation.hpp:1443-1475 (edge_unit_normal, edge), :1491-1498 (edge_length),
    :1577-1589 (center = CGAL::centroid), :1830-1856 (get_area: param "area" else signed area),
    math/coordinates.cpp:100-106 (distance_UTM).  Operation order is kept as written there so the
    device kernel can be compared bit-for-bit.
    """
    T = neigh.shape[0]
    px, py, pz = fv[:, :, 0], fv[:, :, 1], fv[:, :, 2]
    nx = np.empty((3, T))
    ny = np.empty((3, T))
    elen = np.empty((3, T))

    def edge(i):
        # edge(i) = v[cw(i)] - v[ccw(i)], ccw(i)=(i+1)%3, cw(i)=(i+2)%3
        a, b = (i + 1) % 3, (i + 2) % 3
        return px[:T, b] - px[:T, a], py[:T, b] - py[:T, a]

    for i in range(3):
        ex, ey = edge(i)
        e1x, e1y = edge((i + 1) % 3)
        n_x, n_y = ey, -ex
        D = e1x * n_x + e1y * n_y
        flip = D > 0
        n_x = np.where(flip, -n_x, n_x)
        n_y = np.where(flip, -n_y, n_y)
        nrm = np.sqrt(n_x * n_x + n_y * n_y)
        nx[i], ny[i] = n_x / nrm, n_y / nrm
        elen[i] = np.sqrt(ex * ex + ey * ey)

    if area_param is not None:
        area = np.ascontiguousarray(area_param, dtype=np.float64).copy()
    else:
        v1x, v1y = px[:T, 1] - px[:T, 0], py[:T, 1] - py[:T, 0]
        v2x, v2y = px[:T, 2] - px[:T, 0], py[:T, 2] - py[:T, 0]
        area = (v1x * v2y - v1y * v2x) / 2.0
    cx = (px[:, 0] + px[:, 1] + px[:, 2]) / 3.0
    cy = (py[:, 0] + py[:, 1] + py[:, 2]) / 3.0
    cz = (pz[:, 0] + pz[:, 1] + pz[:, 2]) / 3.0
    dx = np.full((3, T), 2.0)
    for j in range(3):
        n = neigh[:, j]
        has = n >= 0
        ddx = cx[:T][has] - cx[n[has]]
        ddy = cy[:T][has] - cy[n[has]]
        dx[j, has] = np.sqrt(ddx * ddx + ddy * ddy)
    return FaceGeometry(nx, ny, elen, area, cx, cy, cz, dx)


# ----------------------------------------------------------------------------- reading
def read_chm_mesh(mesh_path: str, param_paths: Sequence[str] = ()) -> TriMesh:
    """Read a CHM JSON ``.mesh`` (+ ``.param`` files) the way ``triangulation::from_json`` does.

    reference: triangulation.cpp:206-511.  If ``mesh.cell_global_id`` is present the faces are
    permuted by ``reorder_faces`` (:1384-1413): the face at file position ``perm[k]`` becomes
    global id ``k``; parameters were attached before the permutation and travel with the face.
    """
    with open(mesh_path) as f:
        doc = json.loads(_strip_json_comments(f.read()))
    m = doc["mesh"]
    vertex = np.asarray(m["vertex"], dtype=np.float64)
    elem = np.asarray(m["elem"], dtype=np.int32)
    neigh = np.asarray(m["neigh"], dtype=np.int32)
    nelem = int(m["nelem"])
    if vertex.shape[0] != int(m["nvertex"]):
        raise ValueError(f"Expected: {m['nvertex']} vertex, got: {vertex.shape[0]}")
    if elem.shape[0] != nelem:
        raise ValueError(f"Expected: {nelem} elems, got: {elem.shape[0]}")
    if (neigh > nelem).any():
        raise ValueError("Face has out of bound neighbors.")
    params: Dict[str, np.ndarray] = {}
    pdocs = [doc.get("parameters", {})]
    for p in param_paths:
        with open(p) as f:
            pdocs.append(json.loads(_strip_json_comments(f.read())))
    for pd in pdocs:
        for name, vals in pd.items():
            if len(vals) == 0:  # "area": [] is ignored (triangulation.cpp:375-389)
                continue
            if len(vals) > nelem:
                raise ValueError("There are more parameter elements than triangulation elements")
            params[name] = np.asarray(vals, dtype=np.float64)
    local_sizes = np.asarray(m["local_size"], dtype=np.int64) if "local_size" in m else None
    mesh = TriMesh(vertex, elem, neigh, params, n_global=nelem, local_sizes=local_sizes,
                   is_geographic=bool(int(m.get("is_geographic", 0))))
    if "cell_global_id" in m:
        mesh = reorder_faces(mesh, np.asarray(m["cell_global_id"], dtype=np.int64))
    return mesh


def reorder_faces(mesh: TriMesh, permutation: np.ndarray) -> TriMesh:
    """``triangulation::reorder_faces`` (triangulation.cpp:1384-1413): new face k = old face perm[k]."""
    perm = np.asarray(permutation, dtype=np.int64)
    T = mesh.n_local
    if perm.shape[0] != T or mesh.n_ghost:
        raise ValueError("permutation size must equal the number of faces of a global mesh")
    inv = np.empty(T, dtype=np.int64)
    inv[perm] = np.arange(T)
    old_neigh = mesh.neigh[perm]
    new_neigh = np.where(old_neigh >= 0, inv[np.maximum(old_neigh, 0)], -1).astype(np.int32)
    params = {k: v[perm] for k, v in mesh.params.items()}
    return TriMesh(mesh.vertex, mesh.elem[perm], new_neigh, params, n_global=T,
                   local_sizes=mesh.local_sizes, is_geographic=mesh.is_geographic)


# ----------------------------------------------------------------------------- partition
def partition_sizes(n_global: int, n_ranks: int, local_sizes: Optional[np.ndarray] = None) -> np.ndarray:
    """Faces owned by each rank.

    reference: mesh-file ``local_size`` when it matches the rank count
    (preprocessing/partition/main.cpp:204-286, triangulation.cpp:1482-1493), else the balanced
    fallback G/P with the first G%P ranks taking one more (triangulation.cpp:1583-1596).
    """
    if local_sizes is not None and len(local_sizes) == n_ranks:
        ls = np.asarray(local_sizes, dtype=np.int64)
        if ls.sum() != n_global:
            raise ValueError("local_size and partition size mismatch")
        return ls
    sizes = np.full(n_ranks, n_global // n_ranks, dtype=np.int64)
    sizes[: n_global % n_ranks] += 1
    return sizes


def partition_mesh(mesh: TriMesh, rank: int, n_ranks: int,
                   local_sizes: Optional[np.ndarray] = None) -> TriMesh:
    """Rank-local view of a global mesh under CHM's rule.

    Rank r owns the contiguous global-id range ``[start_r, start_r+size_r)``; its NEIGH ghosts are
    the non-owned edge neighbours of owned faces, de-duplicated and sorted by global id
    (triangulation.cpp:1721-1770), so ghosts of one owner are contiguous (:1784-1829).  DIST ghosts
    (max_ghost_distance) are not used by PBSM3D and are not built.
    """
    if mesh.n_ghost or mesh.n_ranks != 1:
        raise ValueError("partition_mesh wants the global mesh")
    G = mesh.n_local
    sizes = partition_sizes(G, n_ranks, local_sizes if local_sizes is not None else mesh.local_sizes)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    s, e = int(starts[rank]), int(starts[rank + 1])
    T = e - s
    nb = mesh.neigh[s:e].astype(np.int64)
    off = (nb >= 0) & ((nb < s) | (nb >= e))
    ghosts = np.unique(nb[off])  # sorted by global id
    owner = (np.searchsorted(starts, ghosts, side="right") - 1).astype(np.int32)
    local = np.full(nb.shape, -1, dtype=np.int64)
    own = (nb >= s) & (nb < e)
    local[own] = nb[own] - s
    local[off] = T + np.searchsorted(ghosts, nb[off])
    gid = np.concatenate([np.arange(s, e, dtype=np.int64), ghosts])
    # keep only the vertices this rank touches (owned + ghost faces)
    elem_g = mesh.elem[gid]
    used, inv = np.unique(elem_g.reshape(-1), return_inverse=True)
    params = {k: v[s:e] for k, v in mesh.params.items()}
    return TriMesh(mesh.vertex[used], inv.reshape(-1, 3).astype(np.int32), local.astype(np.int32), params,
                   n_global=G, global_id=gid, n_ghost=len(ghosts), ghost_owner=owner, rank=rank,
                   n_ranks=n_ranks, local_sizes=sizes, is_geographic=mesh.is_geographic)


def halo_plan(parts: List[TriMesh]):
    """For tests: per rank, per partner, the local indices to send (what
    ``setup_nearest_neighbor_communication`` negotiates, triangulation.cpp:1845-1945)."""
    P = len(parts)
    starts = np.concatenate([[0], np.cumsum(parts[0].local_sizes)])
    plan = [dict() for _ in range(P)]
    for r, p in enumerate(parts):
        gg = p.global_id[p.n_local:]
        for q in np.unique(p.ghost_owner):
            ids = gg[p.ghost_owner == q]
            plan[int(q)][r] = (ids - starts[q]).astype(np.int32)
    return plan


def check_neighbour_symmetry(mesh: TriMesh) -> bool:
    """Shared-edge convention check: neigh[i,j] shares the edge opposite vertex j."""
    T = mesh.n_local
    for i in range(T):
        for j in range(3):
            n = mesh.neigh[i, j]
            if n < 0 or n >= T:
                continue
            if i not in mesh.neigh[n]:
                return False
            a, b = mesh.elem[i, (j + 1) % 3], mesh.elem[i, (j + 2) % 3]
            if a not in mesh.elem[n] or b not in mesh.elem[n]:
                return False
    return True

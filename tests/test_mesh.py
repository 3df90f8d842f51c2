"""Host-side mesh substrate: CHM face order, neighbour convention, geometry, partition rule (CPU)."""
import numpy as np
import pytest

from chm_b200 import synthetic
from chm_b200.mesh import (TriMesh, check_neighbour_symmetry, halo_plan, partition_mesh, partition_sizes, reorder_faces)


def test_granger_fixture_shape(granger):
    # test_data/meshes/granger1m.mesh: 985 triangles, 533 vertices, 75 boundary triangles
    assert granger.n_local == 985 and granger.vertex.shape == (533, 3)
    assert int((granger.neigh < 0).any(axis=1).sum()) == 75
    assert set(granger.params) == {"area", "MS0"}


def test_neighbour_convention(granger, slope):
    # neigh[i][j] is the face across the edge opposite vertex j, adjacency symmetric
    assert check_neighbour_symmetry(granger)
    assert check_neighbour_symmetry(slope)
    assert slope.n_local == 2618


def test_geometry_closed_and_outward(granger):
    g = granger.geometry()
    T = granger.n_local
    assert np.allclose(g.nx ** 2 + g.ny ** 2, 1.0, atol=1e-14)
    # sum_j E_j n_j = 0 for a closed triangle
    assert np.abs((g.elen * g.nx).sum(0)).max() < 1e-8 and np.abs((g.elen * g.ny).sum(0)).max() < 1e-8
    # outward: normal of edge j points away from vertex j
    fv = granger.face_vertices()
    for j in range(3):
        mid = 0.5 * (fv[:T, (j + 1) % 3, :2] + fv[:T, (j + 2) % 3, :2])
        d = mid - fv[:T, j, :2]
        assert ((d[:, 0] * g.nx[j] + d[:, 1] * g.ny[j]) > 0).all()
    # area comes from the "area" parameter when present (triangulation.hpp:1836-1839)
    assert np.array_equal(g.area, granger.params["area"])
    # dx defaults to 2.0 where there is no neighbour (PBSM3D.cpp:1534)
    assert (g.dx[(granger.neigh < 0).T] == 2.0).all()


def test_signed_area_without_param():
    m = synthetic.uniform_mesh(4, 3, 30.0, order="none")
    g = m.geometry()
    assert np.allclose(g.area, 450.0)


def test_reorder_faces_is_chm_permutation(slope, slope_metis):
    # slope.metis.mesh = slope.mesh + cell_global_id permutation: new face k = old face perm[k]
    assert slope_metis.n_local == slope.n_local
    assert check_neighbour_symmetry(slope_metis)
    assert slope_metis.local_sizes is not None and len(slope_metis.local_sizes) == 31
    assert int(slope_metis.local_sizes.sum()) == 2618
    rng = np.random.default_rng(0)
    perm = rng.permutation(slope.n_local)
    r = reorder_faces(slope, perm)
    assert np.array_equal(r.elem, slope.elem[perm])
    assert np.array_equal(r.params["area"], slope.params["area"][perm])
    # adjacency is carried by handle: neighbours of new face k are the renumbered neighbours of old face perm[k]
    inv = np.argsort(perm)
    k = 17
    old = slope.neigh[perm[k]]
    assert np.array_equal(r.neigh[k], np.where(old >= 0, inv[np.maximum(old, 0)], -1))
    assert check_neighbour_symmetry(r)


def test_partition_sizes_rule():
    # balanced fallback: G/P, first G%P ranks one more (triangulation.cpp:1583-1596)
    assert partition_sizes(10, 4).tolist() == [3, 3, 2, 2]
    assert partition_sizes(10, 3, np.array([5, 4, 1])).tolist() == [5, 4, 1]
    with pytest.raises(ValueError):
        partition_sizes(10, 2, np.array([5, 4]))


@pytest.mark.parametrize("P", [2, 3, 8])
def test_partition_ghosts(slope_metis, P):
    parts = [partition_mesh(slope_metis, r, P) for r in range(P)]
    assert sum(p.n_local for p in parts) == slope_metis.n_local
    start = 0
    for r, p in enumerate(parts):
        gid = p.global_id
        assert np.array_equal(gid[:p.n_local], np.arange(start, start + p.n_local))  # contiguous owned range
        gg = gid[p.n_local:]
        assert (np.diff(gg) > 0).all()  # ghosts sorted by global id ...
        assert (np.diff(p.ghost_owner) >= 0).all()  # ... hence contiguous per owner
        assert (p.ghost_owner != r).all()
        # local adjacency reproduces the global one
        loc = p.neigh
        glob = slope_metis.neigh[start:start + p.n_local]
        assert np.array_equal(np.where(loc >= 0, gid[np.maximum(loc, 0)], -1), glob)
        # geometry of owned faces is unchanged by partitioning (bit-exact)
        g0, g1 = slope_metis.geometry(), p.geometry()
        assert np.array_equal(g1.nx, g0.nx[:, start:start + p.n_local])
        assert np.array_equal(g1.dx, g0.dx[:, start:start + p.n_local])
        start += p.n_local
    plan = halo_plan(parts)
    for q in range(P):
        for r, idx in plan[q].items():
            # what q sends to r is exactly r's ghost block owned by q
            gg = parts[r].global_id[parts[r].n_local:][parts[r].ghost_owner == q]
            assert np.array_equal(parts[q].global_id[idx], gg)


def test_partition_uses_mesh_local_sizes(slope_metis):
    p0 = partition_mesh(slope_metis, 0, 31)
    assert p0.n_local == int(slope_metis.local_sizes[0])


def test_uniform_mesh_is_config_c2_shape():
    m = synthetic.uniform_mesh(20, 16)
    assert m.n_local == 2 * 20 * 16
    assert check_neighbour_symmetry(m)
    assert 2 * 708 * 708 == 1002528  # the 1M configuration


def test_variable_mesh_area_range():
    m = synthetic.variable_mesh(4000)
    assert check_neighbour_symmetry(m)
    a = m.geometry().area
    assert (a > 0).all()
    q = np.quantile(a, [0.02, 0.98])
    assert 5.0 < q[1] / q[0] < 20.0  # ≈10:1 area range
    assert abs(m.n_local - 4000) < 400


def test_morton_order_is_local():
    m = synthetic.uniform_mesh(64, 64)
    nb = m.neigh[m.neigh >= 0]
    ii = np.repeat(np.arange(m.n_local), 3)[(m.neigh >= 0).ravel()]
    assert np.median(np.abs(nb - ii)) <= 8  # neighbours sit close in memory

"""CPU: chm_b200/h5mesh.py — CHM's HDF5 mesh / parameter files decoded without libhdf5.

Known answers: the reference ships the same METIS-permuted mesh twice, as HDF5 (functional_tests/mesh_versioning/
slope.metis_mesh.h5 + slope.metis_param.h5, copied byte for byte into tests/golden/ as DATA fixtures) and as JSON (slope.metis.mesh,
re-encoded as tests/golden/slope_metis.npz by make_golden.py); the un-permuted JSON pair slope.mesh / slope.param is
tests/golden/slope.npz.  The HDF5 bytes must decode to exactly those arrays."""
import os

import numpy as np
import pytest

from chm_b200 import h5mesh
from chm_b200.mesh import check_neighbour_symmetry, partition_mesh
from conftest import GOLDEN, load_mesh

MESH = os.path.join(GOLDEN, "slope.metis_mesh.h5")
PARAM = os.path.join(GOLDEN, "slope.metis_param.h5")


def test_file_structure():
    f = h5mesh.H5File(MESH)
    assert f.names("/") == ["mesh"]
    assert f.names("/mesh") == ["cell_global_id", "elem", "local_sizes", "neighbor", "vertex"]      # triangulation.cpp:528-600
    a = f.attributes("/")                                                                             # :604-640, on the root group
    assert a["/mesh/proj4"] == "+proj=utm +zone=8 +datum=NAD83 +units=m +no_defs"
    assert a["/mesh/version"] == "2.0.0" and a["/mesh/partition_method"] == "metis"
    assert int(a["/mesh/is_geographic"]) == 0 and int(a["/mesh/is_partition"]) == 0
    gid = f.dataset("/mesh/cell_global_id")                     # STD_I32BE on disk (:541): byte order honoured
    assert gid.dtype == np.int32 and np.array_equal(gid, np.arange(2618))
    assert h5mesh.H5File(PARAM).names("/parameters") == ["area", "id"]


def test_mesh_equals_its_json_twin():
    m = h5mesh.read_chm_h5(MESH, [PARAM])
    j = load_mesh("slope_metis")
    assert np.array_equal(m.vertex, j.vertex) and np.array_equal(m.elem, j.elem) and np.array_equal(m.neigh, j.neigh)
    assert np.array_equal(m.local_sizes, j.local_sizes) and len(m.local_sizes) == 31 and m.local_sizes.sum() == m.n_local == 2618
    assert not m.is_geographic and check_neighbour_symmetry(m)


def test_parameters_equal_the_json_parameter_file():
    m = h5mesh.read_chm_h5(MESH, [PARAM])
    j = load_mesh("slope")                                      # slope.mesh + slope.param, un-permuted
    cj, cm = j.vertex[j.elem].mean(axis=1), m.vertex[m.elem].mean(axis=1)
    key = lambda c: np.round(c[:, 0] * 1000).astype(np.int64) * 10 ** 7 + np.round(c[:, 1] * 10).astype(np.int64)
    perm = np.empty(m.n_local, dtype=np.int64)
    perm[np.argsort(key(cm))] = np.argsort(key(cj))             # h5 face i is JSON face perm[i]
    assert np.array_equal(cj[perm], cm)
    assert set(m.params) == {"area", "id"}
    for k in ("area", "id"):
        assert np.array_equal(m.params[k], j.params[k][perm]), k
    assert np.array_equal(m.params["area"], np.abs(m.geometry().area))      # get_area() returns the parameter


def test_partitions_from_the_h5_local_sizes():
    m = h5mesh.read_chm_h5(MESH)
    j = load_mesh("slope_metis")
    for r in (0, 7, 30):
        a, b = partition_mesh(m, r, 31), partition_mesh(j, r, 31)
        assert a.n_local == m.local_sizes[r] and np.array_equal(a.global_id, b.global_id) and np.array_equal(a.neigh, b.neigh)
        assert np.array_equal(a.ghost_owner, b.ghost_owner)


def test_refusals(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(h5mesh.H5FormatError, match="not an HDF5"):
        h5mesh.H5File(str(p))
    raw = bytearray(open(MESH, "rb").read())
    raw[8] = 2                                                   # superblock version 2: refused, not guessed at
    p.write_bytes(bytes(raw))
    with pytest.raises(h5mesh.H5FormatError, match="superblock version 2"):
        h5mesh.H5File(str(p))
    with pytest.raises(h5mesh.H5FormatError, match="not a CHM mesh"):
        h5mesh.read_chm_h5(PARAM)
    with pytest.raises(KeyError):
        h5mesh.H5File(MESH).dataset("/mesh/owner")               # absent in a version-2.0.0 file (:844-858)

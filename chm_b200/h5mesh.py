"""Reader for CHM's HDF5 mesh / parameter files without libhdf5 — there is no HDF5 library or h5py in this image,
so the subset of the HDF5 file format that CHM's writer produces is parsed directly (numpy + struct).

    reference writer   src/mesh/triangulation.cpp:514-697   (to_hdf5: /mesh/{local_sizes, cell_global_id, vertex, elem, neighbor}
                                                             [+ owner, ghost_type in partition files], attributes /mesh/proj4,
                                                             /mesh/version, /mesh/partition_method, /mesh/is_geographic,
                                                             /mesh/is_partition on the ROOT group; /parameters/<name> per file)
    reference reader   src/mesh/triangulation.cpp:724-1036  (load_mesh_from_h5), :1238-1380 (load_partition / parameters),
                                                   :1415-1538 (_load_partition: owner / ghost_type)

What is parsed (HDF5 File Format Specification, version 0 superblock as written by the HDF5 1.8-1.14 C++ API with default
properties): superblock v0 -> root symbol-table entry -> version-1 object headers (continuation blocks followed) -> old-style groups
(v1 B-tree "TREE" + symbol nodes "SNOD" + local heap "HEAP") -> datasets with dataspace v1/v2, datatype classes fixed-point /
floating-point / string / array (versions 1-3), CONTIGUOUS or COMPACT layout (v3), and version 1-3 attributes.  Chunked or filtered
datasets, new-style (fractal-heap) groups and superblock v2/v3 are refused with an explicit error, never guessed at.
Pinned on the reference's own fixture: functional_tests/mesh_versioning/slope.metis_mesh.h5 / slope.metis_param.h5 must decode to
exactly what the JSON twin slope.metis.mesh holds (tests/test_h5mesh.py).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .mesh import TriMesh

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(RuntimeError):
    pass


class _Datatype:
    def __init__(self, dtype: Optional[np.dtype], shape: Tuple[int, ...], size: int, is_string: bool = False):
        self.dtype, self.shape, self.size, self.is_string = dtype, shape, size, is_string


class H5File:
    """Just enough of HDF5 to read CHM's mesh and parameter files."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5FormatError(f"{path}: not an HDF5 file")
        ver = b[8]
        if ver not in (0, 1):
            raise H5FormatError(f"{path}: superblock version {ver} is not supported (CHM's writer produces version 0)")
        self.so, self.sl = b[13], b[14]
        if self.so != 8 or self.sl != 8:
            raise H5FormatError("only 8-byte offsets and lengths are supported")
        off = 24 + (4 if ver == 1 else 0)
        self.base = self._u64(off)
        root_entry = off + 32
        self.root_header = self._u64(root_entry + 8)
        self.objects: Dict[str, int] = {"/": self.root_header}
        self._walk_group("", self.root_header)

    # ---- primitives
    def _u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def _u32(self, o): return struct.unpack_from("<I", self.b, o)[0]
    def _u64(self, o): return struct.unpack_from("<Q", self.b, o)[0]

    def _messages(self, addr: int) -> List[Tuple[int, int, int]]:
        """(type, data offset, size) of every message of the version-1 object header at `addr`, continuations included."""
        b = self.b
        if b[addr] != 1:
            if b[addr:addr + 4] == b"OHDR":
                raise H5FormatError("version-2 object headers are not supported (CHM's writer produces version 1)")
            raise H5FormatError(f"bad object header at {addr:#x}")
        nmsg = self._u16(addr + 2)
        hsize = self._u32(addr + 8)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            o, n = blocks.pop(0)
            end = o + n
            while o + 8 <= end and len(out) < nmsg:
                mtype, msize = self._u16(o), self._u16(o + 2)
                data = o + 8
                if mtype == 0x0010:  # continuation
                    blocks.append((self.base + self._u64(data), self._u64(data + 8)))
                out.append((mtype, data, msize))
                o = data + ((msize + 7) & ~7)
        return out

    def _walk_group(self, prefix: str, header: int):
        for mtype, data, _ in self._messages(header):
            if mtype == 0x0011:  # symbol table: B-tree + local heap
                btree, heap = self.base + self._u64(data), self.base + self._u64(data + 8)
                if self.b[heap:heap + 4] != b"HEAP":
                    raise H5FormatError("bad local heap")
                heap_data = self.base + self._u64(heap + 24)
                for name, obj in self._btree_entries(btree, heap_data):
                    path = f"{prefix}/{name}"
                    self.objects[path] = obj
                    if any(t == 0x0011 for t, _, _ in self._messages(obj)):
                        self._walk_group(path, obj)
            elif mtype in (0x0002, 0x0006):
                raise H5FormatError("new-style groups (link info / link messages) are not supported")

    def _btree_entries(self, addr: int, heap_data: int):
        b = self.b
        if b[addr:addr + 4] != b"TREE":
            raise H5FormatError("bad group B-tree node")
        ntype, level, used = b[addr + 4], b[addr + 5], self._u16(addr + 6)
        if ntype != 0:
            raise H5FormatError("expected a group B-tree")
        o = addr + 24
        for k in range(used):
            child = self.base + self._u64(o + 8)  # key k (8 bytes), child k (8 bytes)
            o += 16
            if level > 0:
                yield from self._btree_entries(child, heap_data)
            else:
                if b[child:child + 4] != b"SNOD":
                    raise H5FormatError("bad symbol table node")
                for s in range(self._u16(child + 6)):
                    e = child + 8 + 40 * s
                    name_off, obj = self._u64(e), self.base + self._u64(e + 8)
                    end = b.index(b"\0", heap_data + name_off)
                    yield b[heap_data + name_off:end].decode(), obj

    # ---- datatype / dataspace
    def _datatype(self, o: int) -> Tuple[_Datatype, int]:
        """Parses the datatype message at o; returns (type, bytes consumed)."""
        b = self.b
        cls, ver = b[o] & 0x0F, b[o] >> 4
        bits0 = b[o + 1]
        size = self._u32(o + 4)
        if cls == 0:  # fixed-point
            order = ">" if bits0 & 1 else "<"
            signed = "i" if bits0 & 8 else "u"
            return _Datatype(np.dtype(f"{order}{signed}{size}"), (), size), 8 + 4
        if cls == 1:  # floating-point
            order = ">" if bits0 & 1 else "<"
            if size not in (4, 8):
                raise H5FormatError("unsupported float size")
            return _Datatype(np.dtype(f"{order}f{size}"), (), size), 8 + 12
        if cls == 3:  # string
            return _Datatype(None, (), size, is_string=True), 8
        if cls == 10:  # array
            rank = b[o + 8]
            p = o + 9 + (3 if ver < 3 else 0)
            dims = tuple(self._u32(p + 4 * k) for k in range(rank))
            p += 4 * rank
            if ver < 3:
                p += 4 * rank  # permutation indices
            base, used = self._datatype(p)
            if base.is_string:
                raise H5FormatError("arrays of strings are not supported")
            return _Datatype(base.dtype, dims + base.shape, size), (p - o) + used
        raise H5FormatError(f"datatype class {cls} is not supported")

    def _dataspace(self, o: int) -> Tuple[int, ...]:
        b = self.b
        ver, rank, flags = b[o], b[o + 1], b[o + 2]
        if ver == 1:
            p = o + 8
        elif ver == 2:
            if b[o + 3] == 2:  # null dataspace
                return (0,)
            p = o + 4
        else:
            raise H5FormatError(f"dataspace version {ver} is not supported")
        return tuple(self._u64(p + 8 * k) for k in range(rank))

    # ---- public
    def names(self, group: str = "/") -> List[str]:
        g = group.rstrip("/") + "/"
        return sorted(p[len(g):] for p in self.objects if p.startswith(g) and p != "/" and "/" not in p[len(g):])

    def __contains__(self, path: str) -> bool:
        return path in self.objects

    def dataset(self, path: str) -> np.ndarray:
        if path not in self.objects:
            raise KeyError(path)
        dt = shape = None
        layout = None
        for mtype, data, size in self._messages(self.objects[path]):
            if mtype == 0x0001:
                shape = self._dataspace(data)
            elif mtype == 0x0003:
                dt, _ = self._datatype(data)
            elif mtype == 0x0008:
                ver, cls = self.b[data], self.b[data + 1]
                if ver != 3:
                    raise H5FormatError(f"{path}: data layout version {ver} is not supported")
                if cls == 1:
                    layout = (self._u64(data + 2), self._u64(data + 10))
                elif cls == 0:
                    n = self._u16(data + 2)
                    layout = (data + 4 - self.base, n)
                else:
                    raise H5FormatError(f"{path}: chunked datasets are not supported (CHM's writer produces contiguous ones)")
            elif mtype == 0x000B:
                raise H5FormatError(f"{path}: filtered datasets are not supported")
        if dt is None or shape is None or layout is None:
            raise H5FormatError(f"{path}: not a dataset")
        if dt.is_string:
            raise H5FormatError(f"{path}: string datasets are not supported")
        n = int(np.prod(shape, dtype=np.int64)) * int(np.prod(dt.shape, dtype=np.int64))
        addr, nbytes = layout
        if n == 0:
            return np.zeros(shape + dt.shape, dtype=dt.dtype.newbyteorder("="))
        if addr == UNDEF:
            raise H5FormatError(f"{path}: no storage allocated")
        if nbytes < n * dt.dtype.itemsize:
            raise H5FormatError(f"{path}: storage smaller than the dataspace")
        a = np.frombuffer(self.b, dtype=dt.dtype, count=n, offset=self.base + addr).reshape(shape + dt.shape)
        return a.astype(dt.dtype.newbyteorder("="))

    def attributes(self, path: str = "/") -> Dict[str, object]:
        out = {}
        b = self.b
        for mtype, data, size in self._messages(self.objects[path]):
            if mtype != 0x000C:
                continue
            ver = b[data]
            nsz, tsz, ssz = self._u16(data + 2), self._u16(data + 4), self._u16(data + 6)
            p = data + 8 + (1 if ver == 3 else 0)
            pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
            name = b[p:p + nsz].split(b"\0")[0].decode()
            p += pad(nsz)
            dt, _ = self._datatype(p)
            p += pad(tsz)
            shape = self._dataspace(p)
            p += pad(ssz)
            n = int(np.prod(shape, dtype=np.int64)) if shape else 1
            if dt.is_string:
                out[name] = b[p:p + dt.size].split(b"\0")[0].decode()
            else:
                cnt = n * int(np.prod(dt.shape, dtype=np.int64))
                v = np.frombuffer(b, dtype=dt.dtype, count=cnt, offset=p).astype(dt.dtype.newbyteorder("="))
                out[name] = v.reshape(shape + dt.shape) if cnt > 1 else v.reshape(-1)[0]
        return out


def read_chm_h5(mesh_path: str, param_paths: Sequence[str] = ()) -> TriMesh:
    """What triangulation::load_mesh_from_h5 + load_partition's parameter read build (triangulation.cpp:724-1036, :1238-1380):
    vertices, elements, neighbours (-1 = none), cell_global_id, local_sizes, is_geographic, per-face parameters.
    The faces are returned in FILE order, which for CHM's tools is already the METIS / RCM permutation (cell_global_id == index)."""
    f = H5File(mesh_path)
    attrs = f.attributes("/")
    version = attrs.get("/mesh/version", "")
    if "/mesh/vertex" not in f or "/mesh/elem" not in f or "/mesh/neighbor" not in f:
        raise H5FormatError(f"{mesh_path}: not a CHM mesh file (no /mesh/vertex, /mesh/elem, /mesh/neighbor)")
    if bool(attrs.get("/mesh/is_partition", 0)):
        raise H5FormatError(f"{mesh_path}: per-rank partition files (is_partition) are not handled here; read the whole-mesh h5 and "
                            "partition it with chm_b200.mesh.partition_mesh (same rule, triangulation.cpp:1482-1531)")
    vertex = np.ascontiguousarray(f.dataset("/mesh/vertex"), dtype=np.float64)
    elem = np.ascontiguousarray(f.dataset("/mesh/elem"), dtype=np.int32)
    neigh = np.ascontiguousarray(f.dataset("/mesh/neighbor"), dtype=np.int32)
    T = elem.shape[0]
    if vertex.ndim != 2 or vertex.shape[1] != 3 or elem.shape != (T, 3) or neigh.shape != (T, 3):
        raise H5FormatError(f"{mesh_path}: unexpected dataset shapes")
    gid = f.dataset("/mesh/cell_global_id").astype(np.int64) if "/mesh/cell_global_id" in f else np.arange(T, dtype=np.int64)
    if not np.array_equal(gid, np.arange(T)):
        raise H5FormatError(f"{mesh_path}: cell_global_id is not the file order; re-run CHM's partition tool")
    local_sizes = f.dataset("/mesh/local_sizes").astype(np.int64) if "/mesh/local_sizes" in f else None
    params: Dict[str, np.ndarray] = {}
    for pp in param_paths:
        pf = H5File(pp)
        for name in pf.names("/parameters"):
            a = pf.dataset("/parameters/" + name).astype(np.float64)
            if a.shape != (T,):
                raise H5FormatError(f"{pp}: parameter {name} has {a.shape[0]} entries for {T} faces")
            params[name] = a
    mesh = TriMesh(vertex, elem, neigh, params, local_sizes=local_sizes, is_geographic=bool(attrs.get("/mesh/is_geographic", 0)))
    mesh.h5_attributes = dict(attrs, version=version)
    return mesh

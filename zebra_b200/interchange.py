"""Interchange with the reference's stored values: thin marshalling over the C ABI codecs.

The reference persists an index as two fjall partitions of bincode(legacy) values plus a `.zebra` file
(/root/reference/src/database/index/lsh.rs:63-119, src/database/core.rs:183-190).  The VALUE formats are handled by
libzebra_b200 (zb_tree_blob_*, zb_zebra_file_*, zb_index_import_store, zb_index_export_*); the key-value engine itself
(fjall's LSM files) stays with the host that owns that crate.  For hosts without it this module reads and writes a
flat dump of the two partitions (one file, `write_store` / `read_store`):

    magic "ZBXSTOR1" | u32 dim | u64 zebra_len | zebra bytes (the `.zebra` file, may be empty)
    partition `trees`:       u64 count | count x { key[16] | u64 value_len | value }
    partition `embeddings`:  u64 count | count x { key[16] | u64 value_len (= 4 dim) | value }

all little endian; values are exactly the bytes the reference stores under those keys.
"""
from __future__ import annotations

import ctypes as C
import struct
import uuid as _uuid
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi

MAGIC = b"ZBXSTOR1"


def tree_blob_decode(dim: int, blob: bytes):
    """bincode(legacy) Node<N> -> (nodes [n,4] int32, coef [p,dim] f32, cst [p] f32, leaf_off [l+1] int64, ids [m,16] u8)."""
    buf = np.frombuffer(bytes(blob), dtype=np.uint8)
    sz = np.zeros(4, dtype=np.int64)
    _ffi.check(_ffi.lib().zb_tree_blob_decode(dim, buf.ctypes.data, buf.size, sz.ctypes.data, None, None, None, None, None))
    nn, npl, nl, nm = (int(v) for v in sz)
    nodes = np.zeros((nn, 4), dtype=np.int32)
    coef = np.zeros((max(npl, 1), dim), dtype=np.float32)
    cst = np.zeros(max(npl, 1), dtype=np.float32)
    leaf_off = np.zeros(nl + 1, dtype=np.int64)
    ids = np.zeros((max(nm, 1), 16), dtype=np.uint8)
    _ffi.check(_ffi.lib().zb_tree_blob_decode(dim, buf.ctypes.data, buf.size, sz.ctypes.data, nodes.ctypes.data, coef.ctypes.data,
                                              cst.ctypes.data, leaf_off.ctypes.data, ids.ctypes.data))
    return nodes, coef[:npl], cst[:npl], leaf_off, ids[:nm]


def tree_blob_encode(dim: int, nodes, root: int, coef, cst, leaf_off, member_ids16) -> bytes:
    nodes = np.ascontiguousarray(nodes, dtype=np.int32).reshape(-1, 4)
    coef = np.ascontiguousarray(coef, dtype=np.float32).reshape(-1, dim)
    cst = np.ascontiguousarray(cst, dtype=np.float32)
    leaf_off = np.ascontiguousarray(leaf_off, dtype=np.int64)
    ids = np.ascontiguousarray(member_ids16, dtype=np.uint8).reshape(-1, 16)
    need = C.c_uint64()
    args = (dim, nodes.shape[0], nodes.ctypes.data, int(root), coef.ctypes.data if coef.size else None,
            cst.ctypes.data if cst.size else None, leaf_off.ctypes.data, ids.ctypes.data if ids.size else None)
    _ffi.check(_ffi.lib().zb_tree_blob_encode(*args, None, 0, C.byref(need)))
    out = np.empty(int(need.value), dtype=np.uint8)
    _ffi.check(_ffi.lib().zb_tree_blob_encode(*args, out.ctypes.data, out.size, C.byref(need)))
    return out.tobytes()


def zebra_file_encode(db_uuid: _uuid.UUID, metric, max_node_size: int, num_trees: int) -> bytes:
    """core.rs:183-190: the bytes of `<uuid>.zebra` for a Database<N, metric, Model>."""
    out = np.empty(64, dtype=np.uint8)
    need = C.c_uint64()
    ub = np.frombuffer(db_uuid.bytes, dtype=np.uint8)
    _ffi.check(_ffi.lib().zb_zebra_file_encode(ub.ctypes.data, metric.METRIC, metric.power, max_node_size, num_trees,
                                               out.ctypes.data, out.size, C.byref(need)))
    return out[: int(need.value)].tobytes()


def zebra_file_decode(data: bytes, metric) -> Tuple[_uuid.UUID, int, int, int]:
    """-> (uuid, metric power, max_node_size, num_trees); `metric` says which Met the file was written for
    (the reference knows it from the type parameter, core.rs:92-96)."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    ub = np.zeros(16, dtype=np.uint8)
    power, mns, nt = C.c_int32(), C.c_uint64(), C.c_uint64()
    _ffi.check(_ffi.lib().zb_zebra_file_decode(buf.ctypes.data if buf.size else None, buf.size, metric.METRIC, ub.ctypes.data,
                                               C.byref(power), C.byref(mns), C.byref(nt)))
    return _uuid.UUID(bytes=ub.tobytes()), int(power.value), int(mns.value), int(nt.value)


def store_flatten(dim: int, ids16, tree_blobs: Sequence[bytes]):
    """A whole store -> (report dict, Forest arrays, row_order, orphan_rows): the flat forest zb_index_load_forest takes
    with ordinals in id order (pure host code; zb_index_import_store = this + the load)."""
    ids = np.ascontiguousarray(ids16, dtype=np.uint8).reshape(-1)
    n = ids.size // 16
    bufs = [np.frombuffer(bytes(b), dtype=np.uint8) for b in tree_blobs]
    ptrs = (C.c_void_p * max(1, len(bufs)))(*[b.ctypes.data for b in bufs])
    sizes = np.array([b.size for b in bufs] or [0], dtype=np.uint64)
    h, rep = C.c_void_p(), _ffi.ImportReport()
    L = _ffi.lib()
    _ffi.check(L.zb_store_flatten(dim, n, ids.ctypes.data if n else None, len(bufs), ptrs, sizes.ctypes.data, C.byref(h), C.byref(rep)))
    try:
        sz = np.zeros(4, dtype=np.int64)
        p = [C.c_void_p() for _ in range(8)]
        _ffi.check(L.zb_flat_store_view(h, sz.ctypes.data, *[C.byref(x) for x in p]))
        nn, npl, nl, nm = (int(v) for v in sz)

        def arr(ptr, count, dtype):
            if not count:
                return np.zeros(0, dtype=dtype)
            return np.frombuffer((C.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(ptr.value), dtype=dtype).copy()

        forest = dict(nodes=arr(p[0], nn * 4, np.int32).reshape(-1, 4), roots=arr(p[1], len(bufs), np.int32),
                      coef=arr(p[2], npl * dim, np.float32).reshape(-1, dim), cst=arr(p[3], npl, np.float32),
                      leaf_off=arr(p[4], nl + 1, np.int64), members=arr(p[5], nm, np.uint64))
        return rep.as_dict(), forest, arr(p[6], int(rep.rows_loaded), np.uint32), arr(p[7], int(rep.orphan_rows), np.uint32)
    finally:
        L.zb_flat_store_free(h)


def tree_key(t: int) -> bytes:
    """A UUIDv7-shaped key for tree t of an exported store (the reference mints Uuid::now_v7(), lsh.rs:423)."""
    return _uuid.UUID(int=(0x7 << 76) | (0x2 << 62) | t).bytes


def write_store(path: str, dim: int, zebra: bytes, trees: Sequence[Tuple[bytes, bytes]], ids16: np.ndarray, rows: np.ndarray) -> None:
    ids16 = np.ascontiguousarray(ids16, dtype=np.uint8).reshape(-1, 16)
    rows = np.ascontiguousarray(rows, dtype="<f4").reshape(-1, dim)
    if ids16.shape[0] != rows.shape[0]:
        raise ValueError("one id per row")
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<IQ", dim, len(zebra)) + zebra)
        f.write(struct.pack("<Q", len(trees)))
        for key, blob in trees:
            f.write(bytes(key) + struct.pack("<Q", len(blob)))
            f.write(blob)
        f.write(struct.pack("<Q", rows.shape[0]))
        rec = np.zeros(rows.shape[0], dtype=[("k", "u1", 16), ("n", "<u8"), ("v", "<f4", dim)])
        rec["k"], rec["n"], rec["v"] = ids16, 4 * dim, rows
        rec.tofile(f)


def read_store(path: str):
    """-> (dim, zebra bytes, [(tree key, tree blob)], ids [n,16] uint8, rows [n,dim] float32)."""
    with open(path, "rb") as f:
        head = f.read(20)
        if len(head) != 20 or head[:8] != MAGIC:
            raise ValueError(f"{path} is not a zebra_b200 store dump")
        dim, zlen = struct.unpack("<IQ", head[8:])
        zebra = f.read(zlen)
        (nt,) = struct.unpack("<Q", f.read(8))
        trees = []
        for _ in range(nt):
            key = f.read(16)
            (ln,) = struct.unpack("<Q", f.read(8))
            blob = f.read(ln)
            if len(blob) != ln:
                raise ValueError("truncated tree value")
            trees.append((key, blob))
        (n,) = struct.unpack("<Q", f.read(8))
        rec = np.fromfile(f, dtype=[("k", "u1", 16), ("n", "<u8"), ("v", "<f4", dim)], count=n)
        if rec.shape[0] != n or (n and not np.all(rec["n"] == 4 * dim)):
            raise ValueError("truncated or malformed embeddings partition")
    return dim, zebra, trees, np.ascontiguousarray(rec["k"]), np.ascontiguousarray(rec["v"], dtype=np.float32)

"""ctypes binding of the CPU oracle (oracle/zb_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, bench.py's cpu_baseline / ``--impl reference`` legs and ``__graft_entry__.smoke()`` may import
this module; the product package ``zebra_b200`` never does.  PARITY UNPINNED (see the C file's header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzb_oracle.so")

COSINE, L2SQ, L2 = 0, 1, 2
# the scalar metrics of distance.rs:51-190 (f32 bits zero-extended); Minkowski / p-norm carry the power in bits 8..
CHEBYSHEV, CANBERRA, BRAY_CURTIS, MANHATTAN, L3, L4, HAMMING = 3, 4, 5, 6, 7, 8, 9


def MINKOWSKI(power: int) -> int:
    return 10 | (int(power) << 8)


def PNORM(power: int) -> int:
    return 11 | (int(power) << 8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "zb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u64, i64, i32 = C.c_void_p, C.c_uint64, C.c_int64, C.c_int
        L.zbo_create.restype = vp
        L.zbo_create.argtypes = [i32, i32, u64, i32, u64]
        L.zbo_destroy.argtypes = [vp]
        L.zbo_num_rows.restype = u64
        L.zbo_num_rows.argtypes = [vp]
        L.zbo_num_live.restype = u64
        L.zbo_num_live.argtypes = [vp]
        L.zbo_add.argtypes = [vp, u64, vp, vp]
        L.zbo_remove.argtypes = [vp, u64, vp, vp]
        L.zbo_deduplicate.restype = i64
        L.zbo_deduplicate.argtypes = [vp, vp, u64]
        L.zbo_search.restype = i64
        L.zbo_search.argtypes = [vp, vp, u64, vp, vp]
        L.zbo_search_batch.argtypes = [vp, u64, vp, u64, i32, vp, vp, vp]
        L.zbo_forest_sizes.argtypes = [vp, vp]
        L.zbo_export_forest.argtypes = [vp] + [vp] * 6
        L.zbo_load_forest.argtypes = [vp, u64] + [vp] * 7
        L.zbo_renumber_leaves.argtypes = [vp]
        L.zbo_hash.argtypes = [vp, u64, vp, vp, vp, vp]
        L.zbo_trace.restype = i64
        L.zbo_trace.argtypes = [vp, vp, u64, vp, u64]
        L.zbo_candidates.restype = i64
        L.zbo_candidates.argtypes = [vp, vp, u64, vp, u64]
        L.zbo_dot_f32.restype = C.c_double
        L.zbo_dot_f32.argtypes = [vp, vp, i32]
        L.zbo_l2sq_f32.restype = C.c_double
        L.zbo_l2sq_f32.argtypes = [vp, vp, i32]
        L.zbo_cos_f32.restype = C.c_double
        L.zbo_cos_f32.argtypes = [vp, vp, i32]
        L.zbo_root_p.restype = C.c_float
        L.zbo_root_p.argtypes = [C.c_float, i32]
        L.zbo_distance_bits.restype = u64
        L.zbo_distance_bits.argtypes = [i32, vp, vp, i32]
        L.zbo_distance_bits_batch.argtypes = [i32, u64, vp, vp, i32, vp]
        L.zbo_above_batch.argtypes = [u64, vp, vp, vp, i32, vp]
        L.zbo_point_is_above.restype = i32
        L.zbo_point_is_above.argtypes = [vp, C.c_float, vp, i32]
        L.zbo_make_plane.argtypes = [vp, vp, i32, vp, vp]
        L.zbo_force_scalar.argtypes = [i32]
        L.zbo_using_avx512.restype = i32
        L.zbo_set_build_threads.argtypes = [i32]
        L.zbo_synth_fill.argtypes = [vp, u64, u64, u64, C.c_uint32, u64, C.c_uint32, i32]
        L.zbo_mix64.restype = u64
        L.zbo_mix64.argtypes = [u64]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def dot(a, b) -> float:
    a, b = _f32(a), _f32(b)
    return lib().zbo_dot_f32(_p(a), _p(b), a.size)


def l2sq(a, b) -> float:
    a, b = _f32(a), _f32(b)
    return lib().zbo_l2sq_f32(_p(a), _p(b), a.size)


def cosdist(a, b) -> float:
    a, b = _f32(a), _f32(b)
    return lib().zbo_cos_f32(_p(a), _p(b), a.size)


def root_p(s: float, p: int) -> float:
    """The deterministic p-th root the scalar metrics use in place of cbrtf / powf(., 1/p)."""
    return lib().zbo_root_p(float(s), int(p))


def distance_bits(metric: int, row, query) -> int:
    a, b = _f32(row), _f32(query)
    return lib().zbo_distance_bits(metric, _p(a), _p(b), a.size)


def distance_bits_batch(metric: int, rows, queries) -> np.ndarray:
    a, b = _f32(rows), _f32(queries)
    out = np.empty(a.shape[0], dtype=np.uint64)
    lib().zbo_distance_bits_batch(metric, a.shape[0], _p(a), _p(b), a.shape[1], _p(out))
    return out


def above_batch(coef, cst, x) -> np.ndarray:
    coef, cst, x = _f32(coef), _f32(cst), _f32(x)
    out = np.empty(coef.shape[0], dtype=np.uint8)
    lib().zbo_above_batch(coef.shape[0], _p(coef), _p(cst), _p(x), coef.shape[1], _p(out))
    return out


def point_is_above(coef, constant: float, x) -> bool:
    coef, x = _f32(coef), _f32(x)
    return bool(lib().zbo_point_is_above(_p(coef), float(constant), _p(x), coef.size))


def make_plane(a, b):
    a, b = _f32(a), _f32(b)
    coef = np.empty_like(a)
    cst = np.empty(1, dtype=np.float32)
    lib().zbo_make_plane(_p(a), _p(b), a.size, _p(coef), _p(cst))
    return coef, float(cst[0])


def synth(first_row: int, row_stride: int, n: int, dim: int, seed: int, kind: int = 0, nthreads: int = 0) -> np.ndarray:
    """The synthetic rows of BASELINE.md, generated on the CPU (bit-identical to zb_synth_fill_device)."""
    out = np.empty((n, dim), dtype=np.float32)
    lib().zbo_synth_fill(_p(out), first_row, row_stride, n, dim, seed, kind, nthreads or (os.cpu_count() or 1))
    return out


def set_build_threads(n: int) -> None:
    lib().zbo_set_build_threads(n)


class Forest:
    """Flat forest arrays (the layout zb_index_load_forest / zbo_load_forest take)."""

    def __init__(self, nodes, roots, coef, cst, leaf_off, members):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.int32).reshape(-1, 4)
        self.roots = np.ascontiguousarray(roots, dtype=np.int32)
        self.coef = _f32(coef)
        self.cst = _f32(cst)
        self.leaf_off = np.ascontiguousarray(leaf_off, dtype=np.int64)
        self.members = np.ascontiguousarray(members, dtype=np.uint64)


class OracleIndex:
    """The reference's LSHIndex (lsh.rs:145-566) restated over an in-memory store; ids are row ordinals."""

    def __init__(self, dim: int, metric: int, max_node_size: int = 5, num_trees: int = 15, seed: int = 0):
        self.dim, self.metric, self.max_node_size, self.num_trees, self.seed = dim, metric, max_node_size, num_trees, seed
        self._h = lib().zbo_create(dim, metric, max_node_size, num_trees, seed)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().zbo_destroy(self._h)
            self._h = None

    @property
    def num_rows(self) -> int:
        return lib().zbo_num_rows(self._h)

    @property
    def num_live(self) -> int:
        return lib().zbo_num_live(self._h)

    def add(self, rows) -> np.ndarray:
        rows = _f32(rows).reshape(-1, self.dim)
        ids = np.empty(rows.shape[0], dtype=np.uint64)
        lib().zbo_add(self._h, rows.shape[0], _p(rows), _p(ids))
        return ids

    def remove(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.empty(ids.size, dtype=np.uint8)
        lib().zbo_remove(self._h, ids.size, _p(ids), _p(out))
        return out.astype(bool)

    def deduplicate(self) -> np.ndarray:
        """lsh.rs:270-288: removes every row whose bits equal those of a row with a smaller id; returns the removed ids."""
        out = np.empty(max(1, self.num_rows), dtype=np.uint64)
        n = lib().zbo_deduplicate(self._h, _p(out), out.size)
        return out[:n].copy()

    def search(self, query, top_k: int):
        q = _f32(query)
        ids = np.empty(max(top_k, 1), dtype=np.uint64)
        bits = np.empty(max(top_k, 1), dtype=np.uint64)
        r = lib().zbo_search(self._h, _p(q), top_k, _p(ids), _p(bits))
        return ids[:r].copy(), bits[:r].copy()

    def search_batch(self, queries, top_k: int, nthreads: int = 1):
        q = _f32(queries).reshape(-1, self.dim)
        nq = q.shape[0]
        ids = np.full((nq, max(top_k, 1)), np.iinfo(np.uint64).max, dtype=np.uint64)
        bits = np.full((nq, max(top_k, 1)), np.iinfo(np.uint64).max, dtype=np.uint64)
        counts = np.zeros(nq, dtype=np.uint32)
        lib().zbo_search_batch(self._h, nq, _p(q), top_k, nthreads, _p(ids), _p(bits), _p(counts))
        return ids[:, :top_k], bits[:, :top_k], counts

    def export_forest(self) -> Forest:
        sz = np.zeros(4, dtype=np.int64)
        lib().zbo_forest_sizes(self._h, _p(sz))
        nn, npl, nl, nm = (int(v) for v in sz)
        nodes = np.zeros((nn, 4), dtype=np.int32)
        roots = np.zeros(self.num_trees, dtype=np.int32)
        coef = np.zeros((npl, self.dim), dtype=np.float32)
        cst = np.zeros(npl, dtype=np.float32)
        leaf_off = np.zeros(nl + 1, dtype=np.int64)
        members = np.zeros(nm, dtype=np.uint64)
        lib().zbo_export_forest(self._h, _p(nodes), _p(roots), _p(coef), _p(cst), _p(leaf_off), _p(members))
        return Forest(nodes, roots, coef, cst, leaf_off, members)

    def load_forest(self, rows, forest: Forest):
        rows = _f32(rows).reshape(-1, self.dim)
        f = forest
        lib().zbo_load_forest(self._h, rows.shape[0], _p(rows), _p(f.nodes), _p(f.roots), _p(f.coef), _p(f.cst),
                              _p(f.leaf_off), _p(f.members))

    def hash(self, rows):
        rows = _f32(rows).reshape(-1, self.dim)
        n = rows.shape[0]
        lib().zbo_renumber_leaves(self._h)
        keys = np.zeros((n, self.num_trees), dtype=np.uint64)
        depth = np.zeros((n, self.num_trees), dtype=np.uint32)
        leaf = np.zeros((n, self.num_trees), dtype=np.int32)
        rc = lib().zbo_hash(self._h, n, _p(rows), _p(keys), _p(depth), _p(leaf))
        if rc != 0:
            raise RuntimeError("oracle: hash on an index without trees")
        return keys, depth, leaf

    def trace(self, query, top_k: int) -> np.ndarray:
        """Visit plan of one query: rows of (tree, leaf, nprime, live) in visit order."""
        q = _f32(query)
        lib().zbo_renumber_leaves(self._h)
        cap = 1 << 16
        out = np.zeros((cap, 4), dtype=np.int32)
        n = lib().zbo_trace(self._h, _p(q), top_k, _p(out), cap)
        return out[: min(n, cap)].copy()

    def candidates(self, query, top_k: int) -> np.ndarray:
        q = _f32(query)
        cap = 1 << 20
        out = np.zeros(cap, dtype=np.uint64)
        n = lib().zbo_candidates(self._h, _p(q), top_k, _p(out), cap)
        return out[: min(n, cap)].copy()

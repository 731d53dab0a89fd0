"""Flat tables (north_star (a): the K-bit LSH table as the special case of the reference's tree in which every node at
depth d of a tree shares one hyperplane; /root/reference/src/database/index/lsh.rs:39-43, :46-60, :350-366), CPU side:

* the HOST TWIN of the projection kernel's register tile (tests/project_twin.cpp compiles zebra_b200/csrc/zb_project.cuh
  for the CPU and replays a quad: pj_chunk per thread, then the quad_reduce16 fold) bit for bit against the oracle's dot;
* the ballot -> key bit order;
* the numpy builder of the equivalent forest (complete trees in preorder), which the GPU tests load into the oracle:
  the oracle's own tree walk over it must give exactly the keys the dense sign test gives.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import zb_oracle as zo

HERE = os.path.dirname(os.path.abspath(__file__))
F32 = np.float32


def flat_planes(rng, rows, T, K):
    """T * K hyperplanes through midpoints of row pairs (lsh.rs:222-225); planes are INPUT to both sides (survey Q8)."""
    n, dim = rows.shape
    coef = np.empty((T * K, dim), F32)
    cst = np.empty(T * K, F32)
    for h in range(T * K):
        a, b = rows[rng.choice(n, 2, replace=False)]
        coef[h] = b - a
        cst[h] = -F32(np.dot(coef[h].astype(np.float64), ((a + b) / F32(2)).astype(np.float64)))
    return coef, cst


def flat_keys(rows, coef, cst, T, K):
    """keys[o, t]: the K sign bits of table t, MSB = bit 0, through the oracle's point_is_above."""
    n = rows.shape[0]
    above = np.empty((n, T * K), np.uint8)
    for h in range(T * K):
        above[:, h] = zo.above_batch(np.broadcast_to(coef[h], rows.shape), np.full(n, cst[h], F32), rows)
    keys = np.zeros((n, T), np.uint64)
    for t in range(T):
        for d in range(K):
            keys[:, t] = (keys[:, t] << np.uint64(1)) | above[:, t * K + d].astype(np.uint64)
    return keys


def flat_forest(rows, coef, cst, T, K):
    """The equivalent forest: T complete trees of depth K, preorder numbering (node, left = below, right = above), leaf
    number = t * 2^K + key, members ascending by ordinal."""
    keys = flat_keys(rows, coef, cst, T, K)
    per_tree, lpt = (2 << K) - 1, 1 << K
    nodes = np.full((per_tree * T, 4), -1, np.int32)
    roots = np.zeros(T, np.int32)
    for t in range(T):
        idx = per_tree * t
        roots[t] = idx

        def emit(depth, prefix):
            nonlocal idx
            me = idx
            idx += 1
            if depth == K:
                nodes[me] = (-1, -1, -1, lpt * t + prefix)
                return me
            nodes[me, 0] = t * K + depth
            nodes[me, 1] = emit(depth + 1, prefix << 1)
            nodes[me, 2] = emit(depth + 1, (prefix << 1) | 1)
            return me

        emit(0, 0)
    leaf = (np.arange(T, dtype=np.int64)[None, :] * lpt + keys.astype(np.int64))          # [n, T]
    counts = np.bincount(leaf.ravel(), minlength=lpt * T)
    leaf_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    order = np.lexsort((np.repeat(np.arange(rows.shape[0]), T), leaf.ravel()))              # by leaf, then ordinal
    members = np.repeat(np.arange(rows.shape[0]), T)[order].astype(np.uint64)
    return zo.Forest(nodes, roots, coef, cst, leaf_off, members), keys


@pytest.fixture(scope="module")
def twin(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("pj") / "libproject_twin.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-o", out,
                           os.path.join(HERE, "project_twin.cpp")])
    L = C.CDLL(out)
    L.twin_project.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.twin_key_from_ballots.restype = C.c_uint64
    L.twin_key_from_ballots.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    return L


@pytest.mark.parametrize("n,H,dim", [(1, 1, 16), (5, 3, 20), (37, 18, 100), (64, 16, 384), (33, 7, 768)])
def test_projection_tile_twin_equals_oracle_dot(twin, n, H, dim):
    rng = np.random.default_rng(n * H)
    rows = rng.standard_normal((n, dim)).astype(F32)
    planes = rng.standard_normal((H, dim)).astype(F32)
    rows[0, :] *= F32(1e-20)                       # underflowing products
    planes[-1, :] *= F32(1e20)
    dots = np.zeros((n, H), F32)
    twin.twin_project(rows.ctypes.data, n, planes.ctypes.data, H, dim, dots.ctypes.data)
    exp = np.array([[F32(zo.dot(planes[h], rows[i])) for h in range(H)] for i in range(n)], F32)
    assert np.array_equal(dots.view(np.uint32), exp.view(np.uint32))


def test_ballot_to_key_bit_order(twin):
    rng = np.random.default_rng(0)
    for K in (1, 2, 5, 16, 31, 32, 33, 47, 64):
        for _ in range(50):
            bits = rng.integers(0, 2, K)
            key = 0
            for b in bits:                          # MSB = plane 0 = the root decision (lsh.rs:358-363)
                key = (key << 1) | int(b)
            b0 = sum(int(bits[l]) << l for l in range(min(K, 32)))
            b1 = sum(int(bits[32 + l]) << l for l in range(max(0, K - 32)))
            assert twin.twin_key_from_ballots(b0, b1, K) == key


@pytest.mark.parametrize("T,K,dim,n", [(3, 5, 20, 400), (2, 9, 48, 700), (1, 1, 16, 50)])
def test_equivalent_forest_walks_to_the_dense_keys(T, K, dim, n):
    rng = np.random.default_rng(T * 100 + K)
    rows = rng.standard_normal((n, dim)).astype(F32)
    coef, cst = flat_planes(rng, rows, T, K)
    forest, keys = flat_forest(rows, coef, cst, T, K)
    orc = zo.OracleIndex(dim, zo.L2SQ, 10**9, T, seed=1)
    orc.load_forest(rows, forest)
    back = orc.export_forest()                       # preorder numbering is the oracle's own; an export gives every inner
    assert np.array_equal(back.nodes[:, 1:], forest.nodes[:, 1:])     # node its own copy of the (shared) plane
    assert np.array_equal(back.roots, forest.roots) and np.array_equal(back.leaf_off, forest.leaf_off)
    inner = forest.nodes[:, 0] >= 0
    assert np.array_equal(back.coef.reshape(-1, dim)[back.nodes[inner, 0]], coef[forest.nodes[inner, 0]])
    assert np.array_equal(back.cst[back.nodes[inner, 0]], cst[forest.nodes[inner, 0]])
    fresh = rng.standard_normal((200, dim)).astype(F32)
    for x, want in ((rows, keys), (fresh, flat_keys(fresh, coef, cst, T, K))):
        k, d, l = orc.hash(x)
        assert np.array_equal(k, want) and np.all(d == K)
        assert np.array_equal(l, np.arange(T)[None, :] * (1 << K) + want.astype(np.int64))
    ids, bits = orc.search(rows[3], 5)
    assert ids[0] == 3 and bits[0] == 0

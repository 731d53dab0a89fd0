"""Known-answer and property tests for the oracle's tree walk / search / build / remove.
Hand-derived from /root/reference/src/database/index/lsh.rs:290-348 (tree_result), :544-565 (search),
:250-267 (build_a_tree), :350-382 (insert), :473-503 (remove; divergence D1)."""
import struct

import numpy as np
import pytest

import pyref
from oracle import zb_oracle as zo


def f64bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def kat_forest(dim=4):
    """Root splits on x0, both children split on x1.  Leaves (preorder): A=1 row, B=3, C=2, D=5."""
    coef = np.zeros((3, dim), dtype=np.float32)
    coef[0, 0] = 1.0   # root:  x0 >= 0
    coef[1, 1] = 1.0   # left:  x1 >= 0
    coef[2, 1] = 1.0   # right: x1 >= 0
    cst = np.zeros(3, dtype=np.float32)
    #            plane left right leaf
    nodes = [[0, 1, 4, -1],
             [1, 2, 3, -1], [-1, -1, -1, 0], [-1, -1, -1, 1],
             [2, 5, 6, -1], [-1, -1, -1, 2], [-1, -1, -1, 3]]
    rows = np.zeros((11, dim), dtype=np.float32)
    rows[0, :2] = (-5, -5)                                            # A
    rows[1, :2] = (-5, 5); rows[2, :2] = (-6, 6); rows[3, :2] = (-7, 7)   # B
    rows[4, :2] = (0.5, -1); rows[5, :2] = (1, -1)                      # C
    for i in range(5):
        rows[6 + i, :2] = (5 + i, 5 + i)                                # D
    leaf_off = [0, 1, 4, 6, 11]
    members = np.arange(11, dtype=np.uint64)
    return zo.Forest(nodes, [0], coef, cst, leaf_off, members), rows


def make_kat(metric=zo.L2SQ):
    forest, rows = kat_forest()
    ix = zo.OracleIndex(4, metric, max_node_size=6, num_trees=1)
    ix.load_forest(rows, forest)
    return ix, forest, rows


def test_q1_count_cascade_and_q2_truncation():
    ix, _, _ = make_kat()
    q = np.array([-1, -1, 0, 0], dtype=np.float32)
    # walk: root below -> p1 below -> A(n=4): 1 < 4 -> all, returns 1 -> B(n=3): 3 !< 3 -> top-3, returns 3
    # p1 returns 3 (the 1 from A is DROPPED, lsh.rs:342) -> root: 3 < 4 -> backup subtree with n=1
    # -> p2 below -> C(n=1): 2 !< 1 -> top-1 (row 4), returns 1 -> not < 1 -> done.  D never visited.
    tr = ix.trace(q, 4)
    assert tr.tolist() == [[0, 0, 4, 1], [0, 1, 3, 3], [0, 2, 1, 2]]
    assert ix.candidates(q, 4).tolist() == [0, 1, 2, 3, 4]
    ids, bits = ix.search(q, 4)
    # row 5 (distance 4.0) is closer than rows 0,1,2 but was cut by the leaf truncation (Q2)
    assert ids.tolist() == [4, 0, 1, 2]
    assert bits.tolist() == [f64bits(2.25), f64bits(32.0), f64bits(52.0), f64bits(74.0)]


def test_full_first_leaf_stops_walk():
    ix, _, _ = make_kat()
    q = np.array([4, 4, 0, 0], dtype=np.float32)
    assert ix.trace(q, 4).tolist() == [[0, 3, 4, 5]]
    ids, _ = ix.search(q, 4)
    assert ids.tolist() == [6, 7, 8, 9]


def test_tombstone_whole_main_leaf_gives_backup_full_budget():
    ix, _, _ = make_kat()
    assert ix.remove([0]).tolist() == [True]
    assert ix.remove([0]).tolist() == [False]        # already removed
    q = np.array([-1, -1, 0, 0], dtype=np.float32)
    # A is now empty: k = 0 -> B gets the full n = 4: 3 < 4 -> all 3, returns 3 -> root: 3 < 4 -> C with n=1
    assert ix.trace(q, 4).tolist() == [[0, 0, 4, 0], [0, 1, 4, 3], [0, 2, 1, 2]]
    ids, _ = ix.search(q, 4)
    assert ids.tolist() == [4, 1, 2, 3]
    assert 0 not in ids.tolist()


def test_topk_larger_than_live_rows_returns_short():
    ix, _, _ = make_kat()
    q = np.array([-1, -1, 0, 0], dtype=np.float32)
    ids, bits = ix.search(q, 50)
    assert len(ids) == 11 and sorted(ids.tolist()) == list(range(11))
    assert all(bits[i] <= bits[i + 1] for i in range(len(bits) - 1))
    ids0, _ = ix.search(q, 0)
    assert len(ids0) == 0


def test_duplicate_rows_tie_break_by_id():
    dim = 16
    rows = np.zeros((8, dim), dtype=np.float32)
    rows[:, 0] = [1, 1, 1, 1, 2, 2, 3, 3]
    ix = zo.OracleIndex(dim, zo.L2SQ, max_node_size=100, num_trees=3)
    ix.add(rows)
    q = np.zeros(dim, dtype=np.float32)
    ids, bits = ix.search(q, 5)
    assert ids.tolist() == [0, 1, 2, 3, 4]
    assert bits.tolist() == [f64bits(1.0)] * 4 + [f64bits(4.0)]
    ix.remove([1])
    ids, _ = ix.search(q, 5)
    assert ids.tolist() == [0, 2, 3, 4, 5]


def test_cosine_ranking_quirk_q4():
    # ascending u64 order ranks the SMALLEST non-negative similarity first, negatives last
    dim = 16
    rows = np.zeros((4, dim), dtype=np.float32)
    rows[0, 0] = 1.0                      # same direction as q  -> 1.0
    rows[1, 1] = 1.0                      # orthogonal           -> 0.0
    rows[2, 0] = -1.0                     # opposite             -> -1.0 (sign bit: sorts last)
    rows[3, 0] = 1.0; rows[3, 1] = 1.0    # 45 degrees           -> ~0.7071
    ix = zo.OracleIndex(dim, zo.COSINE, max_node_size=100, num_trees=1)
    ix.add(rows)
    q = np.zeros(dim, dtype=np.float32); q[0] = 1.0
    ids, bits = ix.search(q, 4)
    assert ids.tolist() == [1, 3, 0, 2]
    assert bits[0] == f64bits(0.0) and bits[2] == f64bits(1.0) and bits[3] == f64bits(-1.0)


def _pyforest(ix, rows, tomb):
    f = ix.export_forest()
    return pyref.PyForest(f, rows, tomb,
                          lambda r, q: zo.distance_bits(ix.metric, r, q),
                          lambda c, k, q: zo.point_is_above(c, k, q))


@pytest.mark.parametrize("metric", [zo.COSINE, zo.L2SQ, zo.L2])
@pytest.mark.parametrize("mns,trees,k", [(5, 15, 10), (8, 3, 3), (32, 4, 10), (3, 2, 25)])
def test_c_walk_equals_python_restatement(metric, mns, trees, k):
    rng = np.random.default_rng(mns * 100 + trees)
    dim, n = 32, 400
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    ix = zo.OracleIndex(dim, metric, max_node_size=mns, num_trees=trees, seed=5)
    ix.add(rows)
    dead = rng.choice(n, 60, replace=False)
    ix.remove(dead)
    tomb = np.zeros(n, dtype=bool); tomb[dead] = True
    py = _pyforest(ix, rows, tomb)
    qs = np.concatenate([rng.standard_normal((6, dim)).astype(np.float32), rows[:3]])
    for q in qs:
        ids, bits = ix.search(q, k)
        exp, cand, trace = py.search(q, k)
        assert ids.tolist() == [i for _, i in exp]
        assert bits.tolist() == [b for b, _ in exp]
        assert ix.candidates(q, k).tolist() == cand
        assert [tuple(r) for r in ix.trace(q, k).tolist()] == trace
        assert not set(ids.tolist()) & set(dead.tolist())


def test_build_invariants_and_determinism():
    rng = np.random.default_rng(0)
    dim, n, mns = 48, 1500, 16
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    a = zo.OracleIndex(dim, zo.L2SQ, mns, 4, seed=9); a.add(rows)
    b = zo.OracleIndex(dim, zo.L2SQ, mns, 4, seed=9); b.add(rows)
    fa, fb = a.export_forest(), b.export_forest()
    for name in ("nodes", "roots", "coef", "cst", "leaf_off", "members"):
        assert np.array_equal(getattr(fa, name), getattr(fb, name))
    c = zo.OracleIndex(dim, zo.L2SQ, mns, 4, seed=10); c.add(rows)
    assert not np.array_equal(fa.coef[:4], c.export_forest().coef[:4])
    sizes = np.diff(fa.leaf_off)
    assert sizes.max() < mns                      # lsh.rs:251: leaf iff len < max_node_size
    # every tree partitions all rows exactly once, and members obey every plane on their path
    keys, depth, leaf = a.hash(rows)
    for t in range(4):
        lt = leaf[:, t]
        for lf in np.unique(lt):
            mem = fa.members[fa.leaf_off[lf]:fa.leaf_off[lf + 1]]
            assert sorted(mem.tolist()) == np.nonzero(lt == lf)[0].tolist()
    assert sizes.sum() == 4 * n


def test_incremental_add_splits_overfull_leaves_and_search_sees_new_rows():
    rng = np.random.default_rng(1)
    dim, mns = 32, 8
    rows = rng.standard_normal((300, dim)).astype(np.float32)
    ix = zo.OracleIndex(dim, zo.L2SQ, mns, 3, seed=1)
    ids0 = ix.add(rows[:200])
    ids1 = ix.add(rows[200:])
    assert ids0.tolist() == list(range(200)) and ids1.tolist() == list(range(200, 300))
    f = ix.export_forest()
    assert np.diff(f.leaf_off).max() <= mns          # lsh.rs:368: a leaf may hold max_node_size after a push
    assert np.diff(f.leaf_off).sum() == 3 * 300
    for i in (0, 150, 250, 299):
        ids, bits = ix.search(rows[i], 1)
        assert ids.tolist() == [i] and bits.tolist() == [0]
    py = _pyforest(ix, rows, np.zeros(300, dtype=bool))
    for q in rng.standard_normal((4, dim)).astype(np.float32):
        ids, bits = ix.search(q, 10)
        exp, _, _ = py.search(q, 10)
        assert ids.tolist() == [i for _, i in exp]


def test_batch_search_threads_equal_single():
    rng = np.random.default_rng(2)
    rows = rng.standard_normal((2000, 64)).astype(np.float32)
    ix = zo.OracleIndex(64, zo.COSINE, 5, 15, seed=3)
    ix.add(rows)
    qs = rng.standard_normal((40, 64)).astype(np.float32)
    i1, b1, c1 = ix.search_batch(qs, 10, nthreads=1)
    i4, b4, c4 = ix.search_batch(qs, 10, nthreads=4)
    assert np.array_equal(i1, i4) and np.array_equal(b1, b4) and np.array_equal(c1, c4)
    for j in range(5):
        ids, bits = ix.search(qs[j], 10)
        assert i1[j, :c1[j]].tolist() == ids.tolist()


def test_sharded_oracle_equals_unsharded():
    """Row-sharding invariant: G shards that share the forest, use GLOBAL live leaf counts for
    the plan and merge per-visit top-n' reproduce the unsharded result.  Simulated here with the Python
    restatement: per-visit candidates from each shard are merged per visit, then per query."""
    rng = np.random.default_rng(4)
    dim, n, k = 24, 600, 7
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    ix = zo.OracleIndex(dim, zo.L2SQ, 12, 5, seed=2)
    ix.add(rows)
    dead = rng.choice(n, 50, replace=False); ix.remove(dead)
    tomb = np.zeros(n, dtype=bool); tomb[dead] = True
    f = ix.export_forest()
    py = _pyforest(ix, rows, tomb)
    for G in (2, 4, 8):
        for q in rng.standard_normal((5, dim)).astype(np.float32):
            exp_ids, _ = ix.search(q, k)
            _, _, trace = py.search(q, k)             # plan from global counts
            cand = set()
            for (_, leaf, nprime, _) in trace:
                per_rank = []
                for r in range(G):
                    mem = [i for i in py.members(leaf) if i % G == r]
                    sc = sorted((zo.distance_bits(zo.L2SQ, rows[i], q), i) for i in mem)
                    per_rank += sc[:nprime]           # each rank emits its local top-n' for the visit
                cand.update(i for _, i in sorted(per_rank)[:nprime])   # global per-visit top-n'
            final = sorted((zo.distance_bits(zo.L2SQ, rows[i], q), i) for i in cand)[:k]
            assert [i for _, i in final] == exp_ids.tolist()

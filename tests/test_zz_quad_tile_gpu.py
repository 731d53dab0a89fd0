"""GPU leg of the keys-only leaf-tile scan for cosine / L2 (quad_tile_kernel, knob quad_tile, default on): visits the fused
kernel does not take (n' > 128, or the fused kernel switched off) scored tile by tile instead of pair by pair.  Ids, distance
bits and counts must equal the oracle's, and the one-quad-per-pair path's (knob 0).  On the CPU, tests/test_quadtile.py runs
the kernel's own source thread by thread (tests/quadtile_emu.cpp) and every key equals the oracle's."""
import numpy as np
import pytest

from oracle import zb_oracle as zo

pytestmark = pytest.mark.gpu
F32 = np.float32


def zb():
    import zebra_b200

    return zebra_b200


def clustered(rng, n, dim, centres=8, noise=0.25):
    c = rng.standard_normal((centres, dim)).astype(F32)
    return (c[rng.integers(0, centres, n)] + noise * rng.standard_normal((n, dim))).astype(F32)


def assert_search_equal(ix, orc, queries, k):
    _, ords, bits, counts = ix.search_batch(queries, k, want_ids=False)
    eo, eb, ec = orc.search_batch(queries, k, nthreads=8)
    assert np.array_equal(counts, ec)
    for q in range(queries.shape[0]):
        c = int(ec[q])
        assert np.array_equal(ords[q, :c], eo[q, :c]) and np.array_equal(bits[q, :c], eb[q, :c]), q


@pytest.mark.parametrize("mid,mname,dim,mns,trees,k", [(zo.COSINE, "CosineDistance", 128, 512, 4, 100),
                                                      (zo.COSINE, "CosineDistance", 128, 512, 4, 700),
                                                      (zo.L2SQ, "L2SquaredDistance", 100, 256, 3, 64),      # dim % 16 != 0
                                                      (zo.L2, "L2Distance", 384, 300, 2, 40),
                                                      (zo.L2SQ, "L2SquaredDistance", 64, 5, 15, 50)])       # reference defaults: tiny leaves
def test_quad_tile_scan_equals_oracle_and_default_path(mid, mname, dim, mns, trees, k):
    z = zb()
    rng = np.random.default_rng(k + dim)
    n = 6000
    rows = clustered(rng, n, dim)
    rows[3000:3100] = rows[:100]                              # exact duplicates: equal keys, order by id
    orc = zo.OracleIndex(dim, mid, mns, trees, seed=2)
    orc.add(rows)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(mns, trees), getattr(z, mname)(), seed=2)
    ix.add(rows)
    queries = np.concatenate([rows[:48], rng.standard_normal((48, dim)).astype(F32)])
    ix.set_param("quad_tile", 1)
    ix.set_param("use_tile_scan", 0)                                      # the fused kernel takes n' <= 128 itself: keep it out
    assert_search_equal(ix, orc, queries, k)
    st = ix.stats()
    assert st["last_tiles"] > 0 and st["last_tile_pairs"] == 0           # nothing went through the fused kernel
    dead = rng.choice(n, n // 10, replace=False).astype(np.uint64)
    assert np.array_equal(ix.remove_ordinals(dead), orc.remove(dead))
    assert_search_equal(ix, orc, queries, k)                             # tombstones
    ix.set_param("quad_tile", 0)
    assert_search_equal(ix, orc, queries, k)                             # one quad per pair: same answer


def test_quad_tile_scan_next_to_the_fused_kernel():
    """top_k <= 32: the fused kernel takes the large leaves, the keys-only scan the visits it leaves (small leaves); and
    with the fused kernel switched off, everything."""
    z = zb()
    rng = np.random.default_rng(4)
    dim, n = 96, 5000
    rows = clustered(rng, n, dim, centres=16)
    orc = zo.OracleIndex(dim, zo.COSINE, 100, 5, seed=8)      # leaves of 50..99 rows: some below tile_min_rows = 64
    orc.add(rows)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(100, 5), z.CosineDistance(), seed=8)
    ix.add(rows)
    queries = np.concatenate([rows[:64], rng.standard_normal((64, dim)).astype(F32)])
    ix.set_param("quad_tile", 1)
    assert_search_equal(ix, orc, queries, 10)
    assert ix.stats()["last_tile_pairs"] > 0
    ix.set_param("use_tile_scan", 0)
    assert_search_equal(ix, orc, queries, 10)
    assert ix.stats()["last_tile_pairs"] == 0 and ix.stats()["last_tiles"] > 0


@pytest.mark.parametrize("mid,mname,k,quad", [(zo.MANHATTAN, "ManhattanDistance", 10, 0), (zo.COSINE, "CosineDistance", 100, 1),
                                              (zo.L2SQ, "L2SquaredDistance", 40, 1), (zo.HAMMING, "HammingDistance", 33, 0)])
def test_warp_select_equals_oracle_and_block_select(mid, mname, k, quad):
    """select_variant = 1: the per-visit top-n' with the list in a warp's registers (zb_select_kernel.cuh; its source is run on
    the CPU by tests/test_select.py) instead of the block-wide bitonic sort -- same entries, so the same final answer.
    Hamming gives long runs of equal keys (ties by ordinal)."""
    z = zb()
    rng = np.random.default_rng(k)
    dim, n = 64, 5000
    rows = clustered(rng, n, dim)
    rows[2000:2200] = rows[:200]
    orc = zo.OracleIndex(dim, mid, 400, 3, seed=6)
    orc.add(rows)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(400, 3), getattr(z, mname)(), seed=6)
    ix.add(rows)
    queries = np.concatenate([rows[:40], rng.standard_normal((40, dim)).astype(F32)])
    dead = rng.choice(n, n // 8, replace=False).astype(np.uint64)
    assert np.array_equal(ix.remove_ordinals(dead), orc.remove(dead))
    ix.set_param("use_tile_scan", 0)                        # every visit through the gather path and its select
    ix.set_param("quad_tile", quad)
    ix.set_param("select_variant", 1)
    assert_search_equal(ix, orc, queries, k)
    ix.set_param("select_variant", 0)
    assert_search_equal(ix, orc, queries, k)

"""Single-GPU parity of the fused leaf-tile scan on the HEADLINE shapes (BASELINE configs 2, 3 and 5): dim 768 and 384, leaves
below 2048 rows, cosine / L2 squared / L2, n' in {1, 10, 32}, before and after tombstoning 10 % of the rows -- ids, distance
bits and counts bit-exact against the oracle (lsh.rs:299-331 leaf branch, :557-564 rescoring + sort; distance.rs:19-49,
:103-114).  Every scored pair must have gone through tile_scan_kernel; the gather path is then forced on the same index
(knob use_tile_scan = 0) and must give the same answer with NO pair left to the tile kernel."""
import numpy as np
import pytest

from oracle import zb_oracle as zo

pytestmark = pytest.mark.gpu
F32 = np.float32

METRICS = [(zo.COSINE, "CosineDistance"), (zo.L2SQ, "L2SquaredDistance"), (zo.L2, "L2Distance")]


def zb():
    import zebra_b200

    return zebra_b200


def clustered(rng, n, dim, centres=64, noise=0.25):
    c = rng.standard_normal((centres, dim)).astype(F32)
    return (c[rng.integers(0, centres, n)] + noise * rng.standard_normal((n, dim))).astype(F32)


def make_queries(rng, rows, nq):
    dim = rows.shape[1]
    fresh = clustered(rng, nq // 2, dim)
    pick = rows[rng.integers(0, rows.shape[0], nq - nq // 2)].copy()
    pick[::2] += (1e-3 * rng.standard_normal(pick[::2].shape)).astype(F32)   # every other one stays an exact stored row
    return np.concatenate([fresh, pick]).astype(F32)


def assert_search_equal(ix, orc, queries, k):
    _, ords, bits, counts = ix.search_batch(queries, k, want_ids=False)
    eo, eb, ec = orc.search_batch(queries, k, nthreads=8)
    assert np.array_equal(counts, ec)
    for q in range(queries.shape[0]):
        c = int(ec[q])
        assert np.array_equal(ords[q, :c], eo[q, :c]), f"query {q}: ids differ"
        assert np.array_equal(bits[q, :c], eb[q, :c]), f"query {q}: distance bits differ"


@pytest.mark.parametrize("dim", [768, 384])
@pytest.mark.parametrize("mid,mname", METRICS)
def test_tile_kernel_on_headline_shapes(dim, mid, mname):
    z = zb()
    rng = np.random.default_rng(dim + mid)
    n, mns, trees = 100_000, 2048, 4
    rows = clustered(rng, n, dim)
    rows[50_000:50_200] = rows[:200]                          # exact duplicates: equal keys, order by id (D3)
    orc = zo.OracleIndex(dim, mid, mns, trees, seed=5)
    orc.add(rows)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(mns, trees), getattr(z, mname)(), seed=5)
    ix.add(rows)
    queries = make_queries(rng, rows, 700)                    # ~ 7 queries per leaf: full and partial query tiles
    for phase in ("built", "tombstoned"):
        if phase == "tombstoned":
            dead = rng.choice(n, n // 10, replace=False).astype(np.uint64)
            assert np.array_equal(ix.remove_ordinals(dead), orc.remove(dead))
        for k in (1, 10, 32):
            assert_search_equal(ix, orc, queries, k)
            st = ix.stats()
            # (a few leaves end up below tile_min_rows = 64 rows and stay on the gather path)
            assert st["last_tile_pairs"] > 0.99 * st["last_pairs"], (phase, k, st)
    # the gather path alone on the same (tombstoned) index: same answer, nothing through the tile kernel
    ix.set_param("use_tile_scan", 0)
    for k in (1, 10, 32):
        assert_search_equal(ix, orc, queries[:200], k)
        st = ix.stats()
        assert st["last_tile_pairs"] == 0 and st["last_pairs"] > 0, (k, st)
    ix.set_param("use_tile_scan", 1)
    assert_search_equal(ix, orc, queries[:200], 10)
    assert ix.stats()["last_tile_pairs"] > 0
    # the second-generation kernel (8-query tiles, 128-row stages) stays selectable and must agree
    ix.set_param("scan_gen", 2)
    for k in (1, 10, 32):
        assert_search_equal(ix, orc, queries, k)
        assert ix.stats()["last_tile_pairs"] > 0


@pytest.mark.parametrize("dim,mid,mname", [(384, zo.L2SQ, "L2SquaredDistance"), (768, zo.COSINE, "CosineDistance"),
                                           (100, zo.L2, "L2Distance")])
def test_fused_kernel_with_lists_of_up_to_128_entries(dim, mid, mname):
    """BASELINE config 5 asks for top-100: n' in 33..128 stays in the fused kernel (four list entries per lane, KR = 4);
    above 128 the visits go to the keys-only tile scan + warp select.  10 % tombstones, exact duplicates."""
    z = zb()
    rng = np.random.default_rng(dim * 3 + mid)
    n, mns, trees = 60_000, 2048, 3
    rows = clustered(rng, n, dim)
    rows[30_000:30_300] = rows[:300]
    orc = zo.OracleIndex(dim, mid, mns, trees, seed=11)
    orc.add(rows)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(mns, trees), getattr(z, mname)(), seed=11)
    ix.add(rows)
    dead = rng.choice(n, n // 10, replace=False).astype(np.uint64)
    assert np.array_equal(ix.remove_ordinals(dead), orc.remove(dead))
    queries = make_queries(rng, rows, 500)
    for k in (33, 100, 128):
        assert_search_equal(ix, orc, queries, k)
        st = ix.stats()
        assert st["last_tile_pairs"] > 0.99 * st["last_pairs"], (k, st)
    assert_search_equal(ix, orc, queries[:100], 129)
    assert ix.stats()["last_tile_pairs"] == 0
    # long lists run with eight math warps per team (8 rows of a block each) by default; the four-warp shape must agree
    ix.set_param("long_list_warps", 4)
    for k in (33, 100):
        assert_search_equal(ix, orc, queries, k)
        assert ix.stats()["last_tile_pairs"] > 0.99 * ix.stats()["last_pairs"]
    ix.set_param("long_list_warps", 8)


def test_tile_kernel_crowded_leaves():
    """Far more queries than a tile holds on every leaf (hundreds per leaf): many sibling tiles of one leaf, the shared
    per-query bound (k distinct candidates already found elsewhere) and the dedup across trees all at work."""
    z = zb()
    rng = np.random.default_rng(77)
    dim, n, mns, trees = 384, 20_000, 1024, 6
    rows = clustered(rng, n, dim, centres=4)
    orc = zo.OracleIndex(dim, zo.L2SQ, mns, trees, seed=9)
    orc.add(rows)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(mns, trees), z.L2SquaredDistance(), seed=9)
    ix.add(rows)
    queries = make_queries(rng, rows, 6000)
    for k in (10, 32):
        assert_search_equal(ix, orc, queries, k)
        assert ix.stats()["last_tile_pairs"] > 0.99 * ix.stats()["last_pairs"]


def test_database_open_then_insert_search_remove(tmp_path):
    """core.rs:92-104 open, then :245-254 insert_records on the reopened database: rows reloaded from a store carry their
    stored ids, new rows get minted ids registered next to them."""
    z = zb()
    rng = np.random.default_rng(3)
    dim = 32
    a, b = clustered(rng, 300, dim), clustered(rng, 50, dim)
    db = z.Database(dim, z.L2Distance(), index_options=z.LSHIndexOptions(8, 3))
    ids_a = db.insert_records(a, [bytes([i % 251]) for i in range(300)])
    path = str(tmp_path / "db.zebra")
    db.save_database(path)
    db2 = z.Database.open(path, dim, z.L2Distance())
    ids_b = db2.insert_records(b, [b"new%d" % i for i in range(50)])
    assert len(set(ids_b)) == 50 and not (set(ids_b) & set(ids_a))
    hit = db2.index.search(b[7], 1)
    assert hit[0][0] == ids_b[7] and hit[0][1] == 0
    hit = db2.index.search(a[11], 1)
    assert hit[0][0] == ids_a[11]
    db2.remove([ids_b[7], ids_a[11]])
    assert db2.index.search(b[7], 1)[0][0] != ids_b[7] and db2.index.search(a[11], 1)[0][0] != ids_a[11]
    with pytest.raises(z.ZebraError):                       # a duplicate id inside one batch is rejected as a whole ...
        db2.index.add(b[:2], ids=[ids_b[0], ids_b[0]])
    st = db2.index.stats()
    more = db2.index.add(b[:2])                             # ... and leaves the id bookkeeping untouched
    assert len(more) == 2 and db2.index.stats()["total_rows"] == st["total_rows"] + 2
    assert db2.index.search(b[0], 1)[0][1] == 0


def test_load_forest_rejects_bad_forests_without_touching_the_index():
    z = zb()
    rng = np.random.default_rng(8)
    dim = 16
    rows = clustered(rng, 400, dim)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(16, 2), z.L2SquaredDistance(), seed=1)
    ix.add(rows)
    f = ix.export_forest()
    before = ix.search(rows[3], 3)
    bad = z.index.Forest(f.nodes.copy(), f.roots, f.coef, f.cst, f.leaf_off, f.members.copy())
    bad.members[0] = bad.members[1]                         # a row twice in one leaf, another one missing from the tree
    with pytest.raises(z.ZebraError):
        ix.load_forest(rows, bad)
    bad = z.index.Forest(f.nodes.copy(), f.roots, f.coef, f.cst, f.leaf_off, f.members)
    leafs = np.where(bad.nodes[:, 0] < 0)[0]
    bad.nodes[leafs[1], 3] = bad.nodes[leafs[0], 3]         # a leaf referenced twice, another unreachable
    with pytest.raises(z.ZebraError):
        ix.load_forest(rows, bad)
    assert ix.search(rows[3], 3) == before and ix.stats()["total_rows"] == 400
    ix.load_forest(rows, f)                                 # the good one still loads
    assert [b for _, b in ix.search(rows[3], 3)] == [b for _, b in before]

"""GPU leg of the interchange row: a device index out to the reference's stored values and back, and a
store written by the ORACLE side (oracle/zb_bincode.py: bincode(legacy) Node<N> blobs + raw embedding values,
/root/reference/src/database/index/lsh.rs:91-105) into a device index, searched against the oracle."""
import uuid

import numpy as np
import pytest

from oracle import zb_bincode as zbc
from oracle import zb_oracle as zo

pytestmark = pytest.mark.gpu
F32 = np.float32


def zb():
    import zebra_b200

    return zebra_b200


def clustered(rng, n, dim, centres=32, noise=0.25):
    c = rng.standard_normal((centres, dim)).astype(F32)
    return (c[rng.integers(0, centres, n)] + noise * rng.standard_normal((n, dim))).astype(F32)


def search_ids(ix, queries, k):
    ids, _, bits, counts = ix.search_batch(queries, k)
    return [[(ids[q, j].tobytes(), int(bits[q, j])) for j in range(int(counts[q]))] for q in range(queries.shape[0])]


def test_export_matches_oracle_codec_and_round_trips(tmp_path):
    z = zb()
    rng = np.random.default_rng(4)
    n, dim = 3000, 40
    rows = clustered(rng, n, dim)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(6, 5), z.L2SquaredDistance(), seed=9)
    ix.add(rows)
    dead = rng.choice(n, 300, replace=False).astype(np.uint64)
    assert ix.remove_ordinals(dead).all()
    out_rows, ids, live = ix.export_rows()
    assert np.array_equal(out_rows.view(np.uint32), rows.view(np.uint32))          # the embedding VALUE is the raw row
    assert live.sum() == n - 300 and not live[dead].any()
    id_list = [uuid.UUID(bytes=r.tobytes()) for r in ids]
    assert id_list == sorted(id_list) and all(u.version == 7 for u in id_list[:5])
    # every tree blob == the oracle-side encoder applied to the exported forest (removed rows left out of the leaves)
    blobs = ix.export_tree_blobs()
    want = zbc.forest_to_nodes(ix.export_forest(), id_list, live)
    assert len(blobs) == 5 and ix.export_tree_blob(3) == blobs[3]            # per-tree entry point agrees
    for t in range(5):
        assert blobs[t] == zbc.encode_node(want[t]), t
        assert zbc.nodes_equal(zbc.decode_node(blobs[t], dim), want[t])
    # dump -> fresh index: same answers (ids and distance bits; ordinals are renumbered because removed rows are gone)
    queries = np.concatenate([rows[:60], rng.standard_normal((60, dim)).astype(F32)])
    before = search_ids(ix, queries, 10)
    p = str(tmp_path / "index.store")
    ix.save_store(p)
    ix2 = z.LSHIndex(dim, z.LSHIndexOptions(6, 5), z.L2SquaredDistance(), seed=9)
    rep = ix2.load_store(p)
    assert rep["rows_loaded"] == n - 300 and rep["missing_ids"] == 0 and rep["orphan_rows"] == 0 and rep["orphans"] == []
    assert search_ids(ix2, queries, 10) == before
    st = ix2.stats()
    assert st["rows"] == n - 300 and st["live_rows"] == n - 300
    # the imported index keeps working: hash, remove by id, incremental add with caller ids
    k1, d1, _ = ix.hash(queries[:50])
    k2, d2, _ = ix2.hash(queries[:50])
    assert np.array_equal(k1, k2) and np.array_equal(d1, d2)
    victim = uuid.UUID(bytes=before[0][0][0])
    assert ix2.remove([victim]) == {victim} and ix.remove([victim]) == {victim}
    assert search_ids(ix2, queries, 10) == search_ids(ix, queries, 10)
    assert ix2.export_tree_blobs() == ix.export_tree_blobs()


@pytest.mark.parametrize("mid,mname", [(zo.L2SQ, "L2SquaredDistance"), (zo.COSINE, "CosineDistance"), (zo.MANHATTAN, "ManhattanDistance")])
def test_import_of_an_oracle_written_store(mid, mname):
    """Ids are random (not v7, not in row order), rows arrive shuffled, two leaves name ids whose embedding is gone
    (vectors the reference removed, quirk Q5) and one embedding is missing from one tree (a lost update, quirk Q11).
    The host-side flattening of exactly this store is checked on the CPU in test_interchange.py."""
    from test_interchange import tampered_store

    z = zb()
    rng = np.random.default_rng(17 + mid)
    n, dim, T, k, X = 2000, 48, 5, 10, 1234
    rows, ids, forest, blobs, order, clean = tampered_store(rng, n, dim, T, X, metric=mid)
    perm = rng.permutation(n)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(8, T), getattr(z, mname)(), seed=5)
    rep = ix.import_store([ids[i] for i in perm], rows[perm], blobs)
    assert rep["rows_loaded"] == n - 1 and rep["missing_ids"] == 2 and rep["orphan_rows"] == 1 and rep["orphans"] == [ids[X]]
    assert rep["planes"] == forest.cst.size and rep["leaves"] == forest.leaf_off.size - 1
    # expected: the same forest over the rows kept, renumbered in id order (ordinal = rank of the id)
    exp = zo.OracleIndex(dim, mid, 8, T, seed=5)
    exp.load_forest(rows[order], clean)
    queries = np.concatenate([rows[:100], rng.standard_normal((100, dim)).astype(F32)])
    got_ids, ords, bits, counts = ix.search_batch(queries, k)
    eo, eb, ec = exp.search_batch(queries, k, nthreads=8)
    assert np.array_equal(counts, ec)
    for q in range(queries.shape[0]):
        c = int(ec[q])
        assert np.array_equal(ords[q, :c], eo[q, :c]) and np.array_equal(bits[q, :c], eb[q, :c]), q
        assert [uuid.UUID(bytes=got_ids[q, j].tobytes()) for j in range(c)] == [ids[order[int(o)]] for o in eo[q, :c]]
    # bucket keys follow the imported planes
    ka, da, _ = ix.hash(queries[:64])
    kb, db, _ = exp.hash(queries[:64])
    assert np.array_equal(ka, kb) and np.array_equal(da, db)
    # the orphan goes back in through the ordinary insert path, under its own id
    assert ix.add(rows[X:X + 1], ids=[ids[X]]) == [ids[X]]
    assert ix.stats()["live_rows"] == n
    assert ix.remove([ids[X], uuid.UUID(int=1)]) == {ids[X]}          # it is there; the ghost id never was


def test_import_rejects_inconsistent_stores():
    z = zb()
    rng = np.random.default_rng(2)
    dim = 8
    rows = rng.standard_normal((4, dim)).astype(F32)
    ids = [uuid.UUID(int=10 + i) for i in range(4)]
    leaf = lambda members: zbc.encode_node(("leaf", members))
    ix = z.LSHIndex(dim, z.LSHIndexOptions(5, 2), z.L2Distance())
    with pytest.raises(z.ZebraError, match="trees"):
        ix.import_store(ids, rows, [leaf(ids)])                               # 1 blob for an index of 2 trees
    with pytest.raises(z.ZebraError, match="twice"):
        ix.import_store(ids, rows, [leaf(ids), leaf(ids + [ids[0]])])         # an id twice in one tree
    with pytest.raises(z.ZebraError, match="twice"):
        ix.import_store(ids + [ids[1]], np.concatenate([rows, rows[:1]]), [leaf(ids), leaf(ids)])  # duplicate key
    with pytest.raises(z.ZebraError):
        ix.import_store(ids, rows, [leaf(ids), leaf(ids)[:-3]])               # truncated value
    rep = ix.import_store(ids, rows, [leaf(ids), leaf(ids[::-1])])            # root leaves: fine
    assert rep["rows_loaded"] == 4 and rep["nodes"] == 2 and rep["max_depth"] == 0
    assert [u for u, _ in ix.search(rows[2], 1)] == [ids[2]]
    rep = ix.import_store([], np.zeros((0, dim), F32), [leaf([]), leaf([ids[0]])])   # empty store, one stale id
    assert rep["rows_loaded"] == 0 and rep["missing_ids"] == 1
    assert ix.search(rows[0], 3) == []


def test_database_save_and_open(tmp_path):
    z = zb()
    rng = np.random.default_rng(6)
    rows = clustered(rng, 1500, 32)
    for metric in (z.CosineDistance(), z.MinkowskiDistance(3)):
        db = z.Database(32, metric, index_options=z.LSHIndexOptions(5, 4), seed=2)
        ids = db.insert_records(rows, [b"x"] * 1500)
        db.remove(ids[:100])
        path = str(tmp_path / f"{type(metric).__name__}.zebra")
        db.save_database(path)
        raw = open(path, "rb").read()
        assert raw == zbc.encode_database_inner(db.uuid, 5, 4, metric_power=3 if metric.power else None)
        db2 = z.Database.open(path, 32, type(metric)(*((0,) if metric.power else ())), seed=2)
        assert db2.uuid == db.uuid and db2.index_options == z.LSHIndexOptions(5, 4) and db2.metric.power == metric.power
        q = rows[200:260]
        assert search_ids(db2.index, q, 7) == search_ids(db.index, q, 7)
        assert db2.index.stats()["live_rows"] == 1400

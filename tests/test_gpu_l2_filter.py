"""L2 / L2 squared through the dot-product filter (METRIC 3 of tile_scan3_kernel + refine_visits_kernel) on a
GPU: ids, distance bits and counts bit-exact against the oracle (lsh.rs:299-331 leaf branch, :557-564 rescoring + sort;
distance.rs:38-49, :103-114) with the filter forced on, forced off and adaptive -- on the headline shapes, with tombstones,
with keys so crowded that the 32-entry candidate lists overflow (the visits are rescanned exactly), and with rows whose
squared norms leave the range the error bound covers."""
import numpy as np
import pytest

from oracle import zb_oracle as zo

pytestmark = pytest.mark.gpu
F32 = np.float32


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
@pytest.mark.parametrize("mid,mname", [(zo.L2SQ, "L2SquaredDistance"), (zo.L2, "L2Distance")])
def test_filter_on_headline_shapes(dim, mid, mname):
    z = zb()
    rng = np.random.default_rng(dim * 5 + mid)
    n, mns, trees = 100_000, 2048, 4
    rows = clustered(rng, n, dim)
    rows[50_000:50_200] = rows[:200]                          # exact duplicates: equal keys, order by id (D3)
    orc = zo.OracleIndex(dim, mid, mns, trees, seed=5)
    orc.add(rows)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(mns, trees), getattr(z, mname)(), seed=5)
    ix.add(rows)
    queries = make_queries(rng, rows, 700)
    for phase in ("built", "tombstoned"):
        if phase == "tombstoned":
            dead = rng.choice(n, n // 10, replace=False).astype(np.uint64)
            assert np.array_equal(ix.remove_ordinals(dead), orc.remove(dead))
        for mode in (2, 0, 1):
            ix.set_param("l2_filter", mode)
            for k in (1, 10, 16, 17):
                assert_search_equal(ix, orc, queries, k)
                st = ix.stats()
                assert st["last_tile_pairs"] > 0.99 * st["last_pairs"], (phase, mode, k, st)
                want = 1 if (mode and k <= 16) else 0
                assert st["last_filter_used"] == want, (phase, mode, k, st)
                if want:
                    # the second pass scores a few rows per visit, not the leaves (a handful of visits may be rescanned)
                    assert 0 < st["last_filter_rows"] < 40 * st["last_tile_visits"], (phase, k, st)
                    assert st["last_filter_flagged"] * 16 <= st["last_tile_visits"], (phase, k, st)
                else:
                    assert st["last_filter_rows"] == 0 and st["last_filter_flagged"] == 0


def test_filter_crowded_keys_are_rescanned_exactly_and_back_off():
    """Every row stored 40 times: any visit's n'-th best is tied with more rows than a candidate list holds, so the visits are
    flagged and their leaves scanned exactly (ties by id); the adaptive mode then leaves the filter out of the next batches."""
    z = zb()
    rng = np.random.default_rng(21)
    dim, mns, trees = 256, 1024, 3
    base = clustered(rng, 600, dim, centres=8)
    rows = np.repeat(base, 40, axis=0)[rng.permutation(600 * 40)]
    orc = zo.OracleIndex(dim, zo.L2SQ, mns, trees, seed=3)
    orc.add(rows)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(mns, trees), z.L2SquaredDistance(), seed=3)
    ix.add(rows)
    queries = make_queries(rng, rows, 300)
    ix.set_param("l2_filter", 2)
    for k in (1, 10, 16):
        assert_search_equal(ix, orc, queries, k)
        st = ix.stats()
        assert st["last_filter_used"] == 1 and st["last_filter_flagged"] > 0, (k, st)
    ix.set_param("l2_filter", 1)
    assert_search_equal(ix, orc, queries, 10)
    first = ix.stats()
    assert first["last_filter_used"] == 1 and first["last_filter_flagged"] * 16 > first["last_tile_visits"], first
    assert_search_equal(ix, orc, queries, 10)
    assert ix.stats()["last_filter_used"] == 0                 # backed off: the exact kernel took the batch


def test_filter_rows_outside_the_bound_fall_back():
    """Rows (and queries) whose squared norms exceed 1e37 -- finite, infinite -- next to ordinary ones, tiny rows, zero rows:
    such rows have no usable approximate distance, the visits that meet them are rescanned exactly."""
    z = zb()
    rng = np.random.default_rng(22)
    dim, n, mns, trees = 128, 30_000, 1024, 3
    rows = clustered(rng, n, dim, centres=16)
    big = rng.choice(n, 12, replace=False)
    rows[big[:6]] *= F32(4e17)                                 # |row|^2 ~ 2e37 .. 1e38: finite, above the limit
    rows[big[6:]] *= F32(3e18)                                 # overflows to +inf
    rows[100:110] = 0
    rows[110:120] *= F32(1e-20)
    with np.errstate(over="ignore", invalid="ignore"):
        orc = zo.OracleIndex(dim, zo.L2, mns, trees, seed=8)
        orc.add(rows)
    ix = z.LSHIndex(dim, z.LSHIndexOptions(mns, trees), z.L2Distance(), seed=8)
    ix.load_forest(rows, orc.export_forest())                  # this test is about the scan: the forest comes from the oracle
    queries = make_queries(rng, rows, 400)
    queries[0] = rows[big[0]]
    queries[1] = rows[big[7]]
    queries[2] = 0
    queries[3] = rows[115]
    ix.set_param("l2_filter", 2)
    with np.errstate(over="ignore", invalid="ignore"):
        for k in (1, 10):
            assert_search_equal(ix, orc, queries, k)
            st = ix.stats()
            assert st["last_filter_used"] == 1 and st["last_filter_flagged"] > 0, (k, st)

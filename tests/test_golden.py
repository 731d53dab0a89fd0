"""Frozen vectors (tests/golden/*.npz, generator: tests/golden/make_golden.py).  CPU: the oracle must still reproduce
them.  GPU: the CUDA path, driven through the same add / remove / add sequence over the C ABI, must reproduce them
bit for bit (forest, bucket keys, ids, distance bits, counts)."""
import glob
import os

import numpy as np
import pytest

from oracle import zb_oracle as zo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
NAMES = {zo.COSINE: ("CosineDistance", ()), zo.L2SQ: ("L2SquaredDistance", ()), zo.L2: ("L2Distance", ()),
         zo.MANHATTAN: ("ManhattanDistance", ()), zo.CANBERRA: ("CanberraDistance", ()), zo.HAMMING: ("HammingDistance", ()),
         zo.MINKOWSKI(3): ("MinkowskiDistance", (3,))}
FOREST = ("nodes", "roots", "coef", "cst", "leaf_off", "members")


def test_fixtures_exist():
    assert len(GOLDEN) == 7


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    ix = zo.OracleIndex(g["rows"].shape[1], int(g["metric"]), int(g["mns"]), int(g["trees"]), seed=5)
    ix.add(g["rows"][: int(g["first"])])
    ix.remove(g["dead"])
    ix.add(g["rows"][int(g["first"]):])
    f = ix.export_forest()
    for name in FOREST:
        assert np.array_equal(getattr(f, name), g[name]), name
    ids, bits, counts = ix.search_batch(g["queries"], int(g["k"]))
    assert np.array_equal(ids, g["ids"]) and np.array_equal(bits, g["bits"]) and np.array_equal(counts, g["counts"])
    keys, depth, leaf = ix.hash(g["queries"])
    assert np.array_equal(keys, g["keys"]) and np.array_equal(depth, g["depth"]) and np.array_equal(leaf, g["leaf"])
    nq = g["queries"].shape[0]
    assert np.array_equal(zo.distance_bits_batch(int(g["metric"]), g["rows"][:nq], g["queries"]), g["pair_bits"])


# fixtures of the north-star metrics here; those of the scalar metrics run from tests/test_gpu_scalar_metrics.py
NORTH_STAR = [p for p in GOLDEN if int(np.load(p)["metric"]) <= zo.L2]


@pytest.mark.gpu
@pytest.mark.parametrize("path", NORTH_STAR, ids=[os.path.basename(p) for p in NORTH_STAR])
def test_cuda_reproduces_golden(path):
    cuda_reproduces_golden(path)


def cuda_reproduces_golden(path):
    import zebra_b200 as z

    g = np.load(path)
    cls, args = NAMES[int(g["metric"])]
    metric = getattr(z, cls)(*args)
    ix = z.LSHIndex(g["rows"].shape[1], z.LSHIndexOptions(int(g["mns"]), int(g["trees"])), metric, seed=5)
    ix.add(g["rows"][: int(g["first"])])
    assert ix.remove_ordinals(g["dead"]).all()
    ix.add(g["rows"][int(g["first"]):])
    f = ix.export_forest()
    for name in FOREST:
        assert np.array_equal(getattr(f, name), g[name]), name
    k = int(g["k"])
    _, ords, bits, counts = ix.search_batch(g["queries"], k, want_ids=False)
    assert np.array_equal(counts, g["counts"])
    for q in range(len(counts)):
        c = int(counts[q])
        assert np.array_equal(ords[q, :c], g["ids"][q, :c]) and np.array_equal(bits[q, :c], g["bits"][q, :c])
    keys, depth, leaf = ix.hash(g["queries"])
    assert np.array_equal(keys, g["keys"]) and np.array_equal(depth, g["depth"]) and np.array_equal(leaf, g["leaf"])
    nq = g["queries"].shape[0]
    assert np.array_equal(metric.distance_batch(g["rows"][:nq], g["queries"]), g["pair_bits"])

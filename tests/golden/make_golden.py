"""Generates tests/golden/*.npz: frozen input/output vectors of the query path.

The reference (a Rust crate) cannot be built or run in this environment and ships no vectors of its own, so these
were produced by the CPU oracle (oracle/zb_oracle.c) at the commit that introduced them.  They pin the oracle, and
through it the CUDA path, against drift: any later change of accumulation order, metric epilogue, walk, tie-break
or build sampling shows up as a diff against these files.  They do NOT pin the oracle to the reference ("parity
unpinned").

  python tests/golden/make_golden.py        # rewrites the fixtures (do this only for an intended semantic change)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import zb_oracle as zo  # noqa: E402

CASES = [  # name, metric, dim, n, max_node_size, trees, top_k
    ("cosine_defaults", zo.COSINE, 20, 300, 5, 15, 10),
    ("l2sq_large_leaves", zo.L2SQ, 48, 400, 96, 3, 10),
    ("l2_mid", zo.L2, 33, 350, 16, 4, 25),
    # the scalar metrics of distance.rs:51-190 (f32 bits zero-extended; Minkowski carries its power in bits 8..)
    ("manhattan_defaults", zo.MANHATTAN, 21, 300, 5, 15, 10),
    ("canberra_mid", zo.CANBERRA, 16, 320, 24, 3, 12),
    ("minkowski3_large_leaves", zo.MINKOWSKI(3), 40, 400, 96, 3, 10),
    ("hamming_mid", zo.HAMMING, 24, 300, 16, 4, 8),
]


def build_case(name, metric, dim, n, mns, trees, k):
    rng = np.random.default_rng(sum(map(ord, name)))
    centres = rng.standard_normal((6, dim)).astype(np.float32)
    rows = (centres[rng.integers(0, 6, n)] + 0.25 * rng.standard_normal((n, dim))).astype(np.float32)
    rows[50:58] = rows[0:8]                                    # exact duplicates: ties by id
    queries = np.concatenate([rows[:6], rows[50:52], rng.standard_normal((12, dim)).astype(np.float32),
                              np.zeros((1, dim), np.float32)]).astype(np.float32)
    ix = zo.OracleIndex(dim, metric, mns, trees, seed=5)
    ix.add(rows[: n - 40])
    dead = np.arange(3, n - 40, 11, dtype=np.uint64)
    ix.remove(dead)
    ix.add(rows[n - 40:])                                      # incremental insert after deletes
    f = ix.export_forest()
    ids, bits, counts = ix.search_batch(queries, k)
    keys, depth, leaf = ix.hash(queries)
    pair_bits = zo.distance_bits_batch(metric, rows[: len(queries)], queries)
    return dict(metric=np.int32(metric), mns=np.int32(mns), trees=np.int32(trees), k=np.int32(k), rows=rows, queries=queries,
                first=np.int32(n - 40), dead=dead, nodes=f.nodes, roots=f.roots, coef=f.coef, cst=f.cst, leaf_off=f.leaf_off,
                members=f.members, ids=ids, bits=bits, counts=counts, keys=keys, depth=depth, leaf=leaf, pair_bits=pair_bits)


if __name__ == "__main__":
    only = set(sys.argv[1:])                                   # names to (re)write; none = all
    for c in CASES:
        if only and c[0] not in only:
            continue
        np.savez_compressed(os.path.join(HERE, c[0] + ".npz"), **build_case(*c))
        print("wrote", c[0])

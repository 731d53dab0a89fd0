"""World-size-2 (or more) CPU simulation of the bucket-sharded query path over torch.distributed / gloo.

Mirrors, step for step, what libzebra_b200 does across GPUs (zb_index.cu search_device, sharded branch), with the
oracle's arithmetic standing in for the kernels:
  1. the forest is replicated; rank r walks query slice r only and the visit records are all-gathered;
  2. leaf l is owned by rank l % G: a rank scores only the visits of its leaves (whole leaves, so Q2's per-visit
     top-n' is global) and reduces them to a per-query local top-k with dedup;
  3. all-to-all by query slice, per-slice final merge (union, dedup, sort by (bits, id), take k), all-gather of results.
Rank 0 checks the merged result against the unsharded oracle.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/gloo_shard_sim.py
"""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyref  # noqa: E402
from oracle import zb_oracle as zo  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, G = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(9)                     # same data on every rank
    dim, n, k, nq = 24, 900, 7, 23                     # nq not a multiple of G: the last slice is short
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    rows[100:110] = rows[0:10]                         # exact duplicates: ties by id, dedup across trees
    queries = np.concatenate([rows[:5], rng.standard_normal((nq - 5, dim)).astype(np.float32)])
    ok = True
    for metric, mns, trees in ((zo.L2SQ, 12, 5), (zo.COSINE, 5, 15)):
        orc = zo.OracleIndex(dim, metric, mns, trees, seed=2)
        orc.add(rows)
        dead = rng.choice(n, 60, replace=False)
        orc.remove(dead)
        tomb = np.zeros(n, dtype=bool)
        tomb[dead] = True
        py = pyref.PyForest(orc.export_forest(), rows, tomb, lambda a, b: zo.distance_bits(metric, a, b), zo.point_is_above)
        # 1. sharded plan + all-gather of the visit records
        nqp = (nq + G - 1) // G
        mine = []
        for q in range(rank * nqp, min(nq, (rank + 1) * nqp)):
            _, _, trace = py.search(queries[q], k)
            mine.append([(leaf, nprime) for (_, leaf, nprime, _) in trace])
        plans = [None] * G
        dist.all_gather_object(plans, mine)
        visits = [v for part in plans for v in part]   # visits[q] = [(leaf, n'), ...]
        assert len(visits) == nq
        # 2. score the visits of the leaves this rank owns, per-query local top-k with dedup
        local = []
        for q in range(nq):
            cand = set()
            for leaf, nprime in visits[q]:
                if leaf % G != rank:
                    continue
                mem = py.members(leaf)
                sc = sorted((py.dist_bits(rows[i], queries[q]), i) for i in mem)
                cand.update(sc if len(mem) < nprime else sc[:nprime])
            local.append(sorted(cand)[:k])
        # 3. all-to-all by query slice, final merge of my slice, all-gather of the finished slices
        send = [local[r * nqp:(r + 1) * nqp] for r in range(G)]
        allsend = [None] * G                           # gloo has no all-to-all: every rank picks its column
        dist.all_gather_object(allsend, send)
        recv = [allsend[r][rank] for r in range(G)]
        done = []
        for j in range(len(recv[0])):
            union = set()
            for r in range(G):
                union.update(recv[r][j])
            done.append(sorted(union)[:k])
        slices = [None] * G
        dist.all_gather_object(slices, done)
        final = [x for part in slices for x in part]
        if rank == 0:
            for q in range(nq):
                ids, bits = orc.search(queries[q], k)
                good = [i for _, i in final[q]] == ids.tolist() and [b for b, _ in final[q]] == bits.tolist()
                ok = ok and good
            print(f"[gloo G={G}] metric={metric} leaf<{mns} trees={trees}: {'ok' if ok else 'MISMATCH'}", flush=True)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()

"""Sharded parity: run with torchrun on G >= 2 GPUs of one node.  Every rank builds its shard of the index (row
ordinal % G == rank), the library exchanges leaf counts / per-visit top-n' lists over NCCL, and rank 0 compares the
merged results with the (unsharded) CPU oracle: ids, distance bits and counts bit-exact, plus bucket keys.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zebra_b200 as z  # noqa: E402
from oracle import zb_oracle as zo  # noqa: E402


def clustered(rng, n, dim, centres=32, noise=0.25):
    c = rng.standard_normal((centres, dim)).astype(np.float32)
    return (c[rng.integers(0, centres, n)] + noise * rng.standard_normal((n, dim))).astype(np.float32)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [(zo.COSINE, "CosineDistance", 6000, 384, 5, 15, 10),      # reference defaults: tiny leaves, long count cascades
             (zo.L2, "L2Distance", 20000, 768, 512, 4, 10),            # BASELINE config 2 shape: tile kernel
             (zo.L2SQ, "L2SquaredDistance", 9000, 100, 256, 3, 25),    # dim not a multiple of 16
             (zo.COSINE, "CosineDistance", 24000, 768, 1024, 4, 10),   # BASELINE config 3 shape: cosine, bucket-sharded, tile kernel
             (zo.COSINE, "CosineDistance", 16000, 384, 256, 4, 100),   # BASELINE config 5 shape: 384-dim, deletes, top-100 (generic path)
             (zo.L2SQ, "L2SquaredDistance", 24000, 384, 1024, 4, 16)]  # L2 squared through the dot-product filter at its largest k
    for mid, mname, n, dim, mns, trees, k in cases:
        rng = np.random.default_rng(1234 + n)     # same data on every rank
        rows = clustered(rng, n, dim)
        extra = clustered(rng, n // 10, dim)
        queries = np.concatenate([rows[:64], clustered(rng, 192, dim)])
        ix = z.LSHIndex(dim, z.LSHIndexOptions(mns, trees), getattr(z, mname)(), device=local, seed=7, shard_rank=rank,
                        shard_count=world)
        uid = [z.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ix.comm_init(uid[0])
        if mns == 5:
            ix.set_param("visit_slots", 4)   # force the (collective) grow-and-replan path
        if dim == 768:
            ix.set_param("bm_stage_mb", 1)   # the bucket-major store is exchanged in several leaf groups per tree
        # sliced search, three ways to get every rank's query slice to every rank: pushed into the peers' buffers over NVLink
        # (CUDA IPC, the default), inside the visit-record allgather, in an allgather of their own
        if dim == 100:
            ix.set_param("p2p_queries", 0)
        if mns == 5:
            ix.set_param("p2p_queries", 0)
            ix.set_param("single_exchange", 0)
        ix.add(rows)
        orc = zo.OracleIndex(dim, mid, mns, trees, seed=7) if rank == 0 else None
        if rank == 0:
            orc.add(rows)

        def check(tag):
            nonlocal ok
            _, ords, bits, counts = ix.search_batch(queries, k, want_ids=False)
            keys, depth, leaf = ix.hash(queries[:32])
            # the scalable form: every rank passes only the slice of the batch it fronts and gets that slice's results
            nqs = queries.shape[0] - 3                           # not a multiple of the world size: the last slice is short
            lo, hi = ix.slice_bounds(nqs, rank, world)
            _, s_ords, s_bits, s_counts = ix.search_slice(nqs, queries[lo:hi], k, want_ids=False)
            parts = [None] * world
            dist.all_gather_object(parts, (s_ords, s_bits, s_counts))
            if rank == 0:
                so = np.concatenate([p[0] for p in parts]); sb = np.concatenate([p[1] for p in parts]); sc = np.concatenate([p[2] for p in parts])
                slice_ok = (so.shape[0] == nqs and np.array_equal(sc, counts[:nqs]) and np.array_equal(so, ords[:nqs])
                            and np.array_equal(sb, bits[:nqs]))
            if rank == 0:
                eo, eb, ec = orc.search_batch(queries, k, nthreads=8)
                ek, ed, el = orc.hash(queries[:32])
                good = (np.array_equal(counts, ec) and all(np.array_equal(ords[q, :ec[q]], eo[q, :ec[q]]) and
                                                           np.array_equal(bits[q, :ec[q]], eb[q, :ec[q]]) for q in range(len(ec)))
                        and np.array_equal(keys, ek) and np.array_equal(depth, ed) and np.array_equal(leaf, el) and slice_ok)
                st = ix.stats()
                print(f"[mgpu G={world}] {mname} n={n} dim={dim} leaf<{mns} trees={trees} k={k} {tag} "
                      f"(l2 filter {'on' if st['last_filter_used'] else 'off'}): {'ok' if good else 'MISMATCH'}", flush=True)
                ok = ok and good

        check("bulk build")
        dele = np.arange(0, n, 7, dtype=np.uint64)
        ix.remove_ordinals(dele)
        if rank == 0:
            orc.remove(dele)
        check("after remove")
        ix.add(extra)
        if rank == 0:
            orc.add(extra)
        check("after incremental add")
        if mns == 5:    # sharded deduplicate: copies land on other ranks than their originals
            dup_rows = rows[:300].copy()
            ix.add(dup_rows)
            _, got = ix.deduplicate_raw()
            if rank == 0:
                orc.add(dup_rows)
                exp = orc.deduplicate()
                good = np.array_equal(got, exp) and exp.size > 0
                print(f"[mgpu G={world}] deduplicate removed {got.size}: {'ok' if good else 'MISMATCH'}", flush=True)
                ok = ok and good
            check("after deduplicate")
        ix.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if not int(flag.item()):
        sys.exit(1)


if __name__ == "__main__":
    main()

"""Sharded parity of the rows SURVEY 8(f) 3 and 4 on G >= 2 GPUs (torchrun): a scalar metric through the leaf-tile scan of
the bucket-sharded store, a scalar metric on the reference's default forest, and the import of an oracle-written store
(random ids, shuffled rows, ghosts, an orphan) into a sharded index -- each against the UNSHARDED CPU oracle.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_parity_ext.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import zebra_b200 as z  # noqa: E402
from oracle import zb_oracle as zo  # noqa: E402
from test_interchange import tampered_store  # noqa: E402


def clustered(rng, n, dim, centres=32, noise=0.25):
    c = rng.standard_normal((centres, dim)).astype(np.float32)
    return (c[rng.integers(0, centres, n)] + noise * rng.standard_normal((n, dim))).astype(np.float32)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True

    def make(dim, mns, trees, metric, seed):
        ix = z.LSHIndex(dim, z.LSHIndexOptions(mns, trees), metric, device=local, seed=seed, shard_rank=rank, shard_count=world)
        uid = [z.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ix.comm_init(uid[0])
        return ix

    def compare(tag, ix, orc, queries, k):
        nonlocal ok
        _, ords, bits, counts = ix.search_batch(queries, k, want_ids=False)
        if rank == 0:
            eo, eb, ec = orc.search_batch(queries, k, nthreads=8)
            good = np.array_equal(counts, ec) and all(np.array_equal(ords[q, :ec[q]], eo[q, :ec[q]]) and
                                                      np.array_equal(bits[q, :ec[q]], eb[q, :ec[q]]) for q in range(len(ec)))
            print(f"[mgpu-ext G={world}] {tag}: {'ok' if good else 'MISMATCH'}", flush=True)
            ok = ok and good

    for code, metric, n, dim, mns, trees, k in [(zo.MANHATTAN, z.ManhattanDistance(), 12000, 96, 256, 3, 10),
                                                (zo.CHEBYSHEV, z.ChebyshevDistance(), 3000, 100, 5, 15, 10),
                                                (zo.MINKOWSKI(3), z.MinkowskiDistance(3), 6000, 64, 128, 2, 40)]:
        rng = np.random.default_rng(77 + n)
        rows = clustered(rng, n, dim)
        queries = np.concatenate([rows[:64], clustered(rng, 128, dim)])
        ix = make(dim, mns, trees, metric, 7)
        ix.add(rows)
        orc = zo.OracleIndex(dim, code, mns, trees, seed=7) if rank == 0 else None
        if rank == 0:
            orc.add(rows)
        compare(f"{type(metric).__name__} n={n} leaf<{mns} bulk build", ix, orc, queries, k)
        dele = np.arange(0, n, 5, dtype=np.uint64)
        ix.remove_ordinals(dele)
        if rank == 0:
            orc.remove(dele)
        compare(f"{type(metric).__name__} after remove", ix, orc, queries, k)
        ix.set_param("seq_tile", 0)
        compare(f"{type(metric).__name__} one thread per pair", ix, orc, queries, k)
        ix.close()

    # import of an oracle-written store into the sharded index (every rank passes the same data)
    rng = np.random.default_rng(5)
    n, dim, T, X = 4000, 48, 5, 1234
    rows, ids, forest, blobs, order, clean = tampered_store(rng, n, dim, T, X)
    perm = rng.permutation(n)
    ix = make(dim, 8, T, z.L2SquaredDistance(), 5)
    rep = ix.import_store([ids[i] for i in perm], rows[perm], blobs)
    orc = None
    if rank == 0:
        good = rep["rows_loaded"] == n - 1 and rep["missing_ids"] == 2 and rep["orphans"] == [ids[X]]
        print(f"[mgpu-ext G={world}] import report: {'ok' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
        orc = zo.OracleIndex(dim, zo.L2SQ, 8, T, seed=5)
        orc.load_forest(rows[order], clean)
    compare("search on the imported store", ix, orc, np.concatenate([rows[:100], clustered(rng, 100, dim)]), 10)
    ix.close()

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if not int(flag.item()):
        sys.exit(1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Histogram of visits per leaf / queries per tile for the bench workload (needs a GPU)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zebra_b200 as z
rows, dim, nq, trees, mns = 1_000_000, 768, 10_000, 4, 2048
dev = torch.device("cuda", 0)
ix = z.LSHIndex(dim, z.LSHIndexOptions(mns, trees), z.L2Distance(), device=0, seed=0)
d_rows = torch.empty((rows, dim), dtype=torch.float32, device=dev)
z.synth_fill_device(0, d_rows.data_ptr(), 0, 1, rows, dim, 0, 1)
ix.add_device(d_rows.data_ptr(), rows)
d_q = torch.empty((nq, dim), dtype=torch.float32, device=dev)
z.synth_fill_device(0, d_q.data_ptr(), 3 * nq, 1, nq, dim, 1, 1)
keys, depth, leaf = ix.hash(d_q.cpu().numpy())
f = ix.export_forest()
leaf_len = np.diff(f.leaf_off)
cnt = np.bincount(leaf.reshape(-1), minlength=leaf_len.size)
print("leaves", leaf_len.size, "visited", (cnt > 0).sum(), "len mean", leaf_len.mean(), "min", leaf_len.min(), "max", leaf_len.max())
print("visits/leaf histogram:", np.bincount(np.minimum(cnt, 40)))
# current tiling (zb_scan.cu ts_filltiles_kernel): full tiles of 8 first, the remainder last; a tile costs
# ceil(nqt / 4) query groups x ceil(L / 128) row blocks of FP32 work, and reads its leaf once
tiles = []
for c, L in zip(cnt, leaf_len):
    if c == 0: continue
    full, rem = divmod(int(c), 8)
    tiles += [(8, L)] * full + ([(rem, L)] if rem else [])
t = np.array(tiles)
pairs = (t[:, 0] * t[:, 1]).sum()
padded = (((t[:, 0] + 3) // 4 * 4) * ((t[:, 1] + 127) // 128 * 128)).sum()
padq = (((t[:, 0] + 3) // 4 * 4) * t[:, 1]).sum()
print("tiles", len(t), "pairs", pairs, "padded pairs (query groups and row blocks)", padded, "efficiency", pairs / padded,
      "query padding only", pairs / padq)
print("tile size histogram:", np.bincount(t[:, 0]))
print("moved bytes", (t[:, 1] * dim * 4).sum() / 1e9, "GB; unique", (leaf_len[cnt > 0] * dim * 4).sum() / 1e9, "GB")

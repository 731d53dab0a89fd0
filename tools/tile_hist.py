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
tiles = []
for c, L in zip(cnt, leaf_len):
    if c == 0: continue
    nt = (c + 15) // 16; base, rem = divmod(c, nt)
    tiles += [(base + (1 if j < rem else 0), L) for j in range(nt)]
t = np.array(tiles)
G = (t[:, 0] + 3) // 4
for g in range(1, 5):
    m = G == g
    print(f"G={g}: tiles {m.sum()}, rows {t[m,1].sum()}, pairs {(t[m,0]*t[m,1]).sum()}")
cost = (np.ceil(G / 2) * t[:, 1]).sum()
print("sum L*ceil(G/2) =", cost, " ideal pairs/8 =", (t[:, 0] * t[:, 1]).sum() / 8, " efficiency", (t[:, 0] * t[:, 1]).sum() / 8 / cost)

#!/usr/bin/env python
"""Ablation of the scalar-metric leaf-tile scan on the config-2 shape (1M x 768, 4 trees of <= 2047-row leaves, 10k top-10
queries, Manhattan): kernel time (CUDA events around the launch, library stats) per knob setting.  One index build."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zebra_b200 as z

n, dim, nq, k = 1_000_000, 768, 10_000, 10
ix = z.LSHIndex(dim, z.LSHIndexOptions(2048, 4), z.ManhattanDistance(), seed=0)
d_rows = torch.empty((n, dim), dtype=torch.float32, device="cuda")
z.synth_fill_device(0, d_rows.data_ptr(), 0, 1, n, dim, 0, 1)
ix.add_device(d_rows.data_ptr(), n)
del d_rows
nb = 6
d_q = torch.empty((nb, nq, dim), dtype=torch.float32, device="cuda")
for b in range(nb):
    z.synth_fill_device(0, d_q[b].data_ptr(), b * nq, 1, nq, dim, 1, 1)
d_ord = torch.empty((nq, k), dtype=torch.int64, device="cuda")
d_bits = torch.empty((nq, k), dtype=torch.int64, device="cuda")
d_cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
ref = None
out = []
for pf in [int(x) for x in (sys.argv[1:] or ["0", "2", "4", "8", "16"])]:
    ix.set_param("seq_prefetch", pf)
    ms = []
    for b in range(nb):
        ix.search_batch_device(nq, d_q[b].data_ptr(), k, d_ord.data_ptr(), d_bits.data_ptr(), d_cnt.data_ptr())
        st = ix.stats()
        if b >= 2:
            ms.append(st["last_ms_tile_kernel"])
    torch.cuda.synchronize()
    sig = (int(d_ord.sum()), int(d_bits.sum()))
    ref = ref or sig
    out.append({"seq_prefetch": pf, "kernel_ms": sum(ms) / len(ms), "total_ms": st["last_ms_total"], "select_ms": st["last_ms_select"],
                "moved_gb": st["last_moved_bytes"] / 1e9, "same_result": sig == ref})
    print(json.dumps(out[-1]), flush=True)

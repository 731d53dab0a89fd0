#!/usr/bin/env python
"""Region-level view of an ncu SASS source page: samples, executions and main stall reasons per block of B instructions.
usage: tools/ncu_regions.py x.csv [kernel-index] [B] [min-samples]"""
import csv, sys, collections
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64; mins = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
rows = list(csv.reader(open(path)))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = []; blocks.append(cur)
    elif cur is not None and r: cur.append(r)
b = blocks[which]; hdr = b[0]; data = [r for r in b[1:] if len(r) >= len(hdr) and r[0] != "Address"]
col = {h: i for i, h in enumerate(hdr)}
si = col["# Samples"]; ie = col["Instructions Executed"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
tot = sum(int(r[si] or 0) for r in data)
print("instructions", len(data), "samples", tot)
for s0 in range(0, len(data), B):
    ch = data[s0:s0 + B]; s = sum(int(r[si] or 0) for r in ch); ex = sum(int(r[ie] or 0) for r in ch)
    if s < mins: continue
    ops = collections.Counter(); st = collections.Counter()
    for r in ch:
        t = r[1].split(); op = t[1] if t[0].startswith("@") else t[0]; ops[op.split(".")[0]] += 1
        for h in stall_cols: st[h] += int(r[col[h]] or 0)
    print(f"{s0:5d} samp={s:7d} {100*s/tot:5.1f}% exec={ex/1e6:8.1f}M {ops.most_common(3)} {[(k[6:], v) for k, v in st.most_common(4)]}")

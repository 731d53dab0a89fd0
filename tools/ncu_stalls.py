#!/usr/bin/env python
"""Aggregate the ncu source page (SASS) of one kernel: stall-reason totals and the hottest instructions.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv ; tools/ncu_stalls.py x.csv [kernel-index] [top]"""
import csv, sys, collections
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
# split into kernels
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None and r:
        cur["rows"].append(r)
b = blocks[which]
hdr = b["rows"][0]; data = b["rows"][1:]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for h in stall_cols:
        try: tot[h] += int(r[col[h]])
        except ValueError: pass
alls = sum(tot.values())
print(b["name"]); print("total samples", alls)
for h, v in tot.most_common(): print(f"  {h:28s} {v:9d} {100.0*v/max(1,alls):5.1f}%")
si = col["# Samples"]; ie = col["Instructions Executed"]
data2 = sorted(data, key=lambda r: -int(r[si] or 0))
print("hottest instructions:")
for r in data2[:top]:
    st = {h: int(r[col[h]] or 0) for h in stall_cols}
    main = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"  {r[0]:>6s} {int(r[si]):7d} ex={r[ie]:>10s} {r[1][:70]:70s} {main}")
# opcode histogram by executed instructions
ops = collections.Counter()
for r in data:
    op = r[1].split()[0] if r[1] else ""
    if op.startswith("@"): op = r[1].split()[1]
    try: ops[op.split(".")[0]] += int(r[ie])
    except ValueError: pass
print("executed warp-instructions by opcode:")
tote = sum(ops.values())
for o, v in ops.most_common(16): print(f"  {o:12s} {v:12d} {100.0*v/tote:5.1f}%")

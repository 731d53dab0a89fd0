"""Prints the essentials of bench.py JSON lines (files given on the command line)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable:", e)
        continue
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print(f"{path}: {d.get('value', 0):,.0f} {d.get('unit')}  {d.get('ms_per_step', 0):.3f} ms/step  e2e {e.get('value') or 0:,.0f}"
          f"  frac {r.get('frac')}  unique_frac {r.get('unique_bytes_frac')}  kernel_ms {r.get('kernel_ms_per_launch')}"
          f"  parity {d.get('parity_sample_ok')}  cpu {(d.get('cpu_baseline') or {}).get('value')}")
    if d.get("phases_ms_per_step"):
        print("    phases", {k: round(v, 3) for k, v in d["phases_ms_per_step"].items()}, "clocks", d.get("clocks"))
    if d.get("l2_filter"):
        print("    l2_filter", d["l2_filter"])
    for row in d.get("per_rank_plan_scan_tile_select_merge_ms_pairs_tiles") or []:
        print("     ", row)
    cfg = d.get("config", {})
    extra = {k: cfg[k] for k in ("planes_per_row", "fp32_tflops", "avg_depth") if k in cfg}
    if extra:
        print("    ", extra)

import json,sys
d=json.load(open(sys.argv[1]))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phases_ms_per_step"].items()})
print("e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), "clocks", d["clocks"])
for r in d.get("per_rank_plan_scan_tile_select_merge_ms_pairs_tiles") or []: print("  ", r)

#!/bin/bash
# Profile pass of round 1 (run under gpurun on one B200): launch list of the query step + one full capture of the scan kernel.
# usage: tools/profile_r1.sh <tag> [extra bench args]
set -u
TAG=${1:-r1}; shift || true
OUT=gpurun_out
mkdir -p $OUT
KERN='regex:plan_walk|compact_visits|tile_scan|ts_|score_pairs|select_visits|merge_|DeviceScan|sq_norms|pad_rows'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $OUT/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_scan -s 3 -c 2 -f -o $OUT/scan_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $OUT/scan_$TAG.log 2>&1
ls -la $OUT

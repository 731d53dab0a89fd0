#!/bin/bash
# Round 1, scalar-metric path: one full ncu capture of seq_tile_kernel on the config-2 shape with Manhattan distance.
set -u
OUT=gpurun_out
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:seq_tile_kernel -s 3 -c 1 -f -o $OUT/seq_r01h \
    python bench.py --metric manhattan --steps 2 --warmup 3 --no-cpu-baseline > $OUT/seq_r01h.log 2>&1
ls -la $OUT | tail -5
tail -3 $OUT/seq_r01h.log | cut -c1-300

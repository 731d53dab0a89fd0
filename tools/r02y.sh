#!/bin/bash
# Round 2, GPU call Y (1 GPU): ncu --set full of the cosine kernel and of the long-list kernel (eight math warps per team) with the final binary.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 110 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02y_cos \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --metric cosine > $OUT/scan3_r02y_cos.log 2>&1; echo "ncu cos rc=$?"
timeout 110 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02y_top100 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --topk 100 --metric l2sq --dim 384 > $OUT/scan3_r02y_top100.log 2>&1; echo "ncu top100 rc=$?"
ls -la $OUT/scan3_r02y_*.ncu-rep

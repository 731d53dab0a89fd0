#!/bin/bash
# Round 2, GPU call G (1 GPU): A/B of tile_scan3 build variants (stage depth KC, QH granularity, instruction order), same bench lines.
set -u
OUT=gpurun_out
mkdir -p $OUT
cp zebra_b200/libzebra_b200.so zebra_b200/variants/libzb_a.so
for v in a b c d e; do
  cp zebra_b200/variants/libzb_$v.so zebra_b200/libzebra_b200.so
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r02g_${v}_l2.json 2>> $OUT/r02g.err
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --metric cosine > $OUT/r02g_${v}_cos.json 2>> $OUT/r02g.err
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --metric l2sq --dim 384 > $OUT/r02g_${v}_l2sq384.json 2>> $OUT/r02g.err
done
cp zebra_b200/variants/libzb_a.so zebra_b200/libzebra_b200.so
python tools/show_bench.py $OUT/r02g_*.json | grep -v phases
tail -5 $OUT/r02g.err

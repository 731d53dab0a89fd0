#!/bin/bash
# Round 2, GPU call C (1 GPU): longest-first tile dispatch + epilogue prefetch in tile_scan3, flat projection on the scan3 skeleton.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $OUT/r02c_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $OUT/r02c_gpu_tests.log
tail -8 $OUT/r02c_gpu_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 > $OUT/r02c_bench_l2.json 2>> $OUT/r02c.err; echo "bench l2 rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --metric cosine --cpu-seconds 4 > $OUT/r02c_bench_cos.json 2>> $OUT/r02c.err; echo "bench cos rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --set tile_queries=256 > $OUT/r02c_bench_l2_leaforder.json 2>> $OUT/r02c.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --metric cosine --set tile_queries=256 > $OUT/r02c_bench_cos_leaforder.json 2>> $OUT/r02c.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 100 --metric l2sq --dim 384 > $OUT/r02c_bench_top100.json 2>> $OUT/r02c.err
for cfg in "4 4" "16 8" "16 15"; do
  set -- $cfg
  for fp in 1 2; do
    timeout 150 python bench.py --workload hash --flat-bits $1 --trees $2 --steps 5 --warmup 3 --cpu-seconds 2 --set flat_project=$fp > $OUT/r02c_bench_hash_flat_K$1_T$2_fp$fp.json 2>> $OUT/r02c.err; echo "flat K=$1 T=$2 fp=$fp rc=$?"
  done
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:project3_kernel -s 3 -c 1 -f -o $OUT/project3_r02c \
    python bench.py --workload hash --flat-bits 16 --trees 8 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/project3_r02c.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02c_l2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/scan3_r02c_l2.log 2>&1
python tools/show_bench.py $OUT/r02c_bench_*.json
tail -5 $OUT/r02c.err

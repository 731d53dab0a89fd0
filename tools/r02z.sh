#!/bin/bash
# Round 2, GPU call Z (1 GPU): tile construction launched ahead of the plan's readback (knob early_tiles) -- -m gpu suite, A/B bench.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 100 python -m pytest tests -m gpu -q -x > $OUT/r02z_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $OUT/r02z_gpu_tests.log
tail -3 $OUT/r02z_gpu_tests.log
timeout 45 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r02z_bench_l2.json 2>> $OUT/r02z.err; echo "bench rc=$?"
timeout 45 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --set early_tiles=0 > $OUT/r02z_bench_l2_late.json 2>> $OUT/r02z.err
python tools/show_bench.py $OUT/r02z_bench_l2.json $OUT/r02z_bench_l2_late.json | grep -v "^      \[\|l2_filter"

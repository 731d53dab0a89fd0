#!/bin/bash
# Round 2, GPU call X (1 GPU): the committed final binary -- -m gpu suite, the driver's bench line, the launch list of one step.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/r02x_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $OUT/r02x_gpu_tests.log
tail -3 $OUT/r02x_gpu_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 > $OUT/r02x_bench_l2.json 2>> $OUT/r02x.err; echo "bench l2 rc=$?"
KERN='regex:plan_walk|compact_visits|tile_scan|refine_|n2_|ts_|score_pairs|select_visits|merge_|DeviceScan|DeviceRadix|rinv|pad_rows|plan_totals|quad_tile'
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file $OUT/r02x_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r02x_launches.log 2>&1
python tools/show_bench.py $OUT/r02x_bench_l2.json | grep -v "^      \["

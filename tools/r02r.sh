#!/bin/bash
# Round 2, GPU call R (1 GPU): L2 / L2 squared through the dot-product filter (METRIC 3 of tile_scan3_kernel + refine_visits_kernel):
# its parity tests, the whole -m gpu suite, bench lines with the filter on (default) and off, ncu captures of both passes.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_l2_filter.py -m gpu -q -x --durations=5 > $OUT/r02r_filter_tests.log 2>&1; echo "filter tests rc=$?" | tee -a $OUT/r02r_filter_tests.log
tail -30 $OUT/r02r_filter_tests.log
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $OUT/r02r_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $OUT/r02r_gpu_tests.log
tail -12 $OUT/r02r_gpu_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 > $OUT/r02r_bench_l2.json 2>> $OUT/r02r.err; echo "bench l2 rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --set l2_filter=0 > $OUT/r02r_bench_l2_nofilter.json 2>> $OUT/r02r.err; echo "bench l2 nofilter rc=$?"
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 10 --metric l2sq --dim 384 > $OUT/r02r_bench_l2sq384.json 2>> $OUT/r02r.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 10 --metric l2sq --dim 384 --set l2_filter=0 > $OUT/r02r_bench_l2sq384_nofilter.json 2>> $OUT/r02r.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 1 > $OUT/r02r_bench_l2_top1.json 2>> $OUT/r02r.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 16 > $OUT/r02r_bench_l2_top16.json 2>> $OUT/r02r.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02r_l2f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/scan3_r02r_l2f.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:refine_visits -s 3 -c 1 -f -o $OUT/refine_r02r \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/refine_r02r.log 2>&1
KERN='regex:plan_walk|compact_visits|tile_scan|refine_|n2_|ts_|score_pairs|select_visits|merge_|DeviceScan|DeviceRadix|rinv|pad_rows|plan_totals'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file $OUT/r02r_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r02r_launches.log 2>&1
python tools/show_bench.py $OUT/r02r_bench_*.json
tail -5 $OUT/r02r.err

#!/bin/bash
# Round 2, GPU call B (1 GPU): third-generation leaf-tile scan -- the whole -m gpu suite, bench lines (L2, cosine, second
# generation next to it), launch list and one ncu --set full capture of tile_scan3_kernel (L2 and cosine).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > $OUT/r02b_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $OUT/r02b_gpu_tests.log
tail -15 $OUT/r02b_gpu_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 > $OUT/r02b_bench_l2.json 2>> $OUT/r02b.err; echo "bench l2 rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --metric cosine --cpu-seconds 4 > $OUT/r02b_bench_cos.json 2>> $OUT/r02b.err; echo "bench cos rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --metric l2sq --dim 384 --no-cpu-baseline > $OUT/r02b_bench_l2sq384.json 2>> $OUT/r02b.err; echo "bench l2sq 384 rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --set scan_gen=2 > $OUT/r02b_bench_l2_gen2.json 2>> $OUT/r02b.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --queries 40000 > $OUT/r02b_bench_l2_q40k.json 2>> $OUT/r02b.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --queries 2500 > $OUT/r02b_bench_l2_q2500.json 2>> $OUT/r02b.err
KERN='regex:plan_walk|compact_visits|tile_scan|ts_|score_pairs|select_visits|merge_|DeviceScan|rinv|pad_rows|plan_totals'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file $OUT/r02b_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r02b_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02b_l2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/scan3_r02b_l2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02b_cos \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --metric cosine > $OUT/scan3_r02b_cos.log 2>&1
python tools/show_bench.py $OUT/r02b_bench_*.json
tail -5 $OUT/r02b.err

#!/bin/bash
# Round 2, GPU call L (1 GPU): whole -m gpu suite (fused lists of up to 128 entries, prefetch), bench lines with the default build,
# top-100 through the fused kernel, ncu capture of tile_scan3 (L2, cosine) for the committed traffic figure.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $OUT/r02l_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $OUT/r02l_gpu_tests.log
tail -8 $OUT/r02l_gpu_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 > $OUT/r02l_bench_l2.json 2>> $OUT/r02l.err; echo "bench l2 rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --metric cosine --cpu-seconds 4 > $OUT/r02l_bench_cos.json 2>> $OUT/r02l.err; echo "bench cos rc=$?"
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 100 --metric l2sq --dim 384 > $OUT/r02l_bench_top100.json 2>> $OUT/r02l.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 100 --metric l2sq --dim 384 --delete-frac 0.1 > $OUT/r02l_bench_top100_tomb.json 2>> $OUT/r02l.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 10 --metric l2sq --dim 384 > $OUT/r02l_bench_top10_384.json 2>> $OUT/r02l.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 100 --metric l2sq --dim 384 --set use_tile_scan=0 > $OUT/r02l_bench_top100_quadtile.json 2>> $OUT/r02l.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02l_l2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/scan3_r02l_l2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02l_cos \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --metric cosine > $OUT/scan3_r02l_cos.log 2>&1
KERN='regex:plan_walk|compact_visits|tile_scan|ts_|score_pairs|select_visits|merge_|DeviceScan|DeviceRadix|rinv|pad_rows|plan_totals'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file $OUT/r02l_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r02l_launches.log 2>&1
python tools/show_bench.py $OUT/r02l_bench_*.json
tail -5 $OUT/r02l.err

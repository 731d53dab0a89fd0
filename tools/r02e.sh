#!/bin/bash
# Round 2, GPU call E (1 GPU): tile_scan3 with 2-chunk stages, operands loaded half a chunk ahead across stage boundaries, QH = 1..8.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_headline_shapes.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/r02e_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/r02e_gpu_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --cpu-seconds 4 > $OUT/r02e_bench_l2.json 2>> $OUT/r02e.err; echo "bench l2 rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --metric cosine --no-cpu-baseline > $OUT/r02e_bench_cos.json 2>> $OUT/r02e.err; echo "bench cos rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --metric l2sq --dim 384 --no-cpu-baseline > $OUT/r02e_bench_l2sq384.json 2>> $OUT/r02e.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02e_l2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/scan3_r02e_l2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02e_cos \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --metric cosine > $OUT/scan3_r02e_cos.log 2>&1
python tools/show_bench.py $OUT/r02e_bench_*.json
tail -5 $OUT/r02e.err

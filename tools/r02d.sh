#!/bin/bash
# Round 2, GPU call D (1 GPU): tile_scan3 with gthr loads one block ahead and QH = 1..8 (3-chunk stages).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_headline_shapes.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/r02d_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/r02d_gpu_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --cpu-seconds 4 > $OUT/r02d_bench_l2.json 2>> $OUT/r02d.err; echo "bench l2 rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --metric cosine --no-cpu-baseline > $OUT/r02d_bench_cos.json 2>> $OUT/r02d.err; echo "bench cos rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --metric l2sq --dim 384 --no-cpu-baseline > $OUT/r02d_bench_l2sq384.json 2>> $OUT/r02d.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02d_l2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/scan3_r02d_l2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02d_cos \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --metric cosine > $OUT/scan3_r02d_cos.log 2>&1
python tools/show_bench.py $OUT/r02d_bench_*.json
tail -5 $OUT/r02d.err

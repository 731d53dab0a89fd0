#!/bin/bash
# Round 2, GPU call U (1 GPU): both team shapes in one library (long lists on eight math warps per team), refine pass with twelve
# row chunks in flight, plan-walk tail kernel: whole -m gpu suite, bench lines (config 2, top-100, 384 dims), ncu of the refine pass.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $OUT/r02u_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $OUT/r02u_gpu_tests.log
tail -12 $OUT/r02u_gpu_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 > $OUT/r02u_bench_l2.json 2>> $OUT/r02u.err; echo "bench l2 rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --metric cosine > $OUT/r02u_bench_cos.json 2>> $OUT/r02u.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 100 --metric l2sq --dim 384 > $OUT/r02u_bench_top100.json 2>> $OUT/r02u.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 100 --metric l2sq --dim 384 --set long_list_warps=4 > $OUT/r02u_bench_top100_4warps.json 2>> $OUT/r02u.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 100 --metric cosine > $OUT/r02u_bench_top100_cos768.json 2>> $OUT/r02u.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 100 --metric cosine --set long_list_warps=4 > $OUT/r02u_bench_top100_cos768_4warps.json 2>> $OUT/r02u.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 10 --metric l2sq --dim 384 > $OUT/r02u_bench_l2sq384.json 2>> $OUT/r02u.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:refine_visits -s 3 -c 1 -f -o $OUT/refine_r02u \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/refine_r02u.log 2>&1
python tools/show_bench.py $OUT/r02u_bench_*.json | grep -v "phases"
tail -5 $OUT/r02u.err

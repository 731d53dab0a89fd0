#!/bin/bash
# Round 2, GPU call A (1 GPU): the whole -m gpu suite (knob fix, headline-shape parity), baseline bench lines of the round-1
# kernel (L2, cosine, top-100 both gather variants), flat-table and tree-forest hashing throughput, one ncu capture of the
# projection kernel and of hash_kernel.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/r02a_gpu.txt 2>&1
nproc >> $OUT/r02a_gpu.txt; free -g >> $OUT/r02a_gpu.txt
timeout 900 python -m pytest tests -m gpu -q --durations=12 > $OUT/r02a_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $OUT/r02a_gpu_tests.log
tail -25 $OUT/r02a_gpu_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 > $OUT/r02a_bench_l2.json 2>> $OUT/r02a.err; echo "bench l2 rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --metric cosine --cpu-seconds 4 > $OUT/r02a_bench_cos.json 2>> $OUT/r02a.err; echo "bench cos rc=$?"
for q in 0 1; do timeout 150 python bench.py --topk 100 --metric l2sq --dim 384 --steps 3 --warmup 3 --no-cpu-baseline --set quad_tile=$q --set select_variant=$q > $OUT/r02a_bench_top100_quad$q.json 2>> $OUT/r02a.err; echo "top-100 quad_tile=$q rc=$?"; done
for cfg in "4 4" "16 8" "16 15"; do
  set -- $cfg
  timeout 150 python bench.py --workload hash --flat-bits $1 --trees $2 --steps 5 --warmup 3 --cpu-seconds 3 > $OUT/r02a_bench_hash_flat_K$1_T$2.json 2>> $OUT/r02a.err; echo "flat K=$1 T=$2 rc=$?"
done
timeout 150 python bench.py --workload hash --steps 5 --warmup 3 --cpu-seconds 3 > $OUT/r02a_bench_hash_forest.json 2>> $OUT/r02a.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:project_flat_kernel -s 3 -c 1 -f -o $OUT/project_r02a \
    python bench.py --workload hash --flat-bits 16 --trees 8 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/project_r02a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hash_kernel -s 3 -c 1 -f -o $OUT/hash_r02a \
    python bench.py --workload hash --steps 2 --warmup 3 --no-cpu-baseline > $OUT/hash_r02a.log 2>&1
for v in 0 1; do timeout 100 python bench.py --metric manhattan --steps 3 --warmup 3 --no-cpu-baseline --set select_variant=$v > $OUT/r02a_bench_manhattan_select$v.json 2>> $OUT/r02a.err; echo "manhattan select_variant=$v rc=$?"; done
python tools/show_bench.py $OUT/r02a_bench_*.json 2>/dev/null | tail -40
tail -5 $OUT/r02a.err

#!/bin/bash
# Round 2, GPU call S (2 GPUs): the dot-product filter and the plan-walk tail kernel on a bucket-sharded index -- whole -m gpu suite
# (the 2-GPU tests included), sharded parity scripts, weak-scaling bench with per-rank traces.
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests -m gpu -q -x --durations=5 > $OUT/r02s_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $OUT/r02s_gpu_tests.log
tail -12 $OUT/r02s_gpu_tests.log
timeout 300 $TR --master-port 29551 tests/mgpu_parity.py > $OUT/r02s_mgpu_parity_2gpu.log 2>&1; echo "mgpu_parity rc=$?"
grep -c ": ok" $OUT/r02s_mgpu_parity_2gpu.log; grep -i "mismatch\|error" $OUT/r02s_mgpu_parity_2gpu.log | head -5
timeout 300 $TR --master-port 29552 tests/mgpu_parity_ext.py > $OUT/r02s_mgpu_parity_ext_2gpu.log 2>&1; echo "mgpu_parity_ext rc=$?"
grep -c ": ok" $OUT/r02s_mgpu_parity_ext_2gpu.log; grep -i "mismatch\|error" $OUT/r02s_mgpu_parity_ext_2gpu.log | head -5
ZB_TRACE=2 timeout 240 $TR --master-port 29553 bench.py --gpus 2 --steps 8 --warmup 3 > $OUT/r02s_bench_2gpu.json 2> $OUT/r02s_bench_2gpu.err; echo "bench rc=$?"
python tools/show_bench.py $OUT/r02s_bench_2gpu.json | grep -v "^      \["
grep "zb trace" $OUT/r02s_bench_2gpu.err | grep "plan walk" | tail -6
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r02s_bench_l2_1gpu.json 2>> $OUT/r02s.err; echo "bench 1gpu rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --set plan_tail=0 > $OUT/r02s_bench_l2_1gpu_notail.json 2>> $OUT/r02s.err
python tools/show_bench.py $OUT/r02s_bench_l2_1gpu.json $OUT/r02s_bench_l2_1gpu_notail.json | grep -v "^      \["

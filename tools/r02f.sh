#!/bin/bash
# Round 2, GPU call F (2 GPUs): sharded parity (bulk build, deletes, inserts, deduplicate, sliced search, store exchanged in leaf
# groups), the default weak-scaling bench at N = 2 with the oracle parity sample, BASELINE config 3 at N = 2.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline_shapes.py -m gpu -q -x > $OUT/r02f_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02f_gpu_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --cpu-seconds 4 > $OUT/r02f_bench_l2.json 2>> $OUT/r02f.err; echo "bench l2 rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-prefetch > $OUT/r02f_bench_l2_noprefetch.json 2>> $OUT/r02f.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/mgpu_parity.py > $OUT/r02f_mgpu_parity.log 2>&1; echo "mgpu_parity rc=$?"
grep -c ": ok" $OUT/r02f_mgpu_parity.log; grep -i "mismatch\|error" $OUT/r02f_mgpu_parity.log | head -5
timeout 300 $TR --master-port 29512 tests/mgpu_parity_ext.py > $OUT/r02f_mgpu_parity_ext.log 2>&1; echo "mgpu_parity_ext rc=$?"
grep -c ": ok" $OUT/r02f_mgpu_parity_ext.log
ZB_TRACE=1 timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r02f_bench_2gpu.json 2> $OUT/r02f_bench_2gpu.err; echo "bench 2gpu rc=$?"
timeout 400 $TR --master-port 29514 bench.py --gpus 2 --preset 3 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02f_bench_preset3_2gpu.json 2> $OUT/r02f_bench_preset3_2gpu.err; echo "preset 3 rc=$?"
python tools/show_bench.py $OUT/r02f_bench_*.json
grep "zb trace" $OUT/r02f_bench_2gpu.err | tail -4
tail -3 $OUT/r02f_bench_preset3_2gpu.err

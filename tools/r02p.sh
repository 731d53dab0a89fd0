#!/bin/bash
# Round 2, GPU call P (2 GPUs): sharded parity with the three query-exchange modes (peer push, combined allgather, own allgather), bench at N = 2.
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/mgpu_parity.py > $OUT/r02p_mgpu_parity.log 2>&1; echo "mgpu_parity rc=$?"
grep -c ": ok" $OUT/r02p_mgpu_parity.log; grep -i "mismatch\|error" $OUT/r02p_mgpu_parity.log | head -5
ZB_TRACE=2 timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02p_bench_2gpu.json 2> $OUT/r02p_bench_2gpu.err; echo "bench 2gpu rc=$?"
python tools/show_bench.py $OUT/r02p_bench_*.json
grep "zb trace" $OUT/r02p_bench_2gpu.err | tail -4
grep -v "zb trace" $OUT/r02p_bench_2gpu.err | grep -iv "warn\|OMP_NUM\|\*\*\*" | tail -5

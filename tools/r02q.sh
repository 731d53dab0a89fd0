#!/bin/bash
# Round 2, GPU call Q (8 GPUs): the peer push of the query slices at 8 ranks -- weak-scaling bench with per-rank traces and the oracle
# parity sample, sharded parity script (three exchange modes).
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
ZB_TRACE=2 timeout 240 $TR --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/r02q_bench_8gpu.json 2> $OUT/r02q_bench_8gpu.err; echo "bench rc=$?"
timeout 200 $TR --master-port 29542 tests/mgpu_parity.py > $OUT/r02q_mgpu_parity_8gpu.log 2>&1; echo "mgpu_parity rc=$?"
grep -c ": ok" $OUT/r02q_mgpu_parity_8gpu.log; grep -i "mismatch" $OUT/r02q_mgpu_parity_8gpu.log | head -3
python tools/show_bench.py $OUT/r02q_bench_8gpu.json | grep -v "^      \["
grep "zb trace" $OUT/r02q_bench_8gpu.err | tail -32 | head -16
grep -v "zb trace" $OUT/r02q_bench_8gpu.err | grep -iv "warn\|OMP_NUM\|\*\*\*" | tail -5

#!/bin/bash
# Round 2, GPU call V (8 GPUs): weak-scaling bench (1M rows + 10 k queries per GPU) with the dot-product filter and the plan-walk
# tail kernel, per-rank phase traces, oracle parity sample on rank 0.
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
ZB_TRACE=2 timeout 170 $TR --master-port 29561 bench.py --gpus 8 --steps 8 --warmup 3 > $OUT/r02v_bench_8gpu.json 2> $OUT/r02v_bench_8gpu.err; echo "bench rc=$?"
python tools/show_bench.py $OUT/r02v_bench_8gpu.json | grep -v "^      \["
grep "zb trace" $OUT/r02v_bench_8gpu.err | grep "plan walk" | tail -16 | cut -c1-200
grep -v "zb trace" $OUT/r02v_bench_8gpu.err | grep -iv "warn\|OMP_NUM\|\*\*\*" | tail -5

#!/bin/bash
# Round 2, GPU call M (2 GPUs): sharded parity with the combined visit-record + query-slice exchange (and the separate one), bench at N = 2.
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/mgpu_parity.py > $OUT/r02m_mgpu_parity.log 2>&1; echo "mgpu_parity rc=$?"
grep -c ": ok" $OUT/r02m_mgpu_parity.log; grep -i "mismatch\|error" $OUT/r02m_mgpu_parity.log | head -5
ZB_TRACE=2 timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02m_bench_2gpu.json 2> $OUT/r02m_bench_2gpu.err; echo "bench 2gpu rc=$?"
ZB_TRACE=2 timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --set single_exchange=0 > $OUT/r02m_bench_2gpu_sep.json 2> $OUT/r02m_bench_2gpu_sep.err; echo "bench 2gpu sep rc=$?"
python tools/show_bench.py $OUT/r02m_bench_*.json
grep "zb trace" $OUT/r02m_bench_2gpu.err | tail -4
grep "zb trace" $OUT/r02m_bench_2gpu_sep.err | tail -4

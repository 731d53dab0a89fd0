#!/bin/bash
# Round 2, GPU call N (8 GPUs): where the two pre-scan collectives spend their time -- per-rank phase traces of the default
# weak-scaling bench with the combined exchange (default) and with separate exchanges; NCCL's own log once.
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
ZB_TRACE=2 timeout 200 $TR --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02n_bench_8gpu_one.json 2> $OUT/r02n_bench_8gpu_one.err; echo "one exchange rc=$?"
ZB_TRACE=2 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 200 $TR --master-port 29532 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --set single_exchange=0 > $OUT/r02n_bench_8gpu_sep.json 2> $OUT/r02n_bench_8gpu_sep.err; echo "separate rc=$?"
python tools/show_bench.py $OUT/r02n_bench_*.json | grep -v "^      \["
grep "zb trace" $OUT/r02n_bench_8gpu_one.err | tail -24 | head -16
echo ----
grep "zb trace" $OUT/r02n_bench_8gpu_sep.err | tail -24 | head -8
grep -i "NCCL INFO.*\(NVLS\|P2P\|SHM\|Channel\|Using\|Connected\|algo\)" $OUT/r02n_bench_8gpu_sep.err | sort | uniq -c | sort -rn | head -12

#!/bin/bash
# Round 2, GPU call H (8 GPUs): the BASELINE configs at their stated sizes.  The default weak-scaling line (with the oracle parity
# sample and ZB_TRACE phase times), config 3 (10M x 768 cosine), the north-star size (100M x 768), config 5 (100M x 384, 10 %
# tombstones, top-100), config 4 (hashing of 100M x 768 rows: flat tables H = 128 and the tree forest), sharded parity at 8 ranks.
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/r02h_gpus.txt 2>&1; nproc >> $OUT/r02h_gpus.txt; free -g >> $OUT/r02h_gpus.txt
ZB_TRACE=1 timeout 240 $TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/r02h_bench_8gpu.json 2> $OUT/r02h_bench_8gpu.err; echo "bench 8gpu rc=$?"
timeout 240 $TR --master-port 29523 bench.py --gpus 8 --preset 3 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02h_config3_8gpu.json 2> $OUT/r02h_config3_8gpu.err; echo "config 3 rc=$?"
timeout 330 $TR --master-port 29525 bench.py --gpus 8 --preset 6 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02h_100Mx768_8gpu.json 2> $OUT/r02h_100Mx768_8gpu.err; echo "100M x 768 rc=$?"
timeout 330 $TR --master-port 29524 bench.py --gpus 8 --preset 5 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r02h_config5_8gpu.json 2> $OUT/r02h_config5_8gpu.err; echo "config 5 rc=$?"
timeout 200 $TR --master-port 29526 bench.py --gpus 8 --preset 4 --flat-bits 16 --trees 8 --warmup 3 --no-cpu-baseline > $OUT/r02h_config4_flat128_8gpu.json 2> $OUT/r02h_config4_flat128_8gpu.err; echo "config 4 flat rc=$?"
timeout 200 $TR --master-port 29527 bench.py --gpus 8 --preset 4 --warmup 3 --no-cpu-baseline > $OUT/r02h_config4_forest_8gpu.json 2> $OUT/r02h_config4_forest_8gpu.err; echo "config 4 forest rc=$?"
timeout 240 $TR --master-port 29521 tests/mgpu_parity.py > $OUT/r02h_mgpu_parity_8gpu.log 2>&1; echo "mgpu_parity rc=$?"
grep -c ": ok" $OUT/r02h_mgpu_parity_8gpu.log; grep -i "mismatch" $OUT/r02h_mgpu_parity_8gpu.log | head -3
python tools/show_bench.py $OUT/r02h_*.json
grep "zb trace" $OUT/r02h_bench_8gpu.err | tail -3
for f in $OUT/r02h_*.err; do echo "== $f"; grep -v "zb trace\|Warning\|warn\|OMP_NUM\|\*\*\*\*" $f | tail -3; done

#!/bin/bash
# First GPU call of the next round: the flat-table path (project_flat_kernel, zb_index_load_flat) passed its parity tests at
# the very end of round 1 but its throughput was never measured (budget spent).  Parity again, then its hashing throughput
# at H = 16 / 128 / 240 planes per row next to the tree-forest walker, then one ncu capture of the
# projection kernel.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests/test_zz_flat_tables_gpu.py tests/test_zz_quad_tile_gpu.py -q --durations=5 > $OUT/r02a_flat_tests.log 2>&1; echo "flat + quad-tile tests rc=$?" >> $OUT/r02a_flat_tests.log
tail -12 $OUT/r02a_flat_tests.log
# top-100 (BASELINE config 5 asks for it): one quad per pair vs the keys-only tile scan (never run on a GPU in round 1)
for q in 0 1; do timeout 150 python bench.py --topk 100 --metric cosine --dim 384 --steps 3 --warmup 3 --no-cpu-baseline --set quad_tile=$q > $OUT/r02a_bench_top100_quad$q.json 2>> $OUT/r02a.err; echo "top-100 quad_tile=$q rc=$?"; done
for cfg in "16 1" "16 8" "16 15"; do
  set -- $cfg
  timeout 150 python bench.py --workload hash --flat-bits $1 --trees $2 --steps 5 --warmup 3 --cpu-seconds 3 > $OUT/r02a_bench_hash_flat_K$1_T$2.json 2>> $OUT/r02a.err; echo "flat K=$1 T=$2 rc=$?"
done
timeout 150 python bench.py --workload hash --steps 5 --warmup 3 --cpu-seconds 3 > $OUT/r02a_bench_hash_forest.json 2>> $OUT/r02a.err
ncu --set full --clock-control none --import-source on -k regex:project_flat_kernel -s 3 -c 1 -f -o $OUT/project_r02a \
    python bench.py --workload hash --flat-bits 16 --trees 8 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/project_r02a.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02a_bench_hash_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], d["ms_per_step"], d["roofline"]["frac"], d.get("parity_sample_ok"), {k: d.get(k, d.get("config", {}).get(k)) for k in ("planes_per_row", "fp32_tflops")})
    except Exception as e:
        print(f, "unreadable", e)
PY

# the per-visit select of the gather path: block bitonic vs one warp per visit (scalar-metric step, config-2 shape)
for v in 0 1; do timeout 100 python bench.py --metric manhattan --steps 3 --warmup 3 --no-cpu-baseline --set select_variant=$v > $OUT/r02a_bench_manhattan_select$v.json 2>> $OUT/r02a.err; echo "manhattan select_variant=$v rc=$?"; done

# Later calls of round 2 (multi-GPU, charged N x): the BASELINE configs at their own sizes, strong scaling
#   gpurun --gpus 8 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --preset 3 --steps 5 --warmup 3'
#   gpurun --gpus 8 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --preset 5 --steps 3 --warmup 3 --set quad_tile=1'
# 8-GPU plan exchange: the volume is nqp x T x visit_slots x 8 B per rank (82 MB at 32 slots, ~97 % padding on the config-2 shape).
# Cheapest experiment first -- fewer slots per walker (the plan grows and replans on overflow, so run enough warm-up steps):
#   ... bench.py --gpus 8 --steps 10 --warmup 5 --set visit_slots=8      (and 4), ZB_TRACE=1 for the per-phase times


mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r01h_gpu.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_scalar_metrics.py tests/test_gpu_store_interchange.py -q --durations=5 > gpurun_out/r01h_new_tests.log 2>&1; echo "new tests rc=$?" >> gpurun_out/r01h_new_tests.log
tail -8 gpurun_out/r01h_new_tests.log
timeout 100 python bench.py --metric manhattan --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01h_bench_manhattan.json 2> gpurun_out/r01h_bench_manhattan.err; echo "bench rc=$?"
timeout 100 python bench.py --metric manhattan --steps 3 --warmup 3 --no-cpu-baseline --set seq_tile=0 > gpurun_out/r01h_bench_manhattan_pairs.json 2>> gpurun_out/r01h_bench_manhattan.err; echo "bench rc=$?"
for v in 0 2; do timeout 100 python bench.py --workload hash --steps 5 --warmup 3 --no-cpu-baseline --set hash_variant=$v > gpurun_out/r01h_bench_hash_v$v.json 2>> gpurun_out/r01h_bench_manhattan.err; echo "hash v$v rc=$?"; done
python - <<'PY'
import json
for f in ("manhattan", "manhattan_pairs", "hash_v0", "hash_v2"):
    try:
        d = json.loads(open(f"gpurun_out/r01h_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], d["ms_per_step"], d.get("roofline", {}).get("frac"), d.get("phases_ms_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY

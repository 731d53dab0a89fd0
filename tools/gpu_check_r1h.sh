mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_scalar_metrics.py -q -x --durations=3 > gpurun_out/r01i_new_tests.log 2>&1; echo "scalar tests rc=$?" >> gpurun_out/r01i_new_tests.log
tail -4 gpurun_out/r01i_new_tests.log
timeout 100 python bench.py --metric manhattan --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01i_bench_manhattan.json 2> gpurun_out/r01i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("manhattan",):
    try:
        d = json.loads(open(f"gpurun_out/r01i_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], d["ms_per_step"], d.get("roofline", {}).get("frac"), d.get("phases_ms_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY

mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:seq_tile_kernel -s 3 -c 1 -f -o gpurun_out/seq_r01k \
    python bench.py --metric manhattan --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/seq_r01k.log 2>&1; echo "ncu rc=$?"
timeout 150 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r01k_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r01k_gpu_tests.log
tail -4 gpurun_out/r01k_gpu_tests.log
timeout 100 python bench.py --cpu-seconds 3 > gpurun_out/r01k_bench_default.json 2> gpurun_out/r01k.err; echo "bench rc=$?"
timeout 60 python bench.py --metric manhattan --no-cpu-baseline > gpurun_out/r01k_bench_manhattan.json 2>> gpurun_out/r01k.err; echo "bench rc=$?"
python -c "
import json
for f in ('default','manhattan'):
    d=json.loads(open(f'gpurun_out/r01k_bench_{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('parity_sample_ok'))
"

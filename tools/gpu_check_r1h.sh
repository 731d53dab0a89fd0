mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r01h_gpu.txt 2>&1
timeout 170 python -m pytest tests/test_gpu_scalar_metrics.py tests/test_gpu_store_interchange.py -q -x --durations=5 > gpurun_out/r01h_new_tests.log 2>&1; echo "new tests rc=$?" >> gpurun_out/r01h_new_tests.log
tail -5 gpurun_out/r01h_new_tests.log
timeout 100 python bench.py --metric manhattan --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01h_bench_manhattan.json 2> gpurun_out/r01h_bench_manhattan.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r01h_bench_manhattan.json
timeout 240 python -m pytest tests/test_gpu_parity.py -q -x --durations=5 > gpurun_out/r01h_parity_tests.log 2>&1; echo "parity rc=$?" >> gpurun_out/r01h_parity_tests.log
tail -5 gpurun_out/r01h_parity_tests.log

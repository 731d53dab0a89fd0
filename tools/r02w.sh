#!/bin/bash
# Round 2, GPU call W (1 GPU): final state -- whole -m gpu suite, smoke(), the driver's bench line, the launch list of one step, and
# the L2-ahead prefetch of the scan's producer (knob scan_l2_ahead) at several distances.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $OUT/r02w_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $OUT/r02w_gpu_tests.log
tail -4 $OUT/r02w_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02w_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/r02w_smoke.log
timeout 200 python bench.py --steps 10 --warmup 3 > $OUT/r02w_bench_l2.json 2>> $OUT/r02w.err; echo "bench l2 rc=$?"
for a in 4 8 16; do
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --set scan_l2_ahead=$a > $OUT/r02w_bench_l2_ahead$a.json 2>> $OUT/r02w.err
done
timeout 200 python bench.py --steps 10 --warmup 3 --cpu-seconds 4 --set scan_l2_ahead=8 --metric cosine > $OUT/r02w_bench_cos_ahead8.json 2>> $OUT/r02w.err
KERN='regex:plan_walk|compact_visits|tile_scan|refine_|n2_|ts_|score_pairs|select_visits|merge_|DeviceScan|DeviceRadix|rinv|pad_rows|plan_totals|quad_tile'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file $OUT/r02w_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r02w_launches.log 2>&1
python tools/show_bench.py $OUT/r02w_bench_*.json | grep -v "phases\|l2_filter"
tail -5 $OUT/r02w.err

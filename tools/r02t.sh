#!/bin/bash
# Round 2, GPU call T (1 GPU): A/B of the leaf-tile scan's team shape on the same box -- a = 4 math warps per team x 16 rows (default),
# h = 8 math warps per team x 8 rows (640 threads, four math warps per scheduler, 112 registers each); parity tests with h first.
set -u
OUT=gpurun_out
mkdir -p $OUT
cp zebra_b200/libzebra_b200.so zebra_b200/variants/libzb_a.so
cp zebra_b200/variants/libzb_h.so zebra_b200/libzebra_b200.so
timeout 400 python -m pytest tests/test_gpu_l2_filter.py tests/test_gpu_headline_shapes.py tests/test_zz_flat_tables_gpu.py -m gpu -q -x > $OUT/r02t_gpu_tests_h.log 2>&1; echo "pytest (h) rc=$?"; tail -3 $OUT/r02t_gpu_tests_h.log
for v in h a; do
  cp zebra_b200/variants/libzb_$v.so zebra_b200/libzebra_b200.so
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r02t_${v}_l2.json 2>> $OUT/r02t.err
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --metric cosine > $OUT/r02t_${v}_cos.json 2>> $OUT/r02t.err
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --metric l2sq --dim 384 > $OUT/r02t_${v}_l2sq384.json 2>> $OUT/r02t.err
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --topk 100 --metric l2sq --dim 384 > $OUT/r02t_${v}_top100.json 2>> $OUT/r02t.err
done
cp zebra_b200/variants/libzb_h.so zebra_b200/libzebra_b200.so
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_scan3 -s 3 -c 1 -f -o $OUT/scan3_r02t_h_l2f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/scan3_r02t_h_l2f.log 2>&1
cp zebra_b200/variants/libzb_a.so zebra_b200/libzebra_b200.so
python tools/show_bench.py $OUT/r02t_*.json | grep -v "phases\|l2_filter"
tail -5 $OUT/r02t.err

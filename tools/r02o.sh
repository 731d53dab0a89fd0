#!/bin/bash
# Round 2, GPU call O (1 GPU): dedicated epilogue warp (T3_EPW) -- parity tests, then A/B of the builds: a = EPW with 232 / 40 registers,
# f = lists kept by the math warps (EPW off), g = EPW with 224 / 56 registers.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_headline_shapes.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/r02o_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02o_gpu_tests.log
cp zebra_b200/libzebra_b200.so zebra_b200/variants/libzb_a.so
for v in a f g; do
  cp zebra_b200/variants/libzb_$v.so zebra_b200/libzebra_b200.so
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r02o_${v}_l2.json 2>> $OUT/r02o.err
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --metric cosine > $OUT/r02o_${v}_cos.json 2>> $OUT/r02o.err
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --metric l2sq --dim 384 > $OUT/r02o_${v}_l2sq384.json 2>> $OUT/r02o.err
done
cp zebra_b200/variants/libzb_g.so zebra_b200/libzebra_b200.so
timeout 300 python -m pytest tests/test_gpu_headline_shapes.py -m gpu -q -x > $OUT/r02o_gpu_tests_g.log 2>&1; echo "pytest (g) rc=$?"; tail -2 $OUT/r02o_gpu_tests_g.log
cp zebra_b200/variants/libzb_a.so zebra_b200/libzebra_b200.so
python tools/show_bench.py $OUT/r02o_*.json | grep -v phases
tail -5 $OUT/r02o.err

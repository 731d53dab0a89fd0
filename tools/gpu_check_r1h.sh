mkdir -p gpurun_out
timeout 150 python tools/seq_knob_sweep.py 0 2 4 8 16 > gpurun_out/r01j_seq_prefetch_sweep.txt 2> gpurun_out/r01j.err; echo "sweep rc=$?"
cat gpurun_out/r01j_seq_prefetch_sweep.txt; tail -3 gpurun_out/r01j.err

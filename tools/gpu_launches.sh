#!/bin/bash
# bench + ncu launch list of the same command: two training steps' worth of launches from inside the step loop
# (1497 launches per step at BASELINE config 2; the skip lands in the warm-up steps of the timed loop)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-extras > gpurun_out/plain.log 2>&1 &&
# SRES_CUDA_GRAPHS=0 for the ncu pass only: ncu's per-node graph replay (--graph-profiling node) fails with LaunchFailed on
# the first weight-gradient node of the backward graph whenever that node is inside the profiled window (the same launch
# profiles fine when it is issued eagerly: tools/ncu_wgrad_eager.sh); eager launches are the same kernels with the same
# arguments, and under ncu every kernel runs alone anyway
SRES_CUDA_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 4800 -c 2994 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-extras > gpurun_out/ncu.log 2>&1
tail -n 1 gpurun_out/plain.log | cut -c1-200; tail -n 3 gpurun_out/ncu.log | cut -c1-300; wc -l gpurun_out/launches.csv
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/r02_v4_launches.md; head -34 gpurun_out/r02_v4_launches.md

#!/bin/bash
# ncu --set full capture of the CUDA-core tail-conv backward kernels in a config-5 (x8) training step, one GPU
mkdir -p gpurun_out
python tools/x8_step.py > gpurun_out/plain_x8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv3x3_small_in_kernel|small_out_wgrad_kernel" -c 4 -f -o gpurun_out/prof_tail \
    python tools/x8_step.py > gpurun_out/ncu_tail.log 2>&1
ncu -i gpurun_out/prof_tail.ncu-rep --page details > gpurun_out/prof_tail_details.txt 2>/dev/null
ncu -i gpurun_out/prof_tail.ncu-rep --page source --csv > gpurun_out/prof_tail_source.csv 2>/dev/null
tail -n 2 gpurun_out/plain_x8.log; tail -n 3 gpurun_out/ncu_tail.log; wc -l gpurun_out/prof_tail_details.txt

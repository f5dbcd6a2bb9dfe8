#!/bin/bash
# ncu --set full capture of the split-K weight-gradient kernel (single-kernel command, one GPU)
mkdir -p gpurun_out
python tools/bringup_wgrad.py --B 64 --H 48 --W 48 --iters 4 > gpurun_out/plain_wgrad.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv3x3_wgrad -s 3 -c 2 -f -o gpurun_out/prof_wgrad \
    python tools/bringup_wgrad.py --B 64 --H 48 --W 48 --iters 4 > gpurun_out/ncu_wgrad.log 2>&1
tail -n 2 gpurun_out/plain_wgrad.log gpurun_out/ncu_wgrad.log

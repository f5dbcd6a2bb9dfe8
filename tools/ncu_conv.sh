#!/bin/bash
# ncu --set full capture of the conv kernel, RCAB conv1 flavour (single-kernel command, one GPU)
mkdir -p gpurun_out
python tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 4 --bf16-only --relu > gpurun_out/plain_conv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv3x3_igemm -s 3 -c 2 -f -o gpurun_out/prof_conv \
    python tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 4 --bf16-only --relu > gpurun_out/ncu_conv.log 2>&1
tail -n 3 gpurun_out/plain_conv.log gpurun_out/ncu_conv.log

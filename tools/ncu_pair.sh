#!/bin/bash
# ncu --set full capture of the fused two-convolution kernel (forward pair, B = 64, 48 x 48), one GPU
mkdir -p gpurun_out
python tools/bringup_pair.py > gpurun_out/plain_pair.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair -s 20 -c 2 -f -o gpurun_out/prof_pair \
    python tools/bringup_pair.py > gpurun_out/ncu_pair.log 2>&1
ncu -i gpurun_out/prof_pair.ncu-rep --page raw --csv > gpurun_out/prof_pair_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_pair.ncu-rep --page details > gpurun_out/prof_pair_details.txt 2>/dev/null
python tools/ncu_metrics.py gpurun_out/prof_pair_raw.csv > gpurun_out/r02_pair_ncu_metrics.json
tail -n 3 gpurun_out/plain_pair.log; head -c 1500 gpurun_out/r02_pair_ncu_metrics.json

#!/bin/bash
# ncu --set full capture of the image-resident RCAB chain kernel (B = 64, 48 x 48, 4 blocks per launch), one GPU
mkdir -p gpurun_out
python tools/bringup_chain.py 64 4 > gpurun_out/plain_chain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rcab_chain -s 3 -c 1 -f -o gpurun_out/prof_chain \
    python tools/bringup_chain.py 64 4 > gpurun_out/ncu_chain.log 2>&1
ncu -i gpurun_out/prof_chain.ncu-rep --page raw --csv > gpurun_out/prof_chain_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_chain.ncu-rep --page details > gpurun_out/r02_chain_ncu_details.txt 2>/dev/null
python tools/ncu_metrics.py gpurun_out/prof_chain_raw.csv > gpurun_out/r02_chain_ncu_metrics.json
head -n 2 gpurun_out/plain_chain.log; head -c 1800 gpurun_out/r02_chain_ncu_metrics.json

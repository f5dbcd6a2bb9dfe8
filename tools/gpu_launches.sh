#!/bin/bash
# bench + ncu launch list of the same command
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 9000 -c 3600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -n 2 gpurun_out/plain.log | cut -c1-300; tail -n 3 gpurun_out/ncu.log | cut -c1-300; wc -l gpurun_out/launches.csv

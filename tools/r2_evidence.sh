#!/bin/bash
# Round-2 evidence in one GPU call: default bench line, reference arm, ncu launch list of the bench command,
# ncu --set full of the dominant kernel (fused forward pair).  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_v2_bench.json 2> gpurun_out/r02_v2_bench.err
echo "bench exit=$?"; cut -c1-400 gpurun_out/r02_v2_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_v2_bench_reference.json 2> gpurun_out/r02_v2_bench_reference.err
echo "reference exit=$?"; cut -c1-300 gpurun_out/r02_v2_bench_reference.json
# launch list: two training steps' worth of launches from inside the step loop (1497 launches per step)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-extras > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 4800 -c 2994 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-extras > gpurun_out/ncu.log 2>&1
tail -n 1 gpurun_out/plain.log | cut -c1-200; wc -l gpurun_out/launches.csv
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/r02_v2_launches.md; head -30 gpurun_out/r02_v2_launches.md
bash tools/ncu_pair.sh

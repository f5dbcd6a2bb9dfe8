#!/bin/bash
# Round 2, split-trunk change: full GPU test suite with durations, channel-attention microbench, interleaved A/B of
# SRES_TRUNK_SPLIT on the training-step bench.  Outputs under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -22 gpurun_out/pytest_gpu.log
python tools/bench_ca.py 2>&1 | tee gpurun_out/bench_ca.log
ROUNDS=${ROUNDS:-2} bash tools/r2_ab3.sh "SRES_TRUNK_SPLIT=1" "SRES_TRUNK_SPLIT=0" 2>&1 | tee gpurun_out/ab_split.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-extras 2>/dev/null | tee gpurun_out/bench_split.json | cut -c1-1500

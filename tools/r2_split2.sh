#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "channel_attention" 2>&1 | tail -3
python tools/bench_ca.py 2>&1 | tee gpurun_out/bench_ca.log
ROUNDS=${ROUNDS:-2} bash tools/r2_ab3.sh "SRES_TRUNK_SPLIT=1" "SRES_TRUNK_SPLIT=0" 2>&1 | tee gpurun_out/ab_split.log

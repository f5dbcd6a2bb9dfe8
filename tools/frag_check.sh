#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/frag_check.log
: > $LOG
timeout 300 python tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --mode relu_pool 2>&1 | tail -5 >> $LOG
timeout 300 python tools/bringup_conv.py --B 3 --H 20 --W 24 --iters 3 --mode relu_pool 2>&1 | tail -5 >> $LOG
timeout 300 python tools/bench_conv_flavours.py 2>&1 | tail -14 >> $LOG
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $LOG
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline >> $LOG 2>&1
cat $LOG

#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/bringup5.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 300 python "$@" 2>&1 | tail -8 >> $LOG; echo "exit=${PIPESTATUS[0]}" >> $LOG; }
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --bf16-only
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --mode relu_pool
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --mode resid
run tools/bringup_wgrad.py --B 64 --H 48 --W 48 --iters 30
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $LOG
python bench.py --steps 5 --warmup 3 --no-cpu-baseline >> $LOG 2>&1
cat $LOG

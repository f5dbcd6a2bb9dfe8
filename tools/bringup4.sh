#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/bringup4.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 300 python "$@" 2>&1 | tail -8 >> $LOG; echo "exit=${PIPESTATUS[0]}" >> $LOG; }
run tools/bringup_conv.py --B 3 --H 48 --W 48 --mode relu_pool
run tools/bringup_conv.py --B 2 --H 48 --W 48 --mode dgrad
run tools/bringup_conv.py --B 2 --H 20 --W 24 --mode resid
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --debug-flags 2
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --bf16-only
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --bf16-only --debug-flags 2
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --mode relu_pool
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --mode resid
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --mode resid --debug-flags 2
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $LOG
cat $LOG

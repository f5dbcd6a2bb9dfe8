#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/bringup6.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 300 python "$@" 2>&1 | tail -16 >> $LOG; echo "exit=${PIPESTATUS[0]}" >> $LOG; }
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --bf16-only --timeline
run tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --mode resid --timeline
cat $LOG

#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/bringup2.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 180 python "$@" 2>&1 | tail -16 >> $LOG; echo "exit=${PIPESTATUS[0]}" >> $LOG; }
run tools/bringup_wgrad.py --B 1 --H 48 --W 48
run tools/bringup_wgrad.py --B 3 --H 20 --W 24
run tools/bringup_wgrad.py --B 2 --H 96 --W 96
run tools/bringup_wgrad.py --B 64 --H 48 --W 48 --iters 20
cat $LOG

#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/bringup3.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 600 python "$@" 2>&1 | tail -28 >> $LOG; echo "exit=${PIPESTATUS[0]}" >> $LOG; }
run tools/bringup_net.py --case tiny_x4
run tools/bringup_net.py --case tiny_x4_r16_charb
run tools/bringup_net.py --case tiny_x2_1ch
run tools/bringup_net.py --case tiny_x8_4ch
run tools/bringup_net.py --case tiny_x3
run tools/bringup_net.py --case small_x4
cat $LOG

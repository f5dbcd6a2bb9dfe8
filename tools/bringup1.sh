#!/bin/bash
# first GPU bring-up: conv kernel variants, each isolated in its own process
mkdir -p gpurun_out
LOG=gpurun_out/bringup1.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $LOG 2>&1
run() { echo "=== $*" >> $LOG; timeout 180 python tools/bringup_conv.py "$@" 2>&1 | tail -12 >> $LOG; echo "exit=${PIPESTATUS[0]}" >> $LOG; }
run --B 1 --H 48 --W 48 --debug-flags 0
run --B 1 --H 48 --W 48 --debug-flags 1
run --B 3 --H 48 --W 48 --debug-flags 0 --mode relu_pool
run --B 3 --H 48 --W 48 --debug-flags 1 --mode relu_pool
run --B 2 --H 48 --W 48 --mode dgrad
run --B 2 --H 48 --W 48 --mode resid
run --B 2 --H 24 --W 20 --mode shuffle
run --B 2 --H 96 --W 96 --nout 16
run --B 64 --H 48 --W 48 --iters 20
run --B 64 --H 48 --W 48 --iters 20 --debug-flags 1
tail -60 $LOG

#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/ca_check.log
: > $LOG
timeout 300 python tools/bench_ca.py 2>&1 | tail -3 >> $LOG
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $LOG
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-220 >> $LOG
cat $LOG

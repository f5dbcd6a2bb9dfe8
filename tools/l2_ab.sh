#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/l2_ab.log
: > $LOG
for l in 0 48 64 96; do
  echo "=== l2_persist_MB=$l" >> $LOG
  SRES_L2_PERSIST=$l python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | grep -oE '"value": [0-9.]*, "unit": "tiles/s", "n_gpus": 1, "steps": 8, "warmup": 3, "ms_per_step": [0-9.]*|Error.*|error.*' | head -3 >> $LOG
done
cat $LOG

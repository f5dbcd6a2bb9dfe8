#!/bin/bash
mkdir -p gpurun_out; LOG=gpurun_out/side_ab.log; : > $LOG
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $LOG
for s in 1 0; do
  echo "=== side_stream=$s" >> $LOG
  SRES_SIDE_STREAM=$s python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | grep -oE '"value": [0-9.]*, "unit": "tiles/s", "n_gpus": 1, "steps": 8, "warmup": 3, "ms_per_step": [0-9.]*|Error.*|error.*' | head -3 >> $LOG
done
cat $LOG

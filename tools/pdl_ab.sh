#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/pdl_ab.log
: > $LOG
for g in 1 0; do for p in 0 1 2; do
  echo "=== graphs=$g pdl=$p" >> $LOG
  SRES_CUDA_GRAPHS=$g SRES_PDL=$p python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"value": [0-9.]*, "unit": "tiles/s", "n_gpus": 1, "steps": 8, "warmup": 3, "ms_per_step": [0-9.]*' >> $LOG
done; done
cat $LOG

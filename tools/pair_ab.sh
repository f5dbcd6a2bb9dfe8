#!/bin/bash
# A/B of the CTA-pair (cta_group::2) convolution kernel against the single-CTA one
mkdir -p gpurun_out
LOG=gpurun_out/pair_ab.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 300 "$@" 2>&1 | tail -8 >> $LOG; echo "exit=${PIPESTATUS[0]}" >> $LOG; }
for pair in 1 0; do
  export SRES_CONV_PAIR=$pair
  echo "##### SRES_CONV_PAIR=$pair" >> $LOG
  run python tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --bf16-only
  run python tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --mode relu_pool
  run python tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --mode resid
  run python tools/bringup_conv.py --B 3 --H 20 --W 24 --iters 3 --mode resid
done
export SRES_CONV_PAIR=1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $LOG
for pair in 1 0; do
  echo "##### bench SRES_CONV_PAIR=$pair" >> $LOG
  SRES_CONV_PAIR=$pair timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline >> $LOG 2>&1
done
cat $LOG

#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/pair_tl.log
: > $LOG
for pair in 0; do
  echo "##### SRES_CONV_PAIR=$pair" >> $LOG
  SRES_CONV_PAIR=$pair timeout 300 python tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --bf16-only --timeline 2>&1 | tail -17 >> $LOG
  SRES_CONV_PAIR=$pair timeout 300 python tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --mode relu_pool --timeline 2>&1 | tail -17 >> $LOG
  SRES_CONV_PAIR=$pair timeout 300 python tools/bringup_conv.py --B 64 --H 48 --W 48 --iters 30 --mode resid --timeline 2>&1 | tail -17 >> $LOG
  SRES_CONV_PAIR=$pair timeout 300 python tools/bench_conv_flavours.py 2>&1 | tail -8 >> $LOG
done
cat $LOG

#!/bin/bash
mkdir -p gpurun_out; LOG=gpurun_out/ca_ab.log; : > $LOG
for b in 0 5 10 20 40 76; do SRES_CA_BPI=$b python tools/bench_ca.py >> $LOG 2>&1; done
cat $LOG

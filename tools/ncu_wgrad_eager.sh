#!/bin/bash
# does the first weight-gradient batch of an x4 backward (4 jobs at 96 x 96) survive ncu?  eager launches
mkdir -p gpurun_out
for v in "X=1" "SRES_SIDE_STREAM=0" "SRES_PDL=0 SRES_SIDE_STREAM=0"; do
  echo "== $v"
  env $v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv3x3_wgrad" -c 3 --csv --log-file gpurun_out/wg_eager.csv python tools/x4_step.py 64 2 > gpurun_out/wg_eager.log 2>&1
  echo "exit=$?"; grep -E "ERROR|wgrad_kernel" gpurun_out/wg_eager.csv | cut -c1-40,290-420 | head -6; tail -n 2 gpurun_out/wg_eager.log | cut -c1-200
done
echo "== sanitizer (memcheck), 1 group x 1 block, batch 2"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/x4_step.py 2 1 2>&1 | tail -15

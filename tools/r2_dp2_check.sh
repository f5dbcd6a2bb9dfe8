#!/bin/bash
# DP-2 correctness check (tools/dp_check.py) with the default bucket policy, then the 2-GPU bench line beside one GPU
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dp_check.py 2>&1 | grep -E "RESULT|Error|error|Traceback" | cut -c1-260 | tee gpurun_out/dp2_check.log
bash tools/r2_dp_buckets.sh 2 2 0

#!/bin/bash
# Data-parallel gradient all-reduce bucket policy (SRES_DP_BUCKETS: 0 = one bucket per backward segment, k = k buckets) on N GPUs
# beside the 1-GPU step of the same box:  bash tools/r2_dp_buckets.sh [N] [bucket settings...]
N=${1:-8}; shift
SET=${@:-"0 1 2"}
mkdir -p gpurun_out
OUT=gpurun_out/dp_buckets_$N.log
: > $OUT
python bench.py --gpus 1 --steps 12 --warmup 4 --no-cpu-baseline --skip-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1 GPU: ms/step %.3f  tiles/s %.1f' % (d['ms_per_step'], d['value']))" | tee -a $OUT
port=29520
for b in $SET; do
  port=$((port+1))
  SRES_DP_BUCKETS=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 12 --warmup 4 --no-cpu-baseline --skip-extras 2>/dev/null | grep '"metric"' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$N GPUs, SRES_DP_BUCKETS=$b: ms/step %.3f  tiles/s %.1f  e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))" | tee -a $OUT
done

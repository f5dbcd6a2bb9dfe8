#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
LOG=gpurun_out/dp$N.log
: > $LOG
nvidia-smi -L >> $LOG
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py >> $LOG 2>&1; echo "dp_check exit=$?" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 >> $LOG 2>&1; echo "bench exit=$?" >> $LOG
grep -E "RESULT|exit=|value|Error|error" $LOG | cut -c1-300
grep '"metric"' $LOG > gpurun_out/bench_dp$N.json

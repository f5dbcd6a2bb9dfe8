#!/bin/bash
# N-GPU data-parallel check (tools/dp_check.py) and bench beside the 1-GPU bench on the same box: bash tools/r2_dp8.sh [N]
N=${1:-8}
mkdir -p gpurun_out
LOG=gpurun_out/r2_dp$N.log
: > $LOG
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 tools/dp_check.py >> $LOG 2>&1; echo "dp_check exit=$?" >> $LOG
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --skip-extras 2>/dev/null | tail -1 > gpurun_out/r2_bench_dp1.json
run 29512 bench.py --gpus $N --steps 20 --warmup 5 2>/dev/null | grep '"metric"' > gpurun_out/r2_bench_dp$N.json
grep -E "RESULT|exit=" $LOG | cut -c1-250
for f in gpurun_out/r2_bench_dp1.json gpurun_out/r2_bench_dp$N.json; do
  python -c "import json,sys; d=json.load(open('$f')); print('$f', 'ms/step %.3f' % d['ms_per_step'], 'tiles/s %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], {k: (round(d[k]['value'],1) if k in d else None) for k in ('infer_region','x8')})"
done

#!/bin/bash
# A/B of an environment switch on the training-step bench: bash tools/r2_ab.sh VAR v1 v2 ...
VAR=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $VAR=$v timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --skip-extras 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$VAR=$v', 'ms/step %.3f' % d['ms_per_step'], 'tiles/s %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'conv us %.2f' % d['roofline']['us_per_launch'], 'infer MP/s %.1f' % d['inference']['value'])"
done

#!/bin/bash
# A/B of environment settings on the training-step bench: bash tools/r2_ab2.sh "VAR=a VAR2=b" "VAR=c" ...
for setting in "$@"; do
  env $setting timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --skip-extras 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$setting', 'ms/step %.3f' % d['ms_per_step'], 'tiles/s %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'infer MP/s %.1f' % d['inference']['value'])"
done

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "pair or flavour or conv" 2>&1 | tail -3 | tee gpurun_out/pair_test.log
timeout 300 python tools/bringup_pair_bwd.py 2>&1 | head -3 | tee gpurun_out/bringup_pair_bwd2.log
timeout 300 python tools/bringup_pair.py 2>&1 | head -2
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'infer MP/s', round(d['inference']['value'],1), 'frac', round(d['roofline']['frac'],4))"; done

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "train_stream or controller_api" 2>&1 | tail -8
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train ms', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ratio', round(d['e2e']['value']/d['value'],4))"; done

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q -k "rcab_chain" 2>&1 | tail -5 | tee gpurun_out/chain_test.log
timeout 300 python tools/bringup_chain.py 2>&1 | tee gpurun_out/bringup_chain.log
if grep -q "failed\|error" gpurun_out/chain_test.log; then exit 0; fi
for m in 1 0; do SRES_RCAB_CHAIN=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chain=$m', 'train ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'infer MP/s', round(d['inference']['value'],1), 'launches/step', d['gpu_launches']//d['steps'])"; done 2>&1 | tee gpurun_out/bench_chain.log

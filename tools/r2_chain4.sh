#!/bin/bash
# In-situ A/B of the chain kernel's variants (training step + inference of bench.py), interleaved rounds.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q -k "rcab_chain" 2>&1 | tail -3 | tee gpurun_out/chain_test.log
for round in 1 2; do
for m in "SRES_RCAB_CHAIN=0" "SRES_RCAB_CHAIN=1" "SRES_RCAB_CHAIN=1 SRES_CHAIN_LEND=1" "SRES_RCAB_CHAIN=1 SRES_CHAIN_BULK=1" "SRES_RCAB_CHAIN=1 SRES_CHAIN_LEND=1 SRES_CHAIN_STAGGER=30000"; do env $m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m', 'train ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'infer MP/s', round(d['inference']['value'],1), 'launches/step', d['gpu_launches']//d['steps'])"; done
done 2>&1 | tee gpurun_out/bench_chain4.log

#!/bin/bash
# Overlapped chain variant: parity tests under SRES_CHAIN_OVERLAP=1, its timeline, then the in-network A/B
mkdir -p gpurun_out
SRES_CHAIN_OVERLAP=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q -k "rcab_chain and not variants" 2>&1 | tail -4 | tee gpurun_out/chain_ovl_test.log
if grep -q "failed\|error" gpurun_out/chain_ovl_test.log; then exit 0; fi
SRES_CHAIN_OVERLAP=1 timeout 300 python tools/bringup_chain.py 2>&1 | tee gpurun_out/bringup_chain_ovl.log
for round in ${ROUNDS:-1}; do
for m in "SRES_RCAB_CHAIN=0" "SRES_RCAB_CHAIN=1 SRES_CHAIN_OVERLAP=1" "SRES_RCAB_CHAIN=1"; do env $m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m', 'train ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'infer MP/s', round(d['inference']['value'],1), 'launches/step', d['gpu_launches']//d['steps'])"; done
done 2>&1 | tee gpurun_out/bench_chain5.log

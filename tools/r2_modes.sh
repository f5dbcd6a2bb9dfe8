#!/bin/bash
# full GPU test suite, then the training-step bench with the three conv-kernel modes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.log
echo "pytest exit=${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_gpu.log
for m in 0 1 2; do
  SRES_CONV_N192=$m timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_mode$m.json 2> gpurun_out/r2_bench_mode$m.err
  echo "mode $m exit=$?" >> gpurun_out/r2_bench_mode$m.err
done
cat gpurun_out/r2_pytest_gpu.log; for m in 0 1 2; do cat gpurun_out/r2_bench_mode$m.json; tail -2 gpurun_out/r2_bench_mode$m.err; done

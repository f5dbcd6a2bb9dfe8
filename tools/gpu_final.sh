#!/bin/bash
# round-end style check: GPU tests, smoke, bench (both arms), ncu launch list of the bench command
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/bench.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 9000 -c 3400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cat gpurun_out/bench.log | cut -c1-400; tail -n 2 gpurun_out/bench.err

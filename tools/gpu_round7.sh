#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all, new defaults)"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== kernel variants"; timeout 900 python tools/bench_kernels.py 1.0 > gpurun_out/bench_kernels.log 2>&1; echo "rc=$?"; cat gpurun_out/bench_kernels.log | tail -28
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r7.json 2> gpurun_out/bench_r7.err; echo "rc=$?"; cat gpurun_out/bench_r7.json; tail -3 gpurun_out/bench_r7.err

#!/bin/bash
# First GPU session: smoke, parity suite, bench, launch list. Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench small"; timeout 600 python bench.py --steps 3 --warmup 3 --scale 0.1 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?"; cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err

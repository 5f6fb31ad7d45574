#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu all"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== rmat"; timeout 600 python tools/bench_rmat.py > gpurun_out/bench_rmat.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench_rmat.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r16.json 2> gpurun_out/bench_r16.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r16.json')); print(d['value'], d['ms_per_step'], d['breakdown_ms'])"; tail -3 gpurun_out/bench_r16.err

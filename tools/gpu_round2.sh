#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu1.err; echo "rc=$?"
echo "== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gat_fast_kernel|spmm_kernel" -s 8 -c 2 -o gpurun_out/prof_r1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu2.json 2> gpurun_out/ncu2.err; echo "rc=$?"; tail -3 gpurun_out/ncu2.err
ls -la gpurun_out

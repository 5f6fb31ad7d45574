#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r11.json 2> gpurun_out/bench_r11.err; echo "rc=$?"; cat gpurun_out/bench_r11.json | cut -c1-400; python -c "
import json; d=json.load(open('gpurun_out/bench_r11.json')); print(d['value'], d['ms_per_step'], d['breakdown_ms'], d['roofline']['frac'], d['roofline']['secondary']['frac'], d['e2e'], d['clocks'])"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/ncu1.err; echo "rc=$?"
echo "== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gat_async_kernel|spmm_async_kernel|gemm_tf32x3_kernel" -s 24 -c 6 -o gpurun_out/prof_r1_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/ncu2.err; echo "rc=$?"; tail -3 gpurun_out/ncu2.err

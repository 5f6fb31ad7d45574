#!/bin/bash
# Full single-GPU validation on a B200 box (run through gpurun): smoke, parity suite, bench (+ reference arm), launch list and
# ncu --set full captures of the three hot kernels.  Outputs under gpurun_out/; copy what should be judged into profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
if [ "$1" == "--ncu" ]; then
  echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/ncu1.err; echo "rc=$?"
  echo "== ncu --set full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gat_async_kernel|spmm_async_kernel|gemm_tf32x3_ws_kernel" -s 24 -c 6 -o gpurun_out/prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/ncu2.err; echo "rc=$?"
fi

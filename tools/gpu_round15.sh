#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu all"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== rmat"; timeout 600 python tools/bench_rmat.py > gpurun_out/bench_rmat.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench_rmat.log
echo "== ncu gemm ws"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tf32x3_ws_kernel" -s 8 -c 1 -o gpurun_out/prof_r1_gemm_ws -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/ncu3.err; echo "rc=$?"
ncu -i gpurun_out/prof_r1_gemm_ws.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/prof_r1_gemm_ws_raw.csv; ls -la gpurun_out/prof_r1_gemm_ws_raw.csv

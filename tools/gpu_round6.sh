#!/bin/bash
mkdir -p gpurun_out
echo "== kernel variants"; timeout 900 python tools/bench_kernels.py 1.0 > gpurun_out/bench_kernels.log 2>&1; echo "rc=$?"; cat gpurun_out/bench_kernels.log | grep -v gat_ | tail -24

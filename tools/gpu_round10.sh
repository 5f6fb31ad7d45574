#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tc tests"; TFGK_GEMM_TC=1 timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q --timeout=120 > gpurun_out/pytest_gemm_tc.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gemm_tc.log
echo "== models/golden with tc"; TFGK_GEMM_TC=1 timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_golden.py tests/test_gpu_gat.py -m gpu -q --timeout=300 > gpurun_out/pytest_tc_models.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_tc_models.log
echo "== kernel variants"; timeout 900 python tools/bench_kernels.py 1.0 > gpurun_out/bench_kernels.log 2>&1; echo "rc=$?"; grep -E "gemm|tc vs" gpurun_out/bench_kernels.log

#!/bin/bash
mkdir -p gpurun_out
echo "== spmm tests with stream impl"; TFGK_SPMM_IMPL=async timeout 600 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_full_size.py tests/test_gpu_models.py -m gpu -q -x --timeout=300 > gpurun_out/pytest_stream.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_stream.log
echo "== kernel variants"; timeout 900 python tools/bench_kernels.py 1.0 > gpurun_out/bench_kernels.log 2>&1; echo "rc=$?"; cat gpurun_out/bench_kernels.log | tail -24

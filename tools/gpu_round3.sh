#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tc tests"; timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q -x --timeout=120 > gpurun_out/pytest_gemm_tc.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gemm_tc.log
echo "== pytest gpu (all, new defaults)"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --deselect tests/test_gpu_gemm_tc.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== pytest with tc gemm off"; TFGK_GEMM_TC=0 timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_golden.py -m gpu -q --timeout=600 > gpurun_out/pytest_gpu_notc.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu_notc.log
echo "== kernel variants"; timeout 900 python tools/bench_kernels.py 1.0 > gpurun_out/bench_kernels.log 2>&1; echo "rc=$?"; cat gpurun_out/bench_kernels.log | tail -20
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r3.json 2> gpurun_out/bench_r3.err; echo "rc=$?"; cat gpurun_out/bench_r3.json; tail -3 gpurun_out/bench_r3.err

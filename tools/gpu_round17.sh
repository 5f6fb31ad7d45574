#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_gat.py tests/test_gpu_index.py -m gpu -q --timeout=300 > gpurun_out/pytest_hub.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_hub.log

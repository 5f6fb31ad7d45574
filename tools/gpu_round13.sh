#!/bin/bash
mkdir -p gpurun_out
echo "== sage bwd + full size"; timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_full_size.py -m gpu -q -s --timeout=600 > gpurun_out/pytest_sage.log 2>&1; echo "rc=$?"; grep -E "GraphSAGE|passed|failed|Error" gpurun_out/pytest_sage.log | head
echo "== dist N=2 overlapped"; bash tools/gpu_multi.sh 2 2>&1 | tail -12 | cut -c1-900

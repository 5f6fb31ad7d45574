#!/bin/bash
# GPU check of the training extras / samplers: full GPU suite, then compute-sanitizer memcheck over the new kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_train.py -q -m gpu -x \
    -k "spmm_heads or softmax_bwd or drop_edge or sampler or dropout_mask" > gpurun_out/train_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/train_memcheck.log
tail -8 gpurun_out/train_memcheck.log

#!/bin/bash
# Last GPU call of the round: parity suite on the frozen tree, then an ncu --set full capture of the two GAT backward kernels.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"spmm_heads128_kernel|gat_softmax_bwd128_kernel" -c 4 \
    -o gpurun_out/prof_train -f python tools/bench_train.py --steps 1 > gpurun_out/ncu_train.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_train.log
ls -la gpurun_out/prof_train.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_cfg5.py 1.0 > gpurun_out/cfg5_r2.log 2>&1; echo "rc=$?"; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/cfg5_r2.log | tail -6

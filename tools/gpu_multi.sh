#!/bin/bash
# usage: gpu_multi.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== dist check N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "rc=$?"; grep -E "rank|Error|error" gpurun_out/dist_check_$N.log | head -12
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err

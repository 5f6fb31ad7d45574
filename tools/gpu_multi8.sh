#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$? lines=$(wc -l < gpurun_out/bench_n8.json)"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n8.json').read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'], d['roofline']['frac'])"; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n8.err | tail -3

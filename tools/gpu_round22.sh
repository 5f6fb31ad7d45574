#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench N=2 stdout"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$? lines=$(wc -l < gpurun_out/bench_n2.json)"; head -c 300 gpurun_out/bench_n2.json; echo; python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "rc=$? lines=$(wc -l < gpurun_out/bench_ref_n2.json)"; head -c 200 gpurun_out/bench_ref_n2.json; echo

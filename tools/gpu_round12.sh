#!/bin/bash
mkdir -p gpurun_out
echo "== bench N=1 (pipelined e2e)"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r12.json 2> gpurun_out/bench_r12.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r12.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['cpu_baseline'])"; tail -3 gpurun_out/bench_r12.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_ref.json; python -c "
import json; d=json.load(open('gpurun_out/bench_ref.json')); print(d['value'], d['ms_per_step'], d['cpu_baseline'])"
nproc; free -g | head -2

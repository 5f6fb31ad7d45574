#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tc tests (warp specialised)"; timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q --timeout=120 > gpurun_out/pytest_gemm_tc.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gemm_tc.log
echo "== pytest gpu all"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== gemm timing"; timeout 600 python - <<'PY' 2>&1 | tail -8
import os, sys, torch
sys.path.insert(0, '.')
import bench as B
from tf_geometric_b200 import ops
dev = torch.device('cuda'); n = B.PRODUCTS_NODES
gen = torch.Generator(device=dev); gen.manual_seed(1)
x = torch.randn((n, 100), generator=gen, device=dev); w = B.glorot((100, 128), 2).to(dev); out = torch.empty((n, 128), device=dev)
def timed(label):
    for _ in range(3): ops.gemm(x, w, out=out)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ops.gemm(x, w, out=out)
    b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b) / 10
    print(label, "%.3f ms  %.0f GB/s" % (ms, 4 * (n * 100 + n * 128) / ms / 1e6), flush=True)
    return out.clone()
ref = None
for impl, tc in (("ws", "1"), ("sync", "1"), ("simt", "0")):
    os.environ["TFGK_GEMM_TC"] = tc; os.environ["TFGK_GEMM_TC_IMPL"] = impl
    o = timed(impl)
    if ref is None: ref = o
    else: print("   max rel diff vs ws:", float((o - ref).abs().max() / ref.abs().max()))
PY
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r14.json 2> gpurun_out/bench_r14.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r14.json')); print(d['value'], d['ms_per_step'], d['breakdown_ms'], d['e2e']['ms_per_step'], d['clocks'])"; tail -3 gpurun_out/bench_r14.err

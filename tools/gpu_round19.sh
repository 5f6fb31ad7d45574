#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu all"; timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_index.py -m gpu -q --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== big to_directed"; timeout 300 python - <<'PY'
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import tf_geometric_b200 as tfg
sys.path.insert(0, 'tests')
from oracle import tfg_oracle as o
dev = torch.device('cuda')
gen = torch.Generator(device=dev); gen.manual_seed(0)
n, e = 2449029, 61859140
ei = torch.randint(0, n, (2, e), generator=gen, device=dev, dtype=torch.int32)
w = torch.rand((e,), generator=gen, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
d_i, (d_w,) = tfg.utils.convert_edge_to_directed(ei, [w])
torch.cuda.synchronize(); print("to_directed of %d undirected pairs on device: %.1f ms -> %d directed edges" % (e, (time.perf_counter()-t0)*1e3, d_i.shape[1]))
# spot check against the numpy oracle on a 2M-edge prefix
k = 2000000
want_i, (want_w,) = o.convert_edge_to_directed(ei[:, :k].cpu().numpy(), [w[:k].cpu().numpy()])
got_i, (got_w,) = tfg.utils.convert_edge_to_directed(ei[:, :k].contiguous(), [w[:k].contiguous()])
print("prefix check:", np.array_equal(got_i.cpu().numpy(), want_i), np.array_equal(got_w.cpu().numpy(), want_w))
PY

#!/usr/bin/env python
# coding=utf-8
"""Kernel-variant micro-benchmark at the bench workload's shape (ogbn-products size): times each hot kernel alone with
CUDA events (inputs >> L2), for every implementation variant selectable by environment variable.  Development tool:
the numbers that count are bench.py's."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from tf_geometric_b200 import ops, _structure  # noqa: E402
import tf_geometric_b200 as tfg  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
iters = 5
dev = torch.device("cuda")
n = int(B.PRODUCTS_NODES * scale)
pairs = int(B.PRODUCTS_UNDIRECTED * scale)
ei = B.make_graph_device(n, pairs, 0, dev)
E = ei.shape[1]
gen = torch.Generator(device=dev); gen.manual_seed(1)
x = torch.randn((n, B.FEATURES), generator=gen, device=dev)
h = torch.randn((n, B.UNITS), generator=gen, device=dev)
csr, _ = _structure.csr_for_edge_index(ei, n, add_self_loop=True)
w = torch.rand((csr.nnz,), generator=gen, device=dev)
peak, _ = B.measured_peak_gbs()
results = {}


def timed(fn, label, nbytes=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    rec = {"ms": ms}
    if nbytes:
        rec["GBps"] = nbytes / ms / 1e6
        rec["frac_of_measured_peak"] = rec["GBps"] / peak
    results[label] = rec
    print(label, json.dumps(rec), flush=True)


D = B.UNITS
spmm_bytes = csr.nnz * (4 * D + 8) + n * (4 * D + 8)
out = torch.empty((n, D), device=dev)
ref_out = None
variants = [("ldg", ""), ("async", "4x3"), ("async", "4x4"), ("async", "8x3"), ("gather4", "2"), ("gather4", "3"),
            ("gather4", "4"), ("gather4", "6"), ("gather4", "8")]
if os.environ.get("TFGK_BENCH_QUICK"):
    variants = [("async", "4x3"), ("gather4", "3"), ("gather4", "4")]
for impl, cfg in variants:
    os.environ["TFGK_SPMM_IMPL"] = impl
    os.environ["TFGK_SPMM_ASYNC_CFG"] = cfg
    os.environ["TFGK_SPMM_GATHER4_STAGES"] = cfg if impl == "gather4" else "4"
    tag = impl + ("_" + cfg if cfg else "")
    timed(lambda: ops.spmm(csr, w, h, out=out), "spmm_D128_" + tag, spmm_bytes)
    if ref_out is None:
        ref_out = out.clone()
    else:
        assert torch.equal(out, ref_out), "variant {} changed the bits".format(tag)
    timed(lambda: ops.spmm(csr, None, x, reduce="mean"), "spmm_mean_D100_" + tag,
          csr.nnz * (4 * 100 + 4) + n * (4 * 100 + 8))
os.environ.pop("TFGK_SPMM_IMPL")
os.environ.pop("TFGK_SPMM_ASYNC_CFG")
os.environ.pop("TFGK_SPMM_GATHER4_STAGES")
if os.environ.get("TFGK_BENCH_QUICK"):
    q = torch.randn((n, D), generator=gen, device=dev)
    kv = torch.randn((n, 2 * D), generator=gen, device=dev)
    gat_bytes = csr.nnz * (8 * D + 4) + n * (8 * D + 8)
    ref = None
    for impl in ("async", "gather4:2", "gather4:3", "gather4:4"):
        os.environ["TFGK_GAT_IMPL"] = impl
        timed(lambda: ops.gat_fused(csr, q, kv[:, :D], kv[:, D:], B.HEADS), "gat_" + impl.replace(":", "_").replace("async", "async_2x3"), gat_bytes)
        got = ops.gat_fused(csr, q, kv[:, :D], kv[:, D:], B.HEADS)
        if ref is None:
            ref = got.clone()
        else:
            assert torch.equal(got, ref), "GAT variant {} changed the bits".format(impl)
    os.environ.pop("TFGK_GAT_IMPL", None)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "bench_kernels.json"), "w"), indent=1)
    sys.exit(0)

q = torch.randn((n, D), generator=gen, device=dev)
kv = torch.randn((n, 2 * D), generator=gen, device=dev)
k_sep, v_sep = kv[:, :D].contiguous(), kv[:, D:].contiguous()
gat_bytes = csr.nnz * (8 * D + 4) + n * (8 * D + 8)
att = torch.empty((csr.nnz, B.HEADS), device=dev)
os.environ["TFGK_GAT_IMPL"] = "online"
timed(lambda: ops.gat_fused(csr, q, kv[:, :D], kv[:, D:], B.HEADS), "gat_online_ldg_interleaved", gat_bytes)
ref_gat = ops.gat_fused(csr, q, kv[:, :D], kv[:, D:], B.HEADS).clone()
os.environ.pop("TFGK_GAT_IMPL")
for cfg in ("2x3",):
    os.environ["TFGK_GAT_ASYNC_CFG"] = cfg
    timed(lambda: ops.gat_fused(csr, q, kv[:, :D], kv[:, D:], B.HEADS), "gat_async_interleaved_" + cfg, gat_bytes)
    got = ops.gat_fused(csr, q, kv[:, :D], kv[:, D:], B.HEADS)
    err = float((got - ref_gat).abs().max() / ref_gat.abs().max())
    print("   max rel diff vs online ldg:", err, flush=True)
    assert err < 1e-5
os.environ["TFGK_GAT_ASYNC_CFG"] = "2x3"
timed(lambda: ops.gat_fused(csr, q, k_sep, v_sep, B.HEADS), "gat_async_separate_2x3", gat_bytes)
os.environ.pop("TFGK_GAT_ASYNC_CFG")

wmat = B.glorot((B.FEATURES, B.UNITS), 2).to(dev)
gemm_bytes = 4 * (n * B.FEATURES + B.FEATURES * B.UNITS + n * B.UNITS)
for tc in ("1", "0"):
    os.environ["TFGK_GEMM_TC"] = tc
    timed(lambda: ops.gemm(x, wmat, out=out), "gemm_100x128_" + ("tc" if tc == "1" else "simt"), gemm_bytes)
    if tc == "1":
        tc_out = out.clone()
    else:
        print("   tc vs simt max rel diff:", float((tc_out - out).abs().max() / out.abs().max()), flush=True)
os.environ.pop("TFGK_GEMM_TC")
json.dump(results, open(os.path.join(ROOT, "gpurun_out", "bench_kernels.json"), "w"), indent=1)

#!/usr/bin/env python
# coding=utf-8
"""Timing of the device samplers / pooling selection on the synthetic ogbn-products shape.
Wall clock around synchronised calls (the operators contain their own host synchronisation for the output sizes).
The CPU column is the numpy restatement of the reference's per-node loop (oracle/, utils/graph_utils.py:669-776) on a
bounded slice of the same graph, scaled per row.
    python tools/bench_sampler.py [--scale 1.0]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import tf_geometric_b200 as tfg  # noqa: E402
from oracle import tfg_oracle as o  # noqa: E402


def timed(fn, repeat=3):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(repeat):
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best * 1e3, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    n, pairs = int(bench.PRODUCTS_NODES * args.scale), int(bench.PRODUCTS_UNDIRECTED * args.scale)
    ei = bench.make_graph_device(n, pairs, 0, device)
    E = int(ei.shape[1])
    w = torch.rand((E,), device=device)
    res = {"nodes": n, "edges": E}

    t0 = time.perf_counter()
    sampler = tfg.utils.RandomNeighborSampler(ei, w)
    sampler._structure()
    torch.cuda.synchronize()
    res["random_sampler_build_ms"] = (time.perf_counter() - t0) * 1e3
    for tag, kw in (("k10", {"k": 10}), ("k25", {"k": 25}), ("k10_padding", {"k": 10, "padding": True}), ("ratio_0.2", {"ratio": 0.2})):
        ms, (si, _) = timed(lambda: sampler.sample(seed=1, **kw))
        res["random_sampler_" + tag] = {"ms": ms, "sampled_edges": int(si.shape[1]), "rows_per_s": n / ms * 1e3}
    subset = torch.randperm(n, device=device)[: n // 10].to(torch.int32)
    ms, (si, _) = timed(lambda: sampler.sample(k=10, sampled_node_index=subset, seed=1))
    res["random_sampler_k10_subset10pct"] = {"ms": ms, "sampled_edges": int(si.shape[1])}

    uni = tfg.utils.UniformNeighborSampler(ei, w)
    ms, (si, _) = timed(lambda: uni.sample(0.1, seed=2))
    res["uniform_sampler_p0.1"] = {"ms": ms, "sampled_edges": int(si.shape[1]), "edges_per_s": E / ms * 1e3}
    ms, out = timed(lambda: tfg.nn.drop_edge([ei, w], 0.5, training=True, seed=3))
    res["drop_edge_0.5"] = {"ms": ms, "kept_edges": int(out[0].shape[1]), "edges_per_s": E / ms * 1e3}
    ms, out = timed(lambda: tfg.nn.drop_edge([ei, w], 0.5, force_undirected=True, training=True, seed=3))
    res["drop_edge_0.5_undirected"] = {"ms": ms, "kept_edges": int(out[0].shape[1])}

    graphs = max(n // 250, 1)
    gi = torch.sort(torch.randint(0, graphs, (n,), device=device, dtype=torch.int32)).values
    gi[-1] = graphs - 1
    score = torch.randn((n,), device=device)
    ms, idx = timed(lambda: tfg.nn.topk_pool(gi, score, ratio=0.5))
    res["topk_pool_ratio0.5"] = {"ms": ms, "graphs": graphs, "selected": int(idx.numel()), "nodes_per_s": n / ms * 1e3}
    x = torch.randn((n, 128), device=device)
    layer = tfg.layers.Set2Set(num_iterations=3)
    ms, out = timed(lambda: layer([x, gi]))
    res["set2set_d128_it3"] = {"ms": ms, "graphs": graphs}
    ms, _ = timed(lambda: tfg.nn.set2set(x, gi, lambda h, initial_state=None, training=None: (h[:, :, :128].contiguous(),
                                                                                               initial_state[0], initial_state[1]), 3))
    res["set2set_attention_only_d128_it3"] = {"ms": ms, "algorithmic_gb": 3 * 2 * n * 128 * 4 / 1e9}

    # CPU: the reference-style per-node loop on the first rows of the same graph
    rows_cpu = 20000
    ei_h = ei[:, ei[0] < rows_cpu].cpu().numpy()
    t = time.perf_counter()
    o.random_neighbor_sample(ei_h, None, k=10, seed=1)
    cpu_s = time.perf_counter() - t
    res["cpu_loop_k10"] = {"rows": rows_cpu, "s": cpu_s, "rows_per_s": rows_cpu / cpu_s}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

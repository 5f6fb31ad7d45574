#!/usr/bin/env python
# coding=utf-8
"""Training-step timing (forward + backward, device resident) of the bench layers on the synthetic ogbn-products shape:
GCN(128, relu) and 8-head GAT(128, relu) with trainable weights, loss = sum of the outputs.  Not the headline metric
(bench.py measures the forward hot path); this records where the backward kernels stand.
    python tools/bench_train.py [--scale 1.0] [--steps 5] [--drop 0.0]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import tf_geometric_b200 as tfg  # noqa: E402
from tf_geometric_b200 import _ffi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--drop", type=float, default=0.0)
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    n, pairs = int(bench.PRODUCTS_NODES * args.scale), int(bench.PRODUCTS_UNDIRECTED * args.scale)
    edge_index = bench.make_graph_device(n, pairs, 0, device)
    gen = torch.Generator(device="cpu"); gen.manual_seed(1)
    x = torch.randn((n, bench.FEATURES), generator=gen, dtype=torch.float32).to(device)
    graph = tfg.Graph(x, edge_index)
    layers = {"gcn": tfg.layers.GCN(bench.UNITS, activation=tfg.nn.relu, seed=2, trainable=True, edge_drop_rate=args.drop),
              "gat": tfg.layers.GAT(bench.UNITS, num_heads=bench.HEADS, activation=tfg.nn.relu, seed=3, trainable=True,
                                    edge_drop_rate=args.drop)}
    layers["gcn"].build_cache_for_graph(graph)
    inputs = {"gcn": [graph.x, graph.edge_index, graph.edge_weight], "gat": [graph.x, graph.edge_index]}
    result = {"nodes": n, "edges": int(edge_index.shape[1]), "drop": args.drop, "steps": args.steps}
    for name, layer in layers.items():
        def step():
            for p in layer.parameters():
                p.grad = None
            layer(inputs[name], cache=graph.cache, training=True).sum().backward()
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        trace = _ffi.CallTrace(timed=("tfgk_gat_fused_f32", "tfgk_spmm_f32", "tfgk_gemm_f32", "tfgk_spmm_heads_f32",
                                      "tfgk_gat_softmax_bwd_f32", "tfgk_dropout_f32", "tfgk_gat_fused_stats_f32",
                                      "tfgk_gat_bwd_prepare_f32", "tfgk_gat_bwd_dst_f32", "tfgk_gat_bwd_src_f32",
                                      "tfgk_gemm_proj_f32", "tfgk_colsum_f32"))
        _ffi.set_trace(trace)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(args.steps):
            step()
        ev[1].record()
        torch.cuda.synchronize()
        _ffi.set_trace(None)
        result[name] = {"ms_per_step": ev[0].elapsed_time(ev[1]) / args.steps,
                        "kernels_ms_per_step": {k: sum(trace.elapsed_ms(k)) / args.steps for k in sorted(trace.timed)
                                                if trace.events[k]},
                        "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        with torch.no_grad():
            fw = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            layer(inputs[name], cache=graph.cache)
            fw[0].record()
            for _ in range(args.steps):
                layer(inputs[name], cache=graph.cache)
            fw[1].record()
            torch.cuda.synchronize()
            result[name]["inference_forward_ms"] = fw[0].elapsed_time(fw[1]) / args.steps
    print(json.dumps(result))


if __name__ == "__main__":
    main()

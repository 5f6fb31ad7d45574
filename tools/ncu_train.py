#!/usr/bin/env python
# coding=utf-8
"""Driver for ncu captures of the training kernels: a few forward + backward steps of the bench GAT layer (recompute backward
kernels) and of MeanGraphSage at D = 100 (TMA gather4 aggregation) on the products-shape graph."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import tf_geometric_b200 as tfg  # noqa: E402

dev = torch.device("cuda", 0)
n, pairs = bench.PRODUCTS_NODES, bench.PRODUCTS_UNDIRECTED
ei = bench.make_graph_device(n, pairs, 0, dev)
gen = torch.Generator(device="cpu"); gen.manual_seed(1)
x = torch.randn((n, bench.FEATURES), generator=gen).to(dev)
gat = tfg.layers.GAT(bench.UNITS, num_heads=bench.HEADS, activation=tfg.nn.relu, seed=3, trainable=True)
sage = tfg.layers.MeanGraphSage(2 * bench.UNITS, activation=tfg.nn.relu, concat=True, seed=2, trainable=True)
for _ in range(2):
    for layer, inputs in ((gat, [x, ei]), (sage, [x.detach().requires_grad_(True), ei])):
        layer.zero_grad(set_to_none=True)
        layer(inputs, training=True).sum().backward()
torch.cuda.synchronize()

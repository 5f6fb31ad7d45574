#!/usr/bin/env python
# coding=utf-8
"""Minimal driver for ncu captures of the projection GEMM: three column blocks at the bench shape, a few launches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tf_geometric_b200 import ops  # noqa: E402

m, k, nb = int(os.environ.get("M", 2449029)), int(os.environ.get("K", 100)), int(os.environ.get("NB", 3))
dev = torch.device("cuda")
gen = torch.Generator(device="cpu"); gen.manual_seed(0)
x = torch.randn((m, k), generator=gen).to(dev)
ws = [(torch.randn((k, 128), generator=gen) / 10).to(dev) for _ in range(nb)]
bias = torch.zeros(128, device=dev)
outs = [torch.empty((m, 128), device=dev) for _ in range(nb)]
blocks = [(ws[i], bias, ops.ACT_RELU if i % 2 == 0 else ops.ACT_NONE, outs[i]) for i in range(nb)]
for _ in range(3):
    ops.gemm_proj(x, blocks)
torch.cuda.synchronize()

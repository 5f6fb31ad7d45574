#!/usr/bin/env python
# coding=utf-8
"""Run under torchrun on N GPUs: destination-partitioned GCN / GAT vs the single-GPU result computed on every rank.
The aggregation keeps the per-row edge order, so the rows owned by a rank must be bit-identical to the single-GPU rows
(same GEMM kernel, same row-local arithmetic)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tf_geometric_b200 as tfg  # noqa: E402
from tf_geometric_b200 import dist as tdist, ops  # noqa: E402
from conftest import random_graph, glorot  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

n, f, u, heads = 200003, 100, 128, 8
rs = np.random.RandomState(0)
ei = random_graph(n, 30 * n, seed=1, symmetric=True, isolated=5)
x = rs.randn(n, f).astype(np.float32)
k, b = glorot(rs, f, u), rs.randn(u).astype(np.float32)
wq, wk, wv = glorot(rs, f, u), glorot(rs, f, u), glorot(rs, f, u)
bq, bk = (rs.randn(u) * .1).astype(np.float32), (rs.randn(u) * .1).astype(np.float32)

ei_d, x_d = ops.as_device(ei), ops.as_device(x)
full_gcn = tfg.nn.gcn(x_d, tfg.SparseMatrix(ei_d, None, [n, n]), k, b, tfg.nn.relu)
full_gat = tfg.nn.gat(x_d, ei_d, wq, bq, tfg.nn.relu, wk, bk, tfg.nn.relu, wv, b, tfg.nn.relu, num_heads=heads)

pg = tdist.PartitionedGraph.from_global(ei_d, None, n, rank, world)
p = pg.part
x_loc = x_d[p.lo:p.hi].contiguous()
loc_gcn = tdist.gcn_partitioned(pg, x_loc, k, b, tfg.nn.relu)
loc_gat = tdist.gat_partitioned(pg, x_loc, wq, bq, tfg.nn.relu, wk, bk, tfg.nn.relu, wv, b, tfg.nn.relu, num_heads=heads)
# the same through the public layers (what bench.py --gpus N times): one shared publication, fused projections
gcn_l = tfg.layers.GCN(u, activation=tfg.nn.relu, seed=2)
gat_l = tfg.layers.GAT(u, num_heads=heads, activation=tfg.nn.relu, seed=3)
ref_gcn = gcn_l([x_d, ei_d])
ref_gat = gat_l([x_d, ei_d])
for step in range(3):                      # three publications: both slots and their reuse
    pg.new_step()
    shared = pg.share(x_loc, [gcn_l, gat_l])
    lay_gcn, lay_gat = gcn_l([shared, pg]), gat_l([shared, pg])
    torch.cuda.synchronize()
    ok_layers = torch.equal(lay_gcn, ref_gcn[p.lo:p.hi]) and torch.equal(lay_gat, ref_gat[p.lo:p.hi])
    print("rank {} step {} exchange={} layers bit-identical to single GPU: {}".format(rank, step, pg.exchange, ok_layers), flush=True)
    if not ok_layers:
        print("rank {} max diff gcn {:.3e} gat {:.3e}".format(rank, float((lay_gcn - ref_gcn[p.lo:p.hi]).abs().max()),
                                                             float((lay_gat - ref_gat[p.lo:p.hi]).abs().max())), flush=True)
pg_c = tdist.PartitionedGraph.from_global(ei_d, None, n, rank, world, exchange="collective")
x_c = x_d[pg_c.part.lo:pg_c.part.hi].contiguous()
col_gcn = gcn_l([x_c, pg_c])
ok_coll = torch.equal(col_gcn, ref_gcn[pg_c.part.lo:pg_c.part.hi])
print("rank {} collective path bit-identical: {}".format(rank, ok_coll), flush=True)
torch.cuda.synchronize()
ok_gcn = torch.equal(loc_gcn, full_gcn[p.lo:p.hi])
ok_gat = torch.equal(loc_gat, full_gat[p.lo:p.hi])
err_gcn = float((loc_gcn - full_gcn[p.lo:p.hi]).abs().max())
err_gat = float((loc_gat - full_gat[p.lo:p.hi]).abs().max())
print("rank {} rows [{}, {}): gcn bit-identical={} (max diff {:.2e}), gat bit-identical={} (max diff {:.2e})".format(
    rank, p.lo, p.hi, ok_gcn, err_gcn, ok_gat, err_gat), flush=True)
flag = torch.tensor([int(err_gcn < 1e-5 and err_gat < 1e-5 and ok_layers and ok_coll)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)

# coding=utf-8
"""N>1 path on CPU: world_size-2 `gloo` processes run the destination-partitioned GCN and GAT (host logic +
all-gather exchange) with the kernel layer replaced by the test double, and rank 0 checks that the concatenated
per-rank outputs equal the single-process oracle - bit-exact for the aggregation (per-row edge order is preserved)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Patch(object):
    def setattr(self, obj, name, value):
        setattr(obj, name, value)      # worker processes are short-lived: no undo needed


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, seed, renorm, queue, exchange=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_backend
        fake_backend.install(_Patch())
        from tf_geometric_b200 import dist as tdist, ops
        from conftest import random_graph, glorot
        rs = np.random.RandomState(seed)
        ei = random_graph(n, 14 * n, seed=seed, symmetric=True, isolated=3)
        w = rs.rand(ei.shape[1]).astype(np.float32) + 0.1
        w[len(w) // 2:] = w[:len(w) // 2]
        f, u, heads = 12, 16, 4
        x = rs.randn(n, f).astype(np.float32)
        k, b = glorot(rs, f, u), rs.randn(u).astype(np.float32)
        wq, wk, wv = glorot(rs, f, u), glorot(rs, f, u), glorot(rs, f, u)
        bq, bk = rs.randn(u).astype(np.float32) * .1, rs.randn(u).astype(np.float32) * .1

        pg = tdist.PartitionedGraph.from_global(ei, w, n, exchange=exchange)
        p = pg.part
        if exchange == "p2p":          # 128-row aligned blocks; without CUDA every exchange falls back to the collective
            assert p.block % 128 == 0
        x_local = x[p.lo:p.hi]
        gcn_local = tdist.gcn_partitioned(pg, x_local, k, b, ops.relu, renorm=renorm)
        gat_local = tdist.gat_partitioned(pg, x_local, wq, bq, ops.relu, wk, bk, ops.relu, wv, b, ops.relu, num_heads=heads)
        both = tdist.gcn_gat_overlapped(pg, x_local, k, b, ops.relu, wq, bq, wk, bk, wv, b, ops.relu, heads)
        if renorm:      # the overlapped step uses the default normalisation
            assert np.array_equal(both[0].numpy(), gcn_local.numpy())
        assert np.array_equal(both[1].numpy(), gat_local.numpy())
        # the same through tfg.layers with [x_local, partitioned_graph] inputs, alone and with one shared publication
        import tf_geometric_b200 as tfg
        gcn_l = tfg.layers.GCN(u, activation=ops.relu, renorm=renorm, seed=2)
        gat_l = tfg.layers.GAT(u, num_heads=heads, activation=ops.relu, seed=3)
        xl = torch.from_numpy(x_local)
        alone = gcn_l([xl, pg]), gat_l([xl, pg])
        shared = pg.share(xl, [gcn_l, gat_l])
        assert len(shared.projected) == 2
        together = gcn_l([shared, pg]), gat_l([shared, pg])
        want = (tdist.gcn_partitioned(pg, x_local, gcn_l.kernel, gcn_l.bias, ops.relu, renorm=renorm),
                tdist.gat_partitioned(pg, x_local, gat_l.query_kernel, gat_l.query_bias, ops.relu, gat_l.key_kernel,
                                      gat_l.key_bias, ops.relu, gat_l.kernel, gat_l.bias, ops.relu, num_heads=heads))
        for got in (alone, together):
            assert np.array_equal(got[0].numpy(), want[0].numpy()) and np.array_equal(got[1].numpy(), want[1].numpy())
        # the partitioned path is forward-only: a trainable layer must fail loudly instead of losing its gradients silently,
        # and still runs for inference under no_grad
        trainable = tfg.layers.GCN(u, activation=ops.relu, renorm=renorm, seed=2, trainable=True)
        with pytest.raises(NotImplementedError):
            trainable([xl, pg])
        with pytest.raises(NotImplementedError):
            pg.share(xl, [trainable])
        with torch.no_grad():
            assert np.array_equal(trainable([xl, pg]).numpy(), want[0].numpy())
        queue.put((rank, p.lo, p.hi, gcn_local.numpy(), gat_local.numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,renorm,exchange", [(101, True, None), (64, False, None), (300, True, "p2p")])
def test_partitioned_gcn_and_gat_match_single_process_oracle(n, renorm, exchange):
    from oracle import tfg_oracle as o
    from conftest import random_graph, glorot
    world, seed = 2, 5
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, seed, renorm, queue, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    parts = sorted(queue.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert parts[0][1] == 0 and parts[0][2] == parts[1][1] and parts[1][2] == n      # contiguous cover of the rows

    rs = np.random.RandomState(seed)
    ei = random_graph(n, 14 * n, seed=seed, symmetric=True, isolated=3)
    w = rs.rand(ei.shape[1]).astype(np.float32) + 0.1
    w[len(w) // 2:] = w[:len(w) // 2]
    f, u, heads = 12, 16, 4
    x = rs.randn(n, f).astype(np.float32)
    k, b = glorot(rs, f, u), rs.randn(u).astype(np.float32)
    wq, wk, wv = glorot(rs, f, u), glorot(rs, f, u), glorot(rs, f, u)
    bq, bk = rs.randn(u).astype(np.float32) * .1, rs.randn(u).astype(np.float32) * .1
    want_gcn = o.gcn(x, o.SparseMatrix(ei, w, [n, n]), k, b, o.relu, renorm=renorm)
    want_gat = o.gat(x, ei, wq, bq, o.relu, wk, bk, o.relu, wv, b, o.relu, num_heads=heads)
    got_gcn = np.concatenate([p[3] for p in parts])
    got_gat = np.concatenate([p[4] for p in parts])
    np.testing.assert_allclose(got_gcn, want_gcn, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(got_gat, want_gat, rtol=1e-5, atol=1e-6)

# coding=utf-8
"""GPU parity for the callers of the hot path: gcn / graph_sage / appnp (functional + layer API) and the dense GEMM."""
import numpy as np
import pytest
import torch

import tf_geometric_b200 as tfg
from tf_geometric_b200 import ops
from oracle import tfg_oracle as o
from conftest import random_graph, assert_close, glorot

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    return ops.as_device(a, dtype)


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("m,n,k,ta,tb", [
    (1, 1, 1, False, False), (100, 128, 100, False, False), (257, 16, 1433, False, False), (300, 7, 16, False, False),
    (513, 130, 77, False, False), (64, 200, 300, True, False), (64, 200, 300, False, True), (90, 33, 65, True, True),
    (100, 128, 50000, True, False),      # weight-gradient shape -> split-K
    (2000, 256, 128, False, False)])
def test_gemm(m, n, k, ta, tb):
    rs = np.random.RandomState(m + n + k)
    a = rs.randn(*((k, m) if ta else (m, k))).astype(np.float32)
    b = rs.randn(*((n, k) if tb else (k, n))).astype(np.float32)
    bias = rs.randn(n).astype(np.float32)
    want = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    got = ops.gemm(dev(a), dev(b), trans_a=ta, trans_b=tb)
    assert_close(host(got), want, rtol=1e-4, atol_scale=2e-6, what="gemm")
    got = ops.gemm(dev(a), dev(b), bias=dev(bias), act=ops.ACT_RELU, trans_a=ta, trans_b=tb)
    assert_close(host(got), np.maximum(want + bias, 0), rtol=1e-4, atol_scale=2e-6, what="gemm+bias+relu")
    c0 = rs.randn(m, n).astype(np.float32)
    out = dev(c0.copy())
    ops.gemm(dev(a), dev(b), trans_a=ta, trans_b=tb, beta=1.0, out=out)
    assert_close(host(out), want + c0, rtol=1e-4, atol_scale=2e-6, what="gemm beta")


@pytest.mark.parametrize("norm,loop,sym,renorm,improved,act", [
    ("both", True, True, True, False, "relu"), ("both", True, True, False, False, None),
    ("both", True, False, True, True, "relu"), ("left", True, False, True, False, None),
    ("right", False, False, True, False, "relu")])
def test_gcn_functional(norm, loop, sym, renorm, improved, act):
    rs = np.random.RandomState(1)
    n, f, u = 2500, 60, 32
    ei = random_graph(n, 30000, seed=2, symmetric=sym, isolated=4)
    w = (rs.rand(ei.shape[1]) + .1).astype(np.float32)
    x = rs.randn(n, f).astype(np.float32)
    k, b = glorot(rs, f, u), rs.randn(u).astype(np.float32)
    want = o.gcn(x, o.SparseMatrix(ei, w, [n, n]), k, b, o.relu if act else None, norm, loop, sym, renorm, improved)
    got = tfg.nn.gcn(dev(x), tfg.SparseMatrix(ei, w, [n, n]), dev(k), dev(b), tfg.nn.relu if act else None,
                     norm=norm, add_self_loop=loop, sym=sym, renorm=renorm, improved=improved)
    assert_close(host(got), want, what="gcn")
    # kernel=None propagates x itself; a non-relu activation goes through the generic callable route
    want = o.gcn(x, o.SparseMatrix(ei, w, [n, n]), None, None, np.tanh, norm, loop, sym, renorm, improved)
    got = tfg.nn.gcn(dev(x), tfg.SparseMatrix(ei, w, [n, n]), None, None, torch.tanh, norm=norm, add_self_loop=loop,
                     sym=sym, renorm=renorm, improved=improved)
    assert_close(host(got), want, what="gcn no kernel tanh")


def test_gcn_two_layer_cora_shaped_model_with_cache():
    """BASELINE config 1 wiring (demo/demo_gcn.py:18-47): GCN(16, relu) -> GCN(7), cache pre-built, numpy inputs."""
    rs = np.random.RandomState(0)
    n, f = 2708, 1433
    pairs = rs.randint(0, n, (2, 5278)).astype(np.int32)
    x = (rs.rand(n, f) < 0.0126).astype(np.float32)
    x = x / np.maximum(x.sum(1, keepdims=True), 1)
    g = tfg.Graph(x, pairs).to_directed()
    want_ei, _ = o.convert_edge_to_directed(pairs)
    np.testing.assert_array_equal(np.asarray(g.edge_index), want_ei)
    gd = g.to_device()
    l0, l1 = tfg.layers.GCN(16, activation=tfg.nn.relu, seed=3), tfg.layers.GCN(7, seed=4)
    l0.build_cache_for_graph(gd)
    h = l0([gd.x, gd.edge_index, gd.edge_weight], cache=gd.cache)
    logits = l1([h, gd.edge_index, gd.edge_weight], cache=gd.cache)
    assert len(gd.cache) == 1
    k0, b0, k1, b1 = (host(p) for p in (l0.kernel, l0.bias, l1.kernel, l1.bias))
    adj = o.SparseMatrix(want_ei, np.ones(want_ei.shape[1], np.float32), [n, n])
    want = o.gcn(o.gcn(x, adj, k0, b0, o.relu), adj, k1, b1)
    assert_close(host(logits), want, what="2-layer GCN")
    assert sorted(k for k, _ in l0.named_parameters()) == ["bias", "kernel"]


@pytest.mark.parametrize("concat", [True, False])
@pytest.mark.parametrize("weighted", [True, False])
@pytest.mark.parametrize("kind", ["mean", "sum"])
def test_plain_graph_sage(kind, weighted, concat):
    rs = np.random.RandomState(5)
    n, f, u = 2000, 100, 32
    ei = random_graph(n, 25000, seed=7, isolated=5)
    w = rs.rand(ei.shape[1]).astype(np.float32) if weighted else None
    x = rs.randn(n, f).astype(np.float32)
    ws, wn = glorot(rs, f, u), glorot(rs, f, u)
    b = rs.randn(2 * u if concat else u).astype(np.float32)
    fo = {"mean": o.mean_graph_sage, "sum": o.sum_graph_sage}[kind]
    fg = {"mean": tfg.nn.mean_graph_sage, "sum": tfg.nn.sum_graph_sage}[kind]
    want = fo(x, ei, w, ws, wn, b, o.relu, concat=concat, normalize=True)
    got = fg(dev(x), dev(ei), dev(w), dev(ws), dev(wn), dev(b), tfg.nn.relu, concat=concat, normalize=True)
    assert_close(host(got), want, what=kind + "_graph_sage")


def test_pool_and_gcn_graph_sage_and_layers():
    rs = np.random.RandomState(6)
    n, f, units = 1500, 40, 32
    ei = random_graph(n, 20000, seed=8, symmetric=True)          # every node has in-edges: max-pool stays finite
    w = rs.rand(ei.shape[1]).astype(np.float32)
    x = rs.randn(n, f).astype(np.float32)
    for cls, fo in ((tfg.layers.MeanPoolGraphSage, o.mean_pool_graph_sage), (tfg.layers.MaxPoolGraphSage, o.max_pool_graph_sage)):
        layer = cls(units, seed=11)
        out = layer([dev(x), dev(ei), dev(w)])
        p = {k: host(v) for k, v in layer.named_parameters()}
        mk, mb, nk = layer._names
        want = fo(x, ei, w, p["self_kernel"], p[mk], p[nk], p[mb], p["bias"], o.relu)
        assert_close(host(out), want, what=cls.__name__)
        with pytest.raises(TypeError):
            layer([dev(x), dev(ei)])                               # edge_weight=None crashes in the reference too
    assert sorted(tfg.layers.MaxPoolGraphSage._names) == ["mlp_bias", "mlp_kernel", "neighs_kernel"]
    layer = tfg.layers.GCNGraphSage(units, seed=12)
    p_out = layer([dev(x), dev(ei), dev(w)])
    p = {k: host(v) for k, v in layer.named_parameters()}
    assert_close(host(p_out), o.gcn_graph_sage(x, ei, w, p["kernel"], p["bias"], o.relu), what="GCNGraphSage")
    assert_close(host(layer([dev(x), dev(ei), dev(w)], cache={"warm": 1})),
                 o.gcn_graph_sage(x, ei, w, p["kernel"], p["bias"], o.relu, cache={"warm": 1}), what="GCNGraphSage cache quirk")
    layer = tfg.layers.MeanGraphSage(units, seed=13)
    out = layer([dev(x), dev(ei)])
    p = {k: host(v) for k, v in layer.named_parameters()}
    assert_close(host(out), o.mean_graph_sage(x, ei, None, p["self_kernel"], p["neighbor_kernel"], p["bias"], o.relu),
                 what="MeanGraphSage")
    with pytest.raises(Exception):
        tfg.layers.SumGraphSage(33)


@pytest.mark.parametrize("k,alpha", [(10, 0.1), (1, 0.5), (0, 0.1)])
def test_appnp(k, alpha):
    rs = np.random.RandomState(8)
    n, f = 2000, 50
    ei = random_graph(n, 24000, seed=9, symmetric=True)
    w = (rs.rand(ei.shape[1]) + .2).astype(np.float32)
    x = rs.randn(n, f).astype(np.float32)
    kernels = [glorot(rs, f, 64), glorot(rs, 64, 7)]
    biases = [rs.randn(64).astype(np.float32), rs.randn(7).astype(np.float32)]
    want = o.appnp(x, ei, w, kernels, biases, o.relu, o.relu, k=k, alpha=alpha)
    got = tfg.nn.appnp(dev(x), dev(ei), dev(w), [dev(a) for a in kernels], [dev(a) for a in biases], tfg.nn.relu,
                       tfg.nn.relu, k=k, alpha=alpha)
    assert_close(host(got), want, what="appnp")
    layer = tfg.layers.APPNP([64, 7], k=k, alpha=alpha, seed=2)
    g = tfg.Graph(x, ei, edge_weight=w).to_device()
    out = layer([g.x, g.edge_index, g.edge_weight], cache=g.cache)
    p = {kk: host(v) for kk, v in layer.named_parameters()}
    assert sorted(p) == ["bias_0", "bias_1", "kernel_0", "kernel_1"]
    want = o.appnp(x, ei, w, [p["kernel_0"], p["kernel_1"]], [p["bias_0"], p["bias_1"]], o.relu, None, k=k, alpha=alpha)
    assert_close(host(out), want, what="APPNP layer")
    assert "gcn_normed_adj_both_True_True_True_False" in g.cache


@pytest.mark.parametrize("kind,weighted,concat", [("mean", False, True), ("mean", True, False), ("sum", True, True)])
def test_graph_sage_forward_backward(kind, weighted, concat):
    """BASELINE config 4 semantics at test size: loss = sum(out * G); gradients w.r.t. x, both kernels and the bias
    against torch-CPU autograd over the reference's op sequence (gather -> gcn_mapper -> unsorted_segment_mean -> matmuls),
    which is what TensorFlow's GradientTape differentiates in demo/demo_graph_sage.py:100-106."""
    rs = np.random.RandomState(17)
    n, f, u = 3000, 100, 64
    ei = random_graph(n, 40000, seed=23, isolated=6, hub=(11, 3000))
    w = (rs.rand(ei.shape[1]) + 0.1).astype(np.float32) if weighted else None
    x = rs.randn(n, f).astype(np.float32)
    ws, wn = glorot(rs, f, u), glorot(rs, f, u)
    b = rs.randn(2 * u if concat else u).astype(np.float32)
    G = rs.randn(n, 2 * u if concat else u).astype(np.float32)

    # reference: op-for-op on torch CPU with autograd
    tx, tws, twn, tb = (torch.tensor(a, requires_grad=True) for a in (x, ws, wn, b))
    row, col = torch.from_numpy(ei[0]).long(), torch.from_numpy(ei[1]).long()
    msg = tx.index_select(0, col)
    if weighted:
        msg = msg * torch.from_numpy(w).unsqueeze(1)
    agg = torch.zeros((n, f)).index_add_(0, row, msg)
    if kind == "mean":
        agg = agg / torch.bincount(row, minlength=n).clamp(min=1).to(torch.float32).unsqueeze(1)
    left, right = tx @ tws, agg @ twn
    h = torch.relu((torch.cat([left, right], 1) if concat else left + right) + tb)
    (h * torch.from_numpy(G)).sum().backward()

    dx, dws, dwn, db = (dev(a).requires_grad_(True) for a in (x, ws, wn, b))
    fn = tfg.nn.mean_graph_sage if kind == "mean" else tfg.nn.sum_graph_sage
    out = fn(dx, dev(ei), dev(w), dws, dwn, db, tfg.nn.relu, concat=concat)
    assert out.requires_grad
    assert_close(host(out), h.detach().numpy(), what="sage fwd (autograd path)")
    (out * dev(G)).sum().backward()
    for name, got, want in (("dx", dx.grad, tx.grad), ("dWs", dws.grad, tws.grad), ("dWn", dwn.grad, twn.grad),
                            ("db", db.grad, tb.grad)):
        assert_close(host(got), want.numpy(), what="sage " + name)
    # isolated nodes only receive gradient through the self path
    assert np.isfinite(host(dx.grad)).all()


def test_propagation_layers_sgc_ssgc_tagcn_gin_leconv():
    """the remaining norm(A) @ H convolutions through the layer API, against the oracle."""
    rs = np.random.RandomState(31)
    n, f, u = 2500, 48, 32
    ei = random_graph(n, 30000, seed=33, symmetric=True, isolated=4)
    w = (rs.rand(ei.shape[1]) + 0.2).astype(np.float32)
    w[len(w) // 2:] = w[:len(w) // 2]
    x = rs.randn(n, f).astype(np.float32)
    g = tfg.Graph(x, ei, edge_weight=w).to_device()
    inputs = [g.x, g.edge_index, g.edge_weight]

    layer = tfg.layers.SGC(u, k=3, activation=tfg.nn.relu, seed=1)
    out = layer(inputs, cache=g.cache)
    p = {k: host(v) for k, v in layer.named_parameters()}
    assert_close(host(out), o.sgc(x, ei, w, 3, p["kernel"], p["bias"], o.relu), what="SGC")
    assert "gcn_normed_adj_both_True_True_True_False" in g.cache

    layer = tfg.layers.TAGCN(u, k=2, activation=tfg.nn.relu, seed=2)
    out = layer(inputs)
    p = {k: host(v) for k, v in layer.named_parameters()}
    assert p["kernel"].shape == (f * 3, u)
    assert_close(host(out), o.tagcn(x, ei, w, 2, p["kernel"], p["bias"], o.relu), what="TAGCN")

    layer = tfg.layers.SSGC([64, 7], k=6, alpha=0.15, seed=3)
    out = layer(inputs, cache=g.cache)
    p = {k: host(v) for k, v in layer.named_parameters()}
    want = o.ssgc(x, ei, w, [p["kernel_0"], p["kernel_1"]], [p["bias_0"], p["bias_1"]], k=6, alpha=0.15)
    assert_close(host(out), want, what="SSGC")

    layer = tfg.layers.LEConv(u, activation=tfg.nn.relu, seed=4)
    out = layer(inputs)
    p = {k: host(v) for k, v in layer.named_parameters()}
    assert sorted(p) == ["aggr_neighbor_kernel", "aggr_self_bias", "aggr_self_kernel", "self_bias", "self_kernel"]
    want = o.le_conv(x, ei, w, p["self_kernel"], p["self_bias"], p["aggr_self_kernel"], p["aggr_self_bias"],
                     p["aggr_neighbor_kernel"], None, o.relu)
    assert_close(host(out), want, what="LEConv")

    mlp_w = glorot(rs, f, u)
    mlp_d = dev(mlp_w)
    layer = tfg.layers.GIN(lambda h, training=None: ops.gemm(h, mlp_d, act=ops.ACT_RELU), eps=0.5)
    out = layer([g.x, g.edge_index])
    want = o.gin(x, ei, lambda h: o.relu((h @ mlp_w).astype(np.float32)), eps=0.5)
    assert_close(host(out), want, what="GIN")


@pytest.mark.parametrize("sym", [True, False])
def test_gcn_two_layer_forward_backward(sym):
    """demo/demo_gcn.py wiring with gradients: loss = sum(logits * G) through two GCN layers (relu between), grads w.r.t. x,
    both kernels and biases against torch-CPU autograd over the reference's op sequence."""
    rs = np.random.RandomState(41)
    n, f, hidden, classes = 2000, 60, 16, 7
    ei = random_graph(n, 24000, seed=43, symmetric=sym, isolated=3, hub=(5, 2600))
    w = (rs.rand(ei.shape[1]) + 0.2).astype(np.float32)
    if sym:
        w[len(w) // 2:] = w[:len(w) // 2]
    x = rs.randn(n, f).astype(np.float32)
    k0, b0, k1, b1 = glorot(rs, f, hidden), rs.randn(hidden).astype(np.float32), glorot(rs, hidden, classes), rs.randn(classes).astype(np.float32)
    G = rs.randn(n, classes).astype(np.float32)
    normed = o.gcn_norm_adj(o.SparseMatrix(ei, w, [n, n]), sym=sym)
    row, col = torch.from_numpy(normed.index[0]).long(), torch.from_numpy(normed.index[1]).long()
    val = torch.from_numpy(normed.value)
    tx, tk0, tb0, tk1, tb1 = (torch.tensor(a, requires_grad=True) for a in (x, k0, b0, k1, b1))

    def layer(h, kern, b, relu):
        hw = h @ kern
        out = torch.zeros((n, hw.shape[1])).index_add_(0, row, hw.index_select(0, col) * val.unsqueeze(1)) + b
        return torch.relu(out) if relu else out
    logits = layer(layer(tx, tk0, tb0, True), tk1, tb1, False)
    (logits * torch.from_numpy(G)).sum().backward()

    dx, dk0, db0, dk1, db1 = (dev(a).requires_grad_(True) for a in (x, k0, b0, k1, b1))
    adj = tfg.SparseMatrix(ei, w, [n, n])
    cache = {}
    h = tfg.nn.gcn(dx, adj, dk0, db0, tfg.nn.relu, sym=sym, cache=cache)
    out = tfg.nn.gcn(h, adj, dk1, db1, None, sym=sym, cache=cache)
    assert out.requires_grad and len(cache) == 1
    assert_close(host(out), logits.detach().numpy(), what="2-layer gcn fwd (autograd path)")
    (out * dev(G)).sum().backward()
    for name, got, want in (("dx", dx.grad, tx.grad), ("dW0", dk0.grad, tk0.grad), ("db0", db0.grad, tb0.grad),
                            ("dW1", dk1.grad, tk1.grad), ("db1", db1.grad, tb1.grad)):
        assert_close(host(got), want.numpy(), what="gcn " + name)


def test_trainable_gcn_layer_one_sgd_step_reduces_loss():
    """A GCN layer created with trainable=True: forward, backward and one SGD step through the public layer API."""
    rs = np.random.RandomState(3)
    n, f, c = 1500, 40, 5
    ei = random_graph(n, 18000, seed=4, symmetric=True)
    x = rs.randn(n, f).astype(np.float32)
    y = dev(rs.randn(n, c).astype(np.float32))
    g = tfg.Graph(x, ei).to_device()
    layer = tfg.layers.GCN(c, seed=5, trainable=True)
    layer.build_cache_for_graph(g)
    out = layer([g.x, g.edge_index, g.edge_weight], cache=g.cache)      # first call builds the weights
    opt = torch.optim.SGD(layer.parameters(), lr=0.05)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        out = layer([g.x, g.edge_index, g.edge_weight], cache=g.cache)
        loss = ((out - y) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[2] < losses[1] < losses[0]
    assert layer.kernel.grad is not None and torch.isfinite(layer.kernel.grad).all()


def test_chebynet_layer_static_and_dynamic_lambda():
    rs = np.random.RandomState(51)
    n, f, u, k = 1800, 24, 16, 4
    ei = random_graph(n, 20000, seed=53, symmetric=True)
    w = (rs.rand(ei.shape[1]) + 0.3).astype(np.float32)
    w[len(w) // 2:] = w[:len(w) // 2]
    x = rs.randn(n, f).astype(np.float32)
    g = tfg.Graph(x, ei, edge_weight=w).to_device()
    layer = tfg.layers.ChebyNet(u, k, activation=tfg.nn.relu, seed=7)
    layer.build_cache_for_graph(g)
    out = layer([g.x, g.edge_index, g.edge_weight], cache=g.cache)
    p = {kk: host(v) for kk, v in layer.named_parameters()}
    assert sorted(p) == ["bias", "kernel0", "kernel1", "kernel2", "kernel3"] and "chebynet_normed_edge_sym" in g.cache
    kernels = [p["kernel{}".format(i)] for i in range(k)]
    assert_close(host(out), o.chebynet(x, ei, w, k, kernels, p["bias"], o.relu), what="ChebyNet (lambda_max = 2)")
    # dynamic lambda_max: host-side scipy eigsh in both the product and the reference
    lam = tfg.nn.conv.propagation.laplacian_max_eigenvalue(dev(ei), n, dev(w), "sym")
    got = tfg.nn.chebynet(dev(x), dev(ei), dev(w), k, [dev(a) for a in kernels], dev(p["bias"]), None,
                          use_dynamic_lambda_max=True)
    assert_close(host(got), o.chebynet(x, ei, w, k, kernels, p["bias"], None, lambda_max=lam), what="ChebyNet (dynamic lambda)")
    assert 0.5 < lam < 2.5


def test_degenerate_graphs_empty_edges_single_node():
    """Empty / ragged inputs: an edge list with zero columns (every row empty), a one-node graph, a graph whose only edges are
    pre-existing self loops.  The reference proceeds normally for edge_index of shape [2, 0] (map_reduce.py:57 only tests dim 0)."""
    rs = np.random.RandomState(61)
    for n, ei in ((7, np.zeros((2, 0), np.int32)), (1, np.zeros((2, 0), np.int32)), (1, np.array([[0, 0], [0, 0]], np.int32)),
                  (5, np.array([[2, 2, 4], [2, 2, 4]], np.int32))):
        f, u = 6, 8
        x = rs.randn(n, f).astype(np.float32)
        w = np.ones(ei.shape[1], np.float32)
        k, b = glorot(rs, f, u), rs.randn(u).astype(np.float32)
        what = "n={} e={}".format(n, ei.shape[1])
        got = tfg.nn.aggregate_neighbors(dev(x), dev(ei), None, tfg.nn.identity_mapper, tfg.nn.mean_reducer, tfg.nn.sum_updater)
        np.testing.assert_array_equal(host(got), o.aggregate_neighbors(x, ei, None, o.identity_mapper, o.mean_reducer, o.sum_updater))
        got = tfg.nn.aggregate_neighbors(dev(x), dev(ei), None, tfg.nn.identity_mapper, tfg.nn.max_reducer, tfg.nn.identity_updater)
        np.testing.assert_array_equal(host(got), o.aggregate_neighbors(x, ei, None, o.identity_mapper, o.max_reducer, o.identity_updater))
        assert_close(host(tfg.nn.gcn(dev(x), tfg.SparseMatrix(ei, w, [n, n]), dev(k), dev(b), tfg.nn.relu)),
                     o.gcn(x, o.SparseMatrix(ei, w, [n, n]), k, b, o.relu), what="gcn " + what)
        wq, wk, wv = glorot(rs, f, u), glorot(rs, f, u), glorot(rs, f, u)
        z = np.zeros(u, np.float32)
        assert_close(host(tfg.nn.gat(dev(x), dev(ei), dev(wq), dev(z), tfg.nn.relu, dev(wk), dev(z), tfg.nn.relu, dev(wv), dev(b),
                                     None, num_heads=2)),
                     o.gat(x, ei, wq, z, o.relu, wk, z, o.relu, wv, b, None, num_heads=2), what="gat " + what)
        assert_close(host(tfg.nn.mean_graph_sage(dev(x), dev(ei), None, dev(k), dev(k), dev(np.concatenate([b, b])), tfg.nn.relu)),
                     o.mean_graph_sage(x, ei, None, k, k, np.concatenate([b, b]), o.relu), what="sage " + what)
        assert host(tfg.nn.segment_count(dev(ei[0]), n)).tolist() == np.bincount(ei[0], minlength=n).tolist()


def test_sparse_features_and_column_splits():
    """tf.SparseTensor features (nn/conv/gcn.py:269-272, gat.py:45-70; Cora's bag of words) as SparseMatrix / torch sparse /
    scipy sparse inputs: same result as the dense matrix; num_or_size_splits (gcn.py:274-280) changes no bit."""
    import scipy.sparse as sp
    rs = np.random.RandomState(12)
    n, f, u, heads = 700, 300, 16, 4
    dense = ((rs.rand(n, f) < 0.05) * rs.rand(n, f)).astype(np.float32)
    dense[3] = 0.0                                                  # a node without features
    ei = random_graph(n, 6000, seed=13, symmetric=True, isolated=2)
    k, b = glorot(rs, f, u), rs.randn(u).astype(np.float32)
    wq, wk, wv = glorot(rs, f, u), glorot(rs, f, u), glorot(rs, f, u)
    bq, bk = rs.randn(u).astype(np.float32) * .1, rs.randn(u).astype(np.float32) * .1
    adj = o.SparseMatrix(ei, None, [n, n])
    want_gcn = o.gcn(dense, adj, k, b, o.relu)
    want_gat = o.gat(dense, ei, wq, bq, o.relu, wk, bk, o.relu, wv, b, o.relu, num_heads=heads)
    eid = dev(ei, torch.int32)
    nz = np.nonzero(dense)
    forms = {
        "SparseMatrix": tfg.SparseMatrix(np.stack(nz).astype(np.int32), dense[nz], [n, f]),
        "scipy": sp.csr_matrix(dense),
    }
    if torch.cuda.is_available():
        forms["torch_coo"] = torch.from_numpy(dense).to_sparse().cuda()
    for name, xs in forms.items():
        got = tfg.nn.gcn(xs, tfg.SparseMatrix(eid, None, [n, n]), dev(k), dev(b), tfg.nn.relu)
        assert_close(host(got), want_gcn, rtol=1e-4, atol_scale=1e-4, what="gcn with {} features".format(name))
        got = tfg.nn.gat(xs, eid, dev(wq), dev(bq), tfg.nn.relu, dev(wk), dev(bk), tfg.nn.relu, dev(wv), dev(b), tfg.nn.relu,
                         num_heads=heads)
        assert_close(host(got), want_gat, rtol=1e-4, atol_scale=1e-4, what="gat with {} features".format(name))
    layer = tfg.layers.GCN(u, activation=tfg.nn.relu, seed=4)
    a = layer([forms["SparseMatrix"], eid])
    assert tuple(a.shape) == (n, u)
    full = tfg.nn.gcn(dev(dense), tfg.SparseMatrix(eid, None, [n, n]), dev(k), dev(b), tfg.nn.relu)
    for splits in (2, 4, [5, 11], [16]):
        part = tfg.nn.gcn(dev(dense), tfg.SparseMatrix(eid, None, [n, n]), dev(k), dev(b), tfg.nn.relu, num_or_size_splits=splits)
        assert torch.equal(part, full), splits
    with pytest.raises(ValueError):
        tfg.nn.gcn(dev(dense), tfg.SparseMatrix(eid, None, [n, n]), dev(k), dev(b), num_or_size_splits=3)


def test_training_with_sparse_features_matches_dense_features():
    """demo_gcn.py / demo_gat.py train on tf.SparseTensor features: the weight gradients with a sparse x (projection and
    its transpose on the aggregation kernel) equal the ones the dense training path gives (that path is checked against
    float64 autograd in test_gpu_train)."""
    rs = np.random.RandomState(21)
    n, f, u, heads = 500, 240, 16, 4
    dense = ((rs.rand(n, f) < 0.06) * rs.rand(n, f)).astype(np.float32)
    dense[7] = 0.0
    ei = random_graph(n, 4000, seed=5, symmetric=True, isolated=1)
    eid = dev(ei, torch.int32)
    nz = np.nonzero(dense)
    xs = tfg.SparseMatrix(np.stack(nz).astype(np.int32), dense[nz], [n, f])
    names = ("k", "b", "wq", "bq", "wk", "bk", "wv")
    init = {"k": glorot(rs, f, u), "b": rs.randn(u).astype(np.float32), "wq": glorot(rs, f, u), "wk": glorot(rs, f, u),
            "wv": glorot(rs, f, u), "bq": rs.randn(u).astype(np.float32) * .1, "bk": rs.randn(u).astype(np.float32) * .1}
    gout = dev(rs.randn(n, u).astype(np.float32))

    def grads(x):
        p = {k_: dev(init[k_]).requires_grad_(True) for k_ in names}
        y1 = tfg.nn.gcn(x, tfg.SparseMatrix(eid, None, [n, n]), p["k"], p["b"], tfg.nn.relu)
        y2 = tfg.nn.gat(x, eid, p["wq"], p["bq"], tfg.nn.relu, p["wk"], p["bk"], tfg.nn.relu, p["wv"], p["b"], tfg.nn.relu,
                        num_heads=heads)
        assert y1.requires_grad and y2.requires_grad
        ((y1 + y2) * gout).sum().backward()
        return host(y1), host(y2), {k_: host(p[k_].grad) for k_ in names}

    s1, s2, gs = grads(xs)
    d1, d2, gd = grads(dev(dense))
    assert_close(s1, d1, rtol=1e-4, atol_scale=1e-4, what="gcn forward, sparse features, training path")
    assert_close(s2, d2, rtol=1e-4, atol_scale=1e-4, what="gat forward, sparse features, training path")
    for k_ in names:
        assert np.abs(gd[k_]).sum() > 0, k_
        assert_close(gs[k_], gd[k_], rtol=1e-3, atol_scale=2e-4, what="d " + k_ + " with sparse features")
    layer = tfg.layers.GCN(u, activation=tfg.nn.relu, seed=4, trainable=True)
    out = layer([xs, eid], training=True)
    out.sum().backward()
    assert all(p_.grad is not None and float(p_.grad.abs().sum()) > 0 for n_, p_ in layer.named_parameters() if "bias" not in n_)

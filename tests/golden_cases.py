# coding=utf-8
"""Replays tests/golden/ref_exec_*.npz (outputs of the reference's own Python code, see tools/gen_golden_from_reference.py)
against (a) the CPU oracle and (b) the CUDA product through its public API."""
import numpy as np

from conftest import assert_close

EXACT = dict(rtol=0, atol=0)


def _eq(a, b, what):
    np.testing.assert_array_equal(np.asarray(a), np.asarray(b), err_msg=what)


def _close(a, b, what, rtol=1e-4, atol_scale=1e-4):
    assert_close(np.asarray(a), np.asarray(b), rtol=rtol, atol_scale=atol_scale, what=what)


# ------------------------------------------------------------------------------------------------------------------
# generic driver: `api` is a small adapter with the same function names for the oracle and for the product
# ------------------------------------------------------------------------------------------------------------------

class OracleApi(object):
    exact_float = True

    def __init__(self):
        from oracle import tfg_oracle as o
        self.o = o
        self.relu = o.relu

    def arr(self, a):
        return a

    def out(self, a):
        return np.asarray(a)

    def __getattr__(self, name):
        return getattr(self.o, name)

    def sparse(self, ei, w, n):
        return self.o.SparseMatrix(ei, w, [n, n])

    def gcn_norm(self, adj, norm, loop, sym, renorm, improved):
        m = self.o.gcn_norm_adj(adj, norm, loop, sym, renorm, improved)
        return m.index, m.value

    def gcn_graph_sage(self, *a, **k):
        return self.o.gcn_graph_sage(*a, **k)

    def neighbor_sample(self, ei, w, **kw):
        return self.o.random_neighbor_sample(ei, w, seed=1, **kw)

    def lstm(self, kernel, recurrent_kernel, bias):
        return self.o.numpy_lstm(kernel, recurrent_kernel, bias)

    def subgraph(self, x, ei, w, gi, y, keep):
        return self.o.sample_new_graph_by_node_index(x, ei, w, keep, gi, y)

    def sag_pool_scored(self, x, ei, w, gi, score, **kw):
        return self.o.sag_pool(x, ei, w, gi, lambda inputs: score, score_activation=np.tanh, **kw)

    def batch(self, parts):
        return self.o.batch_graphs(parts)


class ProductApi(object):
    exact_float = False

    def __init__(self):
        import torch
        import tf_geometric_b200 as tfg
        self.tfg, self.torch = tfg, torch
        self.relu = tfg.nn.relu
        for name in ("aggregate_neighbors", "identity_mapper", "gcn_mapper", "neighbor_count_mapper", "sum_reducer",
                     "mean_reducer", "max_reducer", "sum_updater", "identity_updater", "segment_softmax", "segment_count",
                     "gcn", "gat", "mean_graph_sage", "sum_graph_sage", "gcn_graph_sage", "mean_pool_graph_sage",
                     "max_pool_graph_sage", "appnp", "sgc", "ssgc", "tagcn", "gin", "le_conv", "mean_pool", "sum_pool",
                     "max_pool", "min_pool", "chebynet", "chebynet_norm_edge", "topk_pool", "set2set", "sort_pool"):
            setattr(self, name, getattr(tfg.nn, name))
        for name in ("convert_edge_to_directed", "merge_duplicated_edge", "add_self_loop_edge", "remove_self_loop_edge",
                     "adj_norm_edge"):
            setattr(self, name, getattr(tfg.utils, name))

    def arr(self, a):
        if a is None:
            return None
        a = np.asarray(a)
        return self.tfg.ops.as_device(a, {"i": self.torch.int32, "f": self.torch.float32}[a.dtype.kind])

    def out(self, a):
        return a.detach().cpu().numpy() if self.torch.is_tensor(a) else np.asarray(a)

    def sparse(self, ei, w, n):
        return self.tfg.SparseMatrix(ei, w, [n, n])

    def gcn_norm(self, adj, norm, loop, sym, renorm, improved):
        m = self.tfg.nn.gcn_norm_adj(adj, norm, loop, sym, renorm, improved)
        return m.index, m.value

    def neighbor_sample(self, ei, w, **kw):
        return self.tfg.utils.RandomNeighborSampler(ei, w).sample(seed=1, **kw)

    def subgraph(self, x, ei, w, gi, y, keep):
        g = self.tfg.BatchGraph(x, ei, gi, None, y=y, edge_weight=w).sample_new_graph_by_node_index(keep)
        return g.x, g.edge_index, g.edge_weight, g.node_graph_index, g.y

    def sag_pool_scored(self, x, ei, w, gi, score, **kw):
        return self.tfg.nn.sag_pool(x, ei, w, gi, lambda inputs, training=None: score, score_activation=self.torch.tanh, **kw)

    def batch(self, parts):
        bg = self.tfg.BatchGraph.from_graphs([self.tfg.Graph(x, ei, y=y, edge_weight=w) for x, ei, w, y in parts])
        return bg.x, bg.edge_index, bg.edge_weight, bg.y, bg.node_graph_index, bg.edge_graph_index

    def lstm(self, kernel, recurrent_kernel, bias):
        """The same LSTM cell as oracle.numpy_lstm, on torch tensors (it is the caller-supplied ARGUMENT of set2set)."""
        torch = self.torch
        k, r, b = (self.arr(a) for a in (kernel, recurrent_kernel, bias))
        units = r.shape[0]

        def lstm(inputs, initial_state=None, training=None):
            h, c = initial_state
            seq = []
            for t in range(inputs.shape[1]):
                z = inputs[:, t] @ k + h @ r + b
                i, f, g, o = (z[:, j * units:(j + 1) * units] for j in range(4))
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                h = torch.sigmoid(o) * torch.tanh(c)
                seq.append(h)
            return torch.stack(seq, dim=1), h, c
        return lstm


def replay(fname, data, api):
    kind = fname[len("ref_exec_"):-len(".npz")]
    globals()["_replay_" + kind](data, api)


def replay_with_oracle(fname, data):
    replay(fname, data, OracleApi())


def _replay_graph_utils(d, api):
    A, O = api.arr, api.out
    i, (w,) = api.convert_edge_to_directed(A(d["readme_ei"]), [A(d["readme_w"])], ["sum"])
    _eq(O(i), d["readme_directed_index"], "readme to_directed index")
    _eq(O(w), d["readme_directed_w"], "readme to_directed weight")
    for mode in ("sum", "min", "max", "mean"):
        i, (w,) = api.merge_duplicated_edge(A(d["multi_ei"]), [A(d["multi_w"])], [mode])
        _eq(O(i), d["merge_{}_index".format(mode)], "merge index " + mode)
        _eq(O(w), d["merge_{}_w".format(mode)], "merge weight " + mode)
        i, (w,) = api.convert_edge_to_directed(A(d["multi_ei"]), [A(d["multi_w"])], [mode])
        _eq(O(i), d["directed_{}_index".format(mode)], "directed index " + mode)
        _eq(O(w), d["directed_{}_w".format(mode)], "directed weight " + mode)
    i, w = api.add_self_loop_edge(A(d["multi_ei"]), 12, A(d["multi_w"]), fill_weight=2.0)
    _eq(O(i), d["self_loop_index"], "self loop index")
    _eq(O(w), d["self_loop_w"], "self loop weight")
    i, w = api.remove_self_loop_edge(A(d["multi_ei"]), A(d["multi_w"]))
    _eq(O(i), d["no_loop_index"], "remove loop index")
    _eq(O(w), d["no_loop_w"], "remove loop weight")
    i, w = api.adj_norm_edge(A(d["multi_ei"]), 12, A(d["multi_w"]), add_self_loop=True)
    _eq(O(i), d["adj_norm_index"], "adj_norm index")
    _close(O(w), d["adj_norm_w"], "adj_norm weight", rtol=6e-7, atol_scale=0)


def _replay_kernel(d, api):
    A, O = api.arr, api.out
    n = int(d["n"])
    ei, x, w = d["ei"], d["x"], d["w"]
    _close(O(api.segment_softmax(A(d["scores"]), A(ei[0]), n)), d["segment_softmax"], "segment_softmax", rtol=2e-5,
           atol_scale=1e-7)
    cnt = O(api.segment_count(A(ei[0]), n))
    _eq(cnt, d["segment_count"], "segment_count")
    assert cnt.dtype == np.int32
    red = {"sum": api.sum_reducer, "mean": api.mean_reducer, "max": api.max_reducer}
    for rname, fn in red.items():
        got = api.aggregate_neighbors(A(x), A(ei), None, api.identity_mapper, fn, api.sum_updater, num_nodes=n)
        _eq(O(got), d["agg_identity_{}_sumupd".format(rname)], "aggregate identity/" + rname)
        got = api.aggregate_neighbors(A(x), A(ei), A(w), api.gcn_mapper, fn, api.identity_updater, num_nodes=n)
        _eq(O(got), d["agg_gcn_{}_idupd".format(rname)], "aggregate gcn_mapper/" + rname)
    got = api.aggregate_neighbors(A(x), A(ei), None, api.neighbor_count_mapper, api.sum_reducer, api.identity_updater,
                                  num_nodes=n)
    _eq(O(got), d["agg_count"], "neighbor count")


def _replay_gcn(d, api):
    A, O = api.arr, api.out
    n = int(d["n"])
    for i, cfg in enumerate(d["configs"]):
        norm, loop, sym, renorm, improved = str(cfg).split("|")
        loop, sym, renorm, improved = (v == "True" for v in (loop, sym, renorm, improved))
        ei, w = (d["ei_sym"], d["w_sym"]) if sym else (d["ei_dir"], d["w_dir"])
        index, value = api.gcn_norm(api.sparse(ei, w, n), norm, loop, sym, renorm, improved)
        _eq(O(index), d["norm{}_index".format(i)], "gcn_norm index " + str(cfg))          # bit-exact integers
        _close(O(value), d["norm{}_value".format(i)], "gcn_norm value " + str(cfg), rtol=0 if api.exact_float else 6e-7,
               atol_scale=0)
        got = api.gcn(A(d["x"]), api.sparse(ei, w, n), A(d["kernel"]), A(d["bias"]), api.relu, norm, loop, sym, renorm,
                      improved)
        _close(O(got), d["gcn{}_out".format(i)], "gcn out " + str(cfg), rtol=0 if api.exact_float else 1e-4,
               atol_scale=0 if api.exact_float else 1e-4)
    got = api.gcn(A(d["x"]), api.sparse(d["ei_sym"], d["w_sym"], n), None, None)
    _close(O(got), d["gcn_no_kernel"], "gcn no kernel", rtol=0 if api.exact_float else 1e-5, atol_scale=0 if api.exact_float else 1e-6)


def _replay_gat(d, api):
    A, O = api.arr, api.out
    for tag in ("h8", "h4avg", "h1", "demo"):
        g = lambda k: d[tag + "_" + k]
        got = api.gat(A(d["x"]), A(d["ei"]), A(g("wq")), A(g("bq")), api.relu, A(g("wk")), A(g("bk")), api.relu,
                      A(g("wv")), A(g("b")), api.relu, num_heads=int(g("heads")), split_value_heads=bool(g("split")))
        _close(O(got), g("out"), "gat " + tag, rtol=0 if api.exact_float else 1e-4, atol_scale=0 if api.exact_float else 1e-4)


def _replay_graph_sage(d, api):
    A, O = api.arr, api.out
    x, ei, w = A(d["x"]), A(d["ei"]), A(d["w"])
    ws, wn, b2, b1 = A(d["ws"]), A(d["wn"]), A(d["b2"]), A(d["b1"])
    wm, bm, wnk, kernel = A(d["wm"]), A(d["bm"]), A(d["wnk"]), A(d["kernel"])
    tol = dict(rtol=0, atol_scale=0) if api.exact_float else dict(rtol=1e-4, atol_scale=1e-4)
    _close(O(api.mean_graph_sage(x, ei, w, ws, wn, b2, api.relu, True, True)), d["mean_w_concat_norm"], "mean sage",
           **(dict(rtol=1e-6, atol_scale=1e-7) if api.exact_float else tol))
    _close(O(api.mean_graph_sage(x, ei, None, ws, wn, b1, api.relu, False, False)), d["mean_now_add"], "mean sage add", **tol)
    _close(O(api.sum_graph_sage(x, ei, w, ws, wn, b2, None, True, False)), d["sum_w_concat"], "sum sage", **tol)
    _close(O(api.gcn_graph_sage(x, ei, w, kernel, b1, api.relu, False, cache=None)), d["gcn_sage_nocache"], "gcn sage", **tol)
    _close(O(api.gcn_graph_sage(x, ei, w, kernel, b1, api.relu, True, cache={"k": 1})), d["gcn_sage_cache"],
           "gcn sage (cache quirk)", **(dict(rtol=1e-6, atol_scale=1e-7) if api.exact_float else tol))
    _close(O(api.mean_pool_graph_sage(x, ei, w, ws, wm, wnk, bm, b2, api.relu)), d["mean_pool"], "mean pool", **tol)
    _close(O(api.max_pool_graph_sage(x, ei, w, ws, wm, wnk, bm, b2, api.relu)), d["max_pool"], "max pool", **tol)


def _replay_appnp(d, api):
    A, O = api.arr, api.out
    tol = dict(rtol=0, atol_scale=0) if api.exact_float else dict(rtol=1e-4, atol_scale=1e-4)
    ks, bs = [A(d["k0"]), A(d["k1"])], [A(d["b0"]), A(d["b1"])]
    _close(O(api.appnp(A(d["x"]), A(d["ei"]), A(d["w"]), ks, bs, k=10, alpha=0.1)), d["k10"], "appnp k=10", **tol)
    _close(O(api.appnp(A(d["x"]), A(d["ei"]), A(d["w"]), ks, bs, activation=api.relu, k=2, alpha=0.3)), d["k2_relu"],
           "appnp k=2 relu", **tol)


def _replay_propagation(d, api):
    A, O = api.arr, api.out
    tol = dict(rtol=0, atol_scale=0) if api.exact_float else dict(rtol=1e-4, atol_scale=1e-4)
    x, ei, w = A(d["x"]), A(d["ei"]), A(d["w"])
    _close(O(api.sgc(x, ei, w, 2, A(d["kernel"]), A(d["bias"]), api.relu)), d["sgc_k2"], "sgc k=2", **tol)
    _close(O(api.sgc(x, ei, w, 1, A(d["kernel"]), None, None, renorm=True, improved=True)), d["sgc_k1_improved"],
           "sgc improved", **tol)
    ks, bs = [A(d["k0"]), A(d["k1"])], [A(d["b0"]), A(d["b1"])]
    _close(O(api.ssgc(x, ei, w, ks, bs, k=5, alpha=0.2)), d["ssgc_k5"], "ssgc k=5",
           **(dict(rtol=1e-6, atol_scale=1e-7) if api.exact_float else tol))
    _close(O(api.ssgc(x, ei, w, None, None, k=3, alpha=0.1, activation=api.relu)), d["ssgc_nokernel"], "ssgc no kernel",
           **(dict(rtol=1e-6, atol_scale=1e-7) if api.exact_float else tol))
    _close(O(api.tagcn(x, ei, w, 3, A(d["tag_kernel"]), A(d["bias"]), api.relu)), d["tagcn_k3"], "tagcn", **tol)
    mlp_w = A(d["mlp_w"])
    if api.exact_float:
        mlp = lambda h: api.relu((h @ mlp_w).astype(np.float32))        # noqa: E731
        got = api.gin(x, ei, mlp, eps=0.25)
    else:
        mlp = lambda h, training=None: api.tfg.ops.gemm(h, mlp_w, act=api.tfg.ops.ACT_RELU)   # noqa: E731
        got = api.gin(x, ei, mlp, eps=0.25)
    _close(O(got), d["gin_eps"], "gin", **tol)
    _close(O(api.le_conv(x, ei, w, A(d["ws"]), A(d["bs"]), A(d["wa"]), A(d["ba"]), A(d["wn"]), None, api.relu)),
           d["le_conv"], "le_conv", **tol)
    _close(O(api.le_conv(x, ei, None, A(d["ws"]), None, A(d["wa"]), None, A(d["wn"]), None, None)), d["le_conv_now"],
           "le_conv unweighted", **tol)


def _replay_pool(d, api):
    A, O = api.arr, api.out
    g = int(d["g"])
    for name in ("mean_pool", "sum_pool", "max_pool", "min_pool"):
        got = O(getattr(api, name)(A(d["x"]), A(d["gi"]), g))
        _eq(got, d[name], name)                       # bit-exact: sequential fp32 sums in node order
    _eq(O(api.mean_pool(A(d["x"]), A(d["gi"]))), d["mean_pool_auto"], "mean_pool (num_graphs inferred)")


def _replay_chebynet(d, api):
    A, O = api.arr, api.out
    n = int(d["n"])
    tol = dict(rtol=0, atol_scale=0) if api.exact_float else dict(rtol=1e-4, atol_scale=1e-4)
    ks = [A(d["k0"]), A(d["k1"]), A(d["k2"]), A(d["k3"])]
    for tag, kk, nt in (("k1_sym", 1, "sym"), ("k2_sym", 2, "sym"), ("k4_sym", 4, "sym"), ("k3_rw", 3, "rw"), ("k3_none", 3, None)):
        got = api.chebynet(A(d["x"]), A(d["ei"]), A(d["w"]), kk, ks[:kk], A(d["bias"]), api.relu, normalization_type=nt)
        _close(O(got), d["cheb_" + tag], "chebynet " + tag, **tol)
    ni, nw = api.chebynet_norm_edge(A(d["ei"]), n, A(d["w"]), "sym")
    _eq(O(ni), d["norm_index"], "chebynet_norm_edge index")
    _close(O(nw), d["norm_w"], "chebynet_norm_edge weight", rtol=0 if api.exact_float else 6e-7, atol_scale=0)


def _replay_sampler(d, api):
    """RandomNeighborSampler (graph_utils.py:630-776): deterministic branches bit for bit; for the random branches the
    per-row sample counts the reference produced (they do not depend on its generator)."""
    A, O = api.arr, api.out
    ei, w = A(d["ei"]), A(d["w"])
    for tag, kw in (("all", {}), ("k_big", {"k": 1000}), ("subset_all", {"sampled_node_index": d["subset"]}),
                    ("pair_k_big", {"sampled_node_index": (d["rows_sub"], d["cols_sub"]), "k": 1000})):
        si, sw = api.neighbor_sample(ei, w, **kw)
        _eq(O(si), d[tag + "_index"], "sampler " + tag)
        _eq(O(sw), d[tag + "_w"], "sampler weights " + tag)
    for tag, kw in (("k3", {"k": 3}), ("k3_pad", {"k": 3, "padding": True}), ("k9_pad", {"k": 9, "padding": True}),
                    ("ratio", {"ratio": 0.4}), ("subset_k2", {"k": 2, "sampled_node_index": d["subset"]})):
        si, _ = api.neighbor_sample(ei, w, **kw)
        rows = O(si)[0]
        _eq(np.bincount(rows, minlength=int(rows.max()) + 1), d[tag + "_counts"], "sampler counts " + tag)


def _replay_pool2(d, api):
    """topk_pool (integer output: bit-exact) and set2set with a caller-supplied LSTM."""
    A, O = api.arr, api.out
    gi, score = A(d["gi"]), A(d["score"])
    for tag, kw in (("k1", {"k": 1}), ("k5", {"k": 5}), ("k1000", {"k": 1000}), ("r30", {"ratio": 0.3}), ("r100", {"ratio": 1.0})):
        _eq(O(api.topk_pool(gi, score, **kw)), d["topk_" + tag], "topk_pool " + tag)
    _eq(O(api.topk_pool(gi, A(d["score"].reshape(-1, 1)), ratio=0.5)), d["topk_col_r50"], "topk_pool column scores")
    lstm = api.lstm(d["lstm_k"], d["lstm_r"], d["lstm_b"])
    tol = dict(rtol=1e-6, atol_scale=1e-6) if api.exact_float else dict(rtol=1e-4, atol_scale=1e-4)
    _close(O(api.set2set(A(d["x"]), A(d["gi_sorted"]), lstm, 3)), d["set2set_it3"], "set2set 3 iterations", **tol)
    _close(O(api.set2set(A(d["x"]), gi, lstm, 2)), d["set2set_unsorted_it2"], "set2set unsorted graph ids", **tol)


def _replay_graph(d, api):
    """Induced subgraphs, batching, sag_pool and sort_pool as the reference's own data/graph.py and nn/pool code produced
    them (integer outputs and gathered rows: exact; gated features: 1e-6)."""
    A, O = api.arr, api.out
    x, ei, w, gi, y, keep = (A(d[k]) for k in ("x", "ei", "w", "gi", "y", "keep"))
    sx, sei, sw, sgi, sy = api.subgraph(x, ei, w, gi, y, keep)
    for got, key in ((sx, "sub_x"), (sei, "sub_ei"), (sw, "sub_w"), (sgi, "sub_gi"), (sy, "sub_y")):
        _eq(O(got), d[key], "sample_new_graph_by_node_index " + key)
    _eq(d["subnp_ei"], d["sub_ei"], "numpy and tensor containers agree in the reference")
    score = A(d["score"])
    for tag, kw in (("k4", {"k": 4}), ("r50", {"ratio": 0.5})):
        px, pei, pw, pgi = api.sag_pool_scored(x, ei, w, gi, score, **kw)
        _eq(O(pei), d["sag_%s_ei" % tag], "sag_pool edge_index " + tag)
        _eq(O(pgi), d["sag_%s_gi" % tag], "sag_pool node_graph_index " + tag)
        _eq(O(pw), d["sag_%s_w" % tag], "sag_pool edge_weight " + tag)
        _close(O(px), d["sag_%s_x" % tag], "sag_pool x " + tag, rtol=1e-6, atol_scale=1e-6)
        px, pei, pw, pgi = api.sort_pool(x, ei, w, gi, sort_index=1, **kw)
        for got, key in ((px, "x"), (pei, "ei"), (pw, "w"), (pgi, "gi")):
            _eq(O(got), d["sort_%s_%s" % (tag, key)], "sort_pool " + key + " " + tag)
    parts = [(A(d["part%d_x" % i]), A(d["part%d_ei" % i]), A(d["part%d_w" % i]), A(d["part%d_y" % i])) for i in range(3)]
    bx, bei, bw, by, bgi, begi = api.batch(parts)
    for got, key in ((bx, "batch_x"), (bei, "batch_ei"), (bw, "batch_w"), (by, "batch_y"), (bgi, "batch_gi"), (begi, "batch_egi")):
        _eq(O(got), d[key], "BatchGraph.from_graphs " + key)

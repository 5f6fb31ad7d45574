# coding=utf-8
"""Multi-head graph attention with the reference's functional signature (tf_geometric/nn/conv/gat.py:13-122).

The reference materialises Q[row] and K[col] ([E', A] each), builds a virtual graph of H*N nodes, runs a 5-pass
segment_softmax over H*E' scores and finally a SpMM.  Here: three dense projections, then ONE fused kernel
(tfgk_gat_fused_f32) that streams K and V rows once per edge.
"""
import torch

from ... import ops, _structure, _rng, autograd
from ...sparse import as_sparse_features, project_features


def project(x, blocks):
    """Dense projections of the same x: column blocks of at most 128 outputs, at most four per launch."""
    pieces = []
    for w, b, act, out in blocks:
        for c0 in range(0, w.shape[1], 128):
            c1 = min(c0 + 128, w.shape[1])
            pieces.append((w[:, c0:c1], None if b is None else b[c0:c1], act, out[:, c0:c1]))
    for i in range(0, len(pieces), 4):
        ops.gemm_proj(x, pieces[i:i + 4])


def gat(x, edge_index,
        query_kernel, query_bias, query_activation,
        key_kernel, key_bias, key_activation,
        kernel, bias=None, activation=None, num_heads=1,
        split_value_heads=True, edge_drop_rate=0.0, training=False, cache=None, return_attention=False, seed=None):
    """
    :param x: [num_nodes, num_features]
    :param edge_index: [2, num_edges]; self loops are appended (never de-duplicated), reference gat.py:43
    :param query_kernel/key_kernel: [num_features, attention_units]; query_bias/key_bias: [attention_units]
    :param kernel: [num_features, units] (or [num_features, units * num_heads] when split_value_heads=False)
    :param num_heads: heads; attention_units (and units when splitting) must be divisible by it
    :param split_value_heads: True: heads own slices of V and are concatenated; False: every head sees a full V and
        the head outputs are averaged
    :param edge_drop_rate: dropout on the attention coefficients while training (gat.py:85)
    :param cache: optional dict (e.g. graph.cache) memoising the self-looped CSR; an extension of the reference API
    :param seed: optional 64-bit key pinning the dropout mask (extension; default: a fresh key per call)
    :return: [num_nodes, units]
    """
    edge_index = ops.as_device(edge_index, torch.int32)
    dev = edge_index.device
    x_sparse = as_sparse_features(x)             # tf.SparseTensor features (reference gat.py:45-70)
    if x_sparse is None:
        x = ops.as_device(x, torch.float32, device=dev)
    num_nodes = len(x_sparse) if x_sparse is not None else x.shape[0]
    csr, edge_index_used = _structure.csr_for_edge_index(edge_index, num_nodes, add_self_loop=True, cache=cache)
    drop_rate = float(edge_drop_rate) if training else 0.0
    if x_sparse is not None:
        if drop_rate > 0.0 or autograd.needs_grad(query_kernel, query_bias, key_kernel, key_bias, kernel, bias):
            if return_attention:
                raise NotImplementedError("return_attention is an inference-path extension")
            return _gat_training(x_sparse, csr, edge_index_used, query_kernel, query_bias, query_activation, key_kernel,
                                 key_bias, key_activation, kernel, bias, activation, num_heads, split_value_heads, drop_rate,
                                 _rng.resolve(seed) if drop_rate > 0.0 else 0)
        q_act, q_left = ops.activation_code(query_activation)
        k_act, k_left = ops.activation_code(key_activation)
        f32 = lambda t: None if t is None else ops.as_device(t, torch.float32, device=dev)     # noqa: E731
        Q = project_features(x_sparse, query_kernel, bias=f32(query_bias), act=q_act)
        K = project_features(x_sparse, key_kernel, bias=f32(key_bias), act=k_act)
        V = project_features(x_sparse, kernel)
        Q = q_left(Q) if q_left is not None else Q
        K = k_left(K) if k_left is not None else K
        act_code, leftover = ops.activation_code(activation)
        res = ops.gat_fused(csr, Q, K, V, num_heads, split_value_heads=split_value_heads, bias=f32(bias), act=act_code,
                            return_attention=return_attention)
        h, att = res if return_attention else (res, None)
        h = leftover(h) if leftover is not None else h
        return (h, ops.permute(att, csr.perm, inverse=True)) if return_attention else h
    if drop_rate > 0.0 or autograd.needs_grad(x, query_kernel, query_bias, key_kernel, key_bias, kernel, bias):
        if return_attention:
            raise NotImplementedError("return_attention is an inference-path extension")
        return _gat_training(x, csr, edge_index_used, query_kernel, query_bias, query_activation, key_kernel, key_bias,
                             key_activation, kernel, bias, activation, num_heads, split_value_heads, drop_rate,
                             _rng.resolve(seed) if drop_rate > 0.0 else 0)

    q_act, q_left = ops.activation_code(query_activation)
    k_act, k_left = ops.activation_code(key_activation)
    # Q, K and V come out of ONE launch that reads x once (tfgk_gemm_proj_f32).  K and V land in ONE [N, A + U]
    # buffer: the fused kernel then fetches a neighbour's key and value from the same DRAM burst
    wq = ops.as_device(query_kernel, torch.float32, device=dev)
    wk = ops.as_device(key_kernel, torch.float32, device=dev)
    wv = ops.as_device(kernel, torch.float32, device=dev)
    a_units = wk.shape[1]
    Q = torch.empty((num_nodes, wq.shape[1]), dtype=torch.float32, device=dev)
    kv = torch.empty((num_nodes, a_units + wv.shape[1]), dtype=torch.float32, device=dev)
    K, V = kv[:, :a_units], kv[:, a_units:]
    project(x, [(wq, ops.as_device(query_bias, torch.float32, device=dev), q_act, Q),
                (wk, ops.as_device(key_bias, torch.float32, device=dev), k_act, K),
                (wv, None, ops.ACT_NONE, V)])
    if q_left is not None:
        Q = q_left(Q)
    if k_left is not None:
        K.copy_(k_left(K))

    act_code, leftover = ops.activation_code(activation)
    bias = None if bias is None else ops.as_device(bias, torch.float32, device=dev)
    res = ops.gat_fused(csr, Q, K, V, num_heads, split_value_heads=split_value_heads, bias=bias, act=act_code,
                        return_attention=return_attention)
    h, att = res if return_attention else (res, None)
    if leftover is not None:
        h = leftover(h)
    if return_attention:
        return h, ops.permute(att, csr.perm, inverse=True)     # [E', H] in edge_index-with-self-loops order
    return h


def _gat_training(x, csr, edge_index_used, query_kernel, query_bias, query_activation, key_kernel, key_bias,
                  key_activation, kernel, bias, activation, num_heads, split_value_heads, drop_rate, seed):
    """The same layer behind autograd Functions (demo/demo_gat.py trains through tf.GradientTape): dense projections
    with dX/dW/db GEMMs (sparse features: the aggregation kernel over x's pattern and its transpose), then GatAttention."""
    sparse_x = as_sparse_features(x)
    dev = csr.col.device

    def dense(w, b, act):
        code, left = ops.activation_code(act)
        w = ops.as_device(w, torch.float32, device=dev)
        b = None if b is None else ops.as_device(b, torch.float32, device=dev)
        y = autograd.Dense.apply(x, w, b, code) if sparse_x is None else autograd.SparseMatmul.apply(w, b, sparse_x, code)
        return left(y) if left is not None else y

    Q = dense(query_kernel, query_bias, query_activation)
    K = dense(key_kernel, key_bias, key_activation)
    V = dense(kernel, None, None)
    act_code, leftover = ops.activation_code(activation)
    bias = None if bias is None else ops.as_device(bias, torch.float32, device=dev)
    h = autograd.GatAttention.apply(Q, K, V, bias, csr, edge_index_used, int(num_heads), bool(split_value_heads),
                                    act_code, drop_rate, seed)
    return leftover(h) if leftover is not None else h

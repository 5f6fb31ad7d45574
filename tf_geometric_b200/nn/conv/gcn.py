# coding=utf-8
"""GCN normalisation and convolution with the reference's functional signatures (tf_geometric/nn/conv/gcn.py).

    gcn(x, A, W, b) = act( norm(A) @ (x @ W) + b )

norm(A) is computed once per graph by the integer/fp32 preprocessing kernels and memoised in `cache` (the reference
stores a numpy triple under the same key, gcn.py:125-128; here the cached object is a device SparseMatrix that also
carries the destination-sorted CSR, so a warm forward is exactly two launches: the dense projection and tfgk_spmm_f32
with bias + activation fused into its epilogue).
"""
import torch

from ... import ops, autograd
from ...sparse import SparseMatrix, as_sparse_features, project_features
from ..kernel.map_reduce import gcn_mapper  # noqa: F401  (re-exported like the reference)

CACHE_KEY_GCN_NORMED_ADJ_TEMPLATE = "gcn_normed_adj_{}_{}_{}_{}_{}"


def compute_cache_key(norm, add_self_loop, sym, renorm, improved):
    return CACHE_KEY_GCN_NORMED_ADJ_TEMPLATE.format(norm, add_self_loop, sym, renorm, improved)


def gcn_norm_adj(sparse_adj, norm="both", add_self_loop=True, sym=True, renorm=True, improved=False, cache=None):
    """
    Normalised adjacency for GCN (reference gcn.py:32-130).

    :param sparse_adj: SparseMatrix adjacency (row = aggregation target).
    :param norm: "both" (D^-1/2 A D^-1/2), "left" (D^-1 A) or "right" (A D^-1, with ROW sums - as the reference does).
    :param add_self_loop: add I * fill (fill = 2 if improved else 1); appended after the existing entries.
    :param sym: with norm="both", reuse the row degrees on the column side (valid for symmetric A).
    :param renorm: renormalisation trick: add the loops before normalising instead of after.
    :param cache: dict memoising the result under compute_cache_key(...).
    """
    if cache is not None:
        cache_key = compute_cache_key(norm, add_self_loop, sym, renorm, improved)
        cached = cache.get(cache_key, None)
        if cached is not None:
            if isinstance(cached, SparseMatrix):
                return cached
            return SparseMatrix(cached[0], cached[1], cached[2])      # a reference-style (index, value, shape) triple

    fill_weight = 2.0 if improved else 1.0
    if sparse_adj.shape[0] != sparse_adj.shape[1]:
        if add_self_loop:
            raise Exception("cannot set add_self_loop=True for GCN when sparse_adj.shape[0] != sparse_adj.shape[1]")
        if sym:
            raise Exception("cannot set sym=True for GCN when sparse_adj.shape[0] != sparse_adj.shape[1]")

    if add_self_loop and norm != "both":
        sparse_adj = sparse_adj.add_diag(fill_weight)

    if norm == "both":
        if add_self_loop and renorm:
            sparse_adj = sparse_adj.add_diag(fill_weight)
        row_dis = ops.deg_inv(sparse_adj.segment_sum(axis=-1), ops.POW_INV_SQRT)
        col_dis = row_dis if sym else ops.deg_inv(sparse_adj.segment_sum(axis=0), ops.POW_INV_SQRT)
        value = ops.scale_edges(sparse_adj.index[0].contiguous(), sparse_adj.index[1].contiguous(), sparse_adj.value,
                                dl=row_dis, dr=col_dis)
        normed = sparse_adj.with_value(value)
        if add_self_loop and not renorm:
            normed = normed.add_diag(fill_weight)
    elif norm == "left":
        row_inv = ops.deg_inv(sparse_adj.segment_sum(axis=-1), ops.POW_INV)
        normed = sparse_adj.with_value(ops.scale_edges(sparse_adj.index[0].contiguous(), None, sparse_adj.value,
                                                       dl=row_inv))
    elif norm == "right":
        col_inv = ops.deg_inv(sparse_adj.segment_sum(axis=-1), ops.POW_INV)     # row sums, literally as gcn.py:113
        normed = sparse_adj.with_value(ops.scale_edges(None, sparse_adj.index[1].contiguous(), sparse_adj.value,
                                                       dr=col_inv))
    else:
        raise Exception("wrong GCN norm type: {}".format(norm))

    if cache is not None:
        normed.csr, normed.value_csr    # build the CSR now: cached objects are always warm
        cache[cache_key] = normed
    return normed


def gcn_build_cache_by_adj(sparse_adj, norm="both", add_self_loop=True, sym=True, renorm=True, improved=False,
                           override=False, cache=None):
    """Compute norm(A) for this configuration and store it in `cache` (reference gcn.py:133-152)."""
    if cache is None:
        cache = {}
    elif override:
        cache[compute_cache_key(norm, add_self_loop, sym, renorm, improved)] = None
    gcn_norm_adj(sparse_adj, norm, add_self_loop, sym, renorm, improved, cache)
    return cache


def gcn_build_cache_for_graph(graph, norm="both", add_self_loop=True, sym=True, renorm=True, improved=False,
                              override=False):
    """reference gcn.py:155-169."""
    graph.cache = gcn_build_cache_by_adj(graph.adj(), norm=norm, add_self_loop=add_self_loop, sym=sym, renorm=renorm,
                                         improved=improved, override=override, cache=graph.cache)
    return graph.cache


def gcn_norm_edge(edge_index, num_nodes, edge_weight=None, renorm=True, improved=False, cache=None):
    """Deprecated edge-list form (reference gcn.py:180-196)."""
    sparse_adj = SparseMatrix(edge_index, edge_weight, [num_nodes, num_nodes])
    normed = gcn_norm_adj(sparse_adj, renorm=renorm, improved=improved, cache=cache)
    return normed.index, normed.value


def gcn(x, sparse_adj, kernel, bias=None, activation=None,
        norm="both", add_self_loop=True, sym=True, renorm=True, improved=False, edge_drop_rate=0.0,
        num_or_size_splits=None, training=False, cache=None):
    """
    Functional GCN layer (reference gcn.py:225-290).

    :param x: [num_nodes, num_features] float32
    :param sparse_adj: SparseMatrix adjacency
    :param kernel: [num_features, units] or None (propagate x itself)
    :param bias: [units] or None
    :param activation: callable or None; relu is fused into the aggregation epilogue
    :param num_or_size_splits: column chunks of the propagation (gcn.py:274-280): one launch per chunk into slices of one
        output, same bits (the fused kernel has no [E, D] temporary to bound, so this is an API-parity feature)
    :return: [num_nodes, units]
    """
    normed = gcn_norm_adj(sparse_adj, norm=norm, add_self_loop=add_self_loop, sym=sym, renorm=renorm,
                          improved=improved, cache=cache)
    normed = normed.dropout(edge_drop_rate, training=training)
    dev = normed.index.device
    act_code, leftover = ops.activation_code(activation)
    bias = None if bias is None else ops.as_device(bias, torch.float32, device=dev)
    x_sparse = as_sparse_features(x)
    if x_sparse is not None:                   # tf.SparseTensor features (gcn.py:269-272): sparse x dense projection
        if kernel is None:
            raise ValueError("a sparse feature matrix needs a kernel (reference gcn.py:266-272)")
        if autograd.needs_grad(kernel, bias):
            # training (demo/demo_gcn.py:60-75): dW = x^T dH is the same kernel over the transposed pattern of x
            h = autograd.propagate(x_sparse, ops.as_device(kernel, torch.float32, device=dev))
            h = autograd.SparseMatmul.apply(h, bias, normed, act_code)
            return leftover(h) if leftover is not None else h
        h = project_features(x_sparse, kernel)
        h = normed.matmul(h, num_or_size_splits=num_or_size_splits, bias=bias, act=act_code)
        return leftover(h) if leftover is not None else h
    x = ops.as_device(x, torch.float32, device=dev)
    if autograd.needs_grad(x, kernel, bias):
        # training path (demo/demo_gcn.py:60-75 uses tf.GradientTape): same kernels behind autograd Functions
        h = x if kernel is None else autograd.Dense.apply(x, ops.as_device(kernel, torch.float32, device=dev), None,
                                                         ops.ACT_NONE)
        h = autograd.SparseMatmul.apply(h, bias, normed, act_code)
        return leftover(h) if leftover is not None else h
    h = x if kernel is None else ops.gemm(x, ops.as_device(kernel, torch.float32, device=dev))
    h = normed.matmul(h, num_or_size_splits=num_or_size_splits, bias=bias, act=act_code)
    if leftover is not None:
        h = leftover(h)
    return h

# coding=utf-8
"""GraphSAGE aggregators with the reference's functional signatures (tf_geometric/nn/conv/graph_sage.py).

mean/sum aggregate the RAW features (D = num_features) with tfgk_spmm_f32, then project; the pooling variants apply
the neighbour MLP per NODE instead of per EDGE - relu(x[col] @ W + b) == relu(x @ W + b)[col], the same arithmetic
on E/N-times fewer rows - and reduce with the mean / max kernel.  Reference quirks are kept: a provided edge_weight
is replaced by ones in the gcn / pool variants (graph_sage.py:139-140,190-191,253-254), and gcn_graph_sage hands its
`cache` to gcn_norm_edge's `renorm` slot (graph_sage.py:142).
"""
import torch

from ... import ops, _structure, autograd
from .gcn import gcn_norm_edge
from ...sparse import SparseMatrix


def _project_pair(x, agg, self_kernel, neighbor_kernel, bias, activation, concat, normalize):
    """[x @ Ws || agg @ Wn] (+bias, act, l2) with both products written straight into the output columns."""
    dev = x.device
    ws = ops.as_device(self_kernel, torch.float32, device=dev)
    wn = ops.as_device(neighbor_kernel, torch.float32, device=dev)
    act_code, leftover = ops.activation_code(activation)
    b = None if bias is None else ops.as_device(bias, torch.float32, device=dev)
    if concat:
        u = ws.shape[1]
        out = torch.empty((x.shape[0], u + wn.shape[1]), dtype=torch.float32, device=dev)
        ops.gemm(x, ws, bias=None if b is None else b[:u].contiguous(), act=act_code, out=out[:, :u])
        ops.gemm(agg, wn, bias=None if b is None else b[u:].contiguous(), act=act_code, out=out[:, u:])
    else:
        out = ops.gemm(x, ws)
        ops.gemm(agg, wn, bias=b, act=act_code, beta=1.0, out=out)
    if leftover is not None:
        out = leftover(out)
    if normalize:
        out = ops.l2_normalize(out)
    return out


def _plain_sage_autograd(reduce, x, edge_index, edge_weight, ws, wn, bias, activation, concat, normalize):
    """Training path (any input requires grad): the same kernels wrapped in autograd Functions (autograd.py)."""
    act_code, leftover = ops.activation_code(activation)
    if leftover is None and not normalize and not autograd.needs_grad(edge_weight):
        return autograd.SagePair.apply(x, ws, wn, bias, edge_index, edge_weight, reduce, act_code, bool(concat))
    agg = autograd.NeighborAggregate.apply(x, edge_index, edge_weight, reduce, x.shape[0])
    return _project_pair_autograd(x, agg, ws, wn, bias, activation, concat, normalize)


def _project_pair_autograd(x, agg, ws, wn, bias, activation, concat, normalize):
    """Differentiable twin of _project_pair."""
    act_code, leftover = ops.activation_code(activation)
    if concat:
        u = ws.shape[1]
        left = autograd.Dense.apply(x, ws, None if bias is None else bias[:u], act_code)
        right = autograd.Dense.apply(agg, wn, None if bias is None else bias[u:], act_code)
        h = torch.cat([left, right], dim=1)
    else:
        h = autograd.Dense.apply(x, ws, None, ops.ACT_NONE) + autograd.Dense.apply(agg, wn, bias, ops.ACT_NONE)
        if act_code == ops.ACT_RELU:
            h = torch.relu(h)
    if leftover is not None:
        h = leftover(h)
    if normalize:
        h = h * torch.rsqrt(torch.clamp((h * h).sum(dim=-1, keepdim=True), min=1e-12))
    return h


def _plain_sage(reduce, x, edge_index, edge_weight, self_kernel, neighbor_kernel, bias, activation, concat, normalize):
    edge_index = ops.as_device(edge_index, torch.int32)
    dev = edge_index.device
    x = ops.as_device(x, torch.float32, device=dev)
    num_nodes = x.shape[0]
    if autograd.needs_grad(x, self_kernel, neighbor_kernel, bias):
        ew = None if edge_weight is None else ops.as_device(edge_weight, torch.float32, device=dev)
        return _plain_sage_autograd(reduce, x, edge_index, ew, ops.as_device(self_kernel, torch.float32, device=dev),
                                    ops.as_device(neighbor_kernel, torch.float32, device=dev),
                                    None if bias is None else ops.as_device(bias, torch.float32, device=dev),
                                    activation, concat, normalize)
    csr, _ = _structure.csr_for_edge_index(edge_index, num_nodes)
    w_csr = None
    if edge_weight is not None:
        w_csr = _structure.weights_in_csr_order(ops.as_device(edge_weight, torch.float32, device=dev), csr)
    agg = ops.spmm(csr, w_csr, x, reduce=reduce)
    return _project_pair(x, agg, self_kernel, neighbor_kernel, bias, activation, concat, normalize)


def mean_graph_sage(x, edge_index, edge_weight, self_kernel, neighbor_kernel, bias=None, activation=None,
                    concat=True, normalize=False):
    """h = act([x Ws || mean_{j in N(i)} (w_ij x_j) Wn] + b)  (reference graph_sage.py:9-60)."""
    return _plain_sage("mean", x, edge_index, edge_weight, self_kernel, neighbor_kernel, bias, activation, concat,
                       normalize)


def sum_graph_sage(x, edge_index, edge_weight, self_kernel, neighbor_kernel, bias=None, activation=None,
                   concat=True, normalize=False):
    """Sum aggregator (reference graph_sage.py:64-115)."""
    return _plain_sage("sum", x, edge_index, edge_weight, self_kernel, neighbor_kernel, bias, activation, concat,
                       normalize)


def gcn_graph_sage(x, edge_index, edge_weight, kernel, bias=None, activation=None, normalize=False, cache=None):
    """GCN aggregator (reference graph_sage.py:118-161): act((norm(A) x) W + b)."""
    edge_index = ops.as_device(edge_index, torch.int32)
    dev = edge_index.device
    x = ops.as_device(x, torch.float32, device=dev)
    num_nodes = x.shape[0]
    if edge_weight is not None:
        edge_weight = torch.ones([edge_index.shape[1]], dtype=torch.float32, device=dev)
    # reference :142 passes `cache` positionally into gcn_norm_edge(edge_index, num_nodes, edge_weight, renorm=...)
    normed = SparseMatrix(*_norm_edge_as_matrix(edge_index, num_nodes, edge_weight, renorm=bool(cache)))
    if autograd.needs_grad(x, kernel, bias):
        h = autograd.dense(autograd.propagate(normed, x), ops.as_device(kernel, torch.float32, device=dev),
                           None if bias is None else ops.as_device(bias, torch.float32, device=dev), activation)
        if normalize:
            h = h * torch.rsqrt(torch.clamp((h * h).sum(dim=-1, keepdim=True), min=1e-12))
        return h
    reduced = normed.matmul(x)
    act_code, leftover = ops.activation_code(activation)
    h = ops.gemm(reduced, ops.as_device(kernel, torch.float32, device=dev),
                 bias=None if bias is None else ops.as_device(bias, torch.float32, device=dev), act=act_code)
    if leftover is not None:
        h = leftover(h)
    if normalize:
        h = ops.l2_normalize(h)
    return h


def _norm_edge_as_matrix(edge_index, num_nodes, edge_weight, renorm):
    index, value = gcn_norm_edge(edge_index, num_nodes, edge_weight, renorm=renorm)
    return index, value, [num_nodes, num_nodes]


def _pool_sage(reduce, x, edge_index, edge_weight, self_kernel, neighbor_mlp_kernel, neighbor_kernel,
               neighbor_mlp_bias, bias, activation, concat, normalize):
    if edge_weight is None:
        # the reference multiplies by `edge_weight` unconditionally (gcn_mapper) and fails on None
        raise TypeError("edge_weight must not be None for the pooling GraphSAGE variants (reference behaviour)")
    edge_index = ops.as_device(edge_index, torch.int32)
    dev = edge_index.device
    x = ops.as_device(x, torch.float32, device=dev)
    num_nodes = x.shape[0]
    if autograd.needs_grad(x, self_kernel, neighbor_mlp_kernel, neighbor_kernel, neighbor_mlp_bias, bias):
        f32 = lambda t: None if t is None else ops.as_device(t, torch.float32, device=dev)   # noqa: E731
        h_node = autograd.dense(x, f32(neighbor_mlp_kernel), f32(neighbor_mlp_bias), activation)
        if reduce == "mean":
            reduced = autograd.NeighborAggregate.apply(h_node, edge_index, None, "mean", num_nodes)
        else:
            # training only: the per-edge messages are materialised like the reference does (graph_sage.py:262-270), so that
            # the max can route its gradient to the selected neighbours (ties share it, TF's UnsortedSegmentMax gradient)
            messages = autograd.TakeRows.apply(h_node, edge_index[1].contiguous())
            reduced = autograd.SegmentReduce.apply(messages, edge_index[0].contiguous(), num_nodes, "max")
        return _project_pair_autograd(x, reduced, f32(self_kernel), f32(neighbor_kernel), f32(bias), activation, concat,
                                      normalize)
    csr, _ = _structure.csr_for_edge_index(edge_index, num_nodes)
    act_code, leftover = ops.activation_code(activation)
    # per-node neighbour MLP (weights are all ones, so x[col] * w == x[col])
    h_node = ops.gemm(x, ops.as_device(neighbor_mlp_kernel, torch.float32, device=dev),
                      bias=None if neighbor_mlp_bias is None else ops.as_device(neighbor_mlp_bias, torch.float32, device=dev),
                      act=act_code)
    if leftover is not None:
        h_node = leftover(h_node)
    reduced = ops.spmm(csr, None, h_node, reduce=reduce)
    return _project_pair(x, reduced, self_kernel, neighbor_kernel, bias, activation, concat, normalize)


def mean_pool_graph_sage(x, edge_index, edge_weight, self_kernel, neighbor_mlp_kernel, neighbor_kernel,
                         neighbor_mlp_bias=None, bias=None, activation=None, concat=True, normalize=False):
    """Mean-pooling aggregator (reference graph_sage.py:164-225)."""
    return _pool_sage("mean", x, edge_index, edge_weight, self_kernel, neighbor_mlp_kernel, neighbor_kernel,
                      neighbor_mlp_bias, bias, activation, concat, normalize)


def max_pool_graph_sage(x, edge_index, edge_weight, self_kernel, neighbor_mlp_kernel, neighbor_kernel,
                        neighbor_mlp_bias=None, bias=None, activation=None, concat=True, normalize=False):
    """Max-pooling aggregator (reference graph_sage.py:228-287); nodes without in-edges get float32 lowest."""
    return _pool_sage("max", x, edge_index, edge_weight, self_kernel, neighbor_mlp_kernel, neighbor_kernel,
                      neighbor_mlp_bias, bias, activation, concat, normalize)

# coding=utf-8
"""APPNP with the reference's functional signature (tf_geometric/nn/conv/appnp.py:11-92):
an MLP followed by k steps of  out <- norm(A) out (1 - alpha) + h alpha, each step ONE tfgk_spmm_f32 launch with
the teleport term fused into the epilogue."""
import torch

from ... import ops
from ...sparse import SparseMatrix
from .gcn import gcn_norm_adj


def appnp(x, edge_index, edge_weight, kernels, biases,
          dense_activation=ops.relu, activation=None,
          k=10, alpha=0.1,
          dense_drop_rate=0.0, last_dense_drop_rate=0.0, edge_drop_rate=0.0,
          cache=None, training=False):
    """
    :param kernels/biases: weights of the dense layers; every layer but the last is followed by dense_activation
    :param k: number of propagation steps; alpha: teleport probability
    :param cache: dict memoising norm(A) (build it with gcn_build_cache_for_graph, like for GCN)
    """
    if training and (dense_drop_rate > 0.0 or last_dense_drop_rate > 0.0 or edge_drop_rate > 0.0):
        raise NotImplementedError("dropout (TF RNG stream) is outside the forward hot path of this backend")
    edge_index = ops.as_device(edge_index, torch.int32)
    dev = edge_index.device
    x = ops.as_device(x, torch.float32, device=dev)
    num_nodes = x.shape[0]
    normed = gcn_norm_adj(SparseMatrix(edge_index, edge_weight, [num_nodes, num_nodes]), cache=cache)

    h = x
    num_dense = len(kernels)
    for i, (kern, b) in enumerate(zip(kernels, biases)):
        act_code, leftover = ops.activation_code(dense_activation if i < num_dense - 1 else None)
        h = ops.gemm(h, ops.as_device(kern, torch.float32, device=dev),
                     bias=None if b is None else ops.as_device(b, torch.float32, device=dev), act=act_code)
        if leftover is not None:
            h = leftover(h)

    act_code, leftover = ops.activation_code(activation)
    out = h
    bufs = [torch.empty_like(h), torch.empty_like(h)]
    for i in range(k):
        last = i == k - 1
        out = normed.matmul(out, alpha=1.0 - alpha, addend=h, beta=alpha, act=act_code if last else ops.ACT_NONE,
                            out=bufs[i % 2])
    if k == 0 and act_code != ops.ACT_NONE:
        out = torch.relu(out)
    if leftover is not None:
        out = leftover(out)
    return out

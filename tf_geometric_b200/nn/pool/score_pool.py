# coding=utf-8
"""Score-based node pooling: sag_pool (reference nn/pool/sag_pool.py:7-47) and sort_pool (nn/pool/sort_pool.py:7-37).

Both rank the nodes of every graph by one score column, keep the best k (or a ratio) per graph and continue on the induced
subgraph.  On the device that is one radix argsort + one stable CSR build (topk_pool) and one flag/compaction pass
(BatchGraph.sample_new_graph_by_node_index); nothing visits the host except the output sizes."""
import torch

from ... import ops
from ...data.graph import BatchGraph
from .topk_pool import topk_pool


def _keep_best_nodes(features, score, edge_index, edge_weight, node_graph_index, k, ratio):
    """(pooled_x, pooled_edge_index, pooled_edge_weight, pooled_node_graph_index) of the top-scored nodes."""
    chosen = topk_pool(node_graph_index, score, k=k, ratio=ratio)
    batch = BatchGraph(x=features, edge_index=edge_index, node_graph_index=node_graph_index, edge_graph_index=None,
                       edge_weight=edge_weight)
    small = batch.sample_new_graph_by_node_index(chosen)
    return small.x, small.edge_index, small.edge_weight, small.node_graph_index


def _on_device(x, edge_index, node_graph_index):
    edge_index = ops.as_device(edge_index, torch.int32)
    dev = edge_index.device
    return (ops.as_device(x, torch.float32, device=dev), edge_index,
            ops.as_device(node_graph_index, torch.int32, device=dev))


def sag_pool(x, edge_index, edge_weight, node_graph_index, score_gnn, k=None, ratio=None, score_activation=None,
             training=None, cache=None):
    """Self-attention graph pooling.  `score_gnn([x, edge_index, edge_weight], training=..., cache=...)` returns one score
    per node ([num_nodes, 1], e.g. tfg.layers.GCN(1)); the kept nodes' features are gated by `score_activation(score)`."""
    x, edge_index, node_graph_index = _on_device(x, edge_index, node_graph_index)
    extra = {} if cache is None else {"cache": cache}
    score = score_gnn([x, edge_index, edge_weight], training=training, **extra)
    gate = score if score_activation is None else score_activation(score)      # ranking uses the raw score (sag_pool.py:31-34)
    return _keep_best_nodes(x * gate, score, edge_index, edge_weight, node_graph_index, k, ratio)


def sort_pool(x, edge_index, edge_weight, node_graph_index, k=None, ratio=None, sort_index=-1, training=None):
    """SortPool: the score is feature column `sort_index`; features are passed through unchanged."""
    x, edge_index, node_graph_index = _on_device(x, edge_index, node_graph_index)
    return _keep_best_nodes(x, x[:, sort_index].contiguous(), edge_index, edge_weight, node_graph_index, k, ratio)

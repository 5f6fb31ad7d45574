# coding=utf-8
"""SortPool (reference nn/pool/sort_pool.py:7-37): keep the top-k nodes of every graph ranked by one feature column and
take the induced subgraph - topk_pool + BatchGraph.sample_new_graph_by_node_index, both on the device."""
import torch

from ... import ops
from ...data.graph import BatchGraph
from .topk_pool import topk_pool


def sort_pool(x, edge_index, edge_weight, node_graph_index, k=None, ratio=None, sort_index=-1, training=None):
    """
    :param sort_index: the feature column used as the score
    :return: [pooled_x, pooled_edge_index, pooled_edge_weight, pooled_node_graph_index]
    """
    edge_index = ops.as_device(edge_index, torch.int32)
    dev = edge_index.device
    x = ops.as_device(x, torch.float32, device=dev)
    node_graph_index = ops.as_device(node_graph_index, torch.int32, device=dev)
    score = x[:, sort_index].contiguous()
    topk_node_index = topk_pool(node_graph_index, score, k=k, ratio=ratio)
    pooled = BatchGraph(x=x, edge_index=edge_index, node_graph_index=node_graph_index, edge_graph_index=None,
                        edge_weight=edge_weight).sample_new_graph_by_node_index(topk_node_index)
    return pooled.x, pooled.edge_index, pooled.edge_weight, pooled.node_graph_index

# coding=utf-8
"""Self-attention graph pooling (reference nn/pool/sag_pool.py:7-47): score the nodes with a GNN, keep the top-k nodes of
every graph, gate their features with the score and take the induced subgraph - all on the device (topk_pool,
BatchGraph.sample_new_graph_by_node_index)."""
import torch

from ... import ops
from ...data.graph import BatchGraph
from .topk_pool import topk_pool


def sag_pool(x, edge_index, edge_weight, node_graph_index, score_gnn, k=None, ratio=None, score_activation=None,
             training=None, cache=None):
    """
    :param score_gnn: callable [x, edge_index, edge_weight] => node_score [num_nodes, 1] (e.g. tfg.layers.GCN(1))
    :return: [pooled_x, pooled_edge_index, pooled_edge_weight, pooled_node_graph_index]
    """
    edge_index = ops.as_device(edge_index, torch.int32)
    dev = edge_index.device
    x = ops.as_device(x, torch.float32, device=dev)
    if cache is None:
        node_score = score_gnn([x, edge_index, edge_weight], training=training)
    else:
        node_score = score_gnn([x, edge_index, edge_weight], training=training, cache=cache)
    topk_node_index = topk_pool(node_graph_index, node_score, k=k, ratio=ratio)
    if score_activation is not None:
        node_score = score_activation(node_score)
    pooled = BatchGraph(x=x * node_score, edge_index=edge_index, node_graph_index=ops.as_device(node_graph_index, torch.int32,
                                                                                              device=dev),
                        edge_graph_index=None, edge_weight=edge_weight).sample_new_graph_by_node_index(topk_node_index)
    return pooled.x, pooled.edge_index, pooled.edge_weight, pooled.node_graph_index

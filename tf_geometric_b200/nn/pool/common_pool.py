# coding=utf-8
"""Graph-level pooling: mean/sum/max/min_pool over `node_graph_index`, with the signatures of
tf_geometric/nn/pool/common_pool.py.  They are the segment reductions of the K1 kernel keyed by the graph id of every node."""
import torch

from ... import ops
from ..kernel.map_reduce import _segment_reduce


def _num_graphs(node_graph_index, num_graphs):
    if num_graphs is None:
        num_graphs = int(ops.as_device(node_graph_index, torch.int32).max().item()) + 1
    return int(num_graphs)


def mean_pool(x, node_graph_index, num_graphs=None):
    """sum / (count + 1e-8) (reference common_pool.py:7-12).  In fp32 `count + 1e-8 == count` for count >= 1 and an empty
    graph gives 0 / 1e-8 = 0, which is exactly the kernel's MEAN reduce (sum / max(count, 1))."""
    return _segment_reduce(x, node_graph_index, _num_graphs(node_graph_index, num_graphs), "mean")


def sum_pool(x, node_graph_index, num_graphs=None):
    """tf.math.unsorted_segment_sum (reference common_pool.py:15-19)."""
    return _segment_reduce(x, node_graph_index, _num_graphs(node_graph_index, num_graphs), "sum")


def max_pool(x, node_graph_index, num_graphs=None):
    """tf.math.unsorted_segment_max, empty graph -> float32 lowest (reference common_pool.py:39-43)."""
    return _segment_reduce(x, node_graph_index, _num_graphs(node_graph_index, num_graphs), "max")


def min_pool(x, node_graph_index, num_graphs=None):
    """tf.math.unsorted_segment_min, empty graph -> float32 max (reference common_pool.py:45-49): computed in one launch
    as -max(-x) (both negations are exact: edge weight -1 and epilogue scale -1)."""
    return _segment_reduce(x, node_graph_index, _num_graphs(node_graph_index, num_graphs), "min")

# coding=utf-8
"""Pooling layers with the reference's constructors and `inputs` conventions (layers/pool/common_pool.py,
layers/pool/set2set.py:7-38, layers/pool/sag_pool.py:8-44)."""
import torch

from ...nn.pool.common_pool import mean_pool, sum_pool, max_pool, min_pool
from ...nn.pool.set2set import set2set
from ...nn.pool.score_pool import sag_pool, sort_pool
from .._base import Layer


class _CommonPool(torch.nn.Module):
    pool_func = None

    def forward(self, inputs, training=None, mask=None):
        """inputs: [x, node_graph_index] or [x, node_graph_index, num_graphs]."""
        if len(inputs) == 2:
            x, node_graph_index = inputs
            num_graphs = None
        else:
            x, node_graph_index, num_graphs = inputs
        return type(self).pool_func(x, node_graph_index, num_graphs)


class MeanPool(_CommonPool):
    pool_func = staticmethod(mean_pool)


class SumPool(_CommonPool):
    pool_func = staticmethod(sum_pool)


class MaxPool(_CommonPool):
    pool_func = staticmethod(max_pool)


class MinPool(_CommonPool):
    pool_func = staticmethod(min_pool)


class _KerasStyleLSTM(torch.nn.Module):
    """torch.nn.LSTM behind the calling convention of tf.keras.layers.LSTM(units, return_sequences=True,
    return_state=True): lstm(inputs[batch, steps, features], initial_state=[h, c]) -> (sequence, h, c).  The recurrence
    itself is the library's (cuDNN), exactly as the reference leaves it to Keras."""

    def __init__(self, input_size, units, device=None):
        super().__init__()
        self.cell = torch.nn.LSTM(input_size, units, batch_first=True, device=device)

    def forward(self, inputs, initial_state=None, training=None):
        state = None if initial_state is None else (initial_state[0].unsqueeze(0), initial_state[1].unsqueeze(0))
        seq, (h, c) = self.cell(inputs, state)
        return seq, h.squeeze(0), c.squeeze(0)


class Set2Set(Layer):
    """inputs: [x, node_graph_index] -> [num_graphs, 2 * num_features]."""

    def __init__(self, num_iterations=4, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.num_iterations = num_iterations
        self.lstm = None

    def build(self, input_shapes, device=None):
        num_features = input_shapes[0][-1]
        self.__dict__.pop("lstm", None)
        self.lstm = _KerasStyleLSTM(2 * num_features, num_features, device=device)
        for p in self.lstm.parameters():
            p.requires_grad_(self._trainable)

    def call(self, inputs, cache=None, training=None, mask=None):
        x, node_graph_index = inputs
        return set2set(x, node_graph_index, self.lstm, self.num_iterations, training=training)


class SAGPool(torch.nn.Module):
    """inputs: [x, edge_index, edge_weight, node_graph_index] -> [pooled_x, pooled_edge_index, pooled_edge_weight,
    pooled_node_graph_index]."""

    def __init__(self, score_gnn, k=None, ratio=None, score_activation=None):
        super().__init__()
        self.score_gnn = score_gnn
        self.k = k
        self.ratio = ratio
        self.score_activation = score_activation

    def forward(self, inputs, cache=None, training=None, mask=None):
        x, edge_index, edge_weight, node_graph_index = inputs
        return sag_pool(x, edge_index, edge_weight, node_graph_index, self.score_gnn, k=self.k, ratio=self.ratio,
                        score_activation=self.score_activation, training=training, cache=cache)


class SortPool(torch.nn.Module):
    """inputs: [x, edge_index, edge_weight, node_graph_index] -> the pooled graph (layers/pool/sort_pool.py:7-40)."""

    def __init__(self, k=None, ratio=None, sort_index=-1):
        super().__init__()
        self.k = k
        self.ratio = ratio
        self.sort_index = sort_index

    def forward(self, inputs, training=None, mask=None):
        x, edge_index, edge_weight, node_graph_index = inputs
        return sort_pool(x, edge_index, edge_weight, node_graph_index, k=self.k, ratio=self.ratio,
                         sort_index=self.sort_index, training=training)

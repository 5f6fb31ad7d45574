# coding=utf-8
"""tfg.layers.{Mean,Sum,GCN,MeanPool,MaxPool}GraphSage (reference layers/conv/graph_sage.py:8-356).
LSTMGraphSage (Keras LSTM) is not on the kernel hot path and is out of scope."""
from ... import ops
from ...nn.conv.graph_sage import (mean_graph_sage, sum_graph_sage, gcn_graph_sage, mean_pool_graph_sage,
                                   max_pool_graph_sage)
from .._base import Layer


def _unpack(inputs):
    if len(inputs) == 3:
        return inputs
    x, edge_index = inputs
    return x, edge_index, None


class _PairSage(Layer):
    _fn = None

    def __init__(self, units, activation=ops.relu, use_bias=True, concat=True, normalize=False,
                 kernel_regularizer=None, bias_regularizer=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.units = units
        self.activation = activation
        self.use_bias = use_bias
        self.concat = concat
        self.normalize = normalize
        if concat and (units % 2 != 0):
            raise Exception("units must be a event number if concat is True")
        self.kernel_regularizer = kernel_regularizer
        self.bias_regularizer = bias_regularizer
        self.self_kernel = None
        self.neighbor_kernel = None
        self.bias = None

    def build(self, input_shapes, device=None):
        num_features = input_shapes[0][-1]
        kernel_units = self.units // 2 if self.concat else self.units
        self.self_kernel = self.add_weight("self_kernel", [num_features, kernel_units], device=device)
        self.neighbor_kernel = self.add_weight("neighbor_kernel", [num_features, kernel_units], device=device)
        if self.use_bias:
            self.bias = self.add_weight("bias", [self.units], "zeros", device=device)

    def call(self, inputs, cache=None, training=None, mask=None):
        x, edge_index, edge_weight = _unpack(inputs)
        return type(self)._fn(x, edge_index, edge_weight, self.self_kernel, self.neighbor_kernel, bias=self.bias,
                              activation=self.activation, concat=self.concat, normalize=self.normalize)


class MeanGraphSage(_PairSage):
    _fn = staticmethod(mean_graph_sage)


class SumGraphSage(_PairSage):
    _fn = staticmethod(sum_graph_sage)


class GCNGraphSage(Layer):

    def __init__(self, units, activation=ops.relu, use_bias=True, normalize=False,
                 kernel_regularizer=None, bias_regularizer=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.units = units
        self.activation = activation
        self.use_bias = use_bias
        self.normalize = normalize
        self.kernel_regularizer = kernel_regularizer
        self.bias_regularizer = bias_regularizer
        self.kernel = None
        self.bias = None

    def build(self, input_shapes, device=None):
        num_features = input_shapes[0][-1]
        self.kernel = self.add_weight("kernel", [num_features, self.units], device=device)
        if self.use_bias:
            self.bias = self.add_weight("bias", [self.units], "zeros", device=device)

    def call(self, inputs, cache=None, training=None, mask=None):
        x, edge_index, edge_weight = _unpack(inputs)
        return gcn_graph_sage(x, edge_index, edge_weight, self.kernel, self.bias, self.activation, self.normalize,
                              cache=cache)


class _PoolSage(Layer):
    _fn = None
    _names = ("neighbor_mlp_kernel", "neighbor_mlp_bias", "neighbor_kernel")

    def __init__(self, units, activation=ops.relu, use_bias=True, concat=True, normalize=False,
                 kernel_regularizer=None, bias_regularizer=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.units = units
        self.activation = activation
        self.use_bias = use_bias
        self.concat = concat
        if concat and (units % 2 != 0):
            raise Exception("units must be a event number if concat is True")
        self.normalize = normalize
        self.kernel_regularizer = kernel_regularizer
        self.bias_regularizer = bias_regularizer
        self.self_kernel = None
        self.bias = None

    def build(self, input_shapes, device=None):
        num_features = input_shapes[0][-1]
        kernel_units = self.units // 2 if self.concat else self.units
        mlp_k, mlp_b, neigh_k = self._names
        self.self_kernel = self.add_weight("self_kernel", [num_features, kernel_units], device=device)
        self.add_weight(mlp_k, [num_features, kernel_units * 4], device=device)
        if self.use_bias:
            self.add_weight(mlp_b, [kernel_units * 4], "zeros", device=device)
        self.add_weight(neigh_k, [kernel_units * 4, kernel_units], device=device)
        if self.use_bias:
            self.bias = self.add_weight("bias", [self.units], "zeros", device=device)

    def call(self, inputs, cache=None, training=None, mask=None):
        x, edge_index, edge_weight = _unpack(inputs)
        mlp_k, mlp_b, neigh_k = self._names
        return type(self)._fn(x, edge_index, edge_weight, self.self_kernel, getattr(self, mlp_k), getattr(self, neigh_k),
                              neighbor_mlp_bias=getattr(self, mlp_b) if self.use_bias else None, bias=self.bias,
                              activation=self.activation, concat=self.concat, normalize=self.normalize)


class MeanPoolGraphSage(_PoolSage):
    _fn = staticmethod(mean_pool_graph_sage)


class MaxPoolGraphSage(_PoolSage):
    # weight names of the reference's MaxPoolGraphSage.build (layers/conv/graph_sage.py:327-338)
    _names = ("mlp_kernel", "mlp_bias", "neighs_kernel")
    _fn = staticmethod(max_pool_graph_sage)

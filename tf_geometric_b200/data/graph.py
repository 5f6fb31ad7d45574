# coding=utf-8
"""Graph container with the reference's surface (data/graph.py:20-359 of tf_geometric): x, edge_index (int32 [2,E]),
edge_weight (float32, ones by default), y, `cache` dict, adj(), to_directed().  BatchGraph / HeteroGraph are out
of scope.  Data may be numpy (kept as numpy, like the reference) or torch tensors; the kernels
receive CUDA tensors - call `to_device()` (the analogue of `convert_data_to_tensor`) once to avoid per-call copies.
"""
import types

import numpy as np
import torch

from .. import ops
from ..sparse import SparseMatrix
from ..utils.graph_utils import convert_edge_to_directed


def _get_shape(data):
    return None if data is None else tuple(data.shape)


class Graph(object):

    def __init__(self, x, edge_index, y=None, edge_weight=None):
        self._x = Graph.cast_x(x)
        self.edge_index = Graph.cast_edge_index(edge_index)
        self.y = Graph.cast_y(y)
        self.cache = {}
        if edge_weight is not None:
            self.edge_weight = self.cast_edge_weight(edge_weight)
        elif torch.is_tensor(self.edge_index):
            self.edge_weight = torch.ones([self.num_edges], dtype=torch.float32, device=self.edge_index.device)
        else:
            self.edge_weight = np.ones([self.num_edges], dtype=np.float32)

    # ---- casting rules: int32 ids, float32 weights, float64 features demoted to float32 (reference :58-91) ----
    @classmethod
    def cast_edge_index(cls, x):
        if isinstance(x, list):
            x = np.array(x).astype(np.int32)
        elif isinstance(x, np.ndarray):
            x = x.astype(np.int32)
        elif torch.is_tensor(x):
            x = x.to(torch.int32)
        return x

    @classmethod
    def cast_edge_weight(cls, edge_weight):
        if isinstance(edge_weight, list):
            edge_weight = np.array(edge_weight).astype(np.float32)
        elif isinstance(edge_weight, np.ndarray):
            edge_weight = edge_weight.astype(np.float32)
        elif torch.is_tensor(edge_weight) and not isinstance(edge_weight, torch.nn.Parameter):
            edge_weight = edge_weight.to(torch.float32)
        return edge_weight

    @classmethod
    def cast_x(cls, x):
        if isinstance(x, list):
            x = np.array(x)
        if isinstance(x, np.ndarray) and x.dtype == np.float64:
            x = x.astype(np.float32)
        elif torch.is_tensor(x) and not isinstance(x, torch.nn.Parameter) and x.dtype == torch.float64:
            x = x.to(torch.float32)
        return x

    @classmethod
    def cast_y(cls, y):
        if isinstance(y, list):
            y = np.array(y)
        return y

    @property
    def x(self):
        return self._x() if isinstance(self._x, types.FunctionType) else self._x

    @x.setter
    def x(self, value):
        self._x = Graph.cast_x(value)

    @property
    def num_nodes(self):
        return len(self.x)

    @property
    def num_edges(self):
        if len(self.edge_index) == 0:
            return 0
        return int(self.edge_index.shape[1]) if hasattr(self.edge_index, "shape") else len(self.edge_index[0])

    @property
    def num_features(self):
        return self.x.shape[-1]

    def get_shape_desc(self):
        return "Graph Shape: x => {}\tedge_index => {}\ty => {}".format(
            _get_shape(self.x), _get_shape(self.edge_index), _get_shape(self.y))

    def __str__(self):
        return self.get_shape_desc()

    __repr__ = __str__

    def adj(self):
        """SparseMatrix(edge_index, edge_weight, [N, N]) - reference :208-210."""
        n = self.num_nodes
        return SparseMatrix(self.edge_index, self.edge_weight, shape=[n, n])

    def to_device(self, device=None, inplace=False):
        """All graph data as CUDA tensors (the analogue of the reference's convert_data_to_tensor, :212-233)."""
        g = self if inplace else Graph(self._x, self.edge_index, y=self.y, edge_weight=self.edge_weight)
        dev = device if device is not None else ops.default_device()
        g._x = ops.as_device(g.x, torch.float32, device=dev) if not isinstance(g._x, types.FunctionType) else g._x
        g.edge_index = ops.as_device(g.edge_index, torch.int32, device=dev)
        g.edge_weight = ops.as_device(g.edge_weight, torch.float32, device=dev)
        if g.y is not None and not torch.is_tensor(g.y):
            g.y = torch.from_numpy(np.ascontiguousarray(g.y)).to(dev)
        elif g.y is not None:
            g.y = g.y.to(dev)
        if not inplace:
            g.cache = {}
        return g

    convert_data_to_tensor = to_device

    def to_directed(self, merge_mode="sum", inplace=False):
        """(u,v) -> (u,v) and (v,u) with duplicate merging - reference :235-253 / utils/graph_utils.py:155-212."""
        edge_index, [edge_weight] = convert_edge_to_directed(self.edge_index, [self.edge_weight],
                                                             merge_modes=[merge_mode])
        if inplace:
            self.edge_index, self.edge_weight = edge_index, edge_weight
            self.cache = {}
            return self
        return Graph(self.x, edge_index, y=self.y, edge_weight=edge_weight)

    def convert_edge_to_directed(self, merge_mode="sum"):
        return self.to_directed(merge_mode=merge_mode, inplace=True)

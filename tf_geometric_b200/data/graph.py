# coding=utf-8
"""Graph container with the reference's surface (data/graph.py:20-359 of tf_geometric): x, edge_index (int32 [2,E]),
edge_weight (float32, ones by default), y, `cache` dict, adj(), to_directed(), sample_new_graph_by_node_index(), and
BatchGraph (data/graph.py:362-621; what the pooling layers need).  HeteroGraph is out of scope.
Data may be numpy (kept as numpy, like the reference) or torch tensors; the kernels
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

    def sample_new_graph_by_node_index(self, sampled_node_index):
        """The subgraph induced by `sampled_node_index`, nodes relabelled by their position in that list, edge order
        kept (reference :276-359).  Device data: Bernoulli-free edge flags + stable compaction + gathers
        (tfgk_edge_flags_i32 / tfgk_select_flagged_i32); numpy data: numpy, like the reference's eager path."""
        is_batch = isinstance(self, BatchGraph)
        on_device = torch.is_tensor(self.edge_index) or torch.is_tensor(self.x)
        if on_device:
            dev = self.edge_index.device if torch.is_tensor(self.edge_index) else self.x.device
            idx = ops.as_device(sampled_node_index, torch.int32, device=dev).reshape(-1)
            take = lambda d: None if d is None else _take_rows(ops.as_device(d, device=dev), idx)      # noqa: E731
        else:
            idx = np.asarray(sampled_node_index).reshape(-1)
            take = lambda d: None if d is None else np.asarray(d)[idx]                                  # noqa: E731
        x, y = take(self.x), take(self.y)
        node_graph_index = take(self.node_graph_index) if is_batch else None
        edge_index, edge_weight = self.edge_index, self.edge_weight
        edge_graph_index = self.edge_graph_index if is_batch else None
        if edge_index is not None and self.num_edges > 0:
            if on_device:
                ei = ops.as_device(edge_index, torch.int32, device=dev)
                row, col = ei[0].contiguous(), ei[1].contiguous()
                n_ids = max(int(ei.max().item()), int(idx.max().item()) if idx.numel() else 0) + 1
                reverse = torch.full((n_ids,), -1, dtype=torch.int32, device=dev)
                reverse[idx.long()] = torch.arange(idx.numel(), dtype=torch.int32, device=dev)
                kept = ops.select_flagged(ops.edge_flags(row, col, row.numel(), mode=ops.FLAG_MAPPED, row_map=reverse,
                                                         col_map=reverse))
                edge_index = torch.stack([ops.gather_i32(reverse, ops.gather_i32(row, kept)),
                                          ops.gather_i32(reverse, ops.gather_i32(col, kept))])
                by_edge = lambda d: None if d is None else _take_rows(ops.as_device(d, device=dev), kept)  # noqa: E731
            else:
                ei = np.asarray(edge_index)
                n_ids = max(int(ei.max()), int(idx.max()) if idx.size else 0) + 1
                reverse = -np.ones(n_ids, np.int32)
                reverse[idx] = np.arange(idx.size, dtype=np.int32)
                mask = (reverse[ei[0]] >= 0) & (reverse[ei[1]] >= 0)
                edge_index = np.stack([reverse[ei[0][mask]], reverse[ei[1][mask]]]).astype(np.int32)
                by_edge = lambda d: None if d is None else np.asarray(d)[mask]                           # noqa: E731
            edge_weight, edge_graph_index = by_edge(edge_weight), by_edge(edge_graph_index)
        if is_batch:
            return BatchGraph(x=x, edge_index=edge_index, node_graph_index=node_graph_index,
                              edge_graph_index=edge_graph_index, y=y, edge_weight=edge_weight)
        return Graph(x=x, edge_index=edge_index, y=y, edge_weight=edge_weight)


def _take_rows(data, index):
    """data[index] along axis 0 on the device: float32 / int32 through the gather kernel, other dtypes through torch."""
    if data.dtype == torch.float32 and data.dim() <= 2 and data.is_contiguous():
        if data.requires_grad and torch.is_grad_enabled():        # sag_pool / sort_pool inside a trained model
            from .. import autograd
            return autograd.TakeRows.apply(data, index)
        return ops.permute(data, index)
    if data.dtype == torch.int32 and data.dim() == 1 and data.is_contiguous():
        return ops.gather_i32(data, index)
    return data.index_select(0, index.long())


class BatchGraph(Graph):
    """A batch of graphs stored as one graph: every node carries the id of its graph (`node_graph_index`), every edge
    optionally too (`edge_graph_index`); reference data/graph.py:362-621."""

    def __init__(self, x, edge_index, node_graph_index, edge_graph_index, y=None, edge_weight=None, graphs=None):
        super().__init__(x, edge_index, y, edge_weight)
        self.node_graph_index = node_graph_index
        self.edge_graph_index = edge_graph_index
        self.graphs = graphs

    @property
    def num_graphs(self):
        gi = self.node_graph_index
        return int(gi.max().item() if torch.is_tensor(gi) else np.max(gi)) + 1

    def to_device(self, device=None, inplace=False):
        dev = device if device is not None else ops.default_device()
        base = Graph.to_device(self, device=dev, inplace=inplace)
        g = self if inplace else BatchGraph(base._x, base.edge_index, self.node_graph_index, self.edge_graph_index, y=base.y,
                                            edge_weight=base.edge_weight, graphs=self.graphs)
        g.node_graph_index = ops.as_device(g.node_graph_index, torch.int32, device=dev)
        if g.edge_graph_index is not None:
            g.edge_graph_index = ops.as_device(g.edge_graph_index, torch.int32, device=dev)
        return g

    convert_data_to_tensor = to_device

    def reorder(self):
        """Nodes and edges sorted by graph id (reference :396-415), stable; device data uses the radix argsort."""
        def order_of(ids):
            if torch.is_tensor(ids):
                return ops.stable_argsort(ops.as_device(ids, torch.int32)), True
            return np.argsort(np.asarray(ids), kind="stable"), False

        def take(d, order, dev_path):
            if d is None:
                return None
            if dev_path:
                return _take_rows(ops.as_device(d, device=order.device), order)
            return np.asarray(d)[order]

        n_order, n_dev = order_of(self.node_graph_index)
        e_order, e_dev = order_of(self.edge_graph_index)
        if e_dev:
            ei = ops.as_device(self.edge_index, torch.int32, device=e_order.device)
            edge_index = torch.stack([ops.gather_i32(ei[0].contiguous(), e_order), ops.gather_i32(ei[1].contiguous(), e_order)])
        else:
            edge_index = np.asarray(self.edge_index)[:, e_order]
        return BatchGraph(take(self.x, n_order, n_dev), edge_index, take(self.node_graph_index, n_order, n_dev),
                          take(self.edge_graph_index, e_order, e_dev), y=take(self.y, n_order, n_dev),
                          edge_weight=take(self.edge_weight, e_order, e_dev))

    def to_graphs(self):
        """Split back into Graph objects (reference :417-461); offsets come from the per-graph node / edge counts."""
        batch = self.reorder()
        num_graphs = batch.num_graphs

        def offsets(ids, n):
            if torch.is_tensor(ids):
                counts = ops.segment_count(ops.as_device(ids, torch.int32), num_graphs).cpu().numpy()
            else:
                counts = np.bincount(np.asarray(ids), minlength=num_graphs)
            return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64).tolist()

        node_off = offsets(batch.node_graph_index, batch.num_nodes)
        edge_off = offsets(batch.edge_graph_index, batch.num_edges)
        graphs = []
        for i in range(num_graphs):
            n0, n1, e0, e1 = node_off[i], node_off[i + 1], edge_off[i], edge_off[i + 1]
            graphs.append(Graph(x=batch.x[n0:n1], edge_index=batch.edge_index[:, e0:e1] - n0,
                                y=None if batch.y is None else batch.y[n0:n1],
                                edge_weight=None if batch.edge_weight is None else batch.edge_weight[e0:e1]))
        return graphs

    @classmethod
    def from_graphs(cls, graphs):
        """Concatenate graphs, shifting node ids by the number of nodes before each graph (reference :463-560)."""
        on_device = torch.is_tensor(graphs[0].edge_index)
        cat = (lambda xs, axis=0: torch.cat(list(xs), dim=axis)) if on_device else \
            (lambda xs, axis=0: np.concatenate(list(xs), axis=axis))

        def full(n, v):
            if on_device:
                return torch.full((n,), v, dtype=torch.int32, device=graphs[0].edge_index.device)
            return np.full([n], v, dtype=np.int32)

        node_graph_index = cat(full(g.num_nodes, i) for i, g in enumerate(graphs))
        edge_graph_index = cat(full(g.num_edges, i) for i, g in enumerate(graphs))
        shifted, before = [], 0
        for g in graphs:
            shifted.append(g.edge_index + before)
            before += g.num_nodes
        x = cat(g.x for g in graphs)
        y = None if graphs[0].y is None else cat(g.y for g in graphs)
        edge_weight = None if graphs[0].edge_weight is None else cat(g.edge_weight for g in graphs)
        return BatchGraph(x=x, edge_index=cat(shifted, axis=1), node_graph_index=node_graph_index,
                          edge_graph_index=edge_graph_index, graphs=graphs, y=y, edge_weight=edge_weight)

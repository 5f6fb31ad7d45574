# coding=utf-8
"""tfg.layers.MapReduceGNN (reference layers/kernel/map_reduce.py:6-46): a layer defined by three overridable hooks.

    class Mine(tfg.layers.MapReduceGNN):
        def map(self, repeated_x, neighbor_x, edge_weight=None): ...      # per-edge message
        def reduce(self, neighbor_msg, node_index, num_nodes=None): ...   # per-node reduction, e.g. tfg.nn.mean_reducer(...)
        def update(self, x, reduced_neighbor_msg): ...                    # combine with the node's own features

Calling the layer on [x, edge_index, edge_weight] runs aggregate_neighbors with the hooks.  User-defined `map` code takes
the gather route; a stock reducer called from `reduce` still lands on the segment kernel."""
import torch

from ..nn.kernel.map_reduce import aggregate_neighbors


class MapReduceGNN(torch.nn.Module):

    def map(self, repeated_x, neighbor_x, edge_weight=None):
        raise NotImplementedError("override map()")

    def reduce(self, neighbor_msg, node_index, num_nodes=None):
        raise NotImplementedError("override reduce()")

    def update(self, x, reduced_neighbor_msg):
        raise NotImplementedError("override update()")

    # the reference also hands the hooks out as plain callables
    def get_mapper(self):
        return self.map

    def get_reducer(self):
        return self.reduce

    def get_updater(self):
        return self.update

    def forward(self, inputs, training=None, mask=None):
        x, edge_index, edge_weight = inputs
        return aggregate_neighbors(x, edge_index, edge_weight, mapper=self.map, reducer=self.reduce, updater=self.update)

# coding=utf-8
"""tfg.layers.MapReduceGNN (reference layers/kernel/map_reduce.py:6-46): subclass and override map / reduce / update;
the call runs aggregate_neighbors with them.  Stock reducers (tfg.nn.sum_reducer / mean_reducer / max_reducer) returned from
`reduce` still land on the segment kernel; arbitrary `map` callables take the gather route."""
import torch

from ..nn.kernel.map_reduce import aggregate_neighbors


class MapReduceGNN(torch.nn.Module):

    def map(self, repeated_x, neighbor_x, edge_weight=None):
        pass

    def reduce(self, neighbor_msg, node_index, num_nodes=None):
        pass

    def update(self, x, reduced_neighbor_msg):
        pass

    def get_mapper(self):
        def mapper(repeated_x, neighbor_x, edge_weight=None):
            return self.map(repeated_x, neighbor_x, edge_weight)
        return mapper

    def get_reducer(self):
        def reducer(neighbor_msg, node_index, num_nodes=None):
            return self.reduce(neighbor_msg, node_index, num_nodes)
        return reducer

    def get_updater(self):
        def updater(x, reduced_neighbor_msg):
            return self.update(x, reduced_neighbor_msg)
        return updater

    def forward(self, inputs, training=None, mask=None):
        x, edge_index, edge_weight = inputs
        return aggregate_neighbors(x, edge_index, edge_weight, self.get_mapper(), self.get_reducer(), self.get_updater())

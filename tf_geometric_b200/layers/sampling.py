# coding=utf-8
"""tfg.layers.DropEdge (reference layers/sampling/drop_edge.py:7-27)."""
import torch

from ..nn.sampling.drop_edge import drop_edge


class DropEdge(torch.nn.Module):

    def __init__(self, rate=0.5, force_undirected=False):
        """
        :param rate: probability of dropping an edge
        :param force_undirected: keep or drop both directions of an undirected edge together
        """
        super().__init__()
        self.rate = rate
        self.force_undirected = force_undirected
        if self.rate < 0. or self.rate > 1.:
            raise ValueError('Dropout probability has to be between 0 and 1, '
                             'but got {}'.format(self.rate))

    def forward(self, inputs, training=None, mask=None, seed=None):
        """inputs: [edge_index, edge_attr, ...]; identity unless training."""
        return drop_edge(inputs=inputs, rate=self.rate, force_undirected=self.force_undirected, training=training, seed=seed)

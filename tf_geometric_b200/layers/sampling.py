# coding=utf-8
"""tfg.layers.DropEdge: the layer form of tfg.nn.drop_edge (reference layers/sampling/drop_edge.py:7-27)."""
import torch

from ..nn.sampling.drop_edge import drop_edge


class DropEdge(torch.nn.Module):
    """Randomly removes edges while training (Rong et al., "DropEdge", ICLR 2020); identity at inference.
    `layer([edge_index, edge_attr, ...], training=True)` returns the surviving edge_index and attributes."""

    def __init__(self, rate=0.5, force_undirected=False):
        super().__init__()
        if not 0.0 <= rate <= 1.0:
            raise ValueError("Dropout probability has to be between 0 and 1, but got {}".format(rate))
        self.rate, self.force_undirected = rate, force_undirected     # force_undirected: both directions share one draw

    def forward(self, inputs, training=None, mask=None, seed=None):
        return drop_edge(inputs, self.rate, self.force_undirected, training, seed=seed)

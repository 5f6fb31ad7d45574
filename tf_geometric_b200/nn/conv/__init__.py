# coding=utf-8
from . import gcn, gat, graph_sage, appnp, propagation

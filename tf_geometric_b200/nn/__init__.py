# coding=utf-8
"""Functional API (the subset of tf_geometric.nn on the message-passing hot path)."""
from .kernel.map_reduce import (identity_mapper, neighbor_count_mapper, gcn_mapper, sum_reducer, mean_reducer,
                                max_reducer, sum_updater, identity_updater, aggregate_neighbors)
from .kernel.segment import segment_softmax, segment_count
from .conv.gcn import (gcn, gcn_norm_adj, gcn_norm_edge, gcn_build_cache_by_adj, gcn_build_cache_for_graph,
                       compute_cache_key)
from .conv.gat import gat
from .conv.graph_sage import (mean_graph_sage, sum_graph_sage, gcn_graph_sage, mean_pool_graph_sage,
                              max_pool_graph_sage)
from .conv.appnp import appnp
from .conv.propagation import (sgc, ssgc, tagcn, gin, gin_updater, le_conv, chebynet, chebynet_norm_edge,
                               get_laplacian)
from .pool.common_pool import mean_pool, sum_pool, max_pool, min_pool
from .pool.set2set import set2set
from .pool.topk_pool import topk_pool
from .pool.sag_pool import sag_pool
from .pool.sort_pool import sort_pool
from ..ops import relu
from .sampling.drop_edge import drop_edge

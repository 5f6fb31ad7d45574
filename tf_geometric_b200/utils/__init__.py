# coding=utf-8
from . import graph_utils
from .graph_utils import (add_self_loop_edge, remove_self_loop_edge, convert_edge_to_directed, merge_duplicated_edge,
                          convert_edge_to_upper, convert_edge_index_to_edge_hash, convert_edge_hash_to_edge_index,
                          adj_norm_edge, compute_num_or_size_splits)
from .sampling import RandomNeighborSampler, UniformNeighborSampler

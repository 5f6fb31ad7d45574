# coding=utf-8
"""Edge-index integer preprocessing with the reference's semantics (utils/graph_utils.py of tf_geometric).

Device inputs (torch CUDA tensors) run on the GPU kernels - including the hash/unique based helpers
(merge_duplicated_edge, convert_edge_to_upper, convert_edge_to_directed: tfgk_edge_unique / tfgk_directed_edges, bit-exact
with tf.unique's first-occurrence order) - and return device tensors.  numpy / list inputs are processed with numpy and
return numpy, exactly like the reference does for non-tensor inputs (its eager host path).
"""
import numpy as np
import torch

from .. import ops


def _is_device(x):
    return torch.is_tensor(x) and x.is_cuda


def _to_numpy(x):
    if x is None:
        return None
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _like(result, template, dtype):
    """Return `result` (numpy) in the container type of `template`."""
    if torch.is_tensor(template):
        return torch.from_numpy(np.ascontiguousarray(result)).to(device=template.device, dtype=dtype)
    return result


def convert_edge_index_to_edge_hash(edge_index, num_nodes=None):
    """hash = num_nodes * row + col in int64; num_nodes defaults to max id + 1 (reference :14-43)."""
    ei = _to_numpy(edge_index).astype(np.int64)
    if num_nodes is None:
        num_nodes = int(ei.max()) + 1
    edge_hash = np.int64(num_nodes) * ei[0] + ei[1]
    return _like(edge_hash, edge_index, torch.int64), int(num_nodes)


def convert_edge_hash_to_edge_index(edge_hash, num_nodes):
    """reference :46-64."""
    h = _to_numpy(edge_hash).astype(np.int64)
    ei = np.stack([h // num_nodes, h % num_nodes], axis=0).astype(np.int32)
    return _like(ei, edge_hash, torch.int32)


def _first_occurrence_unique(values):
    uniq, first, inverse = np.unique(values, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return uniq[order], rank[inverse]


def _segment_merge(prop, ids, n, mode):
    prop = np.asarray(prop)
    if mode == "sum":
        out = np.zeros((n,) + prop.shape[1:], dtype=prop.dtype)
        np.add.at(out, ids, prop)
    elif mode == "mean":
        out = np.zeros((n,) + prop.shape[1:], dtype=prop.dtype)
        np.add.at(out, ids, prop)
        cnt = np.maximum(np.bincount(ids, minlength=n), 1).astype(prop.dtype)
        out = (out / cnt.reshape((-1,) + (1,) * (prop.ndim - 1))).astype(prop.dtype)
    elif mode == "max":
        out = np.full((n,) + prop.shape[1:], np.finfo(prop.dtype).min if prop.dtype.kind == "f" else np.iinfo(prop.dtype).min,
                      dtype=prop.dtype)
        np.maximum.at(out, ids, prop)
    elif mode == "min":
        out = np.full((n,) + prop.shape[1:], np.finfo(prop.dtype).max if prop.dtype.kind == "f" else np.iinfo(prop.dtype).max,
                      dtype=prop.dtype)
        np.minimum.at(out, ids, prop)
    else:
        raise Exception("wrong merge mode: {}".format(mode))
    return out


def merge_duplicated_edge(edge_index, edge_props=None, merge_modes=None):
    """Duplicates collapse onto their first occurrence; props merged with sum|min|max|mean (reference :67-125)."""
    if edge_props is not None and len(edge_props) > 0:
        if merge_modes is None:
            merge_modes = ["sum"] * len(edge_props)
        elif type(merge_modes) is not list:
            raise Exception("type error: merge_modes should be a list of strings")
    if _is_device(edge_index):
        return _merge_duplicated_edge_device(edge_index, edge_props, merge_modes)
    ei = _to_numpy(edge_index).astype(np.int32)
    edge_hash, hash_n = convert_edge_index_to_edge_hash(ei)
    uniq_hash, uniq_idx = _first_occurrence_unique(edge_hash)
    uniq_ei = convert_edge_hash_to_edge_index(uniq_hash, hash_n)
    out_index = _like(uniq_ei, edge_index, torch.int32)
    if edge_props is None:
        return out_index, None
    out_props = []
    for prop, mode in zip(edge_props, merge_modes):
        if prop is None:
            out_props.append(None)
        else:
            merged = _segment_merge(_to_numpy(prop), uniq_idx, len(uniq_hash), mode)
            out_props.append(_like(merged, prop, prop.dtype if torch.is_tensor(prop) else None))
    return out_index, out_props


def _merge_duplicated_edge_device(edge_index, edge_props, merge_modes):
    from ..nn.kernel.map_reduce import _segment_reduce
    ei = edge_index if edge_index.dtype == torch.int32 else edge_index.to(torch.int32)
    ei = ei.contiguous()
    if ei.shape[1] == 0:
        return ei, (None if edge_props is None else list(edge_props))
    hash_n = int(ei.max().item()) + 1                                 # reference: num_nodes = reduce_max(edge_index) + 1
    uniq_index, of_edge = ops.edge_unique(ei[0].contiguous(), ei[1].contiguous(), hash_n)
    if edge_props is None:
        return uniq_index, None
    out = []
    for prop, mode in zip(edge_props, merge_modes):
        if prop is None:
            out.append(None)
            continue
        if mode not in ("sum", "min", "max", "mean"):
            raise Exception("wrong merge mode: {}".format(mode))
        p = ops.as_device(prop, torch.float32, device=ei.device)
        out.append(_segment_reduce(p, of_edge, uniq_index.shape[1], mode))
    return uniq_index, out


def convert_edge_to_upper(edge_index, edge_props=None, merge_modes=None):
    """(min(u,v), max(u,v)) for every edge, then merge duplicates (reference :128-151)."""
    if _is_device(edge_index):
        ei = edge_index.to(torch.int32)
        upper = torch.stack([torch.minimum(ei[0], ei[1]), torch.maximum(ei[0], ei[1])]).contiguous()
        return merge_duplicated_edge(upper, edge_props, merge_modes)
    ei = _to_numpy(edge_index).astype(np.int32)
    upper = np.stack([ei.min(axis=0), ei.max(axis=0)], axis=0)
    upper_index, upper_props = merge_duplicated_edge(_like(upper, edge_index, torch.int32), edge_props, merge_modes)
    return upper_index, upper_props


def convert_edge_to_directed(edge_index, edge_props=None, merge_modes=None):
    """Undirected -> both directions: upper edges followed by the mirrored non-loop upper edges (reference :155-212)."""
    if edge_props is not None and len(edge_props) > 0 and merge_modes is None:
        merge_modes = ["sum"] * len(edge_props)
    upper_index, upper_props = convert_edge_to_upper(edge_index, edge_props, merge_modes)
    if _is_device(edge_index):
        if upper_index.shape[1] == 0:
            return edge_index, edge_props
        out_index, lower_src = ops.directed_edges(upper_index.contiguous())
        if lower_src.numel() == 0:                                     # only self loops: reference returns the inputs
            return edge_index, edge_props
        if edge_props is None:
            return out_index, None
        out_props = []
        for prop, up_prop in zip(edge_props, upper_props):
            if prop is None:
                out_props.append(None)
            else:
                out_props.append(torch.cat([up_prop, ops.permute(up_prop.contiguous(), lower_src)]))
        return out_index, out_props
    up = _to_numpy(upper_index)
    mask = up[0] != up[1]
    if not mask.any():
        return edge_index, edge_props
    lower = np.stack([up[1][mask], up[0][mask]], axis=0)
    out_index = _like(np.concatenate([up, lower], axis=1), edge_index, torch.int32)
    if edge_props is None:
        return out_index, None
    out_props = []
    for prop, up_prop in zip(edge_props, upper_props):
        if prop is None:
            out_props.append(None)
        else:
            upn = _to_numpy(up_prop)
            out_props.append(_like(np.concatenate([upn, upn[mask]], axis=0), prop,
                                   prop.dtype if torch.is_tensor(prop) else None))
    return out_index, out_props


def remove_self_loop_edge(edge_index, edge_weight=None):
    """reference :252-269."""
    ei = _to_numpy(edge_index)
    mask = ei[0] != ei[1]
    out_w = None
    if edge_weight is not None:
        out_w = _like(_to_numpy(edge_weight)[mask], edge_weight, torch.float32)
    return _like(ei[:, mask], edge_index, torch.int32), out_w


def add_self_loop_edge(edge_index, num_nodes, edge_weight=None, fill_weight=1.0):
    """Append [[0..N-1],[0..N-1]] AFTER the existing edges; weights get `fill_weight`; no dedup (reference :350-366).
    Device tensors run tfgk_self_loops_i32 / tfgk_self_loop_weights_f32 (bit-exact integers)."""
    num_nodes = int(num_nodes)
    if _is_device(edge_index):
        ei = edge_index if edge_index.dtype == torch.int32 else edge_index.to(torch.int32)
        out_index = ops.self_loops(ei.contiguous(), num_nodes)
        out_w = None
        if edge_weight is not None:
            w = ops.as_device(edge_weight, torch.float32, device=ei.device)
            out_w = ops.self_loop_weights(w, ei.shape[1], num_nodes, fill_weight, ei.device)
        return out_index, out_w
    ei = np.asarray(_to_numpy(edge_index), dtype=np.int32).reshape(2, -1)
    diag = np.arange(num_nodes, dtype=np.int32)
    out_index = _like(np.concatenate([ei, np.stack([diag, diag])], axis=1), edge_index, torch.int32)
    out_w = None
    if edge_weight is not None:
        w = np.concatenate([_to_numpy(edge_weight).astype(np.float32), np.full([num_nodes], fill_weight, dtype=np.float32)])
        out_w = _like(w, edge_weight, torch.float32)
    return out_index, out_w


def adj_norm_edge(edge_index, num_nodes, edge_weight=None, add_self_loop=False, cache=None):
    """D^-1/2 A D^-1/2 on the edge list with ROW degrees on both sides (reference :914-943)."""
    cache_key = "adj_normed_edge"
    if cache is not None and cache.get(cache_key) is not None:
        return cache[cache_key]
    from ..sparse import SparseMatrix
    ei = ops.as_device(edge_index, torch.int32)
    adj = SparseMatrix(ei, edge_weight, [num_nodes, num_nodes])
    if add_self_loop:
        adj = adj.add_diag(1.0)
    dis = ops.deg_inv(adj.segment_sum(axis=-1), ops.POW_INV_SQRT)
    normed = ops.scale_edges(adj.index[0].contiguous(), adj.index[1].contiguous(), adj.value, dl=dis, dr=dis)
    if cache is not None:
        cache[cache_key] = adj.index, normed
    return adj.index, normed


def compute_num_or_size_splits(num_h_features, num_splits):
    """utils/tf_sparse_utils.py:71-90 - how GCN(num_splits=...) chunks the feature columns of XW."""
    if num_splits is None or num_splits == 1:
        return None
    if num_h_features % num_splits == 0:
        return num_splits
    split_size = int(np.ceil(num_h_features / num_splits))
    num_pre_splits = int(np.floor(num_h_features / split_size))
    last_split_size = num_h_features % split_size
    sizes = [split_size] * num_pre_splits + ([last_split_size] if last_split_size > 0 else [])
    if len(sizes) != num_splits:
        raise Exception("cannot split H of shape [None, {}] into {} matrices, please provide a valid num_splits"
                        .format(num_h_features, num_splits))
    return sizes


# samplers live in utils/sampling.py; re-exported here because the reference defines them in this module (:630-846)
from .sampling import RandomNeighborSampler, UniformNeighborSampler  # noqa: E402,F401

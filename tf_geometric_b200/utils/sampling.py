# coding=utf-8
"""Neighbour samplers with the reference's interface (utils/graph_utils.py:630-846), on the device.

RandomNeighborSampler: the reference keeps a Python dict of neighbour arrays and loops over every node calling
np.random.choice (the bottleneck of demo/demo_graph_sage.py:53-55); here the dict is the stable row-sorted CSR and the
loop is one thread per row (tfgk_neighbor_sample_*).  UniformNeighborSampler: a Bernoulli flag per edge + compaction.
Randomness is counter-based (csrc/rng.cuh); pass `seed=` for reproducible draws."""
import torch

from .. import ops, _rng


def _split_node_index(sampled_node_index):
    if isinstance(sampled_node_index, tuple):
        return sampled_node_index
    return sampled_node_index, sampled_node_index


def _virtual_mapping(sampled_index, num_nodes, device, strict=True):
    """-1 everywhere, position-in-`sampled_index` at the sampled ids (graph_utils.py:792-800).  Ids outside
    [0, num_nodes) raise like numpy's fancy assignment does, unless strict=False (a symmetric node set applied to the
    narrower side of a rectangular edge list: such nodes simply have no edge there)."""
    idx = ops.as_device(sampled_index, torch.int32, device=device).reshape(-1)
    n_virtual = idx.numel()
    position = torch.arange(n_virtual, dtype=torch.int32, device=device)
    if n_virtual and (int(idx.max().item()) >= num_nodes or int(idx.min().item()) < 0):
        if strict:
            raise IndexError("sampled_node_index holds ids outside [0, {})".format(num_nodes))
        valid = (idx >= 0) & (idx < num_nodes)
        idx, position = idx[valid], position[valid]
    mapping = torch.full((num_nodes,), -1, dtype=torch.int32, device=device)
    mapping[idx.long()] = position
    return mapping, n_virtual


class _SamplerBase(object):

    def __init__(self, edge_index, edge_weight=None):
        self.edge_index = ops.as_device(edge_index, torch.int32)
        dev = self.edge_index.device
        E = self.edge_index.shape[1]
        self.num_edges = E
        self.edge_weight = torch.ones((E,), dtype=torch.float32, device=dev) if edge_weight is None \
            else ops.as_device(edge_weight, torch.float32, device=dev).reshape(-1)
        self.row, self.col = self.edge_index[0].contiguous(), self.edge_index[1].contiguous()
        self.num_row_nodes = int(self.row.max().item()) + 1 if E else 0
        self.num_col_nodes = int(self.col.max().item()) + 1 if E else 0

    def _virtual_edges(self, sampled_node_index, bernoulli=ops.BERNOULLI_NONE, prob=0.0, seed=0):
        """Edges with both ends in the sampled sets, relabelled; optionally thinned by a Bernoulli rule in the same
        pass.  Returns (virtual_row, virtual_col, weight, num_virtual_rows, num_virtual_cols)."""
        rows, cols = _split_node_index(sampled_node_index)
        dev = self.edge_index.device
        row_map, n_vr = _virtual_mapping(rows, self.num_row_nodes, dev)
        if cols is rows and self.num_col_nodes == self.num_row_nodes:
            col_map, n_vc = row_map, n_vr
        else:
            col_map, n_vc = _virtual_mapping(cols, self.num_col_nodes, dev, strict=cols is not rows)
        flag = ops.edge_flags(self.row, self.col, self.num_edges, mode=ops.FLAG_MAPPED, row_map=row_map, col_map=col_map,
                              bernoulli=bernoulli, prob=prob, seed=seed)
        index = ops.select_flagged(flag)
        v_row = ops.gather_i32(row_map, ops.gather_i32(self.row, index))
        v_col = ops.gather_i32(col_map, ops.gather_i32(self.col, index))
        return v_row, v_col, ops.permute(self.edge_weight, index), n_vr, n_vc


class RandomNeighborSampler(_SamplerBase):
    """Per-node fan-out sampling (graph_utils.py:630-776)."""

    def __init__(self, edge_index, edge_weight=None):
        super().__init__(edge_index, edge_weight)
        self._csr = None
        self._w_csr = None

    def _structure(self):
        if self._csr is None:
            self._csr = ops.csr_build(self.row, self.col, self.num_row_nodes, max(self.num_col_nodes, 1))
            self._w_csr = ops.permute(self.edge_weight, self._csr.perm)
        return self._csr, self._w_csr

    def sample(self, k=None, ratio=None, sampled_node_index=None, padding=False, seed=None):
        """
        :param k: neighbours per node (all of them when the node has at most k and padding is False; k draws with
            replacement when padding is True and the node has at most k)
        :param ratio: instead of k, keep ceil(degree * ratio) neighbours per node, without replacement
        :param sampled_node_index: ids (or a (row_ids, col_ids) tuple): restrict to, and relabel into, this node set
        :return: (edge_index int32 [2, S], edge_weight float32 [S]) on the device, or (None, None) when S == 0
        """
        if k is not None and ratio is not None:
            raise Exception("k and ratio cannot be provided simultaneously")
        if sampled_node_index is None:
            csr, w_csr = self._structure()
        else:
            v_row, v_col, w, n_vr, n_vc = self._virtual_edges(sampled_node_index)
            if v_row.numel() == 0:
                return None, None
            csr = ops.csr_build(v_row, v_col, n_vr, max(n_vc, 1))
            w_csr = ops.permute(w, csr.perm)
        if csr.nnz == 0:
            return None, None
        out_row, out_pos, _ = ops.neighbor_sample(csr, k=k, ratio=ratio, padding=padding, seed=_rng.resolve(seed))
        if out_row.numel() == 0:
            return None, None
        return torch.stack([out_row, ops.gather_i32(csr.col, out_pos)]), ops.permute(w_csr, out_pos)


class UniformNeighborSampler(_SamplerBase):
    """Independent Bernoulli(prob) edge sampling (graph_utils.py:775-846)."""

    def sample(self, prob, sampled_node_index=None, seed=None):
        seed = _rng.resolve(seed)
        if sampled_node_index is None:
            flag = ops.edge_flags(None, None, self.num_edges, bernoulli=ops.BERNOULLI_KEEP, prob=float(prob), seed=seed,
                                  device=self.edge_index.device)
            index = ops.select_flagged(flag)
            return (torch.stack([ops.gather_i32(self.row, index), ops.gather_i32(self.col, index)]),
                    ops.permute(self.edge_weight, index))
        v_row, v_col, w, _, _ = self._virtual_edges(sampled_node_index, bernoulli=ops.BERNOULLI_KEEP, prob=float(prob),
                                                    seed=seed)
        return torch.stack([v_row, v_col]), w

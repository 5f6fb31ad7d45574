# coding=utf-8
"""Set2Set read-out (reference nn/pool/set2set.py:8-44).

The attention step of every iteration (set2set.py:35-39: gather h per node, dot with x, segment_softmax by graph, weighted
unsorted_segment_sum) is exactly the fused attention kernel with ONE query row per graph, keys = values = the node features
and the nodes of a graph as that row's "neighbours": a single tfgk_gat_fused_f32 launch over the CSR keyed by
node_graph_index (scale 1, one head) instead of two gathers, a 5-pass softmax and a scatter."""
import torch

from ... import ops, _structure, autograd


def _graph_rows(node_graph_index, num_graphs, num_nodes):
    """CSR with one row per graph whose columns are the node ids of that graph (stable: node order is kept)."""
    seg = _structure.csr_for_segment_ids(node_graph_index, num_graphs)
    hit = _structure._lookup(seg.perm, ("graph_rows",))
    if hit is None:
        view = ops.CSR(seg.rowptr, seg.perm, seg.perm, num_graphs, num_nodes)
        view.plan = seg.plan
        hit = _structure._store(seg.perm, ("graph_rows",), view)
    return hit


def set2set(x, node_graph_index, lstm, num_iterations, training=None):
    """
    :param x: [num_nodes, num_features]
    :param node_graph_index: [num_nodes] graph id of every node
    :param lstm: callable with the Keras calling convention the reference uses:
        lstm(h[1, num_graphs, 2F], initial_state=[state_h, state_c], training=...) -> (sequence[1, num_graphs, F], state_h,
        state_c) (the graphs are the time steps of ONE sequence, like in the reference; tfg.layers.Set2Set supplies one)
    :param num_iterations: attention iterations
    :return: [num_graphs, 2 * num_features]
    """
    node_graph_index = ops.as_device(node_graph_index, torch.int32)
    dev = node_graph_index.device
    x = ops.as_device(x, torch.float32, device=dev)
    num_nodes, units = x.shape
    num_graphs = int(node_graph_index.max().item()) + 1
    rows = _graph_rows(node_graph_index, num_graphs, num_nodes)
    with_grad = autograd.needs_grad(x, *[p for p in getattr(lstm, "parameters", lambda: [])()])
    if with_grad:        # (graph id, node id) pairs: the "edge list" the transposed structure of the backward is built from
        pairs = _structure._lookup(node_graph_index, ("graph_pairs",))
        if pairs is None:
            pairs = _structure._store(node_graph_index, ("graph_pairs",), torch.stack(
                [node_graph_index, torch.arange(num_nodes, dtype=torch.int32, device=dev)]))

    h = torch.zeros((num_graphs, units * 2), dtype=torch.float32, device=dev)
    state = [torch.zeros((1, units), dtype=torch.float32, device=dev), torch.zeros((1, units), dtype=torch.float32, device=dev)]
    for _ in range(num_iterations):
        q, state_h, state_c = lstm(h.unsqueeze(0), initial_state=state, training=training)      # set2set.py:30-33
        state = [state_h, state_c]
        q = q.squeeze(0).contiguous()
        if with_grad:
            att_h = autograd.GatAttention.apply(q, x, x, None, rows, pairs, 1, True, ops.ACT_NONE, 0.0, 0, 1.0)
        else:
            att_h = ops.gat_fused(rows, q, x, x, 1, scale=1.0)                                   # set2set.py:35-39
        h = torch.cat([q, att_h], dim=-1)
    return h

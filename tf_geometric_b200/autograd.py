# coding=utf-8
"""Backward passes (mean/sum_graph_sage forward + backward; 8(f)4: GCN and GAT training); the reference
gets them from TensorFlow autodiff (demo/demo_graph_sage.py:100-106, demo/demo_gcn.py:60-75, demo/demo_gat.py).

    d(unsorted_segment_mean(x[col] * w, row)) / dx  =  scatter-add by col of  (w_e / max(cnt[row_e], 1)) * g[row_e]

i.e. the SAME gather - edge-apply - segment-reduce kernel (tfgk_spmm_f32) run on the TRANSPOSED structure (a CSC = the
destination-sorted CSR of the reversed edges) with rescaled weights; it is built once per edge list and memoised next
to the forward CSR.  Dense layers: dX = dY W^T and dW = X^T dY are tfgk_gemm_f32 with transposes (deterministic split-K
over the node dimension), db = 1^T dY through the same kernel.
"""
import weakref

import numpy as np
import torch

from . import ops, _structure


def _relu_grad(g, y):
    """dL/d(pre-activation) of y = relu(.): g where y > 0, else 0 - one elementwise pass (aten's relu backward) instead of
    compare + cast + multiply."""
    return torch.ops.aten.threshold_backward(g, y, 0.0)


def _transposed_structure(edge_index, num_nodes, edge_weight, mean, csr):
    """(csr_t, w_t): CSR of the reversed edges and w_e / max(cnt[row_e], 1) (or w_e for sum) in its order.
    The structure is memoised per edge list; the weights per (weight tensor, structure) with the weight tensor held by
    weak reference and checked by version (an id() recycled by a new tensor can never alias an old entry)."""
    tag = ("csc", int(num_nodes))
    csr_t = _structure._lookup(edge_index, tag)
    row, col = edge_index[0].contiguous(), edge_index[1].contiguous()
    if csr_t is None:
        csr_t = _structure._store(edge_index, tag, ops.csr_build(col, row, num_nodes, num_nodes))
    owner = edge_weight if edge_weight is not None else edge_index
    wtag = ("csc_w", bool(mean), edge_weight is None, id(csr_t), id(csr))
    hit = _structure._lookup(owner, wtag)
    if hit is not None and hit[0]() is csr_t and hit[1]() is csr:
        return csr_t, hit[2]
    w = edge_weight if edge_weight is not None else torch.ones((edge_index.shape[1],), dtype=torch.float32,
                                                               device=edge_index.device)
    if mean:
        cnt = (csr.rowptr[1:] - csr.rowptr[:-1]).clamp(min=1).to(torch.float32)
        w = ops.scale_edges(row, None, w, dl=torch.reciprocal(cnt))
    w_t = ops.permute(w.contiguous(), csr_t.perm)
    _structure._store(owner, wtag, (weakref.ref(csr_t), weakref.ref(csr), w_t))
    return csr_t, w_t


class NeighborAggregate(torch.autograd.Function):
    """agg = REDUCE_{e: row_e = r} w_e x[col_e]  (sum | mean), differentiable w.r.t. x."""

    @staticmethod
    def forward(ctx, x, edge_index, edge_weight, reduce, num_nodes):
        csr, _ = _structure.csr_for_edge_index(edge_index, num_nodes)
        w_csr = None if edge_weight is None else _structure.weights_in_csr_order(edge_weight, csr)
        ctx.saved = (edge_index, edge_weight, reduce, num_nodes, csr)
        return ops.spmm(csr, w_csr, x.detach(), reduce=reduce)

    @staticmethod
    def backward(ctx, grad_out):
        edge_index, edge_weight, reduce, num_nodes, csr = ctx.saved
        csr_t, w_t = _transposed_structure(edge_index, num_nodes, edge_weight, reduce == "mean", csr)
        grad_x = ops.spmm(csr_t, w_t, grad_out.contiguous(), reduce="sum")
        return grad_x, None, None, None, None


class Dense(torch.autograd.Function):
    """y = act(x @ W + b) with act in {None, relu}; dX, dW, db through tfgk_gemm_f32."""

    @staticmethod
    def forward(ctx, x, weight, bias, act_code):
        y = ops.gemm(x.detach(), weight.detach(), bias=None if bias is None else bias.detach(), act=act_code)
        ctx.save_for_backward(x, weight, y if act_code == ops.ACT_RELU else None)
        ctx.has_bias = bias is not None
        ctx.act_code = act_code
        return y

    @staticmethod
    def backward(ctx, grad_y):
        x, weight, y = ctx.saved_tensors
        g = grad_y.contiguous()
        if ctx.act_code == ops.ACT_RELU:
            g = _relu_grad(g, y)                      # elementwise mask (torch: plumbing, not a hot op)
        grad_x = grad_w = grad_b = None
        if ctx.needs_input_grad[0]:
            grad_x = ops.gemm(g, weight.detach(), trans_b=True)
        if ctx.needs_input_grad[1]:
            grad_w = ops.gemm(x.detach(), g, trans_a=True)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            grad_b = ops.colsum(g)
        return grad_x, grad_w, grad_b, None


class SparseMatmul(torch.autograd.Function):
    """y = act(A @ h + b) for a cached SparseMatrix A (gcn.py:280-288), differentiable w.r.t. h and b.
    dh = A^T dz runs the same kernel on the transposed structure of A (built once per matrix, values permuted into it);
    dz = dy * (y > 0) for relu."""

    @staticmethod
    def forward(ctx, h, bias, adj, act_code):
        y = ops.spmm(adj.csr, adj.value_csr, h.detach(), reduce="sum", bias=None if bias is None else bias.detach(),
                     act=act_code)
        ctx.adj = adj
        ctx.act_code = act_code
        ctx.has_bias = bias is not None
        ctx.save_for_backward(y if act_code == ops.ACT_RELU else None)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        (y,) = ctx.saved_tensors
        g = grad_y.contiguous()
        if ctx.act_code == ops.ACT_RELU:
            g = _relu_grad(g, y)
        adj = ctx.adj
        csr_t = adj._transposed_csr()
        if getattr(adj, "_value_csc", None) is None:
            adj._value_csc = ops.permute(adj.value, csr_t.perm)
        grad_h = ops.spmm(csr_t, adj._value_csc, g, reduce="sum") if ctx.needs_input_grad[0] else None
        grad_b = None
        if ctx.has_bias and ctx.needs_input_grad[1]:
            grad_b = ops.colsum(g)
        return grad_h, grad_b, None, None


def _transposed_of_csr(csr, edge_index_used):
    """(csr_t, emap) for a forward CSR: csr_t has one row per SOURCE node, its columns are the destination rows, and
    emap[p] is the forward-CSR position of transposed slot p (per-edge tables such as the attention coefficients are
    stored in forward-CSR order).  Built once per CSR object."""
    hit = _structure._lookup(csr.col, ("csr_t",))
    if hit is not None and hit[0]() is csr:
        return hit[1], hit[2]
    row_of_pos = ops.gather_i32(edge_index_used[0].contiguous(), csr.perm)        # destination row of every CSR slot
    csr_t = ops.csr_build(csr.col, row_of_pos, csr.n_cols, csr.n_rows)
    _structure._store(csr.col, ("csr_t",), (weakref.ref(csr), csr_t, csr_t.perm))
    return csr_t, csr_t.perm


class GatAttention(torch.autograd.Function):
    """y = act(softmax-attention aggregate(Q, K, V) + b) (gat.py:73-120) with gradients for Q, K, V and b.

    Without attention dropout (the default path):
      forward : tfgk_gat_fused_stats_f32 keeps (max, denominator) per (row, head) - no [E', H] coefficient table;
      backward: tfgk_gat_bwd_prepare/dst/src_f32 recompute the coefficients from those two numbers (dQ over the forward
                CSR, dK and dV over the transposed CSR: gathers, never scatter-adds).
    With attention dropout (gat.py:85), averaged heads, hub-row plans or shapes the streaming kernel does not take:
      forward : tfgk_gat_fused_f32 with the coefficients kept; under dropout they are re-aggregated by
                tfgk_spmm_heads_f32 with a counter-based mask;
      backward: G = dy * act'(y);  ds = tfgk_gat_softmax_bwd_f32(att, G, V);
                dQ = sum_e ds K[col] / scale                       forward CSR
                dK = sum_e ds Q[row] / scale,  dV = sum_e a' G[row] transposed CSR
    The mask is regenerated from (seed, edge, head) in every kernel, never stored."""

    @staticmethod
    def forward(ctx, Q, K, V, bias, csr, edge_index_used, num_heads, split, act_code, drop_rate, seed, scale=None):
        Qd, Kd, Vd = Q.detach(), K.detach(), V.detach()
        b = None if bias is None else bias.detach()
        H = int(num_heads)
        if scale is None:                                    # gat.py:78; set2set.py:37 passes 1 (raw dot products)
            scale = float(np.sqrt(np.float32(Q.shape[1] // H)))
        if drop_rate > 0.0:
            _, att = ops.gat_fused(csr, Qd, Kd, Vd, H, split_value_heads=split, return_attention=True, scale=scale)
            y = ops.spmm_heads(csr, att, Vd, H, mode=ops.HEADS_SPLIT if split else ops.HEADS_REDUCE,
                               drop_rate=drop_rate, seed=seed, alpha=1.0 if split else 1.0 / H, bias=b, act=act_code)
        else:
            # no dropout: keep (max, denominator) per (row, head) instead of the [E', H] coefficients and recompute them in
            # the backward pass; shapes the streaming kernel does not take, and hub-row plans, keep the coefficient table
            res = None
            plan = getattr(csr, "plan", None)
            if split and (plan is None or plan.n_hubs == 0) and csr.n_rows == Qd.shape[0]:
                res = ops.gat_fused_stats(csr, Qd, Kd, Vd, H, bias=b, act=act_code, scale=scale)
            if res is not None:
                y, stats = res
                ctx.save_for_backward(Qd, Kd, Vd, stats, y)
                ctx.meta = (csr, edge_index_used, H, bool(split), act_code, 0.0, seed, bias is not None, float(scale))
                ctx.recompute, ctx.bias = True, b
                return y
            y, att = ops.gat_fused(csr, Qd, Kd, Vd, H, split_value_heads=split, bias=b, act=act_code,
                                   return_attention=True, scale=scale)
        ctx.recompute = False
        ctx.save_for_backward(Qd, Kd, Vd, att, y if act_code == ops.ACT_RELU else None)
        ctx.meta = (csr, edge_index_used, H, bool(split), act_code, float(drop_rate), seed, bias is not None, float(scale))
        return y

    @staticmethod
    def backward(ctx, grad_y):
        Q, K, V, att, y = ctx.saved_tensors
        csr, edge_index_used, H, split, act_code, drop_rate, seed, has_bias, scale = ctx.meta
        g = grad_y.contiguous()
        if ctx.recompute:
            csr_t, _ = _transposed_of_csr(csr, edge_index_used)
            res = ops.gat_backward_recompute(csr, csr_t, Q, K, V, g, y, ctx.bias, act_code, att, H, scale)
            if res is None:
                raise RuntimeError("GatAttention: the recompute backward refused a shape its forward accepted")
            grad_b = None
            if has_bias and ctx.needs_input_grad[3]:
                gm = _relu_grad(g, y) if act_code == ops.ACT_RELU else g
                grad_b = ops.colsum(gm)
            return res[0], res[1], res[2], grad_b, None, None, None, None, None, None, None, None
        if act_code == ops.ACT_RELU:
            g = _relu_grad(g, y)
        inv_scale = 1.0 / scale
        ds = ops.gat_softmax_bwd(csr, att, g, V, H, split_value_heads=split, drop_rate=drop_rate, seed=seed)
        grad_q = grad_k = grad_v = grad_b = None
        if ctx.needs_input_grad[0]:
            grad_q = ops.spmm_heads(csr, ds, K, H, alpha=inv_scale)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            csr_t, emap = _transposed_of_csr(csr, edge_index_used)
            if ctx.needs_input_grad[1]:
                grad_k = ops.spmm_heads(csr_t, ds, Q, H, emap=emap, alpha=inv_scale)
            if ctx.needs_input_grad[2]:
                grad_v = ops.spmm_heads(csr_t, att, g, H, mode=ops.HEADS_SPLIT if split else ops.HEADS_BROADCAST,
                                        emap=emap, drop_rate=drop_rate, seed=seed, alpha=1.0 if split else 1.0 / H)
        if has_bias and ctx.needs_input_grad[3]:
            grad_b = ops.colsum(g)
        return grad_q, grad_k, grad_v, grad_b, None, None, None, None, None, None, None, None


class Dropout(torch.autograd.Function):
    """tf.nn.dropout on dense activations (appnp.py:75-79, ssgc.py:84-88); the backward re-applies the same
    counter-based mask to the incoming gradient."""

    @staticmethod
    def forward(ctx, x, rate, seed):
        ctx.rate, ctx.seed = float(rate), int(seed)
        return ops.dropout(x.detach().contiguous(), ctx.rate, ctx.seed)

    @staticmethod
    def backward(ctx, grad_y):
        return ops.dropout(grad_y.contiguous(), ctx.rate, ctx.seed), None, None


def dropout(x, rate, training, seed=None):
    """Functional helper: identity unless training and rate > 0."""
    if not training or rate <= 0.0:
        return x
    from . import _rng
    return Dropout.apply(x, float(rate), _rng.resolve(seed))


def dense(x, weight, bias=None, activation=None):
    """Differentiable act(x @ W + b): relu is fused into the GEMM epilogue, any other callable runs on the result."""
    code, leftover = ops.activation_code(activation)
    y = Dense.apply(x, weight, bias, code)
    return leftover(y) if leftover is not None else y


def propagate(adj, h, bias=None, act_code=ops.ACT_NONE):
    """Differentiable act(A @ h + b) for a SparseMatrix A (gradient w.r.t. h and b)."""
    return SparseMatmul.apply(h, bias, adj, act_code)


class SegmentReduce(torch.autograd.Function):
    """out[s] = REDUCE_{i: ids_i = s} data[i] for sum | mean | max | min over a plain id vector (the stock reducers of
    nn/kernel/map_reduce.py:15-42 and the graph pooling of nn/pool/common_pool.py), differentiable w.r.t. data.
    Backward: sum/mean are a gather of the upstream rows by segment id (scaled by 1 / max(count, 1) for mean); max/min
    route the gradient to the selected entries, shared equally among ties like TensorFlow's UnsortedSegmentMax gradient."""

    @staticmethod
    def forward(ctx, data, ids, num_segments, reduce):
        csr = _structure.csr_for_segment_ids(ids, int(num_segments))
        d = data.detach()
        if reduce == "min":       # min(x) = -max(-x): weight -1 per message and epilogue scale -1, both exact
            minus = torch.full((csr.nnz,), -1.0, dtype=torch.float32, device=d.device)
            out = ops.spmm(csr, minus, d, reduce="max", alpha=-1.0, col=csr.perm)
        else:
            out = ops.spmm(csr, None, d, reduce=reduce, col=csr.perm)
        ctx.ids, ctx.csr, ctx.reduce = ids, csr, reduce
        ctx.save_for_backward(d if reduce in ("max", "min") else None, out if reduce in ("max", "min") else None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        data, out = ctx.saved_tensors
        ids, csr, reduce = ctx.ids, ctx.csr, ctx.reduce
        g = grad_out.contiguous()
        if reduce == "mean":
            cnt = (csr.rowptr[1:] - csr.rowptr[:-1]).clamp(min=1).to(torch.float32)
            g = g / cnt.unsqueeze(1)
        if reduce in ("sum", "mean"):
            return ops.permute(g, ids), None, None, None
        selected = (data == ops.permute(out, ids)).to(torch.float32)
        n_selected = ops.spmm(csr, None, selected, reduce="sum", col=csr.perm).clamp(min=1.0)
        return ops.permute(g / n_selected, ids) * selected, None, None, None


class TakeRows(torch.autograd.Function):
    """data[index] (tf.gather along axis 0) through the gather kernel; backward = segment sum of the upstream rows by
    index (a deterministic scatter-add on the CSR kernel)."""

    @staticmethod
    def forward(ctx, data, index):
        ctx.index, ctx.n = index, data.shape[0]
        return ops.permute(data.detach(), index)

    @staticmethod
    def backward(ctx, grad_out):
        csr = _structure.csr_for_segment_ids(ctx.index, ctx.n)
        g = grad_out.contiguous()
        flat = g if g.dim() == 2 else g.unsqueeze(1)
        res = ops.spmm(csr, None, flat, reduce="sum", col=csr.perm)
        return (res if g.dim() == 2 else res.squeeze(1)), None


class SagePair(torch.autograd.Function):
    """mean / sum GraphSAGE (nn/conv/graph_sage.py:9-115) as ONE differentiable op:
        out = act([x Ws || agg Wn] + b)   (or x Ws + agg Wn + b),   agg = REDUCE_{e: row_e = r} w_e x[col_e].
    Compared with composing NeighborAggregate and two Dense Functions this writes both projections straight into the
    output (no concat copy), masks the upstream gradient once, and folds dX_self into the epilogue of the transposed
    aggregation (grad_x = A^T d_agg + dX_self in one pass, tfgk_spmm_f32's addend).  dX products run on the tensor-core
    projection kernel (transB), dW products are split-K over the node dimension."""

    @staticmethod
    def forward(ctx, x, ws, wn, bias, edge_index, edge_weight, reduce, act_code, concat):
        n = x.shape[0]
        csr, _ = _structure.csr_for_edge_index(edge_index, n)
        w_csr = None if edge_weight is None else _structure.weights_in_csr_order(edge_weight, csr)
        xd, wsd, wnd = x.detach(), ws.detach(), wn.detach()
        b = None if bias is None else bias.detach()
        agg = ops.spmm(csr, w_csr, xd, reduce=reduce)
        u = wsd.shape[1]
        if concat:
            out = torch.empty((n, u + wnd.shape[1]), dtype=torch.float32, device=xd.device)
            ops.gemm(xd, wsd, bias=None if b is None else b[:u].contiguous(), act=act_code, out=out[:, :u])
            ops.gemm(agg, wnd, bias=None if b is None else b[u:].contiguous(), act=act_code, out=out[:, u:])
        else:
            out = ops.gemm(xd, wsd)
            ops.gemm(agg, wnd, bias=b, act=act_code, beta=1.0, out=out)
        ctx.save_for_backward(x, ws, wn, agg, out if act_code == ops.ACT_RELU else None)
        ctx.meta = (edge_index, edge_weight, reduce, act_code, concat, csr, bias is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, ws, wn, agg, out = ctx.saved_tensors
        edge_index, edge_weight, reduce, act_code, concat, csr, has_bias = ctx.meta
        gm = grad_out.contiguous()
        if act_code == ops.ACT_RELU:
            gm = _relu_grad(gm, out)
        u = ws.shape[1]
        gs, gn = (gm[:, :u], gm[:, u:]) if concat else (gm, gm)
        xd, wsd, wnd = x.detach(), ws.detach(), wn.detach()
        grad_x = grad_ws = grad_wn = grad_b = None
        if ctx.needs_input_grad[0]:
            d_agg = ops.gemm(gn, wnd, trans_b=True)
            dx_self = ops.gemm(gs, wsd, trans_b=True)
            csr_t, w_t = _transposed_structure(edge_index, x.shape[0], edge_weight, reduce == "mean", csr)
            grad_x = ops.spmm(csr_t, w_t, d_agg, reduce="sum", alpha=1.0, addend=dx_self, beta=1.0)
        if ctx.needs_input_grad[1]:
            grad_ws = ops.gemm(xd, gs, trans_a=True)
        if ctx.needs_input_grad[2]:
            grad_wn = ops.gemm(agg, gn, trans_a=True)
        if has_bias and ctx.needs_input_grad[3]:
            grad_b = ops.colsum(gm)
        return grad_x, grad_ws, grad_wn, grad_b, None, None, None, None, None


def needs_grad(*tensors):
    return torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in tensors)

# coding=utf-8
"""TEST DOUBLE for the kernel layer: replaces the thin wrappers of tf_geometric_b200.ops with the CPU oracle so the
HOST logic (argument plumbing, caching, quirks, layer wiring) can be exercised in a GPU-less container.
It lives under tests/ and is injected with monkeypatch; the product has no such path."""
import numpy as np
import torch

from oracle import tfg_oracle as o
from oracle import c_oracle


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def install(monkeypatch):
    from tf_geometric_b200 import ops, _structure
    _structure.clear()
    cpu = torch.device("cpu")
    monkeypatch.setattr(ops, "_require_cuda", lambda: None)
    monkeypatch.setattr(ops, "default_device", lambda: cpu)

    def as_device(x, dtype=None, device=None):
        if x is None:
            return None
        if not torch.is_tensor(x):
            x = torch.from_numpy(np.ascontiguousarray(x))
        if dtype is not None and x.dtype != dtype:
            x = x.to(dtype)
        return x.contiguous()
    monkeypatch.setattr(ops, "as_device", as_device)

    def self_loops(edge_index, num_nodes):
        return _t(o.add_self_loop_edge(_np(edge_index), num_nodes)[0])

    def self_loop_weights(w, num_edges, num_nodes, fill, device):
        base = np.ones(num_edges, np.float32) if w is None else _np(w)
        return _t(np.concatenate([base, np.full(num_nodes, fill, np.float32)]))

    def segment_count(ids, n):
        return _t(np.bincount(_np(ids), minlength=n).astype(np.int32))

    def csr_build(row, col, n_rows, n_cols=None):
        rowptr, cs, perm = c_oracle.csr_build(_np(row), _np(col), n_rows)
        return ops.CSR(_t(rowptr), _t(cs), _t(perm), n_rows, n_rows if n_cols is None else n_cols)

    def permute(src, perm, inverse=False):
        s, p = _np(src), _np(perm)
        if inverse:
            out = np.empty_like(s)
            out[p] = s
            return _t(out)
        return _t(s[p])

    def csr_rowsum(csr, w):
        ids = np.repeat(np.arange(csr.n_rows), np.diff(_np(csr.rowptr)))
        return _t(o.unsorted_segment_sum(_np(w), ids, csr.n_rows))

    def deg_inv(deg, power):
        with np.errstate(divide="ignore", invalid="ignore"):
            return _t(o._remove_inf_and_nan(np.power(_np(deg), np.float32(-0.5 if power == ops.POW_INV_SQRT else -1))))

    def scale_edges(row, col, w, dl=None, dr=None):
        v = _np(w)
        if dl is not None:
            v = _np(dl)[_np(row)] * v
        if dr is not None:
            v = v * _np(dr)[_np(col)]
        return _t(v.astype(np.float32))

    def spmm(csr, w_csr, h, reduce="sum", alpha=1.0, addend=None, beta=0.0, bias=None, act=0, out=None, col=None):
        ids = np.repeat(np.arange(csr.n_rows), np.diff(_np(csr.rowptr))).astype(np.int32)
        cols = _np(csr.col if col is None else col)
        agg = c_oracle.aggregate(ids, cols, _np(w_csr), _np(h), csr.n_rows, reduce if isinstance(reduce, str)
                                 else ["sum", "mean", "max"][reduce])
        if addend is not None:
            agg = agg * np.float32(alpha) + _np(addend) * np.float32(beta)
        elif alpha != 1.0:
            agg = agg * np.float32(alpha)
        if bias is not None:
            agg = agg + _np(bias)
        if act == ops.ACT_RELU:
            agg = np.maximum(agg, 0)
        res = _t(agg.astype(np.float32))
        if out is not None:
            out.copy_(res)
            return out
        return res

    def segment_softmax_csr(csr, score):
        ids = np.repeat(np.arange(csr.n_rows), np.diff(_np(csr.rowptr))).astype(np.int32)
        return _t(c_oracle.segment_softmax(_np(score), ids, csr.n_rows))

    def gat_fused(csr, Q, K, V, num_heads, split_value_heads=True, bias=None, act=0, return_attention=False,
                  att_buffer=None, out=None, scale=None):
        ids = np.repeat(np.arange(csr.n_rows), np.diff(_np(csr.rowptr))).astype(np.int32)
        q = _np(Q)
        if scale is not None:       # the C restatement divides by sqrt(dqk); fold the requested divisor into Q
            q = (q * np.float32(np.sqrt(np.float32(q.shape[1] // num_heads)) / np.float32(scale))).astype(np.float32)
        res, att = c_oracle.gat_core(ids, _np(csr.col), q, _np(K), _np(V), num_heads, split_value_heads, True)
        if bias is not None:
            res = res + _np(bias)
        if act == ops.ACT_RELU:
            res = np.maximum(res, 0)
        return (_t(res), _t(att)) if return_attention else _t(res)

    def gemm(a, b, bias=None, act=0, trans_a=False, trans_b=False, beta=0.0, out=None):
        an, bn = _np(a), _np(b)
        r = (an.T if trans_a else an) @ (bn.T if trans_b else bn)
        if beta != 0.0:
            r = r + np.float32(beta) * _np(out)
        if bias is not None:
            r = r + _np(bias)
        if act == ops.ACT_RELU:
            r = np.maximum(r, 0)
        res = _t(r.astype(np.float32))
        if out is not None:
            out.copy_(res)
            return out
        return res

    def gemm_proj(a, blocks, a_parts=None, part_rows=0, first_part=0, max_ctas=0, num_rows=None):
        assert a_parts is None, "peer-mapped inputs need a GPU"
        outs = []
        for blk in blocks:
            w, bias, act, out = blk[:4]
            res = gemm(a if num_rows is None else a[:num_rows], w, bias=bias, act=act, trans_b=len(blk) > 4 and bool(blk[4]))
            if out is not None:
                out.copy_(res)
                res = out
            outs.append(res)
        return outs

    def gat_fused_stats(csr, Q, K, V, num_heads, bias=None, act=0, scale=None):
        return None          # the recompute-backward kernels are GPU-only: the host logic then keeps the coefficient table

    def colsum(x):
        return _t(_np(x).sum(axis=0, dtype=np.float32))

    def l2_normalize(x, out=None):
        return _t(o.l2_normalize(_np(x)))

    def edge_unique(row, col, num_nodes):
        idx, _ = o.merge_duplicated_edge(np.stack([_np(row), _np(col)]))
        h = _np(row).astype(np.int64) * num_nodes + _np(col)
        _, of_edge = o.tf_unique(h)
        return _t(idx), _t(of_edge.astype(np.int32))

    def directed_edges(upper):
        up = _np(upper)
        mask = up[0] != up[1]
        lower = np.stack([up[1][mask], up[0][mask]])
        return _t(np.concatenate([up, lower], axis=1).astype(np.int32)), _t(np.nonzero(mask)[0].astype(np.int32))

    _MODES = {ops.HEADS_SPLIT: "split", ops.HEADS_BROADCAST: "broadcast", ops.HEADS_REDUCE: "reduce"}

    def dropout(x, rate, seed, rng_stream=0, out=None):
        return _t(o.dropout(_np(x), rate, seed, rng_stream))

    def spmm_heads(csr, w, src, num_heads, mode=ops.HEADS_SPLIT, emap=None, drop_rate=0.0, seed=0, rng_stream=0,
                   alpha=1.0, bias=None, act=0, out=None):
        return _t(o.spmm_heads(_np(csr.rowptr), _np(csr.col), _np(w), _np(src), num_heads, _MODES[mode], _np(emap),
                               drop_rate, seed, alpha, _np(bias), "relu" if act == ops.ACT_RELU else None))

    def gat_softmax_bwd(csr, att, G, V, num_heads, split_value_heads=True, drop_rate=0.0, seed=0, rng_stream=0):
        return _t(o.gat_softmax_bwd(_np(csr.rowptr), _np(csr.col), _np(att), _np(G), _np(V), num_heads,
                                    split_value_heads, drop_rate, seed))

    def edge_flags(row, col, num_edges, mode=ops.FLAG_ALL, row_map=None, col_map=None, bernoulli=ops.BERNOULLI_NONE,
                   prob=0.0, seed=0, rng_stream=1, device=None):
        keep = np.ones(num_edges, bool)
        if mode == ops.FLAG_UPPER:
            keep = _np(row) < _np(col)
        elif mode == ops.FLAG_MAPPED:
            keep = (_np(row_map)[_np(row)] >= 0) & (_np(col_map)[_np(col)] >= 0)
        if bernoulli != ops.BERNOULLI_NONE:
            u = o.random_uniform(seed, rng_stream, np.arange(num_edges, dtype=np.uint64))
            keep = keep & ((u >= np.float32(prob)) if bernoulli == ops.BERNOULLI_DROPOUT else (u <= np.float32(prob)))
        return _t(keep.astype(np.int32))

    def select_flagged(flag):
        return _t(np.nonzero(_np(flag))[0].astype(np.int32))

    def gather_i32(src, index):
        return _t(_np(src)[_np(index)])

    def neighbor_sample(csr, k=None, ratio=None, padding=False, seed=0, rng_stream=1):
        if not isinstance(padding, bool) and padding == ops.SAMPLE_HEAD:
            padding = "head"
        r, p, rp = o.neighbor_sample_csr(_np(csr.rowptr), k, ratio, padding, seed, rng_stream)
        return _t(r), _t(p.astype(np.int32)), _t(rp)

    def sort_keys_f32(score, descending=False):
        u = (_np(score) + np.float32(0)).view(np.uint32)
        u = np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)
        return _t((~u if descending else u).view(np.int32))

    def stable_argsort(keys, key_bits=32):
        mask = np.uint32(0xFFFFFFFF >> (32 - key_bits))
        return _t(np.argsort(_np(keys).view(np.uint32) & mask, kind="stable").astype(np.int32))

    for name, fn in dict(dropout=dropout, spmm_heads=spmm_heads, gat_softmax_bwd=gat_softmax_bwd, edge_flags=edge_flags,
                         select_flagged=select_flagged, gather_i32=gather_i32, neighbor_sample=neighbor_sample,
                         sort_keys_f32=sort_keys_f32, stable_argsort=stable_argsort).items():
        monkeypatch.setattr(ops, name, fn)

    from tf_geometric_b200.utils import graph_utils as gu
    monkeypatch.setattr(gu, "_is_device", lambda t: torch.is_tensor(t))     # CPU tensors take the device code path
    monkeypatch.setattr(ops, "edge_unique", edge_unique)
    monkeypatch.setattr(ops, "directed_edges", directed_edges)

    for name, fn in dict(self_loops=self_loops, self_loop_weights=self_loop_weights, segment_count=segment_count,
                         csr_build=csr_build, permute=permute, csr_rowsum=csr_rowsum, deg_inv=deg_inv,
                         scale_edges=scale_edges, spmm=spmm, segment_softmax_csr=segment_softmax_csr,
                         gat_fused=gat_fused, gemm=gemm, gemm_proj=gemm_proj, colsum=colsum, gat_fused_stats=gat_fused_stats, l2_normalize=l2_normalize).items():
        monkeypatch.setattr(ops, name, fn)

# coding=utf-8
"""Typed wrappers over the C ABI (include/tfgk.h): torch CUDA tensors in, torch CUDA tensors out.

Nothing here computes with PyTorch: tensors are allocated with torch.empty and handed to libtfgk.so by pointer on
the current CUDA stream.  Every function requires CUDA tensors and raises otherwise (no CPU path).
"""
import ctypes

import numpy as np
import torch

from . import _ffi
from ._ffi import (REDUCE_SUM, REDUCE_MEAN, REDUCE_MAX, ACT_NONE, ACT_RELU, POW_INV_SQRT, POW_INV,  # noqa: F401
                   HEADS_SPLIT, HEADS_BROADCAST, HEADS_REDUCE, FLAG_ALL, FLAG_UPPER, FLAG_MAPPED,
                   BERNOULLI_NONE, BERNOULLI_DROPOUT, BERNOULLI_KEEP, SAMPLE_NO_PADDING, SAMPLE_PADDING, SAMPLE_HEAD)

_REDUCE_CODES = {"sum": REDUCE_SUM, "mean": REDUCE_MEAN, "max": REDUCE_MAX}


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("tf_geometric_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")


def default_device():
    _require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def as_device(x, dtype=None, device=None):
    """numpy / list / torch (any device) -> contiguous CUDA tensor of `dtype` (reference casting rules are applied
    by the callers: int32 edge_index, float32 weights/features; data/graph.py:58-86).
    A conversion creates a NEW tensor on every call, and the CSR caches are keyed on tensor identity: for the warm path
    (no re-sort per forward) hand the layers an int32 CUDA edge_index - `Graph.to_device()` produces one - or pass
    `cache=graph.cache`."""
    if x is None:
        return None
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if device is None:
        device = x.device if x.is_cuda else default_device()
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    if x.device != device:
        x = x.to(device, non_blocking=True)
    return x.contiguous()


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check(t, dtype, name):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise TypeError("{} must be a CUDA tensor (got {})".format(name, type(t)))
    if t.dtype != dtype:
        raise TypeError("{} must be {} (got {})".format(name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("{} must be contiguous".format(name))


def _row_major_2d(t, name):
    """Accept [N, D] tensors whose rows are contiguous (stride(1) == 1); returns the leading dimension."""
    if t.dim() != 2:
        raise ValueError("{} must be 2-D".format(name))
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError("{} rows must be contiguous".format(name))
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


HUB_THRESHOLD = 2048     # rows with more edges are cut into slices ...
HUB_CHUNK = 2048         # ... of this many edges (about the work of one ordinary 32-row task)
ROWS_PER_TASK = 32


class Plan(object):
    """Device-side work plan of a CSR with hub rows (struct tfgk_plan): task arrays + a scratch allocator."""

    def __init__(self, arrays, n_tasks, n_hubs, n_slots):
        self.arrays = arrays            # keeps the device tensors alive
        self.n_tasks, self.n_hubs, self.n_slots = n_tasks, n_hubs, n_slots
        self._scratch = None

    def struct(self, floats_per_slot, device):
        # the hub-slice scratch is allocated per call (the caching allocator makes that cheap and stream-ordered): a buffer
        # kept on the plan would be shared by launches on different streams that use the same cached CSR
        need = max(self.n_slots * floats_per_slot, 1)
        self._scratch = torch.empty((need,), dtype=torch.float32, device=device)
        a = self.arrays
        st = _ffi.PlanStruct(self.n_tasks, self.n_hubs, self.n_slots, HUB_CHUNK,
                             a["task_row"].data_ptr(), a["task_nrows"].data_ptr(), a["task_e0"].data_ptr(),
                             a["task_e1"].data_ptr(), a["task_slot"].data_ptr(), a["hub_row"].data_ptr(),
                             a["hub_slot0"].data_ptr(), a["hub_nslots"].data_ptr(), self._scratch.data_ptr(),
                             self._scratch.numel() * 4)
        return st


class CSR(object):
    """Destination-sorted CSR of a COO edge list (stable by row): rowptr int64 [n_rows+1], col int32 [nnz],
    perm int32 [nnz] (position of each CSR slot in the original edge list); `plan` is set when the graph has hub rows."""

    __slots__ = ("rowptr", "col", "perm", "n_rows", "n_cols", "nnz", "plan", "__weakref__")

    def __init__(self, rowptr, col, perm, n_rows, n_cols):
        self.rowptr, self.col, self.perm = rowptr, col, perm
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.nnz = int(col.shape[0])
        self.plan = None

    def degree_i64(self):
        return self.rowptr[1:] - self.rowptr[:-1]


# ---- integer edge preprocessing ---------------------------------------------------------------------------------

def self_loops(edge_index, num_nodes):
    _check(edge_index, torch.int32, "edge_index")
    E = edge_index.shape[1] if edge_index.dim() == 2 else 0
    out = torch.empty((2, E + num_nodes), dtype=torch.int32, device=edge_index.device)
    _ffi.call("tfgk_self_loops_i32", _p(edge_index), E, num_nodes, _p(out), _stream(out))
    return out


def self_loop_weights(edge_weight, num_edges, num_nodes, fill_weight, device):
    if edge_weight is not None:
        _check(edge_weight, torch.float32, "edge_weight")
    out = torch.empty((num_edges + num_nodes,), dtype=torch.float32, device=device)
    _ffi.call("tfgk_self_loop_weights_f32", _p(edge_weight), num_edges, num_nodes, float(fill_weight), _p(out),
              _stream(out))
    return out


def segment_count(ids, num_segments):
    _check(ids, torch.int32, "index")
    out = torch.empty((num_segments,), dtype=torch.int32, device=ids.device)
    _ffi.call("tfgk_segment_count_i32", _p(ids), ids.numel(), num_segments, _p(out), _stream(out))
    return out


def csr_build(row, col, n_rows, n_cols=None):
    _check(row, torch.int32, "row")
    _check(col, torch.int32, "col")
    n_cols = n_rows if n_cols is None else n_cols
    E = row.numel()
    dev = row.device
    need = ctypes.c_size_t()
    _ffi.call("tfgk_csr_workspace_bytes", E, n_rows, ctypes.byref(need))
    ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=dev)
    rowptr = torch.empty((n_rows + 1,), dtype=torch.int64, device=dev)
    col_sorted = torch.empty((E,), dtype=torch.int32, device=dev)
    perm = torch.empty((E,), dtype=torch.int32, device=dev)
    _ffi.call("tfgk_csr_build", _p(row), _p(col), E, n_rows, n_cols, _p(rowptr), _p(col_sorted), _p(perm),
              _p(ws), need.value, _stream(row))
    csr = CSR(rowptr, col_sorted, perm, n_rows, n_cols)
    csr.plan = build_plan(csr)
    return csr


DENSE_ROW_DEGREE = 128   # average row length from which 32-row tasks starve the machine (few, long rows: pooling)
TASK_EDGE_TARGET = 512   # ... and the edges one task should then hold


def build_plan(csr):
    """Work plan: hub rows (> HUB_THRESHOLD edges) are cut into slices, and when the rows are few and long on average
    (graph pooling, set2set: one row per graph) the tasks shrink from 32 rows to about TASK_EDGE_TARGET edges so that
    there are enough warps to fill the GPU.  None when neither applies (the kernels then use their implicit 32-row
    tasks, which is also the fastest path for ordinary graphs)."""
    if csr.nnz == 0 or csr.n_rows == 0:
        return None
    avg = csr.nnz / float(csr.n_rows)
    rows_per_task = ROWS_PER_TASK if avg < DENSE_ROW_DEGREE else max(1, min(ROWS_PER_TASK, int(TASK_EDGE_TARGET // avg)))
    dev = csr.rowptr.device
    cap_t, cap_h = ctypes.c_int64(), ctypes.c_int64()
    _ffi.call("tfgk_plan_capacity", csr.nnz, csr.n_rows, HUB_THRESHOLD, HUB_CHUNK, rows_per_task, ctypes.byref(cap_t),
              ctypes.byref(cap_h))
    need = ctypes.c_size_t()
    _ffi.call("tfgk_plan_workspace_bytes", csr.n_rows, ctypes.byref(need))
    ws = torch.empty((need.value,), dtype=torch.uint8, device=dev)
    arrays = {k: torch.empty((cap_t.value,), dtype=torch.int32, device=dev) for k in ("task_row", "task_nrows", "task_slot")}
    arrays.update({k: torch.empty((cap_t.value,), dtype=torch.int64, device=dev) for k in ("task_e0", "task_e1")})
    arrays.update({k: torch.empty((cap_h.value,), dtype=torch.int32, device=dev) for k in ("hub_row", "hub_slot0", "hub_nslots")})
    counts = (ctypes.c_int32 * 3)()
    _ffi.call("tfgk_plan_build", _p(csr.rowptr), csr.n_rows, HUB_THRESHOLD, HUB_CHUNK, rows_per_task,
              _p(arrays["task_row"]), _p(arrays["task_nrows"]), _p(arrays["task_e0"]), _p(arrays["task_e1"]),
              _p(arrays["task_slot"]), _p(arrays["hub_row"]), _p(arrays["hub_slot0"]), _p(arrays["hub_nslots"]),
              cap_t.value, cap_h.value, counts, _p(ws), need.value, _stream(csr.rowptr))
    n_tasks, n_hubs, n_slots = int(counts[0]), int(counts[1]), int(counts[2])
    if n_hubs == 0 and rows_per_task == ROWS_PER_TASK:
        return None
    arrays = {k: v[:(n_tasks if k.startswith("task") else n_hubs)].clone() for k, v in arrays.items()}
    return Plan(arrays, n_tasks, n_hubs, n_slots)


def edge_unique(row, col, num_nodes):
    """Unique (row, col) pairs in first-occurrence order and, per edge, the index of its representative
    (merge_duplicated_edge's tf.unique step).  Returns (unique_edge_index int32 [2, U], unique_of_edge int32 [E])."""
    _check(row, torch.int32, "row")
    _check(col, torch.int32, "col")
    E = row.numel()
    dev = row.device
    need = ctypes.c_size_t()
    _ffi.call("tfgk_edge_unique_workspace_bytes", E, num_nodes, ctypes.byref(need))
    ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=dev)
    uniq = torch.empty((2, max(E, 1)), dtype=torch.int32, device=dev)
    of_edge = torch.empty((E,), dtype=torch.int32, device=dev)
    n_unique = ctypes.c_int32()
    _ffi.call("tfgk_edge_unique", _p(row), _p(col), E, num_nodes, _p(uniq), _p(of_edge), ctypes.byref(n_unique), _p(ws),
              need.value, _stream(row))
    return uniq[:, :n_unique.value].contiguous(), of_edge


def directed_edges(upper_index):
    """upper edges followed by the mirrored non-self-loop upper edges; also the source column of every mirrored edge."""
    _check(upper_index, torch.int32, "upper_index")
    U = upper_index.shape[1]
    dev = upper_index.device
    need = ctypes.c_size_t()
    _ffi.call("tfgk_directed_workspace_bytes", U, ctypes.byref(need))
    ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=dev)
    out = torch.empty((2, max(2 * U, 1)), dtype=torch.int32, device=dev)
    lower_src = torch.empty((max(U, 1),), dtype=torch.int32, device=dev)
    n_lower = ctypes.c_int32()
    _ffi.call("tfgk_directed_edges", _p(upper_index), U, U, _p(out), max(2 * U, 1), _p(lower_src), ctypes.byref(n_lower),
              _p(ws), need.value, _stream(upper_index))
    return out[:, :U + n_lower.value].contiguous(), lower_src[:n_lower.value]


def permute(src, perm, inverse=False):
    """COO-order values -> CSR order (dst[i] = src[perm[i]]), or back with inverse=True.  src: [E] or [E, W]."""
    _check(src, torch.float32, "src")
    _check(perm, torch.int32, "perm")
    width = 1 if src.dim() == 1 else src.shape[1]
    if inverse:
        if src.shape[0] != perm.numel():
            raise ValueError("unpermute: src must have one row per index")
        dst = torch.empty_like(src)
    else:                                    # a gather: one output row per index (perm may select a subset of src)
        dst = torch.empty((perm.numel(),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    _ffi.call("tfgk_unpermute_f32" if inverse else "tfgk_permute_f32", _p(src), _p(perm), perm.numel(), width, _p(dst),
              _stream(src))
    return dst


def csr_rowsum(csr, w_csr):
    _check(w_csr, torch.float32, "w_csr")
    out = torch.empty((csr.n_rows,), dtype=torch.float32, device=w_csr.device)
    _ffi.call("tfgk_csr_rowsum_f32", _p(csr.rowptr), _p(w_csr), csr.n_rows, _p(out), _stream(out))
    return out


def deg_inv(deg, power):
    _check(deg, torch.float32, "deg")
    out = torch.empty_like(deg)
    _ffi.call("tfgk_deg_inv_f32", _p(deg), deg.numel(), power, _p(out), _stream(out))
    return out


def scale_edges(row, col, w, dl=None, dr=None):
    _check(w, torch.float32, "value")
    out = torch.empty_like(w)
    _ffi.call("tfgk_scale_edges_f32", _p(row), _p(col), _p(w), w.numel(), _p(dl), _p(dr), _p(out), _stream(out))
    return out


# ---- K1 ----------------------------------------------------------------------------------------------------------

def spmm(csr, w_csr, h, reduce="sum", alpha=1.0, addend=None, beta=0.0, bias=None, act=ACT_NONE, out=None, col=None):
    """out = epilogue(REDUCE_{e in row} w[e] * h[col[e]]); see tfgk_spmm_f32.  `col` overrides csr.col (used by the
    generic reducers, which gather message rows through csr.perm)."""
    if h.dtype != torch.float32 or not h.is_cuda:
        raise TypeError("h must be a float32 CUDA tensor")
    ldh = _row_major_2d(h, "h")
    n_dst, D = csr.n_rows, h.shape[1]
    if out is None:
        out = torch.empty((n_dst, D), dtype=torch.float32, device=h.device)
    ldo = _row_major_2d(out, "out")
    lda = 0
    if addend is not None:
        lda = _row_major_2d(addend, "addend")
    if w_csr is not None:
        _check(w_csr, torch.float32, "w_csr")
    if bias is not None:
        _check(bias, torch.float32, "bias")
    code = _REDUCE_CODES[reduce] if isinstance(reduce, str) else reduce
    plan = getattr(csr, "plan", None)
    plan_struct = plan.struct(D, h.device) if plan is not None else None
    _ffi.call("tfgk_spmm_f32", _p(csr.rowptr), _p(csr.col if col is None else col), _p(w_csr), _p(h), ldh, n_dst, D,
              code, float(alpha), _p(addend), lda, float(beta), _p(bias), act, _p(out), ldo,
              ctypes.byref(plan_struct) if plan_struct is not None else None, _stream(h))
    return out


# ---- K3 ----------------------------------------------------------------------------------------------------------

def segment_softmax_csr(csr, score_csr):
    """score_csr: [E] or [E, H] in CSR order."""
    _check(score_csr, torch.float32, "score")
    H = 1 if score_csr.dim() == 1 else score_csr.shape[1]
    out = torch.empty_like(score_csr)
    _ffi.call("tfgk_segment_softmax_f32", _p(csr.rowptr), _p(score_csr), csr.n_rows, H, _p(out), _stream(out))
    return out


def gat_fused(csr, Q, K, V, num_heads, split_value_heads=True, bias=None, act=ACT_NONE, return_attention=False,
              att_buffer=None, out=None, scale=None):
    for t, n in ((Q, "Q"), (K, "K"), (V, "V")):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise TypeError("{} must be a float32 CUDA tensor".format(n))
    N = csr.n_rows
    H = int(num_heads)
    A, VW = Q.shape[1], V.shape[1]
    if A % H or VW % H or K.shape[1] != A:
        raise ValueError("attention units ({}) and value units ({}) must be divisible by num_heads ({})".format(A, VW, H))
    dqk, dv = A // H, VW // H
    out_w = VW if split_value_heads else dv
    if out is None:
        out = torch.empty((N, out_w), dtype=torch.float32, device=Q.device)
    att = att_buffer
    if att is not None and att.numel() < csr.nnz * H:
        att = None
    if att is None and return_attention:
        att = torch.empty((csr.nnz, H), dtype=torch.float32, device=Q.device)
    if bias is not None:
        _check(bias, torch.float32, "bias")
    # gat.py:78  scale = sqrt(cast(shape(Q_)[-1], float32)); set2set.py:37 uses raw dot products (scale = 1)
    scale = float(np.sqrt(np.float32(dqk))) if scale is None else float(scale)

    plan = getattr(csr, "plan", None)
    plan_struct = plan.struct(VW + 64, Q.device) if plan is not None else None

    def launch(att_buf):
        _ffi.call("tfgk_gat_fused_f32", _p(csr.rowptr), _p(csr.col), _p(Q), _row_major_2d(Q, "Q"), _p(K),
                  _row_major_2d(K, "K"), _p(V), _row_major_2d(V, "V"), N, H, dqk, dv, scale,
                  1 if split_value_heads else 0, _p(bias), act, _p(att_buf), 1 if return_attention else 0, _p(out),
                  _row_major_2d(out, "out"), ctypes.byref(plan_struct) if plan_struct is not None else None, _stream(Q))

    try:
        launch(att)
    except _ffi.TfgkError as err:
        # the single-pass kernel needs no scratch; the two-pass / generic kernels ask for an [E, H] score buffer
        if err.code != _ffi.ERR_WORKSPACE or att is not None:
            raise
        att = torch.empty((csr.nnz, H), dtype=torch.float32, device=Q.device)
        launch(att)
    if return_attention:
        return out, att[:csr.nnz]
    return out


def gat_fused_stats(csr, Q, K, V, num_heads, bias=None, act=ACT_NONE, scale=None):
    """Training forward without the [E, H] coefficient table: returns (out, stats[N, 2H]) or None when the shape is not
    taken by the streaming kernel (tfgk_gat_fused_stats_f32)."""
    N, H = csr.n_rows, int(num_heads)
    A = Q.shape[1]
    if A % H or V.shape[1] != A or K.shape[1] != A:
        return None
    dqk = A // H
    scale = float(np.sqrt(np.float32(dqk))) if scale is None else float(scale)
    out = torch.empty((N, A), dtype=torch.float32, device=Q.device)
    stats = torch.empty((N, 2 * H), dtype=torch.float32, device=Q.device)
    plan = getattr(csr, "plan", None)
    plan_struct = plan.struct(A + 64, Q.device) if plan is not None else None
    try:
        _ffi.call("tfgk_gat_fused_stats_f32", _p(csr.rowptr), _p(csr.col), _p(Q), _row_major_2d(Q, "Q"), _p(K),
                  _row_major_2d(K, "K"), _p(V), _row_major_2d(V, "V"), N, H, dqk, dqk, scale, _p(bias), act, _p(out),
                  _row_major_2d(out, "out"), _p(stats), ctypes.byref(plan_struct) if plan_struct is not None else None, _stream(Q))
    except _ffi.TfgkError as err:
        if err.code != _ffi.ERR_UNSUPPORTED:
            raise
        return None
    return out, stats


def gat_backward_recompute(csr, csr_t, Q, K, V, G, Y, bias, act, stats, num_heads, scale):
    """(dQ, dK, dV) of the fused attention aggregation from (max, denominator) per row (tfgk_gat_bwd_*); None when the
    kernels do not take the shape."""
    N, H = csr.n_rows, int(num_heads)
    A = Q.shape[1]
    dqk = A // H
    GS = torch.empty((N, A + 32), dtype=torch.float32, device=Q.device)
    dQ, dK, dV = torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
    try:
        _ffi.call("tfgk_gat_bwd_prepare_f32", _p(G), _row_major_2d(G, "G"), _p(Y), _row_major_2d(Y, "Y"), _p(bias), act,
                  _p(stats), N, H, dqk, _p(GS), A + 32, _stream(Q))
        _ffi.call("tfgk_gat_bwd_dst_f32", _p(csr.rowptr), _p(csr.col), _p(Q), _row_major_2d(Q, "Q"), _p(K),
                  _row_major_2d(K, "K"), _p(V), _row_major_2d(V, "V"), _p(GS), A + 32, N, H, dqk, float(scale), _p(dQ),
                  _row_major_2d(dQ, "dQ"), _stream(Q))
        # the transposed pass walks the SOURCE rows (set2set: nodes, while the forward rows are graphs)
        _ffi.call("tfgk_gat_bwd_src_f32", _p(csr_t.rowptr), _p(csr_t.col), _p(Q), _row_major_2d(Q, "Q"), _p(K),
                  _row_major_2d(K, "K"), _p(V), _row_major_2d(V, "V"), _p(GS), A + 32, csr_t.n_rows, H, dqk, float(scale), _p(dK),
                  _row_major_2d(dK, "dK"), _p(dV), _row_major_2d(dV, "dV"), _stream(Q))
    except _ffi.TfgkError as err:
        if err.code != _ffi.ERR_UNSUPPORTED:
            raise
        return None
    return dQ, dK, dV


# ---- training-mode extras: dropout, per-head aggregation, GAT softmax backward -----------------------------------

RNG_STREAM_DROPOUT, RNG_STREAM_SAMPLER = 0, 1      # rng_stream ids: independent draws for the same (seed, element)


def dropout(x, rate, seed, rng_stream=RNG_STREAM_DROPOUT, out=None):
    """tf.nn.dropout with a counter-based mask: element i is kept iff u(seed, i) >= rate, kept values * 1/(1-rate)."""
    _check(x, torch.float32, "x")
    if out is None:
        out = torch.empty_like(x)
    _ffi.call("tfgk_dropout_f32", _p(x), x.numel(), float(rate), int(seed), int(rng_stream), _p(out), _stream(x))
    return out


def spmm_heads(csr, w, src, num_heads, mode=HEADS_SPLIT, emap=None, drop_rate=0.0, seed=0,
               rng_stream=RNG_STREAM_DROPOUT, alpha=1.0, bias=None, act=ACT_NONE, out=None):
    """Per-(edge, head) weighted aggregation over `csr`; see tfgk_spmm_heads_f32.  w: [E, H] (looked up through `emap`
    when the CSR is a transposed view of the structure w was computed on)."""
    _check(w, torch.float32, "w")
    if not (src.is_cuda and src.dtype == torch.float32):
        raise TypeError("src must be a float32 CUDA tensor")
    if emap is not None:
        _check(emap, torch.int32, "emap")
    H = int(num_heads)
    lds = _row_major_2d(src, "src")
    if mode == HEADS_BROADCAST:
        dh = src.shape[1]
        out_w = H * dh
    else:
        if src.shape[1] % H:
            raise ValueError("spmm_heads: {} source columns are not divisible by {} heads".format(src.shape[1], H))
        dh = src.shape[1] // H
        out_w = dh if mode == HEADS_REDUCE else H * dh
    if out is None:
        out = torch.empty((csr.n_rows, out_w), dtype=torch.float32, device=src.device)
    if bias is not None:
        _check(bias, torch.float32, "bias")
    _ffi.call("tfgk_spmm_heads_f32", _p(csr.rowptr), _p(csr.col), _p(emap), _p(w), _p(src), lds, csr.n_rows, H, dh,
              int(mode), float(drop_rate), int(seed), int(rng_stream), float(alpha), _p(bias), act, _p(out),
              _row_major_2d(out, "out"), _stream(src))
    return out


def gat_softmax_bwd(csr, att, G, V, num_heads, split_value_heads=True, drop_rate=0.0, seed=0,
                    rng_stream=RNG_STREAM_DROPOUT):
    """d loss / d scaled scores [E, H] (CSR order) from G = d loss / d aggregated rows; see tfgk_gat_softmax_bwd_f32."""
    _check(att, torch.float32, "att")
    H = int(num_heads)
    dv = V.shape[1] // H
    ds = torch.empty((csr.nnz, H), dtype=torch.float32, device=att.device)
    _ffi.call("tfgk_gat_softmax_bwd_f32", _p(csr.rowptr), _p(csr.col), _p(att), _p(G), _row_major_2d(G, "G"), _p(V),
              _row_major_2d(V, "V"), csr.n_rows, H, dv, 1 if split_value_heads else 0, float(drop_rate), int(seed),
              int(rng_stream), _p(ds), _stream(att))
    return ds


# ---- device-side edge sampling -----------------------------------------------------------------------------------

def edge_flags(row, col, num_edges, mode=FLAG_ALL, row_map=None, col_map=None, bernoulli=BERNOULLI_NONE, prob=0.0,
               seed=0, rng_stream=RNG_STREAM_SAMPLER, device=None):
    """int32 [E] keep flags: structural rule AND Bernoulli rule (tfgk_edge_flags_i32)."""
    for t, n in ((row, "row"), (col, "col"), (row_map, "row_map"), (col_map, "col_map")):
        if t is not None:
            _check(t, torch.int32, n)
    dev = device if row is None else row.device
    flag = torch.empty((num_edges,), dtype=torch.int32, device=dev)
    _ffi.call("tfgk_edge_flags_i32", _p(row), _p(col), num_edges, int(mode), _p(row_map), _p(col_map), int(bernoulli),
              float(prob), int(seed), int(rng_stream), _p(flag), _stream(flag))
    return flag


def select_flagged(flag):
    """Ascending positions of the non-zero flags (tf.boolean_mask(tf.range(n), flag)), int32 [n_selected]."""
    _check(flag, torch.int32, "flag")
    n = flag.numel()
    need = ctypes.c_size_t()
    _ffi.call("tfgk_select_workspace_bytes", n, ctypes.byref(need))
    ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=flag.device)
    out = torch.empty((max(n, 1),), dtype=torch.int32, device=flag.device)
    n_out = ctypes.c_int64()
    _ffi.call("tfgk_select_flagged_i32", _p(flag), n, _p(out), ctypes.byref(n_out), _p(ws), need.value, _stream(flag))
    return out[:n_out.value]


def gather_i32(src, index):
    """src[index] for int32 vectors: a bit copy through tfgk_permute_f32 (the kernel moves 4-byte words)."""
    _check(src, torch.int32, "src")
    return permute(src.view(torch.float32), index).view(torch.int32)


def sort_keys_f32(score, descending=False):
    """Order-preserving int32 bit patterns of float32 scores (to be sorted as unsigned numbers)."""
    _check(score, torch.float32, "score")
    keys = torch.empty((score.numel(),), dtype=torch.int32, device=score.device)
    _ffi.call("tfgk_sort_keys_f32", _p(score), score.numel(), 1 if descending else 0, _p(keys), _stream(score))
    return keys


def stable_argsort(keys, key_bits=32):
    """Stable argsort of int32 bit patterns read as unsigned numbers (LSD radix, ceil(key_bits / 8) passes)."""
    _check(keys, torch.int32, "keys")
    n = keys.numel()
    need = ctypes.c_size_t()
    _ffi.call("tfgk_argsort_workspace_bytes", n, ctypes.byref(need))
    ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=keys.device)
    perm = torch.empty((n,), dtype=torch.int32, device=keys.device)
    _ffi.call("tfgk_stable_argsort_u32", _p(keys), n, int(key_bits), _p(perm), _p(ws), need.value, _stream(keys))
    return perm


def neighbor_sample(csr, k=None, ratio=None, padding=False, seed=0, rng_stream=RNG_STREAM_SAMPLER):
    """Fan-out sampling over the rows of `csr` (tfgk_neighbor_sample_*).  Returns (row int32 [S], pos int32 [S],
    out_rowptr int64 [n_rows+1]): the row of every sampled edge and the CSR position it was drawn from.
    padding: False | True | SAMPLE_HEAD (deterministic: the first k / ceil(degree*ratio) entries of every row)."""
    dev = csr.rowptr.device
    kk = -1 if k is None else int(k)
    rr = -1.0 if ratio is None else float(ratio)
    if padding == "head" or (not isinstance(padding, bool) and padding == SAMPLE_HEAD):
        padding = SAMPLE_HEAD
    else:
        padding = SAMPLE_PADDING if padding else SAMPLE_NO_PADDING
    need = ctypes.c_size_t()
    _ffi.call("tfgk_neighbor_sample_workspace_bytes", csr.n_rows, ctypes.byref(need))
    ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=dev)
    out_rowptr = torch.empty((csr.n_rows + 1,), dtype=torch.int64, device=dev)
    total = ctypes.c_int64()
    _ffi.call("tfgk_neighbor_sample_count", _p(csr.rowptr), csr.n_rows, kk, rr, padding, _p(out_rowptr),
              ctypes.byref(total), _p(ws), need.value, _stream(out_rowptr))
    S = total.value
    out_row = torch.empty((S,), dtype=torch.int32, device=dev)
    out_pos = torch.empty((S,), dtype=torch.int32, device=dev)
    if S:
        _ffi.call("tfgk_neighbor_sample_fill", _p(csr.rowptr), csr.n_rows, kk, rr, padding, int(seed),
                  int(rng_stream), _p(out_rowptr), _p(out_row), _p(out_pos), _stream(out_rowptr))
    return out_row, out_pos, out_rowptr


# ---- K4 ----------------------------------------------------------------------------------------------------------

def gemm(a, b, bias=None, act=ACT_NONE, trans_a=False, trans_b=False, beta=0.0, out=None):
    """act(op(a) @ op(b) + bias + beta*out) in fp32."""
    for t, n in ((a, "a"), (b, "b")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2):
            raise TypeError("{} must be a 2-D float32 CUDA tensor".format(n))
    M = a.shape[1] if trans_a else a.shape[0]
    Ka = a.shape[0] if trans_a else a.shape[1]
    Kb = b.shape[1] if trans_b else b.shape[0]
    N = b.shape[0] if trans_b else b.shape[1]
    if Ka != Kb:
        raise ValueError("gemm: inner dimensions differ ({} vs {})".format(Ka, Kb))
    if out is None:
        if beta != 0.0:
            raise ValueError("gemm: beta != 0 needs `out`")
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if bias is not None:
        _check(bias, torch.float32, "bias")
    need = ctypes.c_size_t()
    _ffi.call("tfgk_gemm_workspace_bytes", M, N, Ka, ctypes.byref(need))
    ws = torch.empty((need.value,), dtype=torch.uint8, device=a.device) if need.value else None
    _ffi.call("tfgk_gemm_f32", _p(a), _row_major_2d(a, "a"), int(trans_a), _p(b), _row_major_2d(b, "b"), int(trans_b),
              _p(bias), act, float(beta), M, N, Ka, _p(out), _row_major_2d(out, "out"), _p(ws), need.value, _stream(a))
    return out


def gemm_proj(a, blocks, a_parts=None, part_rows=0, first_part=0, max_ctas=0, num_rows=None):
    """Several projections of the same rows in ONE launch (tfgk_gemm_proj_f32): `blocks` is a list of
    (weight [K, n<=128], bias or None, act code, out [M, n] view); returns the list of outputs.
    With `a_parts` (device pointers of the row blocks of A, `part_rows` rows each, e.g. the other ranks' copies of x
    mapped through peer memory) the rows are pulled from where they live; `a` then only supplies lda / K / the stream.
    Shapes the tensor-core kernel does not take fall back to one tfgk_gemm_f32 per block (single-part input only)."""
    if not (a.is_cuda and a.dtype == torch.float32 and a.dim() == 2):
        raise TypeError("a must be a 2-D float32 CUDA tensor")
    lda = _row_major_2d(a, "a")
    M = int(a.shape[0] if num_rows is None else num_rows)
    K = a.shape[1]
    structs = (_ffi.ProjBlock * len(blocks))()
    outs = []
    for i, blk in enumerate(blocks):
        w, bias, act, out = blk[:4]
        trans_b = bool(blk[4]) if len(blk) > 4 else False          # weight given as [n, K]: C = A @ w^T
        k_dim, n_dim = (1, 0) if trans_b else (0, 1)
        if not (w.is_cuda and w.dtype == torch.float32 and w.dim() == 2 and w.shape[k_dim] == K):
            raise TypeError("gemm_proj: weight {} must be a float32 CUDA tensor with inner dimension {}".format(i, K))
        n_cols = w.shape[n_dim]
        if out is None:
            out = torch.empty((M, n_cols), dtype=torch.float32, device=a.device)
        if bias is not None:
            _check(bias, torch.float32, "bias")
        structs[i] = _ffi.ProjBlock(w.data_ptr(), _row_major_2d(w, "weight"), n_cols, 1 if trans_b else 0,
                                    None if bias is None else bias.data_ptr(), int(act), out.data_ptr(),
                                    _row_major_2d(out, "out"))
        outs.append(out)
    if a_parts is None:
        parts = (ctypes.c_void_p * 1)(a.data_ptr())
        n_parts = 1
    else:
        parts = (ctypes.c_void_p * len(a_parts))(*[int(q) for q in a_parts])
        n_parts = len(a_parts)
    try:
        _ffi.call("tfgk_gemm_proj_f32", parts, n_parts, int(part_rows), lda, M, K, structs, len(blocks), int(first_part),
                  int(max_ctas), _stream(a))
    except _ffi.TfgkError as err:
        if err.code != _ffi.ERR_UNSUPPORTED or n_parts != 1:
            raise
        for blk, out in zip(blocks, outs):
            gemm(a[:M], blk[0], bias=blk[1], act=blk[2], trans_b=bool(blk[4]) if len(blk) > 4 else False, out=out)
    return outs


def colsum(x):
    """Column sums of a [N, D] float32 matrix (bias gradients), deterministic."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
        raise TypeError("colsum: x must be a 2-D float32 CUDA tensor")
    ldx = _row_major_2d(x, "x")
    n, d = x.shape
    need = ctypes.c_size_t()
    _ffi.call("tfgk_colsum_workspace_bytes", n, d, ctypes.byref(need))
    ws = torch.empty((max(need.value, 4),), dtype=torch.uint8, device=x.device)
    out = torch.empty((d,), dtype=torch.float32, device=x.device)
    _ffi.call("tfgk_colsum_f32", _p(x), ldx, n, d, _p(out), _p(ws), need.value, _stream(x))
    return out


def l2_normalize(x, out=None):
    if out is None:
        out = torch.empty_like(x)
    _ffi.call("tfgk_l2_normalize_f32", _p(x), _row_major_2d(x, "x"), x.shape[0], x.shape[1], _p(out),
              _row_major_2d(out, "out"), _stream(x))
    return out


def device_info():
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _ffi.call("tfgk_device_info", ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor))
    return {"sm_count": sm.value, "cc": (major.value, minor.value)}


def activation_code(activation):
    """Map an activation callable to a fused epilogue code; returns (code, leftover_callable)."""
    if activation is None:
        return ACT_NONE, None
    if activation in (torch.relu, torch.nn.functional.relu) or getattr(activation, "_tfgk_act", None) == "relu" \
            or isinstance(activation, torch.nn.ReLU):
        return ACT_RELU, None
    return ACT_NONE, activation


def relu(x):
    """Stand-in for tf.nn.relu in the reference's layer defaults (layers/conv/gat.py:15-16, graph_sage.py:13)."""
    return torch.relu(x)


relu._tfgk_act = "relu"

# coding=utf-8
"""ctypes binding of include/tfgk.h (libtfgk.so, sm_100a).

There is deliberately NO fallback: if the shared library is missing, or a call fails, an exception is raised.
PyTorch is only the owner of device memory and streams; every pointer handed to the library is `tensor.data_ptr()`.
"""
import ctypes
import os
import threading

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libtfgk.so")
_lock = threading.Lock()
_lib = None

ABI_VERSION = 5

OK = 0
ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_WORKSPACE, ERR_UNSUPPORTED, ERR_INDEX_OUT_OF_RANGE = 1, 2, 3, 4, 5
REDUCE_SUM, REDUCE_MEAN, REDUCE_MAX = 0, 1, 2
ACT_NONE, ACT_RELU = 0, 1
POW_INV_SQRT, POW_INV = 0, 1
HEADS_SPLIT, HEADS_BROADCAST, HEADS_REDUCE = 0, 1, 2
FLAG_ALL, FLAG_UPPER, FLAG_MAPPED = 0, 1, 2
BERNOULLI_NONE, BERNOULLI_DROPOUT, BERNOULLI_KEEP = 0, 1, 2
SAMPLE_NO_PADDING, SAMPLE_PADDING, SAMPLE_HEAD = 0, 1, 2

_i32, _i64, _f32, _int = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_int
_ptr, _size = ctypes.c_void_p, ctypes.c_size_t
_u64, _u32, _f64 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_double

# name -> argtypes, exactly as declared in include/tfgk.h (tests check every symbol is exported)
SIGNATURES = {
    "tfgk_version": [],
    "tfgk_device_info": [ctypes.POINTER(_int)] * 3,
    "tfgk_self_loops_i32": [_ptr, _i64, _i32, _ptr, _ptr],
    "tfgk_self_loop_weights_f32": [_ptr, _i64, _i32, _f32, _ptr, _ptr],
    "tfgk_segment_count_i32": [_ptr, _i64, _i32, _ptr, _ptr],
    "tfgk_csr_workspace_bytes": [_i64, _i32, ctypes.POINTER(_size)],
    "tfgk_csr_build": [_ptr, _ptr, _i64, _i32, _i32, _ptr, _ptr, _ptr, _ptr, _size, _ptr],
    "tfgk_edge_unique_workspace_bytes": [_i64, _i32, ctypes.POINTER(_size)],
    "tfgk_edge_unique": [_ptr, _ptr, _i64, _i32, _ptr, _ptr, ctypes.POINTER(_i32), _ptr, _size, _ptr],
    "tfgk_directed_workspace_bytes": [_i64, ctypes.POINTER(_size)],
    "tfgk_directed_edges": [_ptr, _i64, _i64, _ptr, _i64, _ptr, ctypes.POINTER(_i32), _ptr, _size, _ptr],
    "tfgk_plan_capacity": [_i64, _i32, _i32, _i32, _i32, ctypes.POINTER(_i64), ctypes.POINTER(_i64)],
    "tfgk_plan_workspace_bytes": [_i32, ctypes.POINTER(_size)],
    "tfgk_plan_build": [_ptr, _i32, _i32, _i32, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64,
                        ctypes.POINTER(_i32), _ptr, _size, _ptr],
    "tfgk_permute_f32": [_ptr, _ptr, _i64, _i32, _ptr, _ptr],
    "tfgk_unpermute_f32": [_ptr, _ptr, _i64, _i32, _ptr, _ptr],
    "tfgk_csr_rowsum_f32": [_ptr, _ptr, _i32, _ptr, _ptr],
    "tfgk_deg_inv_f32": [_ptr, _i32, _int, _ptr, _ptr],
    "tfgk_scale_edges_f32": [_ptr, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _ptr],
    "tfgk_spmm_f32": [_ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _int, _f32, _ptr, _i64, _f32, _ptr, _int, _ptr, _i64,
                      _ptr, _ptr],
    "tfgk_segment_softmax_f32": [_ptr, _ptr, _i32, _i32, _ptr, _ptr],
    "tfgk_gat_fused_f32": [_ptr, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _i32, _f32, _int, _ptr,
                           _int, _ptr, _int, _ptr, _i64, _ptr, _ptr],
    "tfgk_gemm_workspace_bytes": [_i32, _i32, _i32, ctypes.POINTER(_size)],
    "tfgk_gemm_f32": [_ptr, _i64, _int, _ptr, _i64, _int, _ptr, _int, _f32, _i32, _i32, _i32, _ptr, _i64, _ptr, _size,
                      _ptr],
    "tfgk_gemm_proj_f32": [_ptr, _i32, _i64, _i64, _i32, _i32, _ptr, _i32, _i32, _i32, _ptr],
    "tfgk_peer_alloc": [_size, ctypes.POINTER(_ptr)],
    "tfgk_peer_free": [_ptr],
    "tfgk_peer_export": [_ptr, _ptr],
    "tfgk_peer_open": [_ptr, ctypes.POINTER(_ptr)],
    "tfgk_peer_close": [_ptr],
    "tfgk_peer_barrier": [_ptr, _i32, _i32, _u32, _i32, _ptr],
    "tfgk_peer_pull": [_ptr, _ptr, _i64, _i32, _ptr],
    "tfgk_colsum_workspace_bytes": [_i64, _i32, ctypes.POINTER(_size)],
    "tfgk_colsum_f32": [_ptr, _i64, _i64, _i32, _ptr, _ptr, _size, _ptr],
    "tfgk_l2_normalize_f32": [_ptr, _i64, _i32, _i32, _ptr, _i64, _ptr],
    "tfgk_dropout_f32": [_ptr, _i64, _f32, _u64, _u32, _ptr, _ptr],
    "tfgk_spmm_heads_f32": [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _int, _f32, _u64, _u32, _f32, _ptr,
                            _int, _ptr, _i64, _ptr],
    "tfgk_gat_softmax_bwd_f32": [_ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _int, _f32, _u64, _u32,
                                 _ptr, _ptr],
    "tfgk_gat_fused_stats_f32": [_ptr, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _i32, _f32, _ptr, _int,
                                 _ptr, _i64, _ptr, _ptr, _ptr],
    "tfgk_gat_bwd_prepare_f32": [_ptr, _i64, _ptr, _i64, _ptr, _int, _ptr, _i32, _i32, _i32, _ptr, _i64, _ptr],
    "tfgk_gat_bwd_dst_f32": [_ptr, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _f32, _ptr, _i64,
                             _ptr],
    "tfgk_gat_bwd_src_f32": [_ptr, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _f32, _ptr, _i64,
                             _ptr, _i64, _ptr],
    "tfgk_edge_flags_i32": [_ptr, _ptr, _i64, _int, _ptr, _ptr, _int, _f32, _u64, _u32, _ptr, _ptr],
    "tfgk_select_workspace_bytes": [_i64, ctypes.POINTER(_size)],
    "tfgk_select_flagged_i32": [_ptr, _i64, _ptr, ctypes.POINTER(_i64), _ptr, _size, _ptr],
    "tfgk_sort_keys_f32": [_ptr, _i64, _int, _ptr, _ptr],
    "tfgk_argsort_workspace_bytes": [_i64, ctypes.POINTER(_size)],
    "tfgk_stable_argsort_u32": [_ptr, _i64, _int, _ptr, _ptr, _size, _ptr],
    "tfgk_neighbor_sample_workspace_bytes": [_i32, ctypes.POINTER(_size)],
    "tfgk_neighbor_sample_count": [_ptr, _i32, _i32, _f64, _int, _ptr, ctypes.POINTER(_i64), _ptr, _size, _ptr],
    "tfgk_neighbor_sample_fill": [_ptr, _i32, _i32, _f64, _int, _u64, _u32, _ptr, _ptr, _ptr, _ptr],
}


class PlanStruct(ctypes.Structure):
    """struct tfgk_plan of include/tfgk.h."""
    _fields_ = [("n_tasks", _i32), ("n_hubs", _i32), ("n_slots", _i32), ("chunk", _i32),
                ("task_row", _ptr), ("task_nrows", _ptr), ("task_e0", _ptr), ("task_e1", _ptr), ("task_slot", _ptr),
                ("hub_row", _ptr), ("hub_slot0", _ptr), ("hub_nslots", _ptr), ("scratch", _ptr), ("scratch_bytes", _size)]


class ProjBlock(ctypes.Structure):
    """struct tfgk_proj_block of include/tfgk.h."""
    _fields_ = [("B", _ptr), ("ldb", _i64), ("ncols", _i32), ("transB", _i32), ("bias", _ptr), ("act", _int), ("C", _ptr),
                ("ldc", _i64)]


PEER_HANDLE_BYTES = 64


class TfgkError(RuntimeError):
    """A tfgk_* entry point returned a non-zero status."""

    def __init__(self, fn, code, message):
        super().__init__("{} failed with status {}: {}".format(fn, code, message))
        self.code = code


def library_path():
    return _LIB_PATH


def lib():
    """Load libtfgk.so once.  Raises ImportError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_LIB_PATH):
                raise ImportError(
                    "tf_geometric_b200: {} is missing - build it with `python -c 'import __graft_entry__ as g; "
                    "g.build()'` or `make -C tf_geometric_b200/csrc`. There is no CPU/PyTorch fallback.".format(_LIB_PATH))
            handle = ctypes.CDLL(_LIB_PATH)
            for name, argtypes in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.argtypes = argtypes
                fn.restype = _int
            handle.tfgk_last_error.argtypes = []
            handle.tfgk_last_error.restype = ctypes.c_char_p
            if handle.tfgk_version() != ABI_VERSION:
                raise ImportError("libtfgk.so ABI version {} != expected {}".format(handle.tfgk_version(), ABI_VERSION))
            _lib = handle
    return _lib


class CallTrace(object):
    """Optional instrumentation used by bench.py: counts ABI calls by name and, for the names in `timed`, brackets the
    call with CUDA events on the launching stream (torch.cuda.Event on torch's current stream, which is the stream
    handed to the library)."""

    def __init__(self, timed=()):
        self.counts = {}
        self.timed = set(timed)
        self.events = {name: [] for name in self.timed}

    def elapsed_ms(self, name):
        """Per-call device durations (ms); call after synchronising."""
        return [a.elapsed_time(b) for a, b in self.events.get(name, [])]


_trace = None


def set_trace(trace):
    """Install (or remove, with None) a CallTrace; returns the previous one."""
    global _trace
    prev, _trace = _trace, trace
    return prev


def call(name, *args):
    handle = lib()
    trace = _trace
    if trace is None:
        rc = getattr(handle, name)(*args)
    else:
        trace.counts[name] = trace.counts.get(name, 0) + 1
        if name in trace.timed:
            import torch
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            rc = getattr(handle, name)(*args)
            end.record()
            trace.events[name].append((start, end))
        else:
            rc = getattr(handle, name)(*args)
    if rc != OK:
        raise TfgkError(name, rc, handle.tfgk_last_error().decode("utf-8", "replace"))
    return rc

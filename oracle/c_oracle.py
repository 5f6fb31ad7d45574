# coding=utf-8
"""ctypes loader of oracle/_build/libtfg_oracle.so (TEST INFRASTRUCTURE ONLY - see tfg_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtfg_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def aggregate(row, col, w, h, num_nodes, reduce="sum"):
    row = np.ascontiguousarray(row, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    h = np.ascontiguousarray(h, dtype=np.float32)
    w = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
    out = np.empty((num_nodes, h.shape[1]), dtype=np.float32)
    rc = lib().tfgo_aggregate_f32(_ptr(row), _ptr(col), _ptr(w), ctypes.c_int64(len(row)), _ptr(h),
                                  ctypes.c_int32(num_nodes), ctypes.c_int32(h.shape[1]),
                                  ctypes.c_int({"sum": 0, "mean": 1, "max": 2}[reduce]), _ptr(out))
    if rc:
        raise RuntimeError("tfgo_aggregate_f32 failed: {}".format(rc))
    return out


def segment_softmax(data, ids, num_segments):
    data = np.ascontiguousarray(data, dtype=np.float32)
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    H = 1 if data.ndim == 1 else data.shape[1]
    out = np.empty_like(data)
    rc = lib().tfgo_segment_softmax_f32(_ptr(data), _ptr(ids), ctypes.c_int64(len(ids)), ctypes.c_int32(H),
                                        ctypes.c_int32(num_segments), _ptr(out))
    if rc:
        raise RuntimeError("tfgo_segment_softmax_f32 failed: {}".format(rc))
    return out


def gat_core(row, col, Q, K, V, num_heads, split_value_heads=True, return_attention=False):
    row = np.ascontiguousarray(row, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    Q, K, V = (np.ascontiguousarray(a, dtype=np.float32) for a in (Q, K, V))
    N, H = Q.shape[0], num_heads
    dqk, dv = Q.shape[1] // H, V.shape[1] // H
    out = np.empty((N, V.shape[1] if split_value_heads else dv), dtype=np.float32)
    att = np.empty((len(row), H), dtype=np.float32) if return_attention else None
    rc = lib().tfgo_gat_core_f32(_ptr(row), _ptr(col), ctypes.c_int64(len(row)), _ptr(Q), _ptr(K), _ptr(V),
                                 ctypes.c_int32(N), ctypes.c_int32(H), ctypes.c_int32(dqk), ctypes.c_int32(dv),
                                 ctypes.c_int(1 if split_value_heads else 0), _ptr(att), _ptr(out))
    if rc:
        raise RuntimeError("tfgo_gat_core_f32 failed: {}".format(rc))
    return (out, att) if return_attention else out


def csr_build(row, col, num_rows):
    row = np.ascontiguousarray(row, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    rowptr = np.empty(num_rows + 1, dtype=np.int64)
    col_sorted = np.empty(len(row), dtype=np.int32)
    perm = np.empty(len(row), dtype=np.int32)
    rc = lib().tfgo_csr_build(_ptr(row), _ptr(col), ctypes.c_int64(len(row)), ctypes.c_int32(num_rows), _ptr(rowptr),
                              _ptr(col_sorted), _ptr(perm))
    if rc:
        raise RuntimeError("tfgo_csr_build failed: {}".format(rc))
    return rowptr, col_sorted, perm

# coding=utf-8
"""Peer-mapped device buffers for the partitioned path.

One process per GPU.  Every rank allocates a buffer with tfgk_peer_alloc (cudaMalloc, so it can be exported), the
64-byte CUDA IPC handles travel through torch.distributed (host side, once), and every rank maps the other ranks'
buffers.  After that the data path uses no collective at all: ranks pull the owning rank's rows over NVLink with
tfgk_peer_pull (copy engine or a copy kernel; tfgk_gemm_proj_f32 can also read them in place through `a_parts`) and
synchronise with tfgk_peer_barrier, a device-side flag barrier on the stream.  torch only sees these buffers as tensors
created over the raw pointer (no torch allocation behind them).
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _ffi


class _Raw(object):
    """__cuda_array_interface__ holder: lets torch wrap memory it did not allocate."""

    def __init__(self, ptr, nbytes, owner):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "version": 3,
                                         "data": (int(ptr), False), "strides": None}
        self.owner = owner


class PeerBuffer(object):
    """`nbytes` of device memory on every rank of `group`, each mapped into every other rank.

    .ptrs[r]   device pointer (in THIS process) of rank r's buffer
    .local     uint8 tensor over this rank's own buffer
    """

    def __init__(self, nbytes, device, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device(device)
        self.nbytes = int(nbytes)
        torch.cuda.set_device(self.device)
        ptr = ctypes.c_void_p()
        _ffi.call("tfgk_peer_alloc", self.nbytes, ctypes.byref(ptr))
        self._own = ptr.value
        handle = (ctypes.c_uint8 * _ffi.PEER_HANDLE_BYTES)()
        _ffi.call("tfgk_peer_export", ctypes.c_void_p(self._own), handle)
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.device)
        every = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=group)
        self.ptrs = []
        self._opened = []
        for r in range(self.world):
            if r == self.rank:
                self.ptrs.append(self._own)
                continue
            raw = (ctypes.c_uint8 * _ffi.PEER_HANDLE_BYTES)(*every[r].cpu().tolist())
            out = ctypes.c_void_p()
            _ffi.call("tfgk_peer_open", raw, ctypes.byref(out))
            self.ptrs.append(out.value)
            self._opened.append(out.value)
        self.local = torch.as_tensor(_Raw(self._own, self.nbytes, self), device=self.device)
        dist.barrier(group=group)                       # every rank has mapped every buffer before anyone uses them

    def view(self, offset, shape, dtype=torch.float32):
        """Tensor over [offset, offset + prod(shape)*itemsize) of the local buffer."""
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.local[offset:offset + n].view(dtype).view(*shape)

    def close(self):
        for p in self._opened:
            _ffi.call("tfgk_peer_close", ctypes.c_void_p(p))
        self._opened = []
        if self._own is not None:
            self.local = None
            _ffi.call("tfgk_peer_free", ctypes.c_void_p(self._own))
            self._own = None


class RowExchange(object):
    """Publication of this rank's rows of a feature matrix to the other ranks, without a collective.

    Two slots of [block, width] floats per rank alternate between publications.  publish(x_local) copies the rows into
    the current slot and runs the device-side barrier; after it, slot_ptrs() are the addresses (one per rank) a kernel
    later on this stream may read.  The barrier of publication g also certifies that every rank has finished reading
    publication g-1 (the reads were enqueued on the same stream before the barrier - or on a side stream the publishing
    stream has waited for, as dist.PartitionedGraph does with its pull stream), so the slot of g-2 can be overwritten by
    g without further hand-shakes.  next_slot / commit / publish must come from the same stream.
    """

    FLAG_BYTES = 256

    def __init__(self, block_rows, width, device, group=None, timeout_ms=20000):
        self.block, self.width = int(block_rows), int(width)
        self.slot_bytes = ((self.block * self.width * 4 + 255) // 256) * 256
        self.buf = PeerBuffer(self.FLAG_BYTES + 2 * self.slot_bytes, device, group)
        self.rank, self.world = self.buf.rank, self.buf.world
        self.count = 0
        self.timeout_ms = int(timeout_ms)
        self._flag_table = (ctypes.c_void_p * self.world)(*[int(p) for p in self.buf.ptrs])
        self._slots = [self.buf.view(self.FLAG_BYTES + s * self.slot_bytes, (self.block, self.width)) for s in range(2)]
        self._published = None           # (tensor id, version, data_ptr, slot)
        self.nvlink_bytes = 0            # bytes this rank pulled from peers (accounting for bench.py)

    def next_slot(self):
        """Slot the caller may fill for the next publication (it last held publication count-1, which every rank has
        finished reading: see the class comment)."""
        self.count += 1
        self._published = None
        return self.count & 1

    def commit(self, slot):
        """Device-side barrier on the current stream: after it every rank's slot of this publication is complete and may be
        read, and every rank is done with the previous publication."""
        main = torch.cuda.current_stream(self.buf.device)
        _ffi.call("tfgk_peer_barrier", self._flag_table, self.rank, self.world, self.count, self.timeout_ms,
                  ctypes.c_void_p(main.cuda_stream))

    def publish(self, x_local):
        """Make `x_local` ([n_local <= block, width], float32, CUDA) readable by every rank; returns the slot index.
        The same tensor (identity and version) is published once, however many layers ask for it."""
        key = (id(x_local), x_local._version, x_local.data_ptr())
        if self._published is not None and self._published[:3] == key and self._published[4]() is x_local:
            return self._published[3]
        import weakref
        slot = self.next_slot()
        self._slots[slot][:x_local.shape[0]].copy_(x_local)
        self.commit(slot)
        self._published = key + (slot, weakref.ref(x_local))
        return slot

    def pull(self, rank, slot, n_rows, dst, max_ctas=0):
        """Copy the first n_rows rows of `rank`'s slot into `dst` ([n_rows, width], local, contiguous) on the current stream
        (tfgk_peer_pull: wide contiguous loads over the NVLink peer mapping)."""
        if n_rows <= 0:
            return
        nbytes = int(n_rows) * self.width * 4
        src = int(self.buf.ptrs[rank]) + self.FLAG_BYTES + slot * self.slot_bytes
        _ffi.call("tfgk_peer_pull", ctypes.c_void_p(src), ctypes.c_void_p(dst.data_ptr()), nbytes, int(max_ctas),
                  ctypes.c_void_p(torch.cuda.current_stream(self.buf.device).cuda_stream))
        self.nvlink_bytes += nbytes

    def slot_ptrs(self, slot):
        off = self.FLAG_BYTES + slot * self.slot_bytes
        return [int(p) + off for p in self.buf.ptrs]

    def local_slot(self, slot):
        return self._slots[slot]

    def close(self):
        self._slots = None
        self.buf.close()

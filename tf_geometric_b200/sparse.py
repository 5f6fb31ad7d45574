# coding=utf-8
"""SparseMatrix: the subset of tf_sparse.SparseMatrix that tf_geometric's hot path calls
(nn/conv/gcn.py:56,72-119,262,280; gat.py:83-89; appnp.py:51-55,86; data/graph.py:208-210), backed by the
destination-sorted CSR + sm_100a kernels instead of tf.gather / tf.math.unsorted_segment_sum.

COO semantics are kept on the outside: `index` int32 [2, nnz] (row = aggregation target), `value` float32 [nnz]
in the caller's edge order, no sorting or merging is visible.  The CSR (stable sort by row) and the CSR-ordered
values are built lazily, once per object, and reused by every product - this is what `graph.cache` memoises.
"""
import torch

from . import ops, _rng


class SparseMatrix(object):

    def __init__(self, index, value=None, shape=None, _csr=None, _value_csr=None):
        index = ops.as_device(index, torch.int32)
        if index.dim() != 2:
            index = index.reshape(2, -1)
        self.index = index
        nnz = index.shape[1]
        if value is None:
            self.value = torch.ones((nnz,), dtype=torch.float32, device=index.device)
        else:
            self.value = ops.as_device(value, torch.float32, device=index.device)
        if shape is None:
            n = int(index.max().item()) + 1 if nnz else 0
            shape = [n, n]
        self._shape = [int(shape[0]), int(shape[1])]
        self._csr = _csr
        self._value_csr = _value_csr
        self._csc = None
        self._value_csc = None
        self._pattern_of = None           # matrix with the same pattern whose transposed CSR is shared (dropout)

    # ---- structure ----
    @property
    def shape(self):
        return self._shape

    @property
    def row(self):
        return self.index[0]

    @property
    def col(self):
        return self.index[1]

    @property
    def nnz(self):
        return self.index.shape[1]

    @property
    def csr(self):
        if self._csr is None:
            self._csr = ops.csr_build(self.index[0].contiguous(), self.index[1].contiguous(), self._shape[0],
                                      self._shape[1])
        return self._csr

    @property
    def value_csr(self):
        if self._value_csr is None:
            self._value_csr = ops.permute(self.value, self.csr.perm)
        return self._value_csr

    def _transposed_csr(self):
        if self._csc is None and self._pattern_of is not None:
            self._csc = self._pattern_of._transposed_csr()
        if self._csc is None:
            self._csc = ops.csr_build(self.index[1].contiguous(), self.index[0].contiguous(), self._shape[1],
                                      self._shape[0])
        return self._csc

    def with_value(self, value):
        """Same sparsity pattern (and cached CSR), new values in COO order."""
        return SparseMatrix(self.index, value, self._shape, _csr=self._csr)

    # ---- tf_sparse API subset ----
    def segment_sum(self, axis=-1):
        """Row sums (axis=-1/1) or column sums (axis=0/-2), sequential fp32 in edge order."""
        if axis in (-1, 1):
            return ops.csr_rowsum(self.csr, self.value_csr)
        csc = self._transposed_csr()
        return ops.csr_rowsum(csc, ops.permute(self.value, csc.perm))

    def segment_softmax(self, axis=-1):
        if axis not in (-1, 1):
            raise NotImplementedError("segment_softmax is only used with axis=-1 on this path (gat.py:84)")
        soft_csr = ops.segment_softmax_csr(self.csr, self.value_csr)
        out = SparseMatrix(self.index, ops.permute(soft_csr, self.csr.perm, inverse=True), self._shape, _csr=self._csr,
                           _value_csr=soft_csr)
        return out

    def add_diag(self, diag_value):
        """A + diag(diag_value): the diagonal is appended after the existing entries, no merge
        (same order as utils/graph_utils.py:350-366 add_self_loop_edge)."""
        n = min(self._shape)
        index = ops.self_loops(self.index, n)
        value = ops.self_loop_weights(self.value, self.nnz, n, diag_value, self.index.device)
        return SparseMatrix(index, value, self._shape)

    def dropout(self, rate, training=False, seed=None):
        """tf.nn.dropout on the stored values (gcn.py:262, appnp.py:84): the pattern is unchanged, dropped entries
        become explicit zeros and the kept ones are scaled by 1/(1-rate).  `seed` (an extension) pins the mask."""
        if not training or rate <= 0.0:
            return self
        out = SparseMatrix(self.index, ops.dropout(self.value, rate, _rng.resolve(seed)), self._shape, _csr=self._csr)
        out._pattern_of = self            # the transposed structure (backward) is built once, on the cached parent
        return out

    def matmul(self, h, num_or_size_splits=None, **epilogue):
        """A @ h (gcn.py:280).  `num_or_size_splits` is accepted for signature parity; the fused kernel never
        materialises the [E, D] temporary that the split bounds in the reference, so it is a no-op here."""
        h = ops.as_device(h, torch.float32, device=self.index.device)
        return ops.spmm(self.csr, self.value_csr, h, reduce="sum", **epilogue)

    def __matmul__(self, h):
        return self.matmul(h)

    def transpose(self):
        return SparseMatrix(torch.stack([self.index[1], self.index[0]]), self.value, [self._shape[1], self._shape[0]])

    def to_dense(self):
        out = torch.zeros(self._shape, dtype=torch.float32, device=self.index.device)
        out.index_put_((self.index[0].long(), self.index[1].long()), self.value, accumulate=True)
        return out

    def __repr__(self):
        return "SparseMatrix(shape={}, nnz={})".format(self._shape, self.nnz)

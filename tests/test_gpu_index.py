# coding=utf-8
"""GPU parity, integer/index work: BIT-EXACT against the oracle (north_star: edge_index / segment-id outputs)."""
import numpy as np
import pytest
import torch

import tf_geometric_b200 as tfg
from tf_geometric_b200 import ops, _ffi
from oracle import tfg_oracle as o
from oracle import c_oracle
from conftest import random_graph

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    return ops.as_device(a, dtype)


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("n,e", [(5, 4), (1, 3), (7, 0), (1000, 20000), (100000, 300000)])
def test_add_self_loop_edge(n, e):
    ei = random_graph(n, e, seed=n + e) if e else np.zeros((2, 0), np.int32)
    w = np.random.RandomState(0).rand(ei.shape[1]).astype(np.float32)
    want_i, want_w = o.add_self_loop_edge(ei, n, w, fill_weight=2.0)
    got_i, got_w = tfg.utils.add_self_loop_edge(dev(ei, torch.int32), n, dev(w, torch.float32), fill_weight=2.0)
    assert got_i.dtype == torch.int32 and got_w.dtype == torch.float32
    np.testing.assert_array_equal(host(got_i), want_i)
    np.testing.assert_array_equal(host(got_w), want_w)
    got_i2, got_w2 = tfg.utils.add_self_loop_edge(dev(ei, torch.int32), n)
    assert got_w2 is None
    np.testing.assert_array_equal(host(got_i2), want_i)


@pytest.mark.parametrize("n,e", [(6, 9), (1, 5), (4096, 100000), (300, 0)])
def test_segment_count(n, e):
    ids = np.random.RandomState(e).randint(0, n, e).astype(np.int32)
    got = tfg.nn.segment_count(dev(ids), n)
    assert got.dtype == torch.int32
    np.testing.assert_array_equal(host(got), np.bincount(ids, minlength=n).astype(np.int32))
    if e:
        np.testing.assert_array_equal(host(tfg.nn.segment_count(dev(ids))), o.segment_count(ids))


@pytest.mark.parametrize("n,e,kw", [
    (5, 13, {}), (1, 7, {}), (2, 1, {}), (50, 0, {}), (257, 5000, {"isolated": 9}),
    (1000, 40000, {"hub": (3, 9000)}), (70000, 200000, {}), (300000, 2500000, {"hub": (299999, 70000)}),
    (17000000, 4100, {})])
def test_csr_build_is_a_stable_sort(n, e, kw):
    """rowptr / col / perm bit-equal to a stable argsort by row (1, 2, 3 and 4 radix passes are all exercised)."""
    ei = random_graph(n, e, seed=e % 97 + 1, **kw) if e else np.zeros((2, 0), np.int32)
    csr = ops.csr_build(dev(ei[0]), dev(ei[1]), n)
    rowptr, col, perm = c_oracle.csr_build(ei[0], ei[1], n)
    np.testing.assert_array_equal(host(csr.rowptr), rowptr)
    if ei.shape[1]:
        np.testing.assert_array_equal(host(csr.perm), perm)
        np.testing.assert_array_equal(host(csr.col), col)
    assert csr.rowptr.dtype == torch.int64 and csr.col.dtype == torch.int32


def test_csr_build_rejects_out_of_range_ids():
    ei = np.array([[0, 1, 5], [1, 0, 0]], dtype=np.int32)
    with pytest.raises(_ffi.TfgkError) as err:
        ops.csr_build(dev(ei[0]), dev(ei[1]), 3)
    assert err.value.code == 5
    ei = np.array([[0, 1, 2], [1, -1, 0]], dtype=np.int32)
    with pytest.raises(_ffi.TfgkError):
        ops.csr_build(dev(ei[0]), dev(ei[1]), 3)


def test_permute_roundtrip():
    rs = np.random.RandomState(0)
    perm = rs.permutation(10007).astype(np.int32)
    for shape in ((10007,), (10007, 8)):
        src = rs.randn(*shape).astype(np.float32)
        fwd = ops.permute(dev(src), dev(perm))
        np.testing.assert_array_equal(host(fwd), src[perm])
        np.testing.assert_array_equal(host(ops.permute(fwd, dev(perm), inverse=True)), src)


@pytest.mark.parametrize("norm,loop,sym,renorm,improved", [
    ("both", True, True, True, False), ("both", True, True, False, False), ("both", True, False, True, True),
    ("both", False, False, False, False), ("left", True, False, True, False), ("right", True, False, True, False)])
def test_gcn_norm_adj_index_bit_exact_values_close(norm, loop, sym, renorm, improved):
    n = 3000
    ei = random_graph(n, 40000, seed=21, symmetric=sym, isolated=7)
    w = (np.random.RandomState(2).rand(ei.shape[1]) + 0.05).astype(np.float32)
    want = o.gcn_norm_adj(o.SparseMatrix(ei, w, [n, n]), norm, loop, sym, renorm, improved)
    got = tfg.nn.gcn_norm_adj(tfg.SparseMatrix(ei, w, [n, n]), norm, loop, sym, renorm, improved)
    np.testing.assert_array_equal(host(got.index), want.index)                 # bit-exact integers
    assert got.index.dtype == torch.int32
    np.testing.assert_allclose(host(got.value), want.value, rtol=6e-7, atol=0)  # fp32: two correctly rounded rsqrt + two products, <= 4 ulp
    # the CSR the kernels will use is the stable sort of exactly that index
    rowptr, col, perm = o.csr_build(want.index[0], want.index[1], n)
    np.testing.assert_array_equal(host(got.csr.rowptr), rowptr)
    np.testing.assert_array_equal(host(got.csr.perm), perm)
    np.testing.assert_array_equal(host(got.value_csr), host(got.value)[perm])


def test_gcn_norm_derived_kat_and_cache():
    """SURVEY.md 8c known-answer vector through the public API, plus graph.cache semantics (gcn.py:9-20,51-56,125-128)."""
    g = tfg.Graph(np.zeros((5, 2), np.float32), [[0, 0, 1, 3], [1, 2, 2, 1]], edge_weight=[.9, .8, .1, .2]).to_directed()
    assert np.asarray(g.edge_index).tolist() == [[0, 0, 1, 1, 1, 2, 2, 3], [1, 2, 2, 3, 0, 0, 1, 1]]
    g = g.to_device()
    cache = tfg.nn.gcn_build_cache_for_graph(g)
    key = "gcn_normed_adj_both_True_True_True_False"
    assert list(cache.keys()) == [key] and cache is g.cache
    normed = cache[key]
    assert host(normed.index).tolist() == [[0, 0, 1, 1, 1, 2, 2, 3, 0, 1, 2, 3, 4], [1, 2, 2, 3, 0, 0, 1, 1, 0, 1, 2, 3, 4]]
    want = [0.36927447, 0.3532086, 0.0489116, 0.1230915, 0.36927447, 0.3532086, 0.0489116, 0.1230915, 0.37037033,
            0.45454547, 0.52631575, 0.8333334, 1.0]
    np.testing.assert_allclose(host(normed.value), np.array(want, np.float32), rtol=6e-7)
    assert tfg.nn.gcn_norm_adj(g.adj(), cache=g.cache) is normed       # warm hit returns the cached object
    # a reference-style (index, value, shape) numpy triple in the cache is honoured too
    triple_cache = {key: (host(normed.index), host(normed.value), [5, 5])}
    again = tfg.nn.gcn_norm_adj(g.adj(), cache=triple_cache)
    np.testing.assert_array_equal(host(again.value), host(normed.value))


def test_to_directed_and_merge_match_oracle():
    rs = np.random.RandomState(4)
    ei = rs.randint(0, 40, (2, 500)).astype(np.int32)
    w = rs.rand(500).astype(np.float32)
    for mode in ("sum", "min", "max", "mean"):
        want_i, (want_w,) = o.convert_edge_to_directed(ei, [w], [mode])
        got_i, (got_w,) = tfg.utils.convert_edge_to_directed(ei, [w], [mode])
        np.testing.assert_array_equal(got_i, want_i)
        np.testing.assert_array_equal(got_w, want_w)
        d_i, (d_w,) = tfg.utils.convert_edge_to_directed(dev(ei), [dev(w)], [mode])
        np.testing.assert_array_equal(host(d_i), want_i)
        np.testing.assert_array_equal(host(d_w), want_w)

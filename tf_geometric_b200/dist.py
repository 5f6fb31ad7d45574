# coding=utf-8
"""Multi-GPU execution of the hot path: 1-D partition by DESTINATION row block + halo all-gather of source rows.

The reference has no graph partitioning (its two "distributed" demos replicate the whole graph under
tf.distribute.MirroredStrategy, demo/demo_distributed_gcn.py:37-57); this is the new design of SURVEY.md section 8e.

  * rank r owns destination rows [r*B, min((r+1)*B, N)), B = ceil(N / R), all their in-edges and the same rows of x;
  * dense projections are row-local (weights replicated);
  * each aggregation needs the projected rows of every SOURCE: one all-gather (torch.distributed, NCCL over
    NVLink/NVSwitch on the GPU box, gloo in the CPU tests) into a [R*B, D] buffer indexed by global node id;
  * softmax / mean / max are per destination, so nothing is reduced across ranks; the GCN normalisation exchanges only
    the [N] vector of deg^-1/2.
Per-row edge order is the caller's order, so every output row is bit-identical to the single-GPU result.
"""
import torch
import torch.distributed as dist

from . import ops
from .ops import CSR  # noqa: F401


class RowPartition(object):
    """Block partition of node ids: rank r owns [lo(r), hi(r))."""

    def __init__(self, num_nodes, world_size, rank):
        self.num_nodes = int(num_nodes)
        self.world_size = int(world_size)
        self.rank = int(rank)
        self.block = (self.num_nodes + self.world_size - 1) // self.world_size
        self.lo = min(self.rank * self.block, self.num_nodes)
        self.hi = min(self.lo + self.block, self.num_nodes)
        self.n_local = self.hi - self.lo
        self.padded_nodes = self.block * self.world_size

    def owner_of(self, node_ids):
        return node_ids // self.block


class PartitionedGraph(object):
    """The slice of a graph one rank works on: edges whose destination it owns, destinations renumbered locally,
    sources kept as GLOBAL ids (they index the all-gathered buffer)."""

    def __init__(self, partition, local_edge_index, local_edge_weight=None, group=None):
        self.part = partition
        self.edge_index = local_edge_index          # int32 [2, E_local]: row in [0, n_local), col in [0, N)
        self.edge_weight = local_edge_weight
        self.group = group
        self.cache = {}

    @classmethod
    def from_global(cls, edge_index, edge_weight, num_nodes, rank=None, world_size=None, group=None):
        """Select this rank's in-edges from a full edge list (order preserved)."""
        rank = dist.get_rank(group) if rank is None else rank
        world_size = dist.get_world_size(group) if world_size is None else world_size
        part = RowPartition(num_nodes, world_size, rank)
        edge_index = ops.as_device(edge_index, torch.int32)
        row = edge_index[0]
        mask = (row >= part.lo) & (row < part.hi)
        local = torch.stack([row[mask] - part.lo, edge_index[1][mask]]).contiguous()
        w = None
        if edge_weight is not None:
            w = ops.as_device(edge_weight, torch.float32, device=edge_index.device)[mask].contiguous()
        return cls(part, local, w, group)

    # ---- structure ------------------------------------------------------------------------------------------------
    def _with_self_loops(self, weight, fill):
        p = self.part
        dev = self.edge_index.device
        loc = torch.arange(p.n_local, dtype=torch.int32, device=dev)
        index = torch.cat([self.edge_index, torch.stack([loc, loc + p.lo])], dim=1).contiguous()
        if weight is None:
            weight = torch.ones((self.edge_index.shape[1],), dtype=torch.float32, device=dev)
        value = torch.cat([weight, torch.full((p.n_local,), fill, dtype=torch.float32, device=dev)]).contiguous()
        return index, value

    def csr(self, self_loops=False):
        key = "csr_loop" if self_loops else "csr"
        if key not in self.cache:
            index = self._with_self_loops(None, 1.0)[0] if self_loops else self.edge_index
            self.cache[key] = ops.csr_build(index[0].contiguous(), index[1].contiguous(), self.part.n_local,
                                            self.part.padded_nodes)
        return self.cache[key]

    def gcn_normed(self, renorm=True, improved=False):
        """Default GCN normalisation (norm='both', add_self_loop, sym=True; nn/conv/gcn.py:75-98) on the partition:
        row degrees are local, the column side needs deg^-1/2 of every node -> all-gather of one fp32 vector."""
        key = "gcn_normed_{}_{}".format(renorm, improved)
        if key in self.cache:
            return self.cache[key]
        p = self.part
        fill = 2.0 if improved else 1.0
        if renorm:
            index, value = self._with_self_loops(self.edge_weight, fill)
        else:
            index = self.edge_index
            value = self.edge_weight if self.edge_weight is not None else torch.ones(
                (index.shape[1],), dtype=torch.float32, device=index.device)
        csr = ops.csr_build(index[0].contiguous(), index[1].contiguous(), p.n_local, p.padded_nodes)
        deg = ops.csr_rowsum(csr, ops.permute(value, csr.perm))
        dis_local = ops.deg_inv(deg, ops.POW_INV_SQRT)
        dis_full = self.all_gather_rows(dis_local.unsqueeze(1)).squeeze(1).contiguous()
        normed = ops.scale_edges(index[0].contiguous(), index[1].contiguous(), value, dl=dis_local, dr=dis_full)
        if not renorm:
            index, normed = self._append_loops_after(index, normed, fill)
            csr = ops.csr_build(index[0].contiguous(), index[1].contiguous(), p.n_local, p.padded_nodes)
        self.cache[key] = (csr, ops.permute(normed, csr.perm))
        return self.cache[key]

    def _append_loops_after(self, index, value, fill):
        p = self.part
        loc = torch.arange(p.n_local, dtype=torch.int32, device=index.device)
        index = torch.cat([index, torch.stack([loc, loc + p.lo])], dim=1).contiguous()
        value = torch.cat([value, torch.full((p.n_local,), fill, dtype=torch.float32, device=value.device)]).contiguous()
        return index, value

    # ---- the exchange step -----------------------------------------------------------------------------------------
    def all_gather_rows(self, local_rows, out=None, async_op=False):
        """[n_local, D] on every rank -> [R*B, D] indexed by global node id (rows >= N are padding).
        async_op=True returns (buffer, work): the collective runs on the communicator's stream and `work.wait()` makes the
        current stream wait for it - used to run the exchange under independent compute."""
        p = self.part
        d = local_rows.shape[1]
        if out is None:
            out = torch.empty((p.padded_nodes, d), dtype=local_rows.dtype, device=local_rows.device)
        if p.world_size == 1:
            out[:p.n_local].copy_(local_rows)
            return (out, None) if async_op else out
        send = local_rows
        if p.n_local != p.block or not local_rows.is_contiguous():
            send = torch.zeros((p.block, d), dtype=local_rows.dtype, device=local_rows.device)
            send[:p.n_local].copy_(local_rows)
        work = dist.all_gather_into_tensor(out, send, group=self.group, async_op=async_op)
        return (out, work) if async_op else out


def gcn_partitioned(pg, x_local, kernel, bias=None, activation=None, renorm=True, improved=False):
    """tfg.nn.gcn on a PartitionedGraph: returns this rank's rows of act(norm(A) (x W) + b)."""
    dev = pg.edge_index.device
    x_local = ops.as_device(x_local, torch.float32, device=dev)
    csr, value_csr = pg.gcn_normed(renorm=renorm, improved=improved)
    h_local = x_local if kernel is None else ops.gemm(x_local, ops.as_device(kernel, torch.float32, device=dev))
    h_full = pg.all_gather_rows(h_local)
    act_code, leftover = ops.activation_code(activation)
    out = ops.spmm(csr, value_csr, h_full, reduce="sum",
                   bias=None if bias is None else ops.as_device(bias, torch.float32, device=dev), act=act_code)
    return leftover(out) if leftover is not None else out


def gat_partitioned(pg, x_local, query_kernel, query_bias, query_activation, key_kernel, key_bias, key_activation,
                    kernel, bias=None, activation=None, num_heads=1):
    """tfg.nn.gat (split_value_heads=True) on a PartitionedGraph: Q stays local, K and V travel in ONE all-gather of a
    [n_local, A + U] buffer that both projections write into directly."""
    dev = pg.edge_index.device
    x_local = ops.as_device(x_local, torch.float32, device=dev)
    wq, wk, wv = (ops.as_device(t, torch.float32, device=dev) for t in (query_kernel, key_kernel, kernel))
    q_act, q_left = ops.activation_code(query_activation)
    k_act, k_left = ops.activation_code(key_activation)
    if q_left is not None or k_left is not None:
        raise NotImplementedError("partitioned GAT supports relu / None for the query and key activations")
    a, u = wq.shape[1], wv.shape[1]
    n_local = x_local.shape[0]
    Q = ops.gemm(x_local, wq, bias=ops.as_device(query_bias, torch.float32, device=dev), act=q_act)
    kv_local = torch.empty((n_local, a + u), dtype=torch.float32, device=dev)
    ops.gemm(x_local, wk, bias=ops.as_device(key_bias, torch.float32, device=dev), act=k_act, out=kv_local[:, :a])
    ops.gemm(x_local, wv, out=kv_local[:, a:])
    kv_full = pg.all_gather_rows(kv_local)
    act_code, leftover = ops.activation_code(activation)
    out = ops.gat_fused(pg.csr(self_loops=True), Q, kv_full[:, :a], kv_full[:, a:], num_heads,
                        bias=None if bias is None else ops.as_device(bias, torch.float32, device=dev), act=act_code)
    return leftover(out) if leftover is not None else out


def gcn_gat_overlapped(pg, x_local, gcn_kernel, gcn_bias, gcn_activation,
                       query_kernel, query_bias, key_kernel, key_bias, kernel, bias, gat_activation, num_heads):
    """One GCN layer and one GAT layer on the same partitioned graph with the two halo exchanges issued asynchronously:
    all projections run first, the all-gather of h (GCN) and of K|V (GAT) are queued back to back on the communicator
    stream, and the GCN aggregation runs while K|V is still travelling.  Results are identical to calling
    gcn_partitioned / gat_partitioned one after the other (relu query/key activations)."""
    dev = pg.edge_index.device
    x_local = ops.as_device(x_local, torch.float32, device=dev)
    f32 = lambda t: None if t is None else ops.as_device(t, torch.float32, device=dev)   # noqa: E731
    csr, value_csr = pg.gcn_normed()
    h_local = ops.gemm(x_local, f32(gcn_kernel))
    h_full, work_h = pg.all_gather_rows(h_local, async_op=True)
    wq, wk, wv = f32(query_kernel), f32(key_kernel), f32(kernel)
    a, u = wq.shape[1], wv.shape[1]
    Q = ops.gemm(x_local, wq, bias=f32(query_bias), act=ops.ACT_RELU)
    kv_local = torch.empty((x_local.shape[0], a + u), dtype=torch.float32, device=dev)
    ops.gemm(x_local, wk, bias=f32(key_bias), act=ops.ACT_RELU, out=kv_local[:, :a])
    ops.gemm(x_local, wv, out=kv_local[:, a:])
    kv_full, work_kv = pg.all_gather_rows(kv_local, async_op=True)
    if work_h is not None:
        work_h.wait()
    act_gcn, left_gcn = ops.activation_code(gcn_activation)
    out_gcn = ops.spmm(csr, value_csr, h_full, reduce="sum", bias=f32(gcn_bias), act=act_gcn)
    if left_gcn is not None:
        out_gcn = left_gcn(out_gcn)
    if work_kv is not None:
        work_kv.wait()
    act_gat, left_gat = ops.activation_code(gat_activation)
    out_gat = ops.gat_fused(pg.csr(self_loops=True), Q, kv_full[:, :a], kv_full[:, a:], num_heads, bias=f32(bias),
                            act=act_gat)
    if left_gat is not None:
        out_gat = left_gat(out_gat)
    return out_gcn, out_gat


# ---- bench.py --gpus N ------------------------------------------------------------------------------------------------

def bench_partitioned(args, rank, world, device, metric, config):
    """Strong scaling of the bench workload: the same synthetic graph, destination-partitioned over `world` ranks.
    Timed on the device with CUDA events between barriers; the reported time is the max over ranks."""
    import json
    import os
    import numpy as np
    import bench as B
    from . import _ffi

    n = int(B.PRODUCTS_NODES * args.scale)
    pairs = int(B.PRODUCTS_UNDIRECTED * args.scale)
    edge_index = B.make_graph_device(n, pairs, 0, device)      # same seed on every rank -> identical global graph
    E = edge_index.shape[1]
    pg = PartitionedGraph.from_global(edge_index, None, n, rank, world)
    del edge_index
    torch.cuda.empty_cache()
    p = pg.part
    gen = torch.Generator(device="cpu")
    gen.manual_seed(100 + rank)
    x_host = torch.randn((p.n_local, B.FEATURES), generator=gen, dtype=torch.float32).pin_memory()
    x = x_host.to(device)
    wk = B.glorot((B.FEATURES, B.UNITS), 2).to(device)
    wq_, wk_, wv_ = (B.glorot((B.FEATURES, B.UNITS), s).to(device) for s in (3, 4, 5))
    zero = torch.zeros((B.UNITS,), dtype=torch.float32, device=device)
    relu = ops.relu

    def step(xd):
        if os.environ.get("TFGK_DIST_OVERLAP", "1") == "0":
            a = gcn_partitioned(pg, xd, wk, zero, relu)
            b = gat_partitioned(pg, xd, wq_, zero, relu, wk_, zero, relu, wv_, zero, relu, num_heads=B.HEADS)
            return a, b
        return gcn_gat_overlapped(pg, xd, wk, zero, relu, wq_, zero, wk_, zero, wv_, zero, relu, B.HEADS)

    for _ in range(max(args.warmup, 3)):
        step(x)
    trace = _ffi.CallTrace(timed=("tfgk_gat_fused_f32", "tfgk_spmm_f32"))
    _ffi.set_trace(trace)
    sampler = B.ClockSampler(device.index)
    sampler.start()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.steps):
        step(x)
    ev[1].record()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    _ffi.set_trace(None)
    t = torch.tensor([ev[0].elapsed_time(ev[1]) / args.steps], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())

    e2e = None
    if not args.no_e2e:
        local = B.run_e2e(args, device, x_host, step, p.n_local, E, barrier=dist.barrier)
        t2 = torch.tensor([local["ms_per_step"]], dtype=torch.float64, device=device)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = dict(local, value=2.0 * E / (float(t2.item()) * 1e-3), ms_per_step=float(t2.item()),
                   h2d_bytes_per_step=n * B.FEATURES * 4, d2h_bytes_per_step=2 * n * B.UNITS * 4)

    gat_ms = float(np.mean(trace.elapsed_ms("tfgk_gat_fused_f32")))
    e_local = pg.csr(self_loops=True).nnz
    gat_bytes = e_local * (8 * B.UNITS + 4) + p.n_local * (8 * B.UNITS + 8)
    peak, peak_src = B.measured_peak_gbs()
    launching = ("tfgk_gat_fused_f32", "tfgk_spmm_f32", "tfgk_gemm_f32")
    launches = sum(trace.counts.get(k, 0) for k in launching)
    if rank == 0:
        halo = (world - 1) * p.block * (B.UNITS + 2 * B.UNITS) * 4
        line = {"metric": metric, "value": 2.0 * E / (ms_step * 1e-3), "unit": "edges/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": "gat_async_kernel<2,3> (tfgk_gat_fused_f32), rank 0 partition",
                             "achieved": gat_bytes / (gat_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": gat_bytes / (gat_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes": gat_bytes, "kernel_ms": gat_ms},
                "cpu_baseline": None,
                "exchange": {"collective": "all_gather_into_tensor (NCCL), async: K|V exchange overlaps the GCN aggregation", "halo_bytes_in_per_rank_per_step": halo}}
        B.emit(line)
    dist.destroy_process_group()

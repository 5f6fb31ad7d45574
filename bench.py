#!/usr/bin/env python
# coding=utf-8
"""bench.py - the headline measurement (BASELINE.json metric):

    edges/sec of GCN + 8-head GAT forward on a synthetic ogbn-products-shaped graph
    (2,449,029 nodes, 123,718,280 directed edges, 100-dim fp32 features), plus the HBM-roofline fraction of the
    dominant kernel and the reference CPU path timed on the same box.

A "step" = one tfg.layers.GCN(128, relu) forward followed by one tfg.layers.GAT(128, num_heads=8, relu) forward over
the whole graph through the public layer API, warm graph.cache (normalised adjacency + destination-sorted CSR built
once, outside the timed region - the regime of the reference's own harness, demo/demo_gcn.py:47,99-105).
edges/sec = (2 * E) / step time: every layer pass streams all E input edges (appended self loops are NOT counted).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]

--impl reference times the reference's op sequence on the host CPU cores (oracle/torch_cpu_port.py; TensorFlow and
tf_sparse cannot be installed offline) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PRODUCTS_NODES = 2449029
PRODUCTS_UNDIRECTED = 61859140
FEATURES = 100
UNITS = 128
HEADS = 8
METRIC = "edges/sec GCN+GAT fwd on 2.4M-node/123M-edge synthetic; %HBM roofline"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (testing only; 1.0 = BASELINE size)")
    ap.add_argument("--cpu-sample-div", type=int, default=0,
                    help="reference arm: graph scaled down by this factor (0 = auto: about two minutes of CPU work in total)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---- synthetic workload --------------------------------------------------------------------------------------------

def make_graph_device(num_nodes, num_pairs, seed, device):
    """Uniform random undirected pairs u != v, mirrored (SURVEY.md 8d cfg 4 generator), int32 [2, 2*pairs]."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    u = torch.randint(0, num_nodes, (num_pairs,), generator=gen, device=device, dtype=torch.int32)
    v = torch.randint(0, num_nodes - 1, (num_pairs,), generator=gen, device=device, dtype=torch.int32)
    v = v + (v >= u).to(torch.int32)                      # v != u, still uniform
    edge_index = torch.empty((2, 2 * num_pairs), dtype=torch.int32, device=device)
    edge_index[0, :num_pairs] = u
    edge_index[0, num_pairs:] = v
    edge_index[1, :num_pairs] = v
    edge_index[1, num_pairs:] = u
    return edge_index


def glorot(shape, seed):
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    limit = (6.0 / (shape[0] + shape[1])) ** 0.5
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * limit


class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML in a background thread every
    ~2 ms (nvidia-smi -lms cannot resolve a 50 ms region); falls back to one nvidia-smi query if NVML is unavailable."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def _loop(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.samples.append((sm, reasons, power))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        n = self.nvml
        smax = None
        try:
            smax = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        except Exception:
            pass
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        seen = set()
        for _, r, _ in self.samples:
            for k, bit in names.items():
                if r & bit:
                    seen.add(k)
        sm = [x[0] for x in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max([x[2] for x in self.samples]) if self.samples else None,
                "samples": len(sm), "reasons": sorted(seen)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ---- reference arm (CPU) ---------------------------------------------------------------------------------------------

def cpu_reference(num_nodes, num_pairs, steps, warmup, seed=0):
    """The reference's op sequence on the host cores for one step (GCN fwd + GAT fwd) on a bounded graph."""
    from oracle import torch_cpu_port as port
    from oracle import tfg_oracle as o
    torch.set_num_threads(os.cpu_count() or 1)
    ei = make_graph_device(num_nodes, num_pairs, seed, torch.device("cpu"))
    E = ei.shape[1]
    gen = torch.Generator(device="cpu"); gen.manual_seed(1)
    x = torch.randn((num_nodes, FEATURES), generator=gen, dtype=torch.float32)
    # warm cache, like demo_gcn.py:47: normalised adjacency precomputed (numpy oracle), self loops appended for GAT
    normed = o.gcn_norm_adj(o.SparseMatrix(ei.numpy(), None, [num_nodes, num_nodes]))
    n_row = torch.from_numpy(normed.index[0]).long(); n_col = torch.from_numpy(normed.index[1]).long()
    n_val = torch.from_numpy(normed.value)
    wk = glorot((FEATURES, UNITS), 2); b = torch.zeros(UNITS)
    wq_, wk_, wv_ = glorot((FEATURES, UNITS), 3), glorot((FEATURES, UNITS), 4), glorot((FEATURES, UNITS), 5)

    def step():
        port.gcn_forward(x, n_row, n_col, n_val, wk, b)
        port.gat_forward(x, n_row, n_col, wq_, b, wk_, b, wv_, b, HEADS)   # same index: edges + appended self loops

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"edges_per_s": 2.0 * E / dt, "ms_per_step": dt * 1e3, "edges": E, "nodes": num_nodes,
            "cores": torch.get_num_threads()}


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_sample_div(args, passes):
    """Bounded sample for the CPU arm: the op-for-op port sustains ~0.6 M edges/s on the box's host cores (measured), so
    `passes` step executions of the full graph would take hours; scale the graph so the whole arm takes ~2 minutes."""
    if args.cpu_sample_div > 0:
        return args.cpu_sample_div
    budget_edges = 0.6e6 * 120.0                       # edge-layer passes affordable in ~120 s
    per_step_full = 4.0 * PRODUCTS_UNDIRECTED * args.scale
    return max(25, int(np.ceil(per_step_full * passes / budget_edges)))


def run_reference(args, rank, world):
    if rank != 0:
        return
    args.cpu_sample_div = cpu_sample_div(args, args.steps + args.warmup)
    n = max(int(PRODUCTS_NODES * args.scale) // args.cpu_sample_div, 1000)
    pairs = max(int(PRODUCTS_UNDIRECTED * args.scale) // args.cpu_sample_div, 1000)
    res = cpu_reference(n, pairs, args.steps, args.warmup)
    sample = ("same generator and layer shapes at 1/{} scale: {} nodes, {} directed edges; op-for-op torch-CPU port of "
              "the reference op sequence, {}").format(args.cpu_sample_div, res["nodes"], res["edges"], cpu_model_name())
    line = {"metric": METRIC, "value": res["edges_per_s"], "unit": "edges/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": res["edges_per_s"], "unit": "edges/s", "cores": res["cores"], "kind": "port",
                             "sample": sample},
            "e2e": {"value": res["edges_per_s"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(args, world):
    n = int(PRODUCTS_NODES * args.scale)
    e = 2 * int(PRODUCTS_UNDIRECTED * args.scale)
    return {"workload": "GCN(128,relu) fwd + GAT(128, 8 heads, relu) fwd, synthetic ogbn-products shape "
                        "({} nodes, {} directed edges, {} features), uniform random undirected pairs mirrored, "
                        "warm graph.cache".format(n, e, FEATURES),
            "nodes": n, "edges": e, "features": FEATURES, "units": UNITS, "heads": HEADS,
            "edges_per_step": 2 * e, "parallelism": "single GPU" if world == 1 else "dst-partitioned x{}".format(world),
            "l2_policy": "working set (>= 2 GB of gathered rows + CSR) exceeds the 126 MB L2; no explicit flush"}


# ---- our arm -----------------------------------------------------------------------------------------------------------

def run_e2e(args, device, x_host, step_fn, n_out_rows, edges_per_layer, barrier=None):
    """Host-to-host throughput of the same step: every step copies its input features from pinned host memory and lands
    both layer outputs in pinned host memory.  The three engines are pipelined the way a serving loop would do it:
    H2D of step i+1 and D2H of step i's outputs run on their own streams while step i / i+1 compute; device and host
    buffers are double buffered and every dependency is an event.  All copies are inside the timed region."""
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    main = torch.cuda.current_stream(device)
    outs = [[torch.empty((n_out_rows, UNITS), dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(2)]
    x_dev = [torch.empty(x_host.shape, dtype=torch.float32, device=device) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]       # x_dev[slot] no longer read by compute
    ev_out_done = [torch.cuda.Event() for _ in range(2)]   # host output slot drained (previous use)

    def submit(i):
        slot = i & 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[slot])
            x_dev[slot].copy_(x_host, non_blocking=True)
            ev_in[slot].record(s_in)
        main.wait_event(ev_in[slot])
        a, b = step_fn(x_dev[slot])                          # public layer API, GCN then GAT
        ev_free[slot].record(main)
        ev_done = torch.cuda.Event()
        ev_done.record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done)
            s_out.wait_event(ev_out_done[slot])
            outs[slot][0].copy_(a, non_blocking=True)
            outs[slot][1].copy_(b, non_blocking=True)
            ev_out_done[slot].record(s_out)
        a.record_stream(s_out)
        b.record_stream(s_out)

    for slot in range(2):
        ev_free[slot].record(main)
        ev_out_done[slot].record(s_out)
    for i in range(2):
        submit(i)
    torch.cuda.synchronize(device)
    if barrier is not None:
        barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(main)
    for i in range(args.steps):
        submit(i)
    main.wait_stream(s_out)                                  # the last outputs have landed on the host
    main.wait_stream(s_in)
    t1.record(main)
    torch.cuda.synchronize(device)
    ms = t0.elapsed_time(t1) / args.steps
    return {"value": 2.0 * edges_per_layer / (ms * 1e-3), "unit": "edges/s", "ms_per_step": ms,
            "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 2 * n_out_rows * UNITS * 4,
            "pipelining": "H2D / compute / D2H on three streams, double buffered, all inside the timed region"}


def run_ours(args, rank, world, local_rank):
    import tf_geometric_b200 as tfg
    from tf_geometric_b200 import _ffi

    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")     # keep NCCL's version banner off stdout: one JSON line only
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    if world > 1:
        from tf_geometric_b200.dist import bench_partitioned
        return bench_partitioned(args, rank, world, device, METRIC, workload_config(args, world))

    n = int(PRODUCTS_NODES * args.scale)
    pairs = int(PRODUCTS_UNDIRECTED * args.scale)
    edge_index = make_graph_device(n, pairs, 0, device)
    E = edge_index.shape[1]
    gen = torch.Generator(device="cpu"); gen.manual_seed(1)
    x_host = torch.randn((n, FEATURES), generator=gen, dtype=torch.float32).pin_memory()
    x = x_host.to(device)
    graph = tfg.Graph(x, edge_index)

    gcn = tfg.layers.GCN(UNITS, activation=tfg.nn.relu, seed=2)
    gat = tfg.layers.GAT(UNITS, num_heads=HEADS, activation=tfg.nn.relu, seed=3)
    gcn.build_cache_for_graph(graph)                       # normalised adjacency + CSR (one-off, untimed)
    torch.cuda.synchronize()
    t_cache = time.perf_counter()
    gat([graph.x, graph.edge_index], cache=graph.cache)    # builds the self-looped CSR + weights
    gcn([graph.x, graph.edge_index, graph.edge_weight], cache=graph.cache)
    torch.cuda.synchronize()
    t_cache = time.perf_counter() - t_cache

    def step(xd):
        a = gcn([xd, graph.edge_index, graph.edge_weight], cache=graph.cache)
        b = gat([xd, graph.edge_index], cache=graph.cache)
        return a, b

    # ---- device-resident timing ("value") ----
    for _ in range(max(args.warmup, 3)):
        step(x)
    torch.cuda.synchronize()
    trace = _ffi.CallTrace(timed=("tfgk_gat_fused_f32", "tfgk_spmm_f32", "tfgk_gemm_f32"))
    _ffi.set_trace(trace)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(args.steps):
        step(x)
    ev[1].record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    _ffi.set_trace(None)
    ms_step = ev[0].elapsed_time(ev[1]) / args.steps
    value = 2.0 * E / (ms_step * 1e-3)

    gat_ms = float(np.mean(trace.elapsed_ms("tfgk_gat_fused_f32")))
    spmm_ms = float(np.mean(trace.elapsed_ms("tfgk_spmm_f32")))
    gemm_ms = float(np.sum(trace.elapsed_ms("tfgk_gemm_f32"))) / args.steps
    launching = ("tfgk_gat_fused_f32", "tfgk_spmm_f32", "tfgk_gemm_f32")   # each is exactly one kernel launch here
    launches = sum(trace.counts.get(k, 0) for k in launching)

    # roofline of the dominant kernel (K3 fused GAT): algorithmic bytes per launch, DESIGN.md "K3"
    e_loop = E + n
    gat_bytes = e_loop * (4 * UNITS + 4 * UNITS + 4) + n * (4 * UNITS + 4 * UNITS + 8)
    spmm_bytes = e_loop * (4 * UNITS + 4 + 4) + n * (4 * UNITS + 8)
    peak, peak_src = measured_peak_gbs()
    achieved = gat_bytes / (gat_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "gat_async_kernel<2,3> (tfgk_gat_fused_f32)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
                # (profiles/r1_ncu_full_final_kernels.json); only valid for the default full-size workload
                "traffic": 128184537000 if args.scale == 1.0 else None, "peak_source": peak_src,
                "algorithmic_bytes": gat_bytes, "kernel_ms": gat_ms,
                "secondary": {"kernel": "spmm_async_kernel<1,0,4,3> (tfgk_spmm_f32)", "traffic": 63515602000 if args.scale == 1.0 else None, "algorithmic_bytes": spmm_bytes,
                              "kernel_ms": spmm_ms, "achieved": spmm_bytes / (spmm_ms * 1e-3) / 1e9,
                              "frac": spmm_bytes / (spmm_ms * 1e-3) / 1e9 / peak},
                "gemm_ms_per_step": gemm_ms}

    # ---- end to end: host buffers in, host buffers out, through the same public API ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, device, x_host, lambda xd: step(xd), n, E)

    cpu_base = None
    if not args.no_cpu_baseline:
        args.cpu_sample_div = cpu_sample_div(args, 2)
        ns = max(n // args.cpu_sample_div, 1000)
        ps = max(pairs // args.cpu_sample_div, 1000)
        res = cpu_reference(ns, ps, steps=1, warmup=1)
        cpu_base = {"value": res["edges_per_s"], "unit": "edges/s", "cores": res["cores"], "kind": "port",
                    "sample": "same generator and layer shapes at 1/{} scale ({} nodes, {} directed edges), 1 step after "
                              "1 warm-up, op-for-op torch-CPU port of the reference op sequence, {}".format(
                                  args.cpu_sample_div, res["nodes"], res["edges"], cpu_model_name())}

    line = {"metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, 1),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "launches_per_step": launches / args.steps,
            "roofline": roofline, "cpu_baseline": cpu_base,
            "breakdown_ms": {"gcn_spmm": spmm_ms, "gat_fused": gat_ms, "dense_projections": gemm_ms,
                             "cache_build_s": t_cache}}
    emit(line)


class StdoutToStderr(object):
    """Routes file descriptor 1 to stderr while the benchmark runs (NCCL / library banners must not pollute the ONE JSON
    line the driver parses) and restores it for the final print."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def _result_lines():
    # bench.py runs as __main__ and is also imported as `bench` by tf_geometric_b200.dist: share the list through `sys`
    if not hasattr(sys, "_tfgk_bench_lines"):
        sys._tfgk_bench_lines = []
    return sys._tfgk_bench_lines


def emit(line):
    """Collect the JSON line; main() prints it once stdout is restored."""
    _result_lines().append(json.dumps(line))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    with StdoutToStderr():
        if args.impl == "reference":
            run_reference(args, rank, world)
        else:
            if not torch.cuda.is_available():
                raise SystemExit("bench.py needs a CUDA device for --impl ours (there is no CPU fallback); "
                                 "use --impl reference for the CPU arm")
            run_ours(args, rank, world, local_rank)
    for line in _result_lines():
        print(line, flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- edges/sec of GCNConv forward+backward on a synthetic power-law graph.

Metric (BASELINE.json): "edges/sec (GCNConv fwd+bwd) at 1/2/4/8 B200; achieved HBM GB/s vs peak".
A *step* is one pass of the hot path: GCNConv(F, F) forward + backward (out.backward(grad)) on the
whole graph, graph structure cached (GCNConv(cached=True) semantics in both arms).
edges/sec = E_input * layers / t (E_input counted before self-loop insertion).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N=1 workload: N=10M nodes, E=100M edges, F=256, fp32 (the north-star headline shape; inputs of
10 GB >> 126 MB L2, so no explicit L2 flush is needed).  N>1: weak scaling -- every rank owns a
contiguous node range of the same size with the same number of incoming edges, sources outside the
range are halo rows exchanged with one all_to_all per aggregation pass (pytorch_geometric_b200/dist.py).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# --------------------------------------------------------------------------- helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.file = None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=self.file,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        for line in self.file.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.file.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def synth_graph(num_nodes: int, num_edges: int, seed: int, device, lo: int = 0, total_nodes=None,
                p_local: float = 1.0):
    """Seeded synthetic power-law graph: destination in-degrees follow a truncated
    power law P(deg = k) ~ k^-2.1 (inverse CDF of the equivalent rank-frequency law, hub ids scattered
    by a random permutation), sources uniform.  Duplicates and self loops are left in.  For sharded runs destinations fall in
    [lo, lo + num_nodes) and a fraction p_local of the sources too; the rest is uniform over all
    `total_nodes` (the halo)."""
    g = torch.Generator(device=device).manual_seed(seed)
    u = torch.rand(num_edges, device=device, generator=g, dtype=torch.float64)
    gamma = 2.1                       # in-degree distribution P(deg = k) ~ k^-gamma
    s = 1.0 / (gamma - 1.0)           # <=> rank-frequency law f(rank) ~ rank^-s (s = 0.909)
    # inverse CDF of the continuous rank law on [1, num_nodes]: F(r) = (r^(1-s) - 1) / (n^(1-s) - 1)
    rank = (1.0 + u * (float(num_nodes) ** (1.0 - s) - 1.0)) ** (1.0 / (1.0 - s))
    rank = (rank.long() - 1).clamp_(0, num_nodes - 1)
    dst = torch.randperm(num_nodes, device=device, generator=g)[rank] + lo   # scatter hubs over the id range
    total = total_nodes if total_nodes is not None else num_nodes
    src_local = torch.randint(0, num_nodes, (num_edges, ), device=device, generator=g) + lo
    if p_local >= 1.0 or total == num_nodes:
        src = src_local
    else:
        src_any = torch.randint(0, total, (num_edges, ), device=device, generator=g)
        pick = torch.rand(num_edges, device=device, generator=g) < p_local
        src = torch.where(pick, src_local, src_any)
    return torch.stack([src, dst])


def traffic_bytes(args):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), valid for the
    default workload only."""
    if args.traffic_bytes is not None:
        return args.traffic_bytes
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path) and (args.nodes, args.edges, args.feat) == (10_000_000, 100_000_000, 256):
        with open(path) as f:
            return json.load(f)["spmm_csr_bytes_per_launch"]
    return None


def pass_bytes(E_prime: int, N: int, F: int, s: int = 4, b_idx: int = 4, b_w: int = 4) -> int:
    """Algorithmic bytes of one CSR aggregation pass:
    every edge reads one feature row + its column index + its weight, every row is written once,
    rowptr is read once.  No cache-reuse credit."""
    return E_prime * (F * s + b_idx + b_w) + N * F * s + (N + 1) * b_idx


# --------------------------------------------------------------------------- the CPU reference arm
def run_reference(args):
    """`--impl reference`: the reference's own CPU path (ATen call sequence of GCNConv on a [2,E]
    tensor, oracle/ref_aten.py) timed on the host cores on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_aten
    torch.manual_seed(0)
    N, E, F = args.cpu_nodes, args.cpu_edges, args.feat
    ei = synth_graph(N, E, 1, "cpu")
    x = torch.randn(N, F, requires_grad=True)
    weight = torch.nn.Parameter(torch.randn(F, F) / F ** 0.5)
    bias = torch.nn.Parameter(torch.zeros(F))
    ei2, w2 = ref_aten.gcn_norm(ei, None, N)                          # cached=True: outside the loop
    gout = torch.randn(N, F)

    def step():
        x.grad = weight.grad = bias.grad = None
        out = ref_aten.gcn_conv_forward(x, ei2, w2, weight, bias)
        out.backward(gout)

    for _ in range(max(args.warmup, 1) if args.warmup < 2 else 2):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = E / dt
    cores = torch.get_num_threads()
    sample = f"GCNConv({F},{F}) fwd+bwd, N={N}, E={E} power-law, fp32, COO gather->mul->scatter_add_ (reference default path), {steps} steps"
    emit({
        "impl": "reference", "metric": "edges/sec (GCNConv fwd+bwd)", "value": value, "unit": "edges/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": sample, "host_cpus": os.cpu_count(), "torch_threads": cores},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def cpu_baseline_quick(args):
    """Bounded CPU sample for the `cpu_baseline` object of the main line (rank 0, N=1 only)."""
    from oracle import ref_aten
    N, E, F = args.cpu_nodes, args.cpu_edges, args.feat
    ei = synth_graph(N, E, 1, "cpu")
    x = torch.randn(N, F, requires_grad=True)
    weight = torch.nn.Parameter(torch.randn(F, F) / F ** 0.5)
    bias = torch.nn.Parameter(torch.zeros(F))
    ei2, w2 = ref_aten.gcn_norm(ei, None, N)
    gout = torch.randn(N, F)

    def step():
        x.grad = weight.grad = bias.grad = None
        ref_aten.gcn_conv_forward(x, ei2, w2, weight, bias).backward(gout)

    step()
    t0 = time.perf_counter()
    n = 0
    while n < 3 and (n == 0 or time.perf_counter() - t0 < 20):
        step()
        n += 1
    dt = (time.perf_counter() - t0) / n
    cores = torch.get_num_threads()
    return {"value": E / dt, "unit": "edges/s", "cores": cores, "kind": "port",
            "sample": f"oracle/ref_aten.py GCNConv({F},{F}) fwd+bwd on N={N}, E={E} (same generator), {n} steps, "
                      f"{cores} torch threads of {os.cpu_count()} host CPUs"}


# --------------------------------------------------------------------------- the B200 arm
def run_b200(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import pytorch_geometric_b200 as pgb
    from pytorch_geometric_b200 import ops
    from pytorch_geometric_b200 import utils as U
    from pytorch_geometric_b200.nn import GCNConv

    N, E, F = args.nodes, args.edges, args.feat
    ops.set_option("spmm_impl", args.spmm_impl)
    ops.set_option("spmm_tune", args.spmm_tune)
    from pytorch_geometric_b200 import dense
    dense.set_backend(args.dense)
    if args.gemm_bk:
        ops.set_option("gemm_bk", args.gemm_bk)
    if args.gemm_mode >= 0:
        ops.set_option("gemm_mode", args.gemm_mode)
    if args.gemm_prefetch >= 0:
        ops.set_option("gemm_prefetch", args.gemm_prefetch)
    if args.gemm_bsplit >= 0:
        dense.set_b_split(bool(args.gemm_bsplit))
    torch.manual_seed(1234 + rank)
    conv = GCNConv(F, F, cached=True).to(dev)
    with torch.no_grad():
        conv.bias.normal_(0, 0.1)
    if world == 1:
        ei = synth_graph(N, E, 1, dev)
        graph = U.gcn_norm_graph(ei, None, N)
        del ei
        graph.build_transpose()
        halo = None
        E_prime = graph.num_edges
        fwd = lambda xx: conv(xx, graph)                                     # noqa: E731
    else:
        ei = synth_graph(N, E, 1 + rank, dev, lo=rank * N, total_nodes=world * N, p_local=args.p_local)
        shard = None
        if args.dist == "p2p":
            # halo exchange fused into the gather kernel over NVLink peer memory (dist_p2p.py)
            try:
                from pytorch_geometric_b200 import dist_p2p
                shard = dist_p2p.PeerShardedGCNGraph.build(ei, rank * N, N, world * N, F, dist.group.WORLD)
                fwd = lambda xx: dist_p2p.peer_sharded_gcn_conv(conv, xx, shard)    # noqa: E731
            except Exception as exc:                                  # symmetric memory unavailable on this box
                print(f"[bench] symmetric-memory path unavailable ({exc!r}); using the NCCL all_to_all path", file=sys.stderr)
                args.dist = "nccl"
                shard = None
        if shard is None:
            from pytorch_geometric_b200 import dist as pdist
            shard = pdist.ShardedGCNGraph.build(ei, rank * N, N, world * N, dist.group.WORLD)
            fwd = lambda xx: pdist.sharded_gcn_conv(conv, xx, shard)             # noqa: E731
        del ei
        E_prime = shard.num_edges
        graph = shard.graph
    torch.cuda.synchronize()

    x = torch.randn(N, F, device=dev).requires_grad_()
    if world > 1 and args.dist == "p2p":
        gout = shard.gout.normal_()                                   # upstream gradient produced in the symmetric buffer
    else:
        gout = torch.randn(N, F, device=dev)

    def step():
        x.grad = None
        conv.lin.weight.grad = None
        conv.bias.grad = None
        out = fwd(x)
        out.backward(gout)
        if world > 1:                       # data-parallel weight gradients, as DDP would do
            dist.all_reduce(conv.lin.weight.grad)
            dist.all_reduce(conv.bias.grad)
        return out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.PROFILE.reset(enabled=True)
    launches0 = ops.LAUNCHES.count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = ops.LAUNCHES.count - launches0
    kern = ops.PROFILE.summary()                                              # per-kernel CUDA-event times
    ops.PROFILE.reset(enabled=False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * E / (ms_per_step * 1e-3)

    # ---- end to end through the public API with HOST buffers (N=1 path; per rank for N>1)
    e2e = None
    if not args.no_e2e:
        x_host = torch.empty(N, F, dtype=torch.float32, pin_memory=True)
        x_host.normal_()
        # Input prefetch, as a training loop with a pinned-memory loader does it: two device buffers,
        # the H2D copy of step i+1 runs on a copy stream while step i computes.  Every step still
        # pays its own full H2D copy and its own D2H read inside the timed region.
        x_bufs = [torch.empty(N, F, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream(dev)
        gw_host = torch.empty(F, F, dtype=torch.float32, pin_memory=True)
        gb_host = torch.empty(F, dtype=torch.float32, pin_memory=True)
        loss_host = torch.empty(1, dtype=torch.float32, pin_memory=True)
        for ev in freed:
            ev.record(main_stream)

        def prefetch(i):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[b])                              # buffer no longer read by step i-2
                x_bufs[b].copy_(x_host, non_blocking=True)                    # H2D of step i's input
                ready[b].record(copy_stream)

        def e2e_step(i):
            b = i & 1
            prefetch(i + 1)
            main_stream.wait_event(ready[b])
            xin = x_bufs[b].detach().requires_grad_()
            conv.lin.weight.grad = None
            conv.bias.grad = None
            out = fwd(xin)
            loss = (out * gout).sum()                                         # the step's scalar result
            out.backward(gout)
            if world > 1:
                dist.all_reduce(conv.lin.weight.grad)
                dist.all_reduce(conv.bias.grad)
            gw_host.copy_(conv.lin.weight.grad, non_blocking=True)            # D2H of the step's results
            gb_host.copy_(conv.bias.grad, non_blocking=True)
            loss_host.copy_(loss.detach().view(1), non_blocking=True)
            freed[b].record(main_stream)

        prefetch(0)
        for i in range(2):
            e2e_step(i)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(3, min(args.steps, 10))
        e0.record()
        for i in range(2, 2 + n_e2e):
            e2e_step(i)
        e1.record()
        sync_all()
        ems = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ems], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": world * E / (ems / n_e2e * 1e-3), "unit": "edges/s",
               "h2d_bytes_per_step": world * N * F * 4, "d2h_bytes_per_step": world * (F * F + F + 1) * 4,
               "ms_per_step": ems / n_e2e, "steps": n_e2e,
               "what": "pinned-host x -> H2D (prefetched one step ahead on a copy stream, double-buffered) -> "
                       "GCNConv fwd+bwd -> D2H of grad_W, grad_b and the loss scalar, every step; "
                       "bound by the 10.24 GB/step host link"}
        del x_host, x_bufs

    if rank == 0:
        peak, peak_src = measured_peaks()
        bytes_pass = pass_bytes(E_prime, N, F)
        agg = kern.get("spmm_csr", {"ms_total": 0.0, "calls": 0})
        avg_ms = agg["ms_total"] / max(agg["calls"], 1)
        achieved = bytes_pass / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": "csr_reduce_kernel (b200mp_spmm_csr), fwd on CSR + bwd on transposed CSR",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_pass,
                    "avg_launch_ms": avg_ms, "launches_timed": agg["calls"],
                    "share_of_step": agg["ms_total"] / ms if ms > 0 else None,
                    "traffic": traffic_bytes(args), "frac_of_nominal_8TBs": achieved / 8000.0}
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline_quick(args)
        line = {
            "metric": "edges/sec (GCNConv fwd+bwd)", "value": value, "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"GCNConv({F},{F}) fwd+bwd, power-law synthetic graph, N={N} nodes and E={E} edges per GPU"
                                   f" (E'={E_prime} with self loops), fp32, int32 CSR, graph cached (cached=True)",
                       "nodes_per_gpu": N, "edges_per_gpu": E, "feat": F, "layers": 1,
                       "l2_policy": "inputs (x, grad, CSR > 10 GB) are far larger than the 126 MB L2; no explicit flush",
                       "parallelism": "single GPU" if world == 1 else (
                           f"node-range sharding x{world}, p_local={args.p_local}, " +
                           ("halo rows gathered over NVLink peer memory inside the kernel (symmetric memory)"
                            if args.dist == "p2p" else "halo all_to_all (NCCL) overlapped with the local sweep")),
                       "gemm": ("hand-written tcgen05 3xTF32, A operand in TMEM (fp32-accurate, csrc/gemm_tf32x3.cu + gemm_tf32x3_ts.cuh)" if args.dense == "tf32x3"
                                else "torch.nn.functional.linear (cuBLAS fp32, allow_tf32=False)"),
                       "long_rows": graph.plan.n_long, "chunks": graph.plan.n_chunks,
                       "spmm_impl": {0: "default (register-staged lane-group kernel, csrc/csr_reduce.cuh)", 1: "lane-group kernel",
                                     2: "persistent TMA-fed variant (csrc/csr_tma.cuh)"}[args.spmm_impl]},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "kernels": kern, "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """Exactly ONE JSON line on the real stdout (libraries such as NCCL print banners to fd 1,
    so fd 1 is pointed at stderr for the whole run and the result goes to the saved descriptor)."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nodes", type=int, default=10_000_000, help="nodes per GPU")
    ap.add_argument("--edges", type=int, default=100_000_000, help="edges per GPU")
    ap.add_argument("--feat", type=int, default=256)
    ap.add_argument("--dist", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: halo rows gathered over NVLink peer memory inside the kernel, or NCCL all_to_all")
    ap.add_argument("--p-local", type=float, default=0.95, help="fraction of sources inside the owner's range (N>1)")
    ap.add_argument("--cpu-nodes", type=int, default=250_000)
    ap.add_argument("--cpu-edges", type=int, default=2_500_000)
    ap.add_argument("--spmm-impl", type=int, default=0, help="0 auto, 1 lane-group kernel, 2 TMA kernel")
    ap.add_argument("--spmm-tune", type=int, default=0, help="tuning variant of the lane-group kernel (csr_reduce.cuh)")
    ap.add_argument("--dense", default="tf32x3", choices=["tf32x3", "cublas"],
                    help="dense transform: hand-written tcgen05 3xTF32 GEMM (fp32-accurate) or strict-fp32 cuBLAS")
    ap.add_argument("--gemm-bk", type=int, default=0, help="k-block width of the tcgen05 GEMM (16 or 32; 0 = library default)")
    ap.add_argument("--gemm-prefetch", type=int, default=-1, help="TMA L2-prefetch distance of the GEMM in k-blocks")
    ap.add_argument("--gemm-mode", type=int, default=-1, help="0 = SS-mode GEMM, 1 = TS-mode (A in TMEM); -1 = library default")
    ap.add_argument("--gemm-bsplit", type=int, default=-1,
                    help="1 = the GEMM kernel splits W tiles itself (one L2 read of W per tile), 0 = pre-split W_hi / W_lo")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_b200(args)


if __name__ == "__main__":
    main()

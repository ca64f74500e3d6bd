"""world_size-2 gloo tests (CPU) of the node-range sharding host logic (pytorch_geometric_b200/dist.py):
halo plan, forward exchange, relabelling and the backward return leg, checked against the
single-process oracle on the unsharded graph."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_graph(n_total, n_edges, world):
    rng = np.random.default_rng(42)
    n_local = n_total // world
    dst = rng.integers(0, n_total, size=n_edges)
    # 70% of the sources inside the destination's shard, the rest anywhere (the halo)
    local = (dst // n_local) * n_local + rng.integers(0, n_local, size=n_edges)
    anywhere = rng.integers(0, n_total, size=n_edges)
    src = np.where(rng.random(n_edges) < 0.7, local, anywhere)
    src[:5] = dst[:5]                                   # a few self loops
    x = rng.standard_normal((n_total, 6)).astype(np.float32)
    gout = rng.standard_normal((n_total, 6)).astype(np.float32)
    return src.astype(np.int64), dst.astype(np.int64), x, gout


def _worker(rank, world, port, n_total, n_edges, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from pytorch_geometric_b200 import dist as pd
        src, dst, x, gout = _global_graph(n_total, n_edges, world)
        n_local = n_total // world
        lo = rank * n_local
        mine = (dst >= lo) & (dst < lo + n_local)        # owner-computes: I hold the in-edges of my nodes
        s, d = pd.shard_self_loops(torch.from_numpy(src[mine]), torch.from_numpy(dst[mine]), lo, n_local)
        plan, s_rel = pd.build_halo_plan(s, lo, n_local)
        # --- plan invariants
        assert plan.n_halo == len(np.unique(s.numpy()[(s.numpy() < lo) | (s.numpy() >= lo + n_local)]))
        assert sum(plan.recv_counts) == plan.n_halo and sum(plan.send_counts) == plan.n_send
        assert plan.recv_counts[rank] == 0 and plan.send_counts[rank] == 0
        # --- forward exchange returns exactly the remote rows, in halo_ids order
        x_local = torch.from_numpy(x[lo:lo + n_local])
        halo = pd.exchange_halo(plan, x_local)
        assert np.array_equal(halo.numpy(), x[plan.halo_ids.numpy()])
        # --- relabelled aggregation over [local | halo] == my rows of the unsharded aggregation
        # (unsharded reference: add_remaining_self_loops + unit weights, sum)
        r2, c2, _ = O.add_remaining_self_loops(src, dst, None, n_total)
        ref = O.gather_scatter(x, r2, c2, None, n_total, "sum")
        xcat = np.concatenate([x_local.numpy(), halo.numpy()])
        out = O.gather_scatter(xcat, s_rel.numpy(), (d - lo).numpy(), None, n_local, "sum")
        np.testing.assert_allclose(out, ref[lo:lo + n_local], rtol=1e-5, atol=1e-5)
        # --- backward: A^T g on [local | halo] sources, halo part returned to the owners
        g_local_rows = gout[lo:lo + n_local]
        g_cat = O.gather_scatter(g_local_rows, (d - lo).numpy(), s_rel.numpy(), None, n_local + plan.n_halo, "sum")
        g_local = torch.from_numpy(g_cat[:n_local].copy())
        pd.return_halo(plan, torch.from_numpy(g_cat[n_local:].copy()), g_local)
        ref_t = O.gather_scatter(gout, c2, r2, None, n_total, "sum")
        np.testing.assert_allclose(g_local.numpy(), ref_t[lo:lo + n_local], rtol=1e-5, atol=1e-5)
        q.put((rank, "ok"))
    except Exception as exc:  # surface the failure in the parent
        import traceback
        q.put((rank, "FAIL: " + "".join(traceback.format_exception(exc))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_halo_plan_and_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 400, 6000, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_single_rank_plan_has_no_halo():
    """world == 1: no collective is issued and nothing is relabelled to the halo segment."""
    sys.path.insert(0, ROOT)
    port = _free_port()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        from pytorch_geometric_b200 import dist as pd
        src = torch.tensor([0, 3, 2, 2, 1])
        plan, rel = pd.build_halo_plan(src, 0, 4)
        assert plan.n_halo == 0 and plan.n_send == 0 and torch.equal(rel, src)
    finally:
        dist.destroy_process_group()


def _sampled_worker(rank, world, port, q):
    """bench.py's sampled-row checker (oracle/sampled.py) on 2 gloo ranks: every rank holds the in-edges of its node
    range; rows / degrees of remote nodes travel through the checker's own small collectives.  Fed with the
    single-process oracle's results on the unsharded graph it must report a tiny error -- and catch a corrupted shard."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from oracle import sampled
        n_total, n_edges, Fi, Fo = 4000, 50000, 16, 24
        rng = np.random.default_rng(7)
        n_local = n_total // world
        dst = ((rng.random(n_edges) ** 3) * (n_total - 1)).astype(np.int64)
        dst[: n_total // 2] = rng.integers(0, n_total, size=n_total // 2)
        src = rng.integers(0, n_total, size=n_edges)
        x = rng.standard_normal((n_total, Fi)).astype(np.float32)
        W = (rng.standard_normal((Fo, Fi)) / 4).astype(np.float32)
        b = rng.standard_normal(Fo).astype(np.float32)
        gout = rng.standard_normal((n_total, Fo)).astype(np.float32)
        out = O.gcn_conv(x, src, dst, None, W, b)
        gx, gw, gb = O.gcn_conv_backward(gout, x, src, dst, None, W)
        lo = rank * n_local
        mine = (dst >= lo) & (dst < lo + n_local)
        t = torch.from_numpy
        ei = t(np.stack([src[mine], dst[mine]]))
        sl = slice(lo, lo + n_local)
        args = (ei, lo, n_local, t(x[sl]), t(W), t(b), t(gout[sl]))
        r = sampled.gcn_check(*args, t(out[sl]), t(gx[sl]), t(gw), t(gb), n_rows=256, max_edges=20000, seed=1)
        assert r["ok"] and r["max_rel"] < 2e-6, r
        bad = out[sl].copy()
        if rank == 1:
            bad += 1e-3
        r2 = sampled.gcn_check(*args, t(bad), t(gx[sl]), t(gw), t(gb), n_rows=256, max_edges=20000, seed=1)
        assert not r2["ok"], r2                     # the max over ranks is what every rank reports
        q.put((rank, "ok"))
    except Exception as exc:
        import traceback
        q.put((rank, "FAIL: " + "".join(traceback.format_exception(exc))))
    finally:
        dist.destroy_process_group()


def test_sampled_row_checker_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sampled_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"

"""Sampled-row parity check of a GCNConv forward+backward at FULL benchmark size against the CPU oracle.

TEST / BENCH INFRASTRUCTURE ONLY (like everything under oracle/): used by bench.py's `parity_check` object and by
tests/; never imported by pytorch_geometric_b200/.

The C oracle is serial, so at N = 10 M / E = 100 M / F = 256 it cannot recompute a whole layer in minutes.  What it
can do is recompute a seeded SAMPLE of rows exactly as `oracle.gcn_conv` / `gcn_conv_backward` would
(nn/conv/gcn_conv.py:95-113,241-274 restated in oracle/mp_oracle.c), restricted to those rows:

  forward   out[S]    = sum_{e: dst_e in S} dinv[src_e] dinv[dst_e] (x W^T)[src_e] + b     S = sampled destinations
  backward  gx[T]     = ( sum_{e: src_e in T} dinv[src_e] dinv[dst_e] gout[dst_e] ) W      T = sampled sources
            grad_b    = sum_i gout[i]
            u^T grad_W v = sum_e w_e (gout[dst_e].u)(x[src_e].v)   for random probes u, v (grad_W is a reduction over
                        ALL nodes, so it is checked through bilinear probes evaluated edge by edge in fp64)

Everything the check needs from the GPU -- degrees, the in/out-edges of the sampled rows, feature rows -- is taken
from the RAW [2, E] edge list with plain ATen ops (bincount, boolean masks, index_select), never from the engine's
CSR structures or kernels, so a wrong sort / plan / weight in the engine cannot hide.  The arithmetic of the sampled
rows runs in the C oracle (fp32, edge order) and numpy fp64 for the dense products.

Sharded runs (world > 1): ranks own contiguous node ranges [lo, lo + n); `ei` holds the in-edges of the owned
destinations with GLOBAL source ids.  Rows and degrees of remote nodes are fetched with small torch.distributed
collectives (independent of the engine's peer-memory path), and the per-rank partial sums of the backward rows are
all-reduced in fp64.

Tolerance: |got - want| <= tol * sum|terms| elementwise (the bar DESIGN.md states: 1e-5 relative to the sum of the
absolute values of the summed terms; order of summation differs from the serial oracle).
"""
from __future__ import annotations

import numpy as np
import torch

from . import oracle as O


def _pick(deg: torch.Tensor, n_rows: int, max_edges: int, seed: int) -> torch.Tensor:
    """Seeded sample of row ids: hub rows (> 512 edges: the chunked path), the heaviest row that fits, rows with
    1 and 2 edges, rows with 0 edges (empty), the rest uniform -- capped at `max_edges` edges in total."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    deg_c = deg.cpu()
    n = deg_c.numel()
    chosen = []

    def take(mask, k):
        ids = torch.nonzero(mask).view(-1)
        if ids.numel() == 0 or k <= 0:
            return
        sel = ids[torch.randperm(ids.numel(), generator=g)[:k]]
        chosen.append(sel)

    budget = max_edges
    hubs = torch.nonzero(deg_c > 512).view(-1)
    if hubs.numel():
        order = hubs[torch.argsort(deg_c[hubs], descending=True)]
        top = [int(i) for i in order[:64] if int(deg_c[i]) <= budget // 2][:1]          # heaviest hub that fits
        budget -= sum(int(deg_c[i]) for i in top)
        rest = order[torch.randperm(order.numel(), generator=g)]
        acc, keep = 0, []
        for i in rest.tolist():
            if i in top:
                continue
            d = int(deg_c[i])
            if acc + d > budget // 2 or len(keep) >= 48:
                continue
            keep.append(i)
            acc += d
        chosen.append(torch.tensor(top + keep, dtype=torch.long))
    take(deg_c == 0, n_rows // 8)
    take(deg_c == 1, n_rows // 4)
    take(deg_c == 2, n_rows // 8)
    have = sum(c.numel() for c in chosen)
    take((deg_c > 2) & (deg_c <= 512), max(n_rows - have, n_rows // 4))
    ids = torch.unique(torch.cat(chosen)) if chosen else torch.zeros(0, dtype=torch.long)
    # enforce the edge budget (drop the largest non-hub rows first if needed)
    tot = int(deg_c[ids].sum())
    if tot > max_edges:
        order = torch.argsort(deg_c[ids])
        csum = torch.cumsum(deg_c[ids][order], 0)
        ids = ids[order[csum <= max_edges]]
    return torch.sort(ids)[0]


class _Dist:
    """Small-collective helpers; no-ops in a single process."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.group = group
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0

    def fetch_rows(self, ids_global: torch.Tensor, n_local: int, *tables: torch.Tensor):
        """rows `ids_global` of row-sharded tables (every rank owns rows [rank*n_local, (rank+1)*n_local)).
        Collective: every rank calls it with its own request list."""
        if not self.on:
            return [t.index_select(0, ids_global) for t in tables]
        dist, dev = self.dist, ids_global.device
        owner = torch.div(ids_global, n_local, rounding_mode="floor")
        order = torch.argsort(owner, stable=True)
        req = ids_global[order]
        counts = torch.bincount(owner, minlength=self.world)
        rc = torch.empty_like(counts)
        dist.all_to_all_single(rc, counts, group=self.group)
        send_counts, recv_counts = counts.tolist(), rc.tolist()
        asked = torch.empty(int(sum(recv_counts)), dtype=ids_global.dtype, device=dev)
        dist.all_to_all_single(asked, req, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=self.group)
        local = asked - self.rank * n_local
        outs = []
        for t in tables:
            ans = t.index_select(0, local).contiguous()
            got = torch.empty((req.numel(), ) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            dist.all_to_all_single(got, ans, output_split_sizes=send_counts, input_split_sizes=recv_counts, group=self.group)
            inv = torch.empty_like(order)
            inv[order] = torch.arange(order.numel(), device=dev)
            outs.append(got.index_select(0, inv))
        return outs

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        if not self.on:
            return t.unsqueeze(0)
        parts = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t.contiguous(), group=self.group)
        return torch.stack(parts)

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.on:
            self.dist.all_reduce(t, group=self.group)
        return t


def _rel(got: np.ndarray, want: np.ndarray, scale: np.ndarray) -> float:
    """max over elements of |got - want| / sum|terms| (elements with no terms must match exactly -> 0 or inf)."""
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    s = np.maximum(scale.astype(np.float64), 1e-30)
    r = err / s
    r[(scale <= 0) & (err == 0)] = 0.0
    return float(r.max()) if r.size else 0.0


@torch.no_grad()
def gcn_check(ei: torch.Tensor, lo: int, n_local: int, x: torch.Tensor, weight: torch.Tensor, bias, gout: torch.Tensor,
              out: torch.Tensor, gx: torch.Tensor, gw: torch.Tensor, gb, *, n_rows: int = 4096,
              max_edges: int = 3_000_000, seed: int = 0, tol: float = 1e-5, group=None, n_probes: int = 4) -> dict:
    """Sampled-row oracle check of one GCNConv(add_self_loops=True, normalize=True, unweighted) fwd+bwd.

    ei [2,E] raw edge list (GLOBAL ids, dst in [lo, lo+n_local)); x, gout, out, gx: this rank's rows; gw, gb: the
    (all-reduced) weight / bias gradients.  Returns {"rows", "edges", "max_rel", "ok", per-quantity max_rel}."""
    D = _Dist(group)
    dev = ei.device
    src, dst = ei[0], ei[1]
    nonloop = src != dst
    dst_l = dst - lo
    # ---- degrees as gcn_norm defines them (in-degree over non-loop edges + the one inserted loop), ATen only
    deg_in = torch.bincount(dst_l[nonloop], minlength=n_local) + 1
    dinv = deg_in.to(torch.float32).pow(-0.5)                      # fp32 like the reference (deg.pow_(-0.5))
    W64 = weight.detach().double().cpu().numpy()
    res = {}

    # ================= forward rows S
    S = _pick(deg_in - 1, n_rows, max_edges, seed * 7919 + 11 + D.rank).to(dev)       # sampled on the RAW in-degree
    sel = torch.zeros(n_local, dtype=torch.bool, device=dev)
    sel[S] = True
    m = sel[dst_l] & nonloop
    e_src = torch.cat([src[m], S + lo])                                              # reference order: non-loops, then loops
    e_dst = torch.cat([dst_l[m], S])
    U, e_src_rel = torch.unique(e_src, return_inverse=True)
    xU, dinvU = D.fetch_rows(U, n_local, x.detach(), dinv)
    pos = torch.full((n_local, ), -1, dtype=torch.long, device=dev)
    pos[S] = torch.arange(S.numel(), device=dev)
    e_dst_rel = pos[e_dst]
    w_e = (dinvU[e_src_rel] * 1.0 * dinv[e_dst]).cpu().numpy()                       # dinv[row] * w * dinv[col], w = 1
    xwU = (xU.double().cpu().numpy() @ W64.T)                                        # dense transform of the fetched rows, fp64
    xwU32 = xwU.astype(np.float32)
    ref = O.gather_scatter(xwU32, e_src_rel.cpu().numpy(), e_dst_rel.cpu().numpy(), w_e, S.numel(), "sum")
    scale = O.gather_scatter(np.abs(xwU32), e_src_rel.cpu().numpy(), e_dst_rel.cpu().numpy(), np.abs(w_e), S.numel(), "sum")
    # the GEMM's own rounding is relative to sum_k |x_k W_k|: fold it into the scale
    absxw = (np.abs(xU.double().cpu().numpy()) @ np.abs(W64).T).astype(np.float32)
    scale = scale + O.gather_scatter(absxw, e_src_rel.cpu().numpy(), e_dst_rel.cpu().numpy(), np.abs(w_e), S.numel(), "sum")
    if bias is not None:
        b = bias.detach().float().cpu().numpy()
        ref = ref + b
        scale = scale + np.abs(b)
    res["out"] = _rel(out.detach().index_select(0, S).float().cpu().numpy(), ref, scale)
    n_edges = int(e_src.numel())

    # ================= backward rows T (sampled sources of every rank; partial sums from every rank's edges)
    deg_out_partial = torch.bincount(src[nonloop], minlength=D.world * n_local)      # out-edges landing in MY destinations
    deg_out = D.all_reduce_sum(deg_out_partial.clone())[lo:lo + n_local]
    T = _pick(deg_out, n_rows // 2, max_edges // 2, seed * 104729 + 5 + D.rank).to(dev) + lo   # global ids
    n_t = torch.tensor([T.numel()], device=dev)
    n_t_max = int(D.all_gather(n_t).max())
    T_pad = torch.full((n_t_max, ), -1, dtype=torch.long, device=dev)
    T_pad[:T.numel()] = T
    T_all = D.all_gather(T_pad)                                                      # [world, n_t_max]
    dinvT_pad = torch.zeros(n_t_max, dtype=torch.float32, device=dev)
    dinvT_pad[:T.numel()] = dinv[T - lo]
    dinvT_all = D.all_gather(dinvT_pad)
    F_out = gout.size(1)
    partial = torch.zeros((D.world, n_t_max, F_out), dtype=torch.float64, device=dev)
    pscale = torch.zeros_like(partial)
    gout32 = gout.detach().float()
    tpos = torch.full((D.world * n_local, ), -1, dtype=torch.long, device=dev)
    for r in range(D.world):
        Tr = T_all[r]
        valid = Tr >= 0
        tpos.fill_(-1)
        tpos[Tr[valid]] = torch.arange(int(valid.sum()), device=dev)
        mm = (tpos[src] >= 0) & nonloop
        es, ed = tpos[src[mm]], dst_l[mm]
        if r == D.rank:                                                              # the inserted self loops of my own rows
            es = torch.cat([es, tpos[Tr[valid]]])
            ed = torch.cat([ed, Tr[valid] - lo])
        if es.numel() == 0:
            continue
        DU, ed_rel = torch.unique(ed, return_inverse=True)
        w = (dinvT_all[r][es] * 1.0 * dinv[ed]).cpu().numpy()
        gU = gout32.index_select(0, DU).cpu().numpy()
        p = O.gather_scatter(gU, ed_rel.cpu().numpy(), es.cpu().numpy(), w, int(valid.sum()), "sum")
        ps = O.gather_scatter(np.abs(gU), ed_rel.cpu().numpy(), es.cpu().numpy(), np.abs(w), int(valid.sum()), "sum")
        partial[r, :p.shape[0]] = torch.from_numpy(p).double().to(dev)
        pscale[r, :p.shape[0]] = torch.from_numpy(ps).double().to(dev)
        n_edges += int(es.numel())
    D.all_reduce_sum(partial)
    D.all_reduce_sum(pscale)
    g_xw_T = partial[D.rank, :T.numel()].cpu().numpy()                               # (A^T gout)[T], fp64
    ref_gx = g_xw_T @ W64
    scale_gx = pscale[D.rank, :T.numel()].cpu().numpy() @ np.abs(W64)
    res["grad_x"] = _rel(gx.detach().index_select(0, T - lo).float().cpu().numpy(), ref_gx, scale_gx)

    # ================= grad_b and grad_W (reductions over all nodes): fp64 ATen sums and bilinear probes
    if gb is not None:
        s = D.all_reduce_sum(gout.detach().double().sum(0))
        sa = D.all_reduce_sum(gout.detach().double().abs().sum(0))
        res["grad_b"] = _rel(gb.detach().cpu().numpy(), s.cpu().numpy(), sa.cpu().numpy())
    gen = torch.Generator(device="cpu").manual_seed(seed + 99)
    worst = 0.0
    dinv_src = D.all_gather(dinv).view(-1) if D.on else dinv                         # dinv of every node (4 B each)
    for _ in range(n_probes):
        u = torch.randn(F_out, generator=gen, dtype=torch.float64).to(dev)
        v = torch.randn(x.size(1), generator=gen, dtype=torch.float64).to(dev)
        gu = gout.detach().double() @ u                                              # [n_local]
        xv_local = x.detach().double() @ v
        xv = D.all_gather(xv_local).view(-1) if D.on else xv_local                  # [world * n_local]
        w64 = dinv_src.double()[src[nonloop]] * dinv.double()[dst_l[nonloop]]
        terms = w64 * gu[dst_l[nonloop]] * xv[src[nonloop]]
        loops = dinv.double() ** 2 * gu * xv_local
        rhs = D.all_reduce_sum((terms.sum() + loops.sum()).view(1))
        rhs_abs = D.all_reduce_sum((terms.abs().sum() + loops.abs().sum()).view(1))
        lhs = u @ gw.detach().double() @ v
        worst = max(worst, float((lhs - rhs).abs() / rhs_abs.clamp(min=1e-30)))
    res["grad_W_probe"] = worst

    mx = max(res.values())
    mx_all = mx
    if D.on:
        t = torch.tensor([mx], device=dev, dtype=torch.float64)
        D.dist.all_reduce(t, op=D.dist.ReduceOp.MAX, group=group)
        mx_all = float(t.item())
    rows = int(S.numel() + T.numel())
    return {"rows": rows * D.world if D.on else rows, "dst_rows": int(S.numel()), "src_rows": int(T.numel()),
            "edges_recomputed": n_edges, "max_rel": mx_all, "tol": tol, "ok": bool(mx_all <= tol),
            "per_quantity": res,
            "how": "oracle.gather_scatter (C, fp32, edge order) on the in-/out-edges of a seeded row sample taken from the "
                   "raw edge list with ATen ops; dense products in numpy fp64; grad_W by bilinear probes over all edges; "
                   "error relative to sum|terms|"}

"""Node-range sharding of the aggregation path over the GPUs of one box.

One process per GPU.  Rank r owns the contiguous node range [r*n, (r+1)*n): its feature rows, its
output rows, and the CSR slice of its DESTINATIONS (the reference's closest notion is
`EdgeIndex.sparse_narrow`, edge_index.py:1028-1133, which nothing in the reference calls).  Sources
outside the range are *halo* rows.  Per aggregation pass there is exactly ONE exchange:

    forward :  pack the rows my peers need (gather) -> all_to_all_single (NCCL over NVLink)
               -> one gather-reduce over [local rows | halo rows] (two-segment source, no concat)
    backward:  transposed gather-reduce -> the halo part of the result goes back with the mirror
               all_to_all_single and is added into the owners' rows (index_add).

Everything else (degrees, gcn_norm, softmax, max, mean) is per destination, hence local.  The plan
(unique remote ids grouped by owner, send lists, relabelled columns) is integer set-up work done
once per graph with device-agnostic torch ops, so the same code is exercised on CPU with the gloo
backend by tests/test_dist_gloo.py.  The reference has no counterpart of this module; correctness
is shard-vs-unsharded equality against the single-process oracle.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


@dataclass
class HaloPlan:
    rank: int
    world: int
    lo: int                 # first owned node id
    n_local: int            # nodes per rank (uniform ranges)
    halo_ids: Tensor        # [n_halo] global ids of remote sources, sorted => grouped by owner
    recv_counts: List[int]  # rows received from each rank
    send_index: Tensor      # [n_send] LOCAL row ids to send, grouped by requesting rank
    send_counts: List[int]  # rows sent to each rank

    @property
    def n_halo(self) -> int:
        return int(self.halo_ids.numel())

    @property
    def n_send(self) -> int:
        return int(self.send_index.numel())


def _all_to_all_rows(out: Tensor, inp: Tensor, out_counts: List[int], in_counts: List[int], group) -> Tensor:
    if dist.get_world_size(group) == 1:
        return out
    dist.all_to_all_single(out, inp, output_split_sizes=out_counts, input_split_sizes=in_counts, group=group)
    return out


def build_halo_plan(src_global: Tensor, lo: int, n_local: int, group=None):
    """Returns (plan, src_relabelled): local sources -> [0, n_local), remote -> n_local + halo slot."""
    group = group if group is not None else dist.group.WORLD
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = src_global.device
    is_remote = (src_global < lo) | (src_global >= lo + n_local)
    halo_ids = torch.unique(src_global[is_remote])            # sorted => grouped by owner rank
    owner = torch.div(halo_ids, n_local, rounding_mode="floor")
    recv_counts = torch.bincount(owner, minlength=world).tolist()
    # tell every owner how many (then which) of its rows I need
    rc = torch.tensor(recv_counts, dtype=torch.int64, device=dev)
    sc = torch.empty_like(rc)
    if world > 1:
        dist.all_to_all_single(sc, rc, group=group)
    else:
        sc.copy_(rc)
    send_counts = sc.tolist()
    want = torch.empty(int(sum(send_counts)), dtype=halo_ids.dtype, device=dev)
    _all_to_all_rows(want, halo_ids, send_counts, recv_counts, group)
    send_index = want - lo                                     # local row ids, grouped by requester
    if want.numel():
        assert int(send_index.min()) >= 0 and int(send_index.max()) < n_local, "halo request outside the owner's range"
    # relabel
    slot = torch.searchsorted(halo_ids, src_global.clamp(min=0)) if halo_ids.numel() else torch.zeros_like(src_global)
    src_rel = torch.where(is_remote, slot + n_local, src_global - lo)
    plan = HaloPlan(rank, world, lo, n_local, halo_ids, recv_counts, send_index, send_counts)
    return plan, src_rel


def _pack(x_local: Tensor, index: Tensor) -> Tensor:
    if x_local.is_cuda:
        from . import ops
        return ops.gather_rows(x_local, index)
    return x_local.index_select(0, index)


def exchange_halo(plan: HaloPlan, x_local: Tensor, group=None) -> Tensor:
    """Forward exchange: returns the halo rows [n_halo, F] in halo_ids order."""
    group = group if group is not None else dist.group.WORLD
    send = _pack(x_local, plan.send_index)
    recv = torch.empty((plan.n_halo, ) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    return _all_to_all_rows(recv, send, plan.recv_counts, plan.send_counts, group)


def return_halo(plan: HaloPlan, g_halo: Tensor, g_local: Tensor, group=None) -> Tensor:
    """Backward exchange: sends per-halo-row contributions back to their owners and adds them into
    g_local in place (rows listed in send_index)."""
    group = group if group is not None else dist.group.WORLD
    recv = torch.empty((plan.n_send, ) + tuple(g_halo.shape[1:]), dtype=g_halo.dtype, device=g_halo.device)
    _all_to_all_rows(recv, g_halo.contiguous(), plan.send_counts, plan.recv_counts, group)
    if plan.n_send:
        if g_local.is_cuda and g_local.dtype == torch.float32 and g_local.dim() == 2:
            from . import ops
            ops.index_add_rows(g_local, plan.send_index, recv)
        else:
            g_local.index_add_(0, plan.send_index, recv)
    return g_local


def shard_self_loops(src_global: Tensor, dst_global: Tensor, lo: int, n_local: int):
    """add_remaining_self_loops restricted to the owned destinations: drop (i,i), append one loop
    per owned node (the global rule of utils/loop.py:623-657 applied shard by shard)."""
    keep = src_global != dst_global
    loops = torch.arange(lo, lo + n_local, device=src_global.device, dtype=src_global.dtype)
    return torch.cat([src_global[keep], loops]), torch.cat([dst_global[keep], loops])


class ShardedGCNGraph:
    """One rank's slice of a gcn_norm'ed graph: CSR over the owned destinations, columns relabelled
    to [local | halo], D^-1/2 (A+I) D^-1/2 weights computed with the degrees of remote sources
    fetched by one halo exchange at build time."""

    def __init__(self, graph, plan: HaloPlan, group):
        self.graph, self.plan, self.group = graph, plan, group

    @classmethod
    def build(cls, edge_index_global: Tensor, lo: int, n_local: int, n_total: int, group=None,
              add_self_loops: bool = True, improved: bool = False):
        from . import ops
        from .graph import CSRGraph
        group = group if group is not None else dist.group.WORLD
        src, dst = edge_index_global[0], edge_index_global[1]
        if add_self_loops:
            src, dst = shard_self_loops(src, dst, lo, n_local)
        plan, src_rel = build_halo_plan(src, lo, n_local, group)
        g = CSRGraph(src_rel, dst - lo, n_local + plan.n_halo, n_local)
        # degrees are sums over incoming edges => local; remote sources' dinv comes by halo exchange
        deg = g.in_degree().to(torch.float32)
        if improved and add_self_loops:
            deg = deg + 1.0                                    # loop weight 2 instead of 1
        dinv = deg.pow(-0.5)
        dinv.masked_fill_(dinv == float("inf"), 0.0)
        dinv_halo = exchange_halo(plan, dinv.view(-1, 1), group).view(-1)
        dinv_cat = torch.cat([dinv, dinv_halo])
        w = ops.gather_rows(dinv_cat.view(-1, 1), g.col).view(-1) * ops.gather_rows(dinv.view(-1, 1), g.dst_csr).view(-1)
        if improved and add_self_loops:
            is_loop = g.col.long() == g.dst_csr.long()
            w = torch.where(is_loop, w * 2.0, w)
        g.val = w
        g.build_transpose()
        return cls(g, plan, group)


class _ShardedAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local: Tensor, bias: Optional[Tensor], shard: ShardedGCNGraph):
        from . import ops
        g, plan = shard.graph, shard.plan
        x_local = x_local.contiguous()
        halo = exchange_halo(plan, x_local, shard.group)
        ctx.shard = shard
        ctx.has_bias = bias is not None
        return ops.spmm_csr(g.rowptr, g.col, g.val, x_local, g.num_dst, "sum", g.plan, bias=bias,
                            x_halo=halo if plan.n_halo else None)

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        from . import ops
        shard = ctx.shard
        g, plan = shard.graph, shard.plan
        grad_out = grad_out.contiguous()
        gx = gb = None
        if ctx.needs_input_grad[0]:
            g_cat = ops.spmm_csr(g.rowptr_t, g.col_t, g.val_t, grad_out, g.num_src, "sum", g.plan_t)
            g_local = g_cat[:plan.n_local]
            return_halo(plan, g_cat[plan.n_local:], g_local, shard.group)
            gx = g_local
        if ctx.has_bias and ctx.needs_input_grad[1]:
            gb = grad_out.sum(0, dtype=torch.float32)
        return gx, gb, None


def sharded_aggregate(x_local: Tensor, shard: ShardedGCNGraph, bias: Optional[Tensor] = None) -> Tensor:
    return _ShardedAggregate.apply(x_local, bias, shard)


def sharded_gcn_conv(conv, x_local: Tensor, shard: ShardedGCNGraph) -> Tensor:
    """GCNConv.forward on one shard: local dense transform, halo exchange of the transformed rows,
    fused aggregate (+ bias).  Weight gradients are per-rank partial sums (all-reduce them like
    DDP does; bench.py includes that all_reduce in the timed step)."""
    return sharded_aggregate(conv.lin(x_local), shard, conv.bias)

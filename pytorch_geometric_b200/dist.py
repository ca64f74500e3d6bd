"""Node-range sharding of the aggregation path over the GPUs of one box.

One process per GPU.  Rank r owns the contiguous node range [r*n, (r+1)*n): its feature rows, its
output rows, and the CSR slice of its DESTINATIONS (the reference's closest notion is
`EdgeIndex.sparse_narrow`, edge_index.py:1028-1133, which nothing in the reference calls).  Sources
outside the range are *halo* rows.  Per aggregation pass there is exactly ONE exchange, and it is
hidden behind the local part of the sweep:

    forward :  pack the rows my peers need (gather) -> all_to_all_single (NCCL over NVLink, async)
               || gather-reduce over the LOCAL-source edges (~p_local of the work)
               -> gather-reduce over the HALO-source edges, accumulated into the same rows
    backward:  transposed gather-reduce of the HALO sources first (small) -> mirror all_to_all (async)
               || transposed gather-reduce of the LOCAL sources
               -> returned rows added into their owners' rows (red.global.add.v4)

Everything else (degrees, gcn_norm, softmax, max, mean) is per destination, hence local.  The plan
(unique remote ids grouped by owner, send lists, relabelled columns) is integer set-up work done
once per graph with device-agnostic torch ops, so the same code is exercised on CPU with the gloo
backend by tests/test_dist_gloo.py.  The reference has no counterpart of this module; correctness
is shard-vs-unsharded equality against the single-process engine and oracle.
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


class _Done:
    def wait(self):
        return None


def _all_to_all_rows(out: Tensor, inp: Tensor, out_counts: List[int], in_counts: List[int], group, async_op=False):
    """Variable-size row exchange; returns a handle with .wait() (a no-op handle for world == 1)."""
    if dist.get_world_size(group) == 1:
        return _Done()
    work = dist.all_to_all_single(out, inp, output_split_sizes=out_counts, input_split_sizes=in_counts, group=group,
                                  async_op=async_op)
    return work if async_op else _Done()


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
    slot = torch.searchsorted(halo_ids, src_global.clamp(min=0)) if halo_ids.numel() else torch.zeros_like(src_global)
    src_rel = torch.where(is_remote, slot + n_local, src_global - lo)
    plan = HaloPlan(rank, world, lo, n_local, halo_ids, recv_counts, send_index, send_counts)
    return plan, src_rel


def _pack(x_local: Tensor, index: Tensor) -> Tensor:
    if x_local.is_cuda:
        from . import ops
        return ops.gather_rows(x_local, index)
    return x_local.index_select(0, index)


def exchange_halo_start(plan: HaloPlan, x_local: Tensor, group=None):
    """Starts the forward exchange; returns (halo_rows_buffer, handle).  handle.wait() makes the
    current stream wait for the rows to have arrived."""
    group = group if group is not None else dist.group.WORLD
    send = _pack(x_local, plan.send_index)
    recv = torch.empty((plan.n_halo, ) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    work = _all_to_all_rows(recv, send, plan.recv_counts, plan.send_counts, group, async_op=x_local.is_cuda)
    return recv, work, send   # `send` is returned so it stays alive until the exchange completes


def exchange_halo(plan: HaloPlan, x_local: Tensor, group=None) -> Tensor:
    """Blocking forward exchange: the halo rows [n_halo, F] in halo_ids order."""
    recv, work, _ = exchange_halo_start(plan, x_local, group)
    work.wait()
    return recv


def return_halo_start(plan: HaloPlan, g_halo: Tensor, group=None):
    group = group if group is not None else dist.group.WORLD
    g_halo = g_halo.contiguous()
    recv = torch.empty((plan.n_send, ) + tuple(g_halo.shape[1:]), dtype=g_halo.dtype, device=g_halo.device)
    work = _all_to_all_rows(recv, g_halo, plan.send_counts, plan.recv_counts, group, async_op=g_halo.is_cuda)
    return recv, work, g_halo


def return_halo_finish(plan: HaloPlan, recv: Tensor, g_local: Tensor) -> Tensor:
    if plan.n_send:
        if g_local.is_cuda and g_local.dtype == torch.float32 and g_local.dim() == 2:
            from . import ops
            ops.index_add_rows(g_local, plan.send_index, recv)
        else:
            g_local.index_add_(0, plan.send_index, recv)
    return g_local


def return_halo(plan: HaloPlan, g_halo: Tensor, g_local: Tensor, group=None) -> Tensor:
    """Blocking backward exchange: per-halo-row contributions go back to their owners and are added
    into g_local in place (rows listed in send_index)."""
    recv, work, _ = return_halo_start(plan, g_halo, group)
    work.wait()
    return return_halo_finish(plan, recv, g_local)


def shard_self_loops(src_global: Tensor, dst_global: Tensor, lo: int, n_local: int):
    """add_remaining_self_loops restricted to the owned destinations: drop (i,i), append one loop
    per owned node (the global rule of utils/loop.py:623-657 applied shard by shard)."""
    keep = src_global != dst_global
    loops = torch.arange(lo, lo + n_local, device=src_global.device, dtype=src_global.dtype)
    return torch.cat([src_global[keep], loops]), torch.cat([dst_global[keep], loops])


class ShardedGCNGraph:
    """One rank's slice of a gcn_norm'ed graph, split by source locality:
      g_local : CSR over the owned destinations with the LOCAL-source edges (+ its transpose)
      g_halo  : CSR over the owned destinations with the HALO-source edges (+ its transpose, whose
                rows are the halo slots)
    D^-1/2 (A+I) D^-1/2 weights use the degrees of remote sources fetched by one halo exchange at
    build time.  `graph` is kept as an alias of g_local for reporting."""

    def __init__(self, g_local, g_halo, plan: HaloPlan, group):
        self.g_local, self.g_halo, self.plan, self.group = g_local, g_halo, plan, group
        self.graph = g_local
        self.num_edges = g_local.num_edges + (g_halo.num_edges if g_halo is not None else 0)

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
        dst_l = dst - lo
        # degrees are sums over incoming edges => local; remote sources' dinv comes by halo exchange
        deg = ops.degree(dst_l, n_local).to(torch.float32)
        # `improved` only changes the fill value of the inserted loops, and the reference ignores the fill
        # value when edge_weight is None (gcn_conv.py:98-104 with loop.py:623-657): unweighted graphs -- the
        # only kind this builder takes -- get loop weight 1 either way, exactly like utils.gcn_norm_graph.
        dinv = deg.pow(-0.5)
        dinv.masked_fill_(dinv == float("inf"), 0.0)
        dinv_halo = exchange_halo(plan, dinv.view(-1, 1), group).view(-1)
        dinv_cat = torch.cat([dinv, dinv_halo])
        w = ops.gather_rows(dinv_cat.view(-1, 1), src_rel).view(-1) * ops.gather_rows(dinv.view(-1, 1), dst_l).view(-1)
        is_halo = src_rel >= n_local
        keep = ~is_halo
        g_local = CSRGraph(src_rel[keep], dst_l[keep], n_local, n_local, w[keep])
        g_local.build_transpose()
        g_halo = None
        if plan.n_halo:
            g_halo = CSRGraph(src_rel[is_halo] - n_local, dst_l[is_halo], plan.n_halo, n_local, w[is_halo])
            g_halo.build_transpose()
        return cls(g_local, g_halo, plan, group)


class _ShardedAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local: Tensor, bias: Optional[Tensor], shard: ShardedGCNGraph):
        from . import ops
        gl, gh, plan = shard.g_local, shard.g_halo, shard.plan
        x_local = x_local.contiguous()
        ctx.shard = shard
        ctx.has_bias = bias is not None
        if gh is None:
            return ops.spmm_csr(gl.rowptr, gl.col, gl.val, x_local, gl.num_dst, "sum", gl.plan, bias=bias)
        halo, work, _keep = exchange_halo_start(plan, x_local, shard.group)          # async over NVLink
        out = ops.spmm_csr(gl.rowptr, gl.col, gl.val, x_local, gl.num_dst, "sum", gl.plan, bias=bias)   # overlaps
        work.wait()
        ops.spmm_csr(gh.rowptr, gh.col, gh.val, halo, gh.num_dst, "sum", gh.plan, out=out, accumulate=True)
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        from . import ops
        shard = ctx.shard
        gl, gh, plan = shard.g_local, shard.g_halo, shard.plan
        grad_out = grad_out.contiguous()
        gx = gb = None
        if ctx.needs_input_grad[0]:
            if gh is None:
                gx = ops.spmm_csr(gl.rowptr_t, gl.col_t, gl.val_t, grad_out, gl.num_src, "sum", gl.plan_t)
            else:
                g_halo = ops.spmm_csr(gh.rowptr_t, gh.col_t, gh.val_t, grad_out, gh.num_src, "sum", gh.plan_t)
                recv, work, _keep = return_halo_start(plan, g_halo, shard.group)     # async over NVLink
                gx = ops.spmm_csr(gl.rowptr_t, gl.col_t, gl.val_t, grad_out, gl.num_src, "sum", gl.plan_t)  # overlaps
                work.wait()
                return_halo_finish(plan, recv, gx)
        if ctx.has_bias and ctx.needs_input_grad[1]:
            gb = grad_out.sum(0, dtype=torch.float32)
        return gx, gb, None


def sharded_aggregate(x_local: Tensor, shard: ShardedGCNGraph, bias: Optional[Tensor] = None) -> Tensor:
    return _ShardedAggregate.apply(x_local, bias, shard)


def sharded_gcn_conv(conv, x_local: Tensor, shard: ShardedGCNGraph) -> Tensor:
    """GCNConv.forward on one shard: local dense transform, halo exchange of the transformed rows
    hidden behind the local-edge sweep, fused aggregate (+ bias).  Weight gradients are per-rank
    partial sums (all-reduce them like DDP does; bench.py includes that all_reduce in the step)."""
    return sharded_aggregate(conv.lin(x_local), shard, conv.bias)

"""Differentiable front ends of the hot path (custom autograd over the C-ABI kernels).

`aggregate(graph, x, reduce, edge_weight)` is the fused replacement for the reference's
collect -> message -> aggregate sequence (nn/conv/message_passing.py:421-563) and for
`EdgeIndex.matmul` / `spmm` (edge_index.py:1925-1970, utils/_spmm.py:12-136).
Backward follows `_TorchSPMM.backward` (edge_index.py:1860-1900): the same kernel on the
transposed CSR; min/max use the ATen tie rule (oracle_scatter_backward in oracle/mp_oracle.c).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops
from .graph import CSRGraph


class _Aggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, edge_weight: Optional[Tensor], graph: CSRGraph, reduce: str):
        val = graph.val
        if edge_weight is not None:
            val = graph.to_csr_order(edge_weight.detach().float().contiguous().view(-1))
        out = ops.spmm_csr(graph.rowptr, graph.col, val, x, graph.num_dst, reduce, graph.plan)
        ctx.graph, ctx.reduce = graph, reduce
        ctx.has_ew = edge_weight is not None
        need_x = reduce in ("min", "max") or (ctx.has_ew and edge_weight.requires_grad)
        ctx.save_for_backward(x if need_x else None, out if reduce in ("min", "max") else None, val)
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        graph, reduce = ctx.graph, ctx.reduce
        x, out, val = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        graph.build_transpose()
        gx = gw = None
        val_t = None
        if val is not None:
            val_t = graph.val_t if (not ctx.has_ew and graph.val_t is not None) else \
                graph.to_csc_order(graph.from_csr_order(val))
        if ctx.needs_input_grad[0]:
            if reduce in ("sum", "add"):
                gx = ops.spmm_csr(graph.rowptr_t, graph.col_t, val_t, grad_out, graph.num_src, "sum", graph.plan_t)
            elif reduce == "mean":
                mv = graph.mean_val_t()
                if val_t is not None:
                    mv = mv * val_t
                gx = ops.spmm_csr(graph.rowptr_t, graph.col_t, mv, grad_out, graph.num_src, "sum", graph.plan_t)
            else:  # min / max
                ties = ops.minmax_ties(graph.rowptr, graph.col, val, x, out, count_self_zero=True)
                gx = ops.minmax_backward(graph.rowptr_t, graph.col_t, val_t, x, out, grad_out, ties)
        if ctx.has_ew and ctx.needs_input_grad[1]:
            if reduce not in ("sum", "add", "mean"):
                raise NotImplementedError("gradient wrt edge_weight is implemented for sum/mean only")
            g = grad_out
            dot = ops.sddmm_csr(graph.rowptr, graph.col, g, x)          # CSR order
            if reduce == "mean":
                inv = 1.0 / graph.in_degree().clamp(min=1).to(torch.float32)
                dot = dot * ops.gather_rows(inv.view(-1, 1), graph.dst_csr).view(-1)
            gw = graph.from_csr_order(dot)
        return gx, gw, None, None


def aggregate(graph: CSRGraph, x: Tensor, reduce: str = "sum", edge_weight: Optional[Tensor] = None) -> Tensor:
    """out[i] = REDUCE_{(j -> i)} w_ji * x[j]; x: [num_src, F] -> [num_dst, F].

    `edge_weight` (original edge order, may require grad) overrides the static values cached in
    the graph (e.g. gcn_norm weights).  Empty destinations give 0 for every reduce.
    """
    if reduce not in ("sum", "add", "mean", "min", "max"):
        raise ValueError(f"Encountered invalid `reduce` argument '{reduce}'")
    if x.dim() == 1:
        return aggregate(graph, x.view(-1, 1), reduce, edge_weight).view(-1)
    if x.dim() > 2:
        shape = x.shape[1:]
        return aggregate(graph, x.reshape(x.size(0), -1), reduce, edge_weight).view((graph.num_dst, ) + shape)
    if x.size(0) != graph.num_src:
        raise ValueError(f"x has {x.size(0)} rows but the graph has {graph.num_src} source nodes")
    return _Aggregate.apply(x, edge_weight, graph, reduce)


class _Segment(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src: Tensor, ptr: Tensor, reduce: str):
        ctx.plan = ops.segment_plan(ptr, src.size(0))      # hub segments are cut into chunks (csr_reduce.cuh)
        out = ops.segment_csr(src, ptr, reduce, ctx.plan)
        ctx.reduce, ctx.n_src = reduce, src.size(0)
        ctx.save_for_backward(ptr, src if reduce in ("min", "max") else None, out if reduce in ("min", "max") else None)
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        ptr, src, out = ctx.saved_tensors
        reduce = ctx.reduce
        n_src = ctx.n_src                                   # == ptr[-1], known on the host: no D2H read
        index = ops.ptr2index(ptr, n_src)
        g2 = grad_out.contiguous().view(grad_out.size(0), -1)
        if reduce in ("sum", "add"):
            g = ops.gather_rows(g2, index)
        elif reduce == "mean":
            inv = 1.0 / (ptr[1:] - ptr[:-1]).clamp(min=1).to(torch.float32)
            g = ops.gather_rows(g2, index, ops.gather_rows(inv.view(-1, 1), index).view(-1))
        else:
            # _segment_reduce backward: even split among ties (no zero-initialised self here)
            s2, o2 = src.view(src.size(0), -1), out.view(out.size(0), -1)
            eq = (s2 == ops.gather_rows(o2, index)).to(s2.dtype)
            ties = ops.segment_csr(eq, ptr, "sum", ctx.plan)
            g = eq * ops.gather_rows(g2 / ties.clamp(min=1), index)
        return g.view((n_src, ) + tuple(grad_out.shape[1:])), None, None


def segment(src: Tensor, ptr: Tensor, reduce: str = "sum") -> Tensor:
    return _Segment.apply(src, ptr, reduce)


class _ScatterCOO(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src: Tensor, index: Tensor, dim_size: int, reduce: str):
        out = ops.scatter_coo(src, index, dim_size, reduce)
        ctx.reduce, ctx.dim_size = reduce, dim_size
        keep = reduce in ("min", "max", "mul")
        ctx.save_for_backward(index, src if keep else None, out if keep else None)
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        index, src, out = ctx.saved_tensors
        reduce = ctx.reduce
        g2 = grad_out.contiguous().view(grad_out.size(0), -1)
        if reduce in ("sum", "add"):
            g = ops.gather_rows(g2, index)                       # scatter_add_ backward == gather
        elif reduce == "mean":
            cnt = ops.degree(index, ctx.dim_size).clamp(min=1).to(torch.float32)
            g = ops.gather_rows(g2, index, ops.gather_rows((1.0 / cnt).view(-1, 1), index).view(-1))
        elif reduce in ("min", "max"):
            s2, o2 = src.view(src.size(0), -1), out.view(out.size(0), -1)
            eq = (s2 == ops.gather_rows(o2, index)).to(torch.float32)
            # ATen scatter_reduce rule incl. the zero-initialised `self` tie (oracle_scatter_backward)
            ties = ops.scatter_coo(eq, index, ctx.dim_size, "sum") + (o2 == 0).to(torch.float32)
            g = eq * ops.gather_rows(g2 / ties, index)
        else:
            # ATen's scatter_reduce_('prod') rule (FunctionsManual.cpp, scatter_reduce_backward; `self` is all ones here):
            # a value that is the ONLY zero of its group gets grad * (product of the others), every other value gets
            # grad * result / value (0 when its group holds a zero elsewhere, 0 for every member of a group with >= 2 zeros)
            s2, o2 = src.view(src.size(0), -1).float(), out.view(out.size(0), -1).float()
            zero = s2 == 0
            n_zero = ops.gather_rows(ops.scatter_coo(zero.to(torch.float32), index, ctx.dim_size, "sum"), index)
            single = zero & (n_zero == 1)
            others = ops.scatter_coo(torch.where(single, torch.ones_like(s2), s2), index, ctx.dim_size, "mul")
            g = torch.where(single, ops.gather_rows(g2 * others, index), ops.gather_rows(g2 * o2, index) / torch.where(zero, torch.ones_like(s2), s2))
        return g.view((index.numel(), ) + tuple(grad_out.shape[1:])), None, None, None


def scatter_coo(src: Tensor, index: Tensor, dim_size: int, reduce: str = "sum") -> Tensor:
    return _ScatterCOO.apply(src, index, dim_size, reduce)


class _ScatterAny(torch.autograd.Function):
    """scatter(reduce='any') (utils/_scatter.py:75-77: `src.new_zeros(size).scatter_(dim, index, src)`): one member of
    every group.  Which one is unspecified on CUDA; here it is the LAST one in index order -- what the reference's CPU
    kernel produces -- found by an integer amax over the positions, then one row gather.  Backward = scatter_'s:
    every member receives its group's gradient."""

    @staticmethod
    def forward(ctx, src: Tensor, index: Tensor, dim_size: int):
        pos = torch.arange(index.numel(), device=index.device, dtype=torch.int64)
        last = torch.full((dim_size, ), -1, dtype=torch.int64, device=index.device).scatter_reduce_(
            0, index.long(), pos, "amax", include_self=True)
        s2 = src.contiguous().view(src.size(0), -1)
        if s2.size(0) == 0:
            out = s2.new_zeros((dim_size, s2.size(1)))
        else:
            out = ops.gather_rows(s2, ops.convert_index(last.clamp(min=0), index.dtype))
            out = out * (last >= 0).to(out.dtype).view(-1, 1)                    # empty groups stay 0
        ctx.save_for_backward(index)
        return out.view((dim_size, ) + tuple(src.shape[1:]))

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        index, = ctx.saved_tensors
        g = ops.gather_rows(grad_out.contiguous().view(grad_out.size(0), -1), index)
        return g.view((index.numel(), ) + tuple(grad_out.shape[1:])), None, None


def scatter_any(src: Tensor, index: Tensor, dim_size: int) -> Tensor:
    return _ScatterAny.apply(src, index, dim_size)


class _SoftmaxCSR(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src: Tensor, ptr: Tensor, index: Optional[Tensor]):
        ctx.plan = ops.segment_plan(ptr, src.size(0))        # hub groups: chunked path (ops.softmax_csr)
        if ctx.plan is not None and index is None:
            index = ops.ptr2index(ptr, src.size(0))
        out = ops.softmax_csr(src, ptr, ctx.plan, index)
        ctx.save_for_backward(out, ptr, index)
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        out, ptr, index = ctx.saved_tensors
        return ops.softmax_csr_backward(out, grad_out, ptr, ctx.plan, index), None, None


def softmax_csr(src: Tensor, ptr: Tensor, index: Optional[Tensor] = None) -> Tensor:
    return _SoftmaxCSR.apply(src, ptr, index)


class _GATFused(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xh: Tensor, a_src: Tensor, a_dst: Tensor, graph: CSRGraph, heads: int, chan: int, slope: float,
                want_alpha: bool):
        out, row_max, row_den, alpha = ops.gat_fused_csr(graph.rowptr, graph.col, xh, a_src, a_dst, heads, chan,
                                                         slope, want_alpha, plan=graph.plan,
                                                         dst_of_edge=graph.dst_csr if want_alpha else None)
        ctx.graph, ctx.dims = graph, (heads, chan, slope)
        ctx.save_for_backward(xh, a_src, a_dst, row_max, row_den, out)
        if alpha is None:
            alpha = xh.new_empty(0)
        ctx.mark_non_differentiable(alpha)
        return out, alpha

    @staticmethod
    def backward(ctx, grad_out: Tensor, _grad_alpha):
        xh, a_src, a_dst, row_max, row_den, out = ctx.saved_tensors
        graph = ctx.graph
        heads, chan, slope = ctx.dims
        graph.build_transpose()
        gxh, gas, gad = ops.gat_fused_csr_backward(graph.rowptr, graph.col, graph.dst_csr, graph.rowptr_t, graph.col_t,
                                                   graph.t2csr, xh, a_src.float().contiguous(),
                                                   a_dst.float().contiguous(), row_max, row_den, out, grad_out,
                                                   heads, chan, slope, plan=graph.plan)
        return gxh, gas.to(a_src.dtype), gad.to(a_dst.dtype), None, None, None, None, None


def gat_attention(graph: CSRGraph, xh: Tensor, a_src: Tensor, a_dst: Tensor, heads: int, chan: int,
                  negative_slope: float = 0.2, return_alpha: bool = False):
    """out[i,h,:] = sum_e softmax_i(leaky_relu(a_src[j,h] + a_dst[i,h]))_e * xh[j,h,:]
    (GATConv.edge_update + message + aggregate, gat_conv.py:387-409) in one fused sweep.
    xh: [num_src, H*C]; returns out [num_dst, H*C] (and alpha [E, H] in CSR order)."""
    if ops.attn_supported(heads, chan, xh.dtype) and xh.data_ptr() % 16 == 0:
        return attention("gat", graph, heads, chan, v=xh, s_src=a_src, s_dst=a_dst, negative_slope=negative_slope,
                         return_alpha=return_alpha)
    # head widths off the vector path: the scalar kernels of csrc/gat.cu (no padding copy needed for GAT)
    out, alpha = _GATFused.apply(xh, a_src, a_dst, graph, heads, chan, float(negative_slope), return_alpha)
    return (out, alpha) if return_alpha else out


class _AttnFused(torch.autograd.Function):
    """Fused attention family (csrc/attention.cu): GAT / GATv2 / dot-product scores, edge softmax and the weighted
    aggregation in one sweep; backward = destination sweep + source sweep with the attention recomputed."""

    @staticmethod
    def forward(ctx, mode: str, graph: CSRGraph, heads: int, chan: int, slope: float, scale: float, want_alpha: bool,
                v: Tensor, k: Optional[Tensor], q: Optional[Tensor], s_src: Optional[Tensor], s_dst: Optional[Tensor],
                att: Optional[Tensor], s_edge: Optional[Tensor], kv: Optional[Tensor], dropout: tuple = (0.0, 0),
                e_feat: Optional[Tensor] = None):
        hc = heads * chan
        if kv is not None:                                   # keys | values as the two halves of one [N, 2HC] product
            k, v = kv[:, :hc], kv[:, hc:]
        f32 = lambda t: None if t is None else t.detach().float().contiguous()   # noqa: E731
        s_src32, s_dst32, att32 = f32(s_src), f32(s_dst), (None if att is None else f32(att).view(-1))
        s_edge_csr = None if s_edge is None else graph.to_csr_order_rows(f32(s_edge))
        if q is not None and q.stride(1) != 1:
            q = q.contiguous()
        ref = kv if kv is not None else v
        e_csr = None if e_feat is None else graph.to_csr_order_rows(e_feat.detach().to(ref.dtype).contiguous())
        out, row_max, row_den, alpha = ops.attn_forward(mode, graph.rowptr, graph.col, v, heads, chan, k=k, q=q, s_src=s_src32,
                                                        s_dst=s_dst32, att=att32, s_edge=s_edge_csr, slope=slope, scale=scale,
                                                        want_alpha=want_alpha, plan=graph.plan, dropout_p=dropout[0],
                                                        dropout_seed=dropout[1], edge_feat=e_csr)
        ctx.mode, ctx.graph, ctx.dims, ctx.fused_kv = mode, graph, (heads, chan, slope, scale), kv is not None
        ctx.dropout = dropout
        ctx.dt = tuple(None if t is None else t.dtype for t in (s_src, s_dst, att, s_edge, e_feat))
        ctx.save_for_backward(kv if kv is not None else v, None if kv is not None else k, q, s_src32, s_dst32, att32, s_edge_csr,
                              row_max, row_den, out, e_csr)
        if alpha is None:
            alpha = out.new_empty(0)
        ctx.mark_non_differentiable(alpha)
        return out, alpha

    @staticmethod
    def backward(ctx, grad_out: Tensor, _grad_alpha):
        v, k, q, s_src, s_dst, att, s_edge, row_max, row_den, out, e_csr = ctx.saved_tensors
        graph = ctx.graph
        heads, chan, slope, scale = ctx.dims
        hc = heads * chan
        graph.build_transpose()
        grad_kv = grad_v = grad_k = None
        if ctx.fused_kv:
            kv = v
            k, v = kv[:, :hc], kv[:, hc:]
            grad_kv = torch.empty_like(kv)
            grad_k, grad_v = grad_kv[:, :hc], grad_kv[:, hc:]
        r = ops.attn_backward(ctx.mode, graph.rowptr, graph.col, graph.rowptr_t, graph.col_t, graph.t2csr, v, heads, chan,
                              row_max, row_den, out, grad_out, k=k, q=q, s_src=s_src, s_dst=s_dst, att=att, s_edge=s_edge,
                              slope=slope, scale=scale, plan=graph.plan, plan_t=graph.plan_t, grad_v=grad_v, grad_k=grad_k,
                              dropout_p=ctx.dropout[0], dropout_seed=ctx.dropout[1], edge_feat=e_csr)
        cast = lambda t, d: None if (t is None or d is None) else t.to(d)       # noqa: E731
        g_edge = None
        if s_edge is not None:
            g_edge = cast(graph.from_csr_order_rows(r["grad_s_edge"].contiguous()), ctx.dt[3])
        g_att = None if r["grad_att"] is None else cast(r["grad_att"], ctx.dt[2])
        g_ef = None
        if e_csr is not None and ctx.needs_input_grad[16]:
            g_ef = cast(graph.from_csr_order_rows(r["grad_edge_feat"]), ctx.dt[4])
        return (None, None, None, None, None, None, None,
                None if ctx.fused_kv else r["grad_v"], None if ctx.fused_kv else r["grad_k"], r["grad_q"],
                cast(r["grad_s_src"], ctx.dt[0]), cast(r["grad_s_dst"], ctx.dt[1]), g_att, g_edge, grad_kv, None, g_ef)


def _vector_shape(heads: int, chan: int, dtype: torch.dtype):
    """(chan_padded, heads_per_group) that put [*, heads*chan] rows on the vector path of csrc/attention.cu:
    a head is a power-of-two number of 16-byte vectors and a row group is at most 1 KB."""
    epv = 8 if dtype == torch.bfloat16 else 4
    vec = -(-chan // epv)
    lph = 1
    while lph < vec:
        lph *= 2
    if lph > 32:
        raise NotImplementedError(f"attention heads wider than 512 bytes (C = {chan}) are not on the fused path")
    chan_p = lph * epv
    per_group = max(1, 64 // lph)
    return chan_p, min(heads, per_group)


def attention(mode: str, graph: CSRGraph, heads: int, chan: int, *, v: Optional[Tensor] = None, k: Optional[Tensor] = None,
              q: Optional[Tensor] = None, kv: Optional[Tensor] = None, s_src: Optional[Tensor] = None,
              s_dst: Optional[Tensor] = None, att: Optional[Tensor] = None, s_edge: Optional[Tensor] = None,
              negative_slope: float = 0.2, scale: float = 1.0, return_alpha: bool = False, dropout_p: float = 0.0,
              dropout_seed: Optional[int] = None, e_feat: Optional[Tensor] = None):
    """out[i,h,:] = sum_e softmax_i(score_e,h) v[j,h,:] for mode in {"gat", "gatv2", "dot"} (see csrc/attention.cu).
    dropout_p > 0: attention dropout (F.dropout on the normalised coefficients, gat_conv.py:404) inside the sweep; the
    seed is drawn from torch's CPU generator (reproducible under torch.manual_seed, no device sync) unless given.
    e_feat [E, H*C] (caller's edge order; "gatv2" and "dot"): lin_edge(edge_attr) of `edge_dim` layers -- added inside
    GATv2's leaky_relu (gatv2_conv.py:358-360) / to the key and the value of the dot mode (transformer_conv.py:258-272).
    All feature operands are [n, H*C]; `kv` = [n_src, 2*H*C] (keys | values from one fused product) instead of k, v;
    s_edge [E, H] in the caller's edge order.  Returns out (and alpha [E, H] in CSR order).

    Head widths off the kernel's vector path (C not a power-of-two number of 16-byte vectors, rows above 1 KB) are
    mapped onto it: channels are zero-padded per head (zeros change neither a score nor a sum) and heads -- which are
    independent -- are processed in groups; the kernels are the same."""
    ref = kv if kv is not None else v
    dtype = ref.dtype
    if dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"attention operands must be float32 or bfloat16, got {dtype}")
    chan_p, hpg = _vector_shape(heads, chan, dtype)
    if not 0.0 <= dropout_p < 1.0:
        raise ValueError(f"dropout probability has to be in [0, 1), got {dropout_p}")
    if dropout_p > 0.0 and dropout_seed is None:
        dropout_seed = int(torch.randint(0, 2 ** 62, (1, ), device="cpu"))
    drop = (float(dropout_p), int(dropout_seed or 0))
    if chan_p == chan and hpg == heads and ops.attn_supported(heads, chan, dtype):
        out, alpha = _AttnFused.apply(mode, graph, heads, chan, float(negative_slope), float(scale), return_alpha, v, k, q,
                                      s_src, s_dst, att if att is None else att.reshape(-1), s_edge, kv, drop, e_feat)
        return (out, alpha) if return_alpha else out
    if kv is not None:
        k, v = kv[:, :heads * chan], kv[:, heads * chan:]

    def group(t, h0, h1, feat: bool):
        if t is None:
            return None
        if not feat:                                        # [n, H] scalars per head
            return t[:, h0:h1]
        t3 = t.reshape(t.size(0), heads, chan)[:, h0:h1]
        if chan_p != chan:
            t3 = torch.nn.functional.pad(t3, (0, chan_p - chan))
        return t3.reshape(t.size(0), (h1 - h0) * chan_p).contiguous()

    outs, alphas = [], []
    for h0 in range(0, heads, hpg):
        h1 = min(h0 + hpg, heads)
        a_g = None if att is None else group(att.reshape(1, heads * chan), h0, h1, True).view(-1)
        o, al = _AttnFused.apply(mode, graph, h1 - h0, chan_p, float(negative_slope), float(scale), return_alpha,
                                 group(v, h0, h1, True), group(k, h0, h1, True), group(q, h0, h1, True),
                                 group(s_src, h0, h1, False), group(s_dst, h0, h1, False), a_g, group(s_edge, h0, h1, False),
                                 None, (drop[0], drop[1] + 0x51ED27 * h0),          # a different stream per head group
                                 group(e_feat, h0, h1, True))
        outs.append(o.view(o.size(0), h1 - h0, chan_p)[:, :, :chan])
        alphas.append(al)
    out = torch.cat(outs, dim=1).reshape(outs[0].size(0), heads * chan)
    if return_alpha:
        return out, torch.cat(alphas, dim=1)
    return out


class _MultiAggregate(torch.autograd.Function):
    """k aggregations of one neighbourhood in one sweep (FusedAggregation, nn/aggr/fused.py:191-336).
    `where` is a CSRGraph (gather mode: x is [num_src, F]) or a (ptr, index) pair (segment mode: x is the
    destination-sorted [E, F] message matrix)."""

    @staticmethod
    def forward(ctx, x: Tensor, where, names: tuple, semi_grad: bool, count_self_zero: bool):
        gather = isinstance(where, CSRGraph)
        if gather:
            rowptr, col, n_rows, plan = where.rowptr, where.col, where.num_dst, where.plan
        else:
            rowptr, col, n_rows, plan = where[0], None, where[0].numel() - 1, where[2]
        need_grad = x.requires_grad
        # the var / std gradients need the group mean even when it is not an output
        extra = ("mean", ) if need_grad and "mean" not in names and ("var" in names or "std" in names) else ()
        res = ops.multi_aggr_csr(rowptr, col, x, n_rows, tuple(names) + extra, plan, with_ties=need_grad,
                                 count_self_zero=count_self_zero)
        ctx.where, ctx.names, ctx.gather, ctx.semi_grad = where, tuple(names), gather, semi_grad
        ctx.save_for_backward(x, res.get("min"), res.get("max"), res.get("ties_min"), res.get("ties_max"),
                              res.get("mean"), res.get("std"), res.get("hit_mask"))
        return tuple(res[n] for n in names)

    @staticmethod
    def backward(ctx, *grads):
        x, omin, omax, tmin, tmax, mean, std, hit_mask = ctx.saved_tensors
        where = ctx.where
        g = {n: (None if gr is None else gr.to(x.dtype)) for n, gr in zip(ctx.names, grads)}
        if all(v is None for v in g.values()):
            return None, None, None, None, None
        rowptr = where.rowptr if ctx.gather else where[0]
        term_a, term_b, gmin, gmax = ops.multi_aggr_prepare_backward(rowptr, g, mean, std, tmin, tmax, ctx.semi_grad)
        omin = omin if gmin is not None else None
        omax = omax if gmax is not None else None
        if ctx.gather:
            where.build_transpose()
            gx = ops.multi_aggr_backward(where.rowptr_t, where.col_t, x, term_a, term_b, omin, gmin, omax, gmax, False,
                                         hit_mask, None if hit_mask is None else where.t2csr)
        else:
            gx = ops.multi_aggr_backward(None, where[1], x, term_a, term_b, omin, gmin, omax, gmax, True)
        return gx, None, None, None, None


def multi_aggregate(where, x: Tensor, aggrs, semi_grad: bool = False, count_self_zero: bool = True):
    """[aggr(x) for aggr in aggrs] with aggrs from {sum, mean, min, max, var, std}, one sweep over the edges.
    where: CSRGraph (x: [num_src, F]) or (ptr, index[, plan]) for a destination-sorted [E, F] message matrix."""
    names = tuple({"add": "sum"}.get(a, a) for a in aggrs)
    if len(set(names)) != len(names):
        raise ValueError("duplicate aggregation in the fused list")
    for n in names:
        if n not in ops.MULTI_AGGRS:
            raise ValueError(f"cannot fuse aggregation '{n}' (supported: {ops.MULTI_AGGRS})")
    if not isinstance(where, CSRGraph):
        where = tuple(where) + (None, ) * (3 - len(where))
    return list(_MultiAggregate.apply(x, where, names, semi_grad, count_self_zero))

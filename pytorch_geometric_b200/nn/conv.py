"""GNN layers on the fused aggregation path.

Two layers of code:
  * functional cores (`gcn_conv`, `sage_conv`, `graph_conv`, `gin_aggregate`, `gat_conv`, `gatv2_conv`,
    `transformer_conv`, `rgcn_conv`): the layer arithmetic on explicit parameters -- shared by the standalone modules
    below and by the subclasses of the reference's own layer classes in `pytorch_geometric_b200.plugin.conv`;
  * standalone `torch.nn.Module`s whose constructor arguments, parameter names and `state_dict` layout follow the
    reference layers so checkpoints are interchangeable:

      GCNConv         nn/conv/gcn_conv.py:116-274        lin.weight [out,in], bias [out]
      SAGEConv        nn/conv/sage_conv.py:19-156        lin_l.weight/bias, lin_r.weight
      GraphConv       nn/conv/graph_conv.py:13-115       lin_rel.weight/bias, lin_root.weight
      GINConv         nn/conv/gin_conv.py:18-105         nn.*, eps [1]
      RGCNConv        nn/conv/rgcn_conv.py:40-300        weight [R,in,out] (or bases/blocks + comp), root, bias
      FastRGCNConv    nn/conv/rgcn_conv.py:302-374       same parameters
      GATConv         nn/conv/gat_conv.py:27-413         lin (or lin_src/lin_dst), att_src/att_dst, lin_edge/att_edge, res, bias
      GATv2Conv       nn/conv/gatv2_conv.py:24-385       lin_l, lin_r, att [1,H,C], res, bias
      TransformerConv nn/conv/transformer_conv.py:17-285 lin_key/lin_query/lin_value/lin_skip(/lin_beta)

`forward(x, edge_index, ...)` accepts a `[2, E]` tensor or a prebuilt `CSRGraph` -- the counterpart of handing the
reference a `SparseTensor adj_t`.  Graph structures built from a `[2, E]` tensor are cached by the identity of that
tensor (`graph.cached_graph`), so a training loop that passes the same edge_index every step sorts once.
Dense transforms: hand-written tcgen05 3xTF32 GEMMs where the shape allows (dense.py), a library GEMM otherwise.
"""
from __future__ import annotations

import math
from typing import Optional, Union

import torch
import torch.nn.functional as F
from torch import Tensor

from .. import dense
from .. import functional as Fn
from .. import ops
from .. import utils as U
from ..graph import CSRGraph, cached_graph

Adj = Union[Tensor, CSRGraph]


def glorot_(w: Tensor) -> Tensor:
    a = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
    with torch.no_grad():
        return w.uniform_(-a, a)


class _Lin(torch.nn.Module):
    """torch_geometric.nn.dense.linear.Linear (nn/dense/linear.py:121-127): x W^T + b."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels)) if bias else None
        glorot_(self.weight)

    def forward(self, x: Tensor) -> Tensor:
        return dense.linear(x, self.weight, self.bias)


def _w(mod):
    return None if mod is None else mod.weight


class _BiasAggregate(torch.autograd.Function):
    """aggregate(graph, x, 'sum') + bias with the bias add fused into the kernel epilogue."""

    @staticmethod
    def forward(ctx, x: Tensor, bias: Tensor, graph: CSRGraph):
        ctx.graph = graph
        return ops.spmm_csr(graph.rowptr, graph.col, graph.val, x, graph.num_dst, "sum", graph.plan, bias=bias)

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        graph = ctx.graph
        grad_out = grad_out.contiguous()
        gx = gb = None
        if ctx.needs_input_grad[0]:
            graph.build_transpose()
            gx = ops.spmm_csr(graph.rowptr_t, graph.col_t, graph.val_t, grad_out, graph.num_src, "sum", graph.plan_t)
        if ctx.needs_input_grad[1]:
            gb = ops.column_sum(grad_out)
        return gx, gb, None


# ================================================================================================ functional cores
def gcn_conv(x: Tensor, graph: CSRGraph, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """GCNConv.forward after gcn_norm (gcn_conv.py:241-268): aggregate(x W^T) + b, bias fused into the sweep."""
    xw = dense.linear(x, weight)
    if bias is not None:
        return _BiasAggregate.apply(xw, bias, graph)
    return Fn.aggregate(graph, xw, "sum")


class _SageFused(torch.autograd.Function):
    """One SAGEConv layer on a non-bipartite graph as ONE autograd node (sage_conv.py:120-152):
        agg = aggr_j x_j;  y = act(agg W_l^T + x W_r^T + b)
    forward  = gather-reduce sweep + one pair GEMM (two A streams into one TMEM accumulator, bias / ReLU in the epilogue);
    backward = one pair GEMM for both input gradients (one read of g), the two split-K weight gradients, the bias column
               sum, and the transposed sweep ACCUMULATING into the root gradient (out += A^T g_agg in the kernel epilogue)
    -- no elementwise pass of size N x F exists in either direction except the ReLU mask."""

    @staticmethod
    def forward(ctx, x: Tensor, w_l: Tensor, b_l: Optional[Tensor], w_r: Tensor, graph: CSRGraph, aggr: str, relu: bool,
                input_is_relu: bool, grad_masked_by_consumer: bool):
        x = x.contiguous()
        agg = ops.spmm_csr(graph.rowptr, graph.col, graph.val, x, graph.num_dst, aggr, graph.plan)
        w_hi, w_lo = dense.split_tf32(torch.cat([w_l.detach(), w_r.detach()], dim=1))
        y, _ = dense.gemm_pair(agg, x, w_hi, w_lo, 0, w_l.size(0), bias=None if b_l is None else b_l.detach(), relu=relu)
        ctx.graph, ctx.aggr, ctx.relu, ctx.has_bias = graph, aggr, relu, b_l is not None
        ctx.input_is_relu, ctx.premasked = input_is_relu, grad_masked_by_consumer
        ctx.save_for_backward(x, agg, w_hi, w_lo, y if (relu and not grad_masked_by_consumer) else None)
        return y

    @staticmethod
    def backward(ctx, g: Tensor):
        x, agg, w_hi, w_lo, y = ctx.saved_tensors
        graph = ctx.graph
        g = g.contiguous()
        if ctx.relu and not ctx.premasked:
            g = g * (y > 0)
        k = x.size(1)
        gx = gwl = gwr = gb = None
        if ctx.needs_input_grad[0]:
            ga, gx = dense.gemm_pair(g, None, w_hi, w_lo, 1, k, k)                   # g . [W_l | W_r]
            graph.build_transpose()
            val_t = graph.mean_val_t() if ctx.aggr == "mean" else graph.val_t
            # gx += A^T ga in the sweep's epilogue; when this layer's input is a ReLU output (x = relu(pre)), the same
            # epilogue applies that ReLU's backward mask (x > 0), so the producing layer needs no elementwise pass
            ops.spmm_csr(graph.rowptr_t, graph.col_t, val_t, ga, graph.num_src, "sum", graph.plan_t, out=gx, accumulate=True,
                         relu_mask=x if ctx.input_is_relu else None)
        if ctx.needs_input_grad[1]:
            gwl = dense._mm_tn(g, agg)
        if ctx.needs_input_grad[3]:
            gwr = dense._mm_tn(g, x)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = ops.column_sum(g)
        return gx, gwl, gb, gwr, None, None, None, None, None


def _sage_fusable(x_src: Tensor, x_dst: Optional[Tensor], graph: CSRGraph, aggr: str, w_l: Tensor, w_r: Optional[Tensor]) -> bool:
    k, n = x_src.size(1), w_l.size(0)
    return (x_dst is x_src and w_r is not None and aggr in ("mean", "sum", "add") and graph.val is None
            and graph.num_src == graph.num_dst and x_src.is_cuda and x_src.dtype == torch.float32 and w_l.dtype == torch.float32
            and dense.get_backend() == "tf32x3" and k % 128 == 0 and n % 128 == 0 and w_l.size(1) == k and w_r.size(1) == k
            and x_src.size(0) > 0)


def sage_conv(x_src: Tensor, x_dst: Optional[Tensor], graph: CSRGraph, aggr: str, w_l: Tensor, b_l: Optional[Tensor],
              w_r: Optional[Tensor], normalize: bool = False, relu: bool = False, input_is_relu: bool = False,
              grad_masked_by_consumer: bool = False) -> Tensor:
    """SAGEConv.forward (sage_conv.py:120-152): act(lin_l(aggr_j x_j) + lin_r(x_i)).  Non-bipartite fp32 layers with
    widths on the GEMM kernel's grid run as ONE autograd node (`_SageFused`); otherwise the two products still
    accumulate into one output (dense.linear_pair).

    Stacking hints (both default False and are only honoured by the fused node; results are identical either way):
    `input_is_relu` -- x is the output of a ReLU: the gradient this layer returns is masked by (x > 0) in the sweep's
    epilogue; `grad_masked_by_consumer` -- this layer's `relu=True` output feeds ONLY a layer called with
    `input_is_relu=True`, so its own backward skips the (idempotent) mask pass."""
    if _sage_fusable(x_src, x_dst, graph, aggr, w_l, w_r):
        out = _SageFused.apply(x_src, w_l, b_l, w_r, graph, "sum" if aggr == "add" else aggr, relu, input_is_relu,
                               grad_masked_by_consumer and relu)
    else:
        agg = Fn.aggregate(graph, x_src, aggr)
        if w_r is not None and x_dst is not None:
            out = dense.linear_pair(agg, w_l, x_dst, w_r, b_l, relu=relu)
        else:
            out = dense.linear(agg, w_l, b_l, relu=relu)
    if normalize:
        out = F.normalize(out, p=2.0, dim=-1)
    return out


def graph_conv(x_src: Tensor, x_dst: Optional[Tensor], graph: CSRGraph, aggr: str, w_rel: Tensor, b_rel: Optional[Tensor],
               w_root: Tensor, edge_weight: Optional[Tensor] = None) -> Tensor:
    """GraphConv.forward (graph_conv.py:78-112): lin_rel(aggr_j e_ji x_j) + lin_root(x_i)."""
    agg = Fn.aggregate(graph, x_src, aggr, edge_weight)
    if x_dst is not None:
        return dense.linear_pair(agg, w_rel, x_dst, w_root, b_rel)
    return dense.linear(agg, w_rel, b_rel)


def gin_aggregate(x_src: Tensor, x_dst: Optional[Tensor], graph: CSRGraph, eps) -> Tensor:
    """GINConv before its MLP (gin_conv.py:86-92): sum_j x_j + (1 + eps) x_i."""
    out = Fn.aggregate(graph, x_src, "sum")
    if x_dst is not None:
        out = out + (1 + eps) * x_dst
    return out


class _HeadDot(torch.autograd.Function):
    """(s_a, s_b) = ((xh * att_a).sum(-1), (xh * att_b).sum(-1)) from ONE read of xh, and a one-pass backward
    (csrc/head_dot.cu) instead of the ten broadcast-multiply / reduce launches autograd makes of gat_conv.py:330-331."""

    @staticmethod
    def forward(ctx, xh, att_a, att_b, H, C):
        s_a, s_b = ops.head_dot(xh, att_a, att_b, H, C)
        ctx.save_for_backward(xh, att_a, att_b)
        ctx.hc = (H, C)
        return s_a, s_b                                                   # s_b is None without att_b

    @staticmethod
    def backward(ctx, g_a, g_b):
        xh, att_a, att_b = ctx.saved_tensors
        H, C = ctx.hc
        if g_a is None:
            g_a = torch.zeros(xh.size(0), H, dtype=torch.float32, device=xh.device)
        if att_b is not None and g_b is None:
            g_b = torch.zeros(xh.size(0), H, dtype=torch.float32, device=xh.device)
        gx, ga, gb = ops.head_dot_backward(xh, att_a, att_b, g_a, g_b if att_b is not None else None, None, H, C,
                                           ctx.needs_input_grad[0])
        ga = ga.view_as(att_a).to(att_a.dtype) if ctx.needs_input_grad[1] else None
        gb = gb.view_as(att_b).to(att_b.dtype) if att_b is not None and ctx.needs_input_grad[2] else None
        return gx, ga, gb, None, None


def _head_dot(xh: Tensor, att: Tensor, H: int, C: int, att2: Optional[Tensor] = None):
    """(xh.view(-1,H,C) * att).sum(-1) in fp32: the node-level attention terms (gat_conv.py:330-331).  With att2 the
    pair of terms for two attention vectors over the same features."""
    if ops.head_dot_supported(xh, H, C):
        s_a, s_b = _HeadDot.apply(xh, att, att2, H, C)
        return s_a if att2 is None else (s_a, s_b)
    x3 = xh.view(-1, H, C).float()
    s_a = (x3 * att.view(1, H, C).float()).sum(dim=-1)
    return s_a if att2 is None else (s_a, (x3 * att2.view(1, H, C).float()).sum(dim=-1))


def _finish_heads(out: Tensor, H: int, C: int, concat: bool, res: Optional[Tensor], bias: Optional[Tensor]) -> Tensor:
    if not concat:
        out = out.view(-1, H, C).mean(dim=1)
    if res is not None:
        out = out + res
    if bias is not None:
        out = out + bias.to(out.dtype)
    return out


def gat_conv(xh_src: Tensor, xh_dst: Optional[Tensor], graph: CSRGraph, att_src: Tensor, att_dst: Optional[Tensor], H: int,
             C: int, negative_slope: float = 0.2, concat: bool = True, res: Optional[Tensor] = None,
             bias: Optional[Tensor] = None, s_edge: Optional[Tensor] = None, return_alpha: bool = False,
             dropout_p: float = 0.0):
    """GATConv after its linear maps (gat_conv.py:330-385): xh_* = lin(x) [n, H*C]; xh_dst None = the sources are the
    destinations; att_dst None = no destination term (x = (x_src, None)).  s_edge [E, H] =
    (lin_edge(edge_attr) * att_edge).sum(-1) aligned with the edges the graph was built from (edge_dim)."""
    if att_dst is None:
        a_src = _head_dot(xh_src, att_src, H, C)
        a_dst = a_src.new_zeros(graph.num_dst, H)
    elif xh_dst is None:
        a_src, a_dst = _head_dot(xh_src, att_src, H, C, att_dst)         # both terms from one read of the features
    else:
        a_src = _head_dot(xh_src, att_src, H, C)
        a_dst = _head_dot(xh_dst, att_dst, H, C)
    r = Fn.attention("gat", graph, H, C, v=xh_src, s_src=a_src, s_dst=a_dst, s_edge=s_edge, negative_slope=negative_slope,
                     return_alpha=return_alpha, dropout_p=dropout_p)
    out, alpha = r if return_alpha else (r, None)
    out = _finish_heads(out, H, C, concat, res, bias)
    return (out, alpha) if return_alpha else out


def gatv2_conv(x_l: Tensor, x_r: Tensor, graph: CSRGraph, att: Tensor, H: int, C: int, negative_slope: float = 0.2,
               concat: bool = True, res: Optional[Tensor] = None, bias: Optional[Tensor] = None, return_alpha: bool = False,
               dropout_p: float = 0.0, e_feat: Optional[Tensor] = None):
    """GATv2Conv after lin_l / lin_r (gatv2_conv.py:300-331, 356-378): x_l [n_src, H*C], x_r [n_dst, H*C];
    e_feat [E, H*C] = lin_edge(edge_attr) aligned with the edges the graph was built from (edge_dim)."""
    r = Fn.attention("gatv2", graph, H, C, v=x_l, q=x_r, att=att.reshape(-1), negative_slope=negative_slope,
                     return_alpha=return_alpha, dropout_p=dropout_p, e_feat=e_feat)
    out, alpha = r if return_alpha else (r, None)
    out = _finish_heads(out, H, C, concat, res, bias)
    return (out, alpha) if return_alpha else out


def transformer_conv(query: Tensor, kv: Tensor, graph: CSRGraph, H: int, C: int, concat: bool = True,
                     x_skip: Optional[Tensor] = None, w_beta: Optional[Tensor] = None, return_alpha: bool = False,
                     dropout_p: float = 0.0, e_feat: Optional[Tensor] = None):
    """TransformerConv after its linear maps (transformer_conv.py:222-275): query [n_dst, H*C], kv [n_src, 2*H*C]
    (keys | values from ONE product with the concatenated lin_key / lin_value weights); x_skip = lin_skip(x_dst)."""
    r = Fn.attention("dot", graph, H, C, q=query, kv=kv, scale=1.0 / math.sqrt(C), return_alpha=return_alpha,
                     dropout_p=dropout_p, e_feat=e_feat)            # e_feat = lin_edge(edge_attr): added to key_j and value_j
    out, alpha = r if return_alpha else (r, None)
    if not concat:
        out = out.view(-1, H, C).mean(dim=1)
    if x_skip is not None:
        if w_beta is not None:
            beta = F.linear(torch.cat([out, x_skip, out - x_skip], dim=-1), w_beta).sigmoid()
            out = beta * x_skip + (1 - beta) * out
        else:
            out = out + x_skip
    return (out, alpha) if return_alpha else out


def rgcn_weight(weight: Tensor, comp: Optional[Tensor], num_relations: int, in_channels: int, out_channels: int,
                num_blocks: Optional[int]) -> Tensor:
    """The [R, F_in, F_out] relation weights from the basis / block-diagonal parametrisations (rgcn_conv.py:204-222)."""
    if comp is not None:                                               # basis decomposition
        return (comp @ weight.view(weight.size(0), -1)).view(num_relations, in_channels, out_channels)
    if num_blocks is not None:                                         # block-diagonal: weight [R, B, in/B, out/B]
        R, B, ci, co = weight.shape
        return torch.stack([torch.block_diag(*weight[r]) for r in range(R)])
    return weight


def rgcn_conv(x: Tensor, graph: CSRGraph, weight: Tensor, root: Optional[Tensor], bias: Optional[Tensor], aggr: str = "mean"):
    """RGCNConv with the per-relation semantics of the reference's loop path (rgcn_conv.py:257-280):
    out_i = sum_r aggr_{j in N_r(i)} x_j W_r + x_i root + bias.

    B200 mapping: edges are keyed by the virtual destination `dst * R + r`, so ONE gather-reduce sweep produces
    H [N, R*F_in] (per-relation mean/sum for every node) and the R small GEMMs of the reference collapse into ONE
    product with K = R * F_in against [W_1; ...; W_R] (+ the root product accumulated into the same output) -- a true
    GEMM, on the tcgen05 3xTF32 kernel (no cuBLAS on this path when the widths are supported)."""
    N = x.size(0)
    R, Fi, Fo = weight.shape
    h = Fn.aggregate(graph, x, aggr).view(N, R * Fi)                   # [N*R, F_in] -> [N, R*F_in]
    w = weight.reshape(R * Fi, Fo)
    if root is not None:
        return dense.matmul_pair(h, w, x, root, bias)
    out = dense.matmul(h, w)
    return out if bias is None else out + bias.to(out.dtype)


# ================================================================================================ standalone modules
def _src_dst(edge_index: Tensor, flow: str):
    return (edge_index[0], edge_index[1]) if flow == "source_to_target" else (edge_index[1], edge_index[0])


def _plain_graph(edge_index: Adj, num_src: int, num_dst: int, flow: str = "source_to_target") -> CSRGraph:
    if isinstance(edge_index, CSRGraph):
        return edge_index
    return cached_graph(edge_index, num_src, num_dst, flow=flow)


class GCNConv(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops: Optional[bool] = None, normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        if add_self_loops is None:
            add_self_loops = normalize
        if add_self_loops and not normalize:
            raise ValueError(f"'{self.__class__.__name__}' does not support adding self-loops to the graph when no "
                             f"on-the-fly normalization is applied")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self.flow = kwargs.get("flow", "source_to_target")
        self.lin = _Lin(in_channels, out_channels, bias=False)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels)) if bias else None
        self._cached_graph: Optional[CSRGraph] = None

    def reset_parameters(self):
        glorot_(self.lin.weight)
        if self.bias is not None:
            torch.nn.init.zeros_(self.bias)
        self._cached_graph = None

    def graph_for(self, edge_index: Adj, edge_weight: Optional[Tensor], num_nodes: int) -> CSRGraph:
        if isinstance(edge_index, CSRGraph):
            return edge_index
        if self._cached_graph is not None:
            return self._cached_graph
        if self.normalize:
            g = U.gcn_norm_graph(edge_index, edge_weight, num_nodes, self.improved, self.add_self_loops, self.flow)
        else:
            src, dst = _src_dst(edge_index, self.flow)
            g = CSRGraph(src, dst, num_nodes, num_nodes, edge_weight)
        if self.cached:
            self._cached_graph = g
        return g

    def forward(self, x: Tensor, edge_index: Adj, edge_weight: Optional[Tensor] = None) -> Tensor:
        if isinstance(x, (tuple, list)):
            raise ValueError(f"'{self.__class__.__name__}' received a tuple of node features as input while "
                             "this layer does not support bipartite message passing. Please try other layers "
                             "such as 'SAGEConv' or 'GraphConv' instead")
        return gcn_conv(x, self.graph_for(edge_index, edge_weight, x.size(0)), self.lin.weight, self.bias)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels})"


def _pair(x):
    return (x, x) if isinstance(x, Tensor) else (x[0], x[1])


def _num_dst(x, size):
    if x[1] is not None:
        return x[1].size(0)
    return size[1] if size is not None and size[1] is not None else x[0].size(0)


class SAGEConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels: int, aggr: str = "mean", normalize: bool = False,
                 root_weight: bool = True, project: bool = False, bias: bool = True, **kwargs):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        if aggr not in ("mean", "sum", "add", "max", "min"):
            raise ValueError(f"aggr='{aggr}' is not on the fused path")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.aggr, self.normalize, self.root_weight, self.project = aggr, normalize, root_weight, project
        self.flow = kwargs.get("flow", "source_to_target")
        if project:
            self.lin = _Lin(in_channels[0], in_channels[0], bias=True)
        self.lin_l = _Lin(in_channels[0], out_channels, bias=bias)
        if root_weight:
            self.lin_r = _Lin(in_channels[1], out_channels, bias=False)

    def forward(self, x, edge_index: Adj, size=None) -> Tensor:
        x = _pair(x)
        if self.project and hasattr(self, "lin"):
            x = (self.lin(x[0]).relu(), x[1])
        graph = _plain_graph(edge_index, x[0].size(0), _num_dst(x, size), self.flow)
        return sage_conv(x[0], x[1], graph, self.aggr, self.lin_l.weight, self.lin_l.bias,
                         self.lin_r.weight if self.root_weight else None, self.normalize)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, aggr={self.aggr})"


class GraphConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels: int, aggr: str = "add", bias: bool = True, **kwargs):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        if aggr not in ("mean", "sum", "add", "max", "min"):
            raise ValueError(f"aggr='{aggr}' is not on the fused path")
        self.in_channels, self.out_channels, self.aggr = in_channels, out_channels, aggr
        self.flow = kwargs.get("flow", "source_to_target")
        self.lin_rel = _Lin(in_channels[0], out_channels, bias=bias)
        self.lin_root = _Lin(in_channels[1], out_channels, bias=False)

    def forward(self, x, edge_index: Adj, edge_weight: Optional[Tensor] = None, size=None) -> Tensor:
        x = _pair(x)
        graph = _plain_graph(edge_index, x[0].size(0), _num_dst(x, size), self.flow)
        return graph_conv(x[0], x[1], graph, self.aggr, self.lin_rel.weight, self.lin_rel.bias, self.lin_root.weight,
                          edge_weight)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels[0]}, {self.out_channels})"


class GINConv(torch.nn.Module):
    def __init__(self, nn: torch.nn.Module, eps: float = 0.0, train_eps: bool = False, **kwargs):
        super().__init__()
        self.nn = nn
        self.initial_eps = eps
        self.flow = kwargs.get("flow", "source_to_target")
        if train_eps:
            self.eps = torch.nn.Parameter(torch.full((1, ), float(eps)))     # shape [1] as gin_conv.py:63-65
        else:
            self.register_buffer("eps", torch.full((1, ), float(eps)))

    def forward(self, x, edge_index: Adj, size=None) -> Tensor:
        x = _pair(x)
        graph = _plain_graph(edge_index, x[0].size(0), _num_dst(x, size), self.flow)
        return self.nn(gin_aggregate(x[0], x[1], graph, self.eps))

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(nn={self.nn})"


class RGCNConv(torch.nn.Module):
    """See `rgcn_conv`.  `num_bases` / `num_blocks` (rgcn_conv.py:140-160) are parametrisations of the same
    [R, F_in, F_out] weights and are expanded before the single K = R F_in (+ root) product."""

    def __init__(self, in_channels: int, out_channels: int, num_relations: int, num_bases: Optional[int] = None,
                 num_blocks: Optional[int] = None, aggr: str = "mean", root_weight: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        if num_bases is not None and num_blocks is not None:
            raise ValueError("Can not apply both basis-decomposition and block-diagonal-decomposition at the same time.")
        if aggr not in ("mean", "sum", "add", "max", "min"):
            raise ValueError(f"aggr='{aggr}' is not on the fused path")
        if isinstance(in_channels, (tuple, list)):
            in_channels = in_channels[0]
        self.in_channels, self.out_channels, self.num_relations, self.aggr = in_channels, out_channels, num_relations, aggr
        self.num_bases, self.num_blocks = num_bases, num_blocks
        if num_bases is not None:
            self.weight = torch.nn.Parameter(torch.empty(num_bases, in_channels, out_channels))
            self.comp = torch.nn.Parameter(torch.empty(num_relations, num_bases))
            glorot_(self.comp)
        elif num_blocks is not None:
            assert in_channels % num_blocks == 0 and out_channels % num_blocks == 0
            self.weight = torch.nn.Parameter(torch.empty(num_relations, num_blocks, in_channels // num_blocks,
                                                         out_channels // num_blocks))
            self.register_parameter("comp", None)
        else:
            self.weight = torch.nn.Parameter(torch.empty(num_relations, in_channels, out_channels))
            self.register_parameter("comp", None)
        self.root = torch.nn.Parameter(torch.empty(in_channels, out_channels)) if root_weight else None
        self.bias = torch.nn.Parameter(torch.zeros(out_channels)) if bias else None
        glorot_(self.weight)
        if self.root is not None:
            glorot_(self.root)

    def relation_graph(self, edge_index: Tensor, edge_type: Tensor, num_nodes: int) -> CSRGraph:
        return cached_graph(edge_index, num_nodes, num_nodes * self.num_relations, edge_type=edge_type,
                            num_relations=self.num_relations)

    def forward(self, x: Tensor, edge_index: Adj, edge_type: Optional[Tensor] = None) -> Tensor:
        if isinstance(edge_index, CSRGraph):
            graph = edge_index
        else:
            assert edge_type is not None
            graph = self.relation_graph(edge_index, edge_type, x.size(0))
        w = rgcn_weight(self.weight, self.comp, self.num_relations, self.in_channels, self.out_channels, self.num_blocks)
        return rgcn_conv(x, graph, w, self.root, self.bias, self.aggr)

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, "
                f"num_relations={self.num_relations})")


class FastRGCNConv(RGCNConv):
    """rgcn_conv.py:302-374 trades memory for speed with one [E, F_in] x W[edge_type] product per edge; the fused
    path above is already one sweep + one GEMM without any [E, *] tensor, so the fast variant IS the same code."""


def _attention_graph(edge_index: Adj, num_src: int, num_dst: int, add_self_loops: bool, flow: str = "source_to_target"):
    if isinstance(edge_index, CSRGraph):
        return edge_index
    return cached_graph(edge_index, num_src, num_dst, flow=flow, loops="gat" if add_self_loops else None,
                        loop_nodes=min(num_src, num_dst))


def edge_attr_with_loops(edge_index: Tensor, edge_attr: Tensor, num_nodes: int, fill_value, flow: str = "source_to_target"):
    """remove_self_loops + add_self_loops on the edge features, in the order of the graph built with loops='gat'
    (gat_conv.py:342-346; utils/loop.py:382-492: fill_value 'mean' = scatter-mean of the incoming edge features)."""
    if edge_attr.dim() == 1:
        edge_attr = edge_attr.view(-1, 1)
    keep = edge_index[0] != edge_index[1]
    ea = edge_attr[keep]
    dst = (edge_index[1] if flow == "source_to_target" else edge_index[0])[keep]
    if isinstance(fill_value, str):
        loop = U.scatter(ea.float(), dst, 0, num_nodes, fill_value).to(ea.dtype)
    elif isinstance(fill_value, Tensor):
        loop = fill_value.to(ea.dtype).view(1, -1).expand(num_nodes, ea.size(1))
    else:
        loop = ea.new_full((num_nodes, ea.size(1)), float(fill_value))
    return torch.cat([ea, loop], dim=0)


class GATConv(torch.nn.Module):
    """Mirror of torch_geometric.nn.GATConv (nn/conv/gat_conv.py:27-413) incl. bipartite inputs and `edge_dim`;
    attention + aggregation run in the fused kernel (csrc/attention.cu), attention dropout (training-time,
    gat_conv.py:404) included: the kept (edge, head) pairs come from a counter-based hash seeded from torch's CPU generator
    -- the same distribution as F.dropout, not the same random stream."""

    def __init__(self, in_channels, out_channels: int, heads: int = 1, concat: bool = True, negative_slope: float = 0.2,
                 dropout: float = 0.0, add_self_loops: bool = True, edge_dim: Optional[int] = None, fill_value="mean",
                 bias: bool = True, residual: bool = False, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negative_slope, self.dropout, self.add_self_loops, self.residual = negative_slope, dropout, add_self_loops, residual
        self.edge_dim, self.fill_value = edge_dim, fill_value
        self.flow = kwargs.get("flow", "source_to_target")
        self.lin = self.lin_src = self.lin_dst = None
        if isinstance(in_channels, int):
            self.lin = _Lin(in_channels, heads * out_channels, bias=False)
        else:
            self.lin_src = _Lin(in_channels[0], heads * out_channels, bias=False)
            self.lin_dst = _Lin(in_channels[1], heads * out_channels, bias=False)
        self.att_src = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        glorot_(self.att_src)
        glorot_(self.att_dst)
        if edge_dim is not None:
            self.lin_edge = _Lin(edge_dim, heads * out_channels, bias=False)
            self.att_edge = torch.nn.Parameter(torch.empty(1, heads, out_channels))
            glorot_(self.att_edge)
        else:
            self.lin_edge = None
            self.register_parameter("att_edge", None)
        total = heads * out_channels if concat else out_channels
        self.res = _Lin(in_channels if isinstance(in_channels, int) else in_channels[1], total, bias=False) if residual else None
        self.bias = torch.nn.Parameter(torch.zeros(total)) if bias else None

    def graph_for(self, edge_index: Adj, num_nodes: int, num_dst: Optional[int] = None) -> CSRGraph:
        return _attention_graph(edge_index, num_nodes, num_nodes if num_dst is None else num_dst, self.add_self_loops, self.flow)

    def forward(self, x, edge_index: Adj, edge_attr: Optional[Tensor] = None, size=None,
                return_attention_weights: Optional[bool] = None):
        drop = float(self.dropout) if self.training else 0.0        # attention dropout runs inside the sweep
        H, C = self.heads, self.out_channels
        att_dst = self.att_dst
        if isinstance(x, Tensor):
            assert x.dim() == 2, "Static graphs not supported in 'GATConv'"
            res = self.res(x) if self.res is not None else None
            if self.lin is not None:
                xh_src, xh_dst = self.lin(x), None
            else:
                xh_src, xh_dst = self.lin_src(x), self.lin_dst(x)
            n_src = n_dst = x.size(0)
        else:
            xs, xd = x
            assert xs.dim() == 2, "Static graphs not supported in 'GATConv'"
            res = self.res(xd) if (xd is not None and self.res is not None) else None
            lin_s, lin_d = (self.lin, self.lin) if self.lin is not None else (self.lin_src, self.lin_dst)
            xh_src, n_src = lin_s(xs), xs.size(0)
            if xd is not None:
                xh_dst, n_dst = lin_d(xd), xd.size(0)
            else:                                                        # alpha_i is absent (gat_conv.py:331)
                xh_dst, n_dst, att_dst = None, (size[1] if size is not None else n_src), None
        graph = self.graph_for(edge_index, n_src, n_dst)
        s_edge = None
        if edge_attr is not None and self.lin_edge is not None:
            if isinstance(edge_index, CSRGraph):
                raise NotImplementedError("edge_attr needs the [2, E] edge_index it is aligned with")
            ea = edge_attr if not self.add_self_loops else edge_attr_with_loops(
                edge_index, edge_attr, min(n_src, n_dst), self.fill_value, self.flow)
            if ea.dim() == 1:
                ea = ea.view(-1, 1)
            s_edge = _head_dot(self.lin_edge(ea), self.att_edge, H, C)
        want = return_attention_weights is not None
        r = gat_conv(xh_src, xh_dst, graph, self.att_src, att_dst, H, C, self.negative_slope, self.concat, res, self.bias,
                     s_edge, want, drop)
        if want:
            out, alpha = r
            ei = torch.stack([graph.col.long(), graph.dst_csr.long()])         # alpha is in the engine's CSR order
            return out, (ei, alpha)
        return r

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"


class GATv2Conv(torch.nn.Module):
    """Mirror of torch_geometric.nn.GATv2Conv (nn/conv/gatv2_conv.py:24-385) incl. `edge_dim`: the score
    att . leaky_relu(x_l[j] + x_r[i] (+ lin_edge(e_ij))), the edge softmax and the aggregation are ONE sweep (csrc/attention.cu)."""

    def __init__(self, in_channels, out_channels: int, heads: int = 1, concat: bool = True, negative_slope: float = 0.2,
                 dropout: float = 0.0, add_self_loops: bool = True, edge_dim: Optional[int] = None, fill_value="mean",
                 bias: bool = True, residual: bool = False, share_weights: bool = False, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negative_slope, self.dropout, self.add_self_loops = negative_slope, dropout, add_self_loops
        self.edge_dim, self.fill_value = edge_dim, fill_value
        self.residual, self.share_weights = residual, share_weights
        self.flow = kwargs.get("flow", "source_to_target")
        ic = (in_channels, in_channels) if isinstance(in_channels, int) else in_channels
        self.lin_l = _Lin(ic[0], heads * out_channels, bias=bias)
        self.lin_r = self.lin_l if share_weights else _Lin(ic[1], heads * out_channels, bias=bias)
        self.att = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        glorot_(self.att)
        self.lin_edge = _Lin(edge_dim, heads * out_channels, bias=False) if edge_dim is not None else None
        total = heads * out_channels if concat else out_channels
        self.res = _Lin(ic[1], total, bias=False) if residual else None
        self.bias = torch.nn.Parameter(torch.zeros(total)) if bias else None

    def forward(self, x, edge_index: Adj, edge_attr=None, return_attention_weights: Optional[bool] = None):
        drop = float(self.dropout) if self.training else 0.0        # attention dropout runs inside the sweep
        H, C = self.heads, self.out_channels
        if isinstance(x, Tensor):
            res = self.res(x) if self.res is not None else None
            x_l = self.lin_l(x)
            x_r = x_l if self.share_weights else self.lin_r(x)
        else:
            res = self.res(x[1]) if (x[1] is not None and self.res is not None) else None
            x_l = self.lin_l(x[0])
            x_r = self.lin_r(x[1])
        graph = _attention_graph(edge_index, x_l.size(0), x_r.size(0), self.add_self_loops, self.flow)
        e_feat = None
        if edge_attr is not None and self.lin_edge is not None:         # gatv2_conv.py:318-325, 358-360
            if isinstance(edge_index, CSRGraph):
                raise ValueError("edge_attr needs the [2, E] edge_index it is aligned with")
            ea = edge_attr
            if self.add_self_loops:
                ea = edge_attr_with_loops(edge_index, edge_attr, min(x_l.size(0), x_r.size(0)), self.fill_value, self.flow)
            e_feat = self.lin_edge(ea.view(-1, 1) if ea.dim() == 1 else ea)
        want = return_attention_weights is not None
        r = gatv2_conv(x_l, x_r, graph, self.att, H, C, self.negative_slope, self.concat, res, self.bias, want, drop, e_feat)
        if want:
            out, alpha = r
            return out, (torch.stack([graph.col.long(), graph.dst_csr.long()]), alpha)
        return r

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"


class TransformerConv(torch.nn.Module):
    """Mirror of torch_geometric.nn.TransformerConv (nn/conv/transformer_conv.py:17-285) incl. `edge_dim` (lin_edge(e_ij) added
    to the key and to the value): q.k / sqrt(C) scores, edge softmax and the value aggregation in ONE sweep; keys and values come from one GEMM
    with the concatenated lin_key / lin_value weights and are read as the two halves of one [N, 2HC] matrix."""

    def __init__(self, in_channels, out_channels: int, heads: int = 1, concat: bool = True, beta: bool = False,
                 dropout: float = 0.0, edge_dim: Optional[int] = None, bias: bool = True, root_weight: bool = True, **kwargs):
        super().__init__()
        self.edge_dim = edge_dim
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.beta, self.root_weight, self.dropout = beta and root_weight, root_weight, dropout
        self.flow = kwargs.get("flow", "source_to_target")
        ic = (in_channels, in_channels) if isinstance(in_channels, int) else in_channels
        hc = heads * out_channels
        self.lin_key = _Lin(ic[0], hc, bias=bias)
        self.lin_query = _Lin(ic[1], hc, bias=bias)
        self.lin_value = _Lin(ic[0], hc, bias=bias)
        self.lin_edge = _Lin(edge_dim, hc, bias=False) if edge_dim is not None else None
        total = hc if concat else out_channels
        self.lin_skip = _Lin(ic[1], total, bias=bias)
        self.lin_beta = _Lin(3 * total, 1, bias=False) if self.beta else None

    def forward(self, x, edge_index: Adj, edge_attr=None, return_attention_weights: Optional[bool] = None):
        drop = float(self.dropout) if self.training else 0.0        # attention dropout runs inside the sweep
        H, C = self.heads, self.out_channels
        x = _pair(x)
        e_feat = None
        if self.lin_edge is not None:                                   # transformer_conv.py:258-261
            assert edge_attr is not None
            if isinstance(edge_index, CSRGraph):
                raise ValueError("edge_attr needs the [2, E] edge_index it is aligned with")
            e_feat = self.lin_edge(edge_attr.view(-1, 1) if edge_attr.dim() == 1 else edge_attr)
        query = self.lin_query(x[1])
        w_kv = torch.cat([self.lin_key.weight, self.lin_value.weight], dim=0)
        b_kv = None if self.lin_key.bias is None else torch.cat([self.lin_key.bias, self.lin_value.bias], dim=0)
        kv = dense.linear(x[0], w_kv, b_kv)
        graph = _plain_graph(edge_index, x[0].size(0), x[1].size(0), self.flow)
        x_skip = self.lin_skip(x[1]) if self.root_weight else None
        want = isinstance(return_attention_weights, bool)
        r = transformer_conv(query, kv, graph, H, C, self.concat, x_skip, _w(self.lin_beta), want, drop, e_feat)
        if want:
            out, alpha = r
            return out, (torch.stack([graph.col.long(), graph.dst_csr.long()]), alpha)
        return r

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"


class HeteroLinear(torch.nn.Module):
    """Mirror of torch_geometric.nn.HeteroLinear (nn/dense/linear.py:174-340): x_k W_k + b_k per type k.  The per-type
    products are ONE launch of the grouped tcgen05 kernel (dense.segment_matmul); unsorted type vectors are sorted
    with the engine's stable radix sort and the result is un-permuted, as the reference does."""

    def __init__(self, in_channels: int, out_channels: int, num_types: int, is_sorted: bool = False, bias: bool = True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels, self.num_types, self.is_sorted = in_channels, out_channels, num_types, is_sorted
        self.weight = torch.nn.Parameter(torch.empty(num_types, in_channels, out_channels))
        self.bias = torch.nn.Parameter(torch.zeros(num_types, out_channels)) if bias else None
        bound = 1.0 / math.sqrt(in_channels)                          # reset_weight_ (linear.py:42-60): kaiming_uniform(a=sqrt(5))
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def forward(self, x: Tensor, type_vec: Tensor) -> Tensor:
        perm = None
        if not self.is_sorted:
            type_vec, perm = U.index_sort(type_vec, self.num_types - 1)
            x = ops.gather_rows(x, perm) if x.dtype in (torch.float32, torch.bfloat16) and not x.requires_grad else x[perm]
        ptr = ops.index2ptr(type_vec, self.num_types)
        out = dense.segment_matmul(x, ptr, self.weight)
        if self.bias is not None:
            out = out + self.bias[type_vec.long()]
        if perm is not None:
            out_unsorted = torch.empty_like(out)
            out_unsorted[perm] = out
            out = out_unsorted
        return out

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, num_types={self.num_types}, "
                f"bias={self.bias is not None})")

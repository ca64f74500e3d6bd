"""GNN layers on the fused aggregation path.  Constructor arguments, parameter names and
`state_dict` layout follow the reference layers so checkpoints are interchangeable:

  GCNConv   nn/conv/gcn_conv.py:116-274     lin.weight [out,in], bias [out]
  SAGEConv  nn/conv/sage_conv.py:19-156     lin_l.weight/bias, lin_r.weight
  GINConv   nn/conv/gin_conv.py:18-105      nn.*, eps
  RGCNConv  nn/conv/rgcn_conv.py:40-300     weight [R,in,out], root [in,out], bias [out]
  GATConv   nn/conv/gat_conv.py:27-413      lin.weight [H*C,in], att_src/att_dst [1,H,C], bias

`forward(x, edge_index, ...)` accepts either a `[2, E]` tensor (the graph structure is then built
on the fly, and kept if `cached=True`, cf. GCNConv.cached gcn_conv.py:150-158) or a prebuilt
`CSRGraph` -- the counterpart of handing the reference a `SparseTensor adj_t`.
The dense transforms are plain library GEMMs (torch.nn.functional.linear -> cuBLAS).
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
from ..graph import CSRGraph

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


class GCNConv(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops: bool = True, normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__()
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
            src, dst = (edge_index[0], edge_index[1]) if self.flow == "source_to_target" else (edge_index[1], edge_index[0])
            g = CSRGraph(src, dst, num_nodes, num_nodes, edge_weight)
        if self.cached:
            self._cached_graph = g
        return g

    def forward(self, x: Tensor, edge_index: Adj, edge_weight: Optional[Tensor] = None) -> Tensor:
        if isinstance(x, (tuple, list)):
            raise ValueError(f"'{self.__class__.__name__}' received a tuple of node features as input while "
                             "this layer does not support bipartite message passing. Please try other layers "
                             "such as 'SAGEConv' or 'GraphConv' instead")
        graph = self.graph_for(edge_index, edge_weight, x.size(0))
        xw = self.lin(x)
        if self.bias is not None:
            return _BiasAggregate.apply(xw, self.bias, graph)
        return Fn.aggregate(graph, xw, "sum")

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels})"


def _plain_graph(edge_index: Adj, num_src: int, num_dst: int, cache_owner=None) -> CSRGraph:
    if isinstance(edge_index, CSRGraph):
        return edge_index
    return CSRGraph(edge_index[0], edge_index[1], num_src, num_dst)


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
        if project:
            self.lin = _Lin(in_channels[0], in_channels[0], bias=True)
        self.lin_l = _Lin(in_channels[0], out_channels, bias=bias)
        if root_weight:
            self.lin_r = _Lin(in_channels[1], out_channels, bias=False)

    def forward(self, x, edge_index: Adj, size=None) -> Tensor:
        if isinstance(x, Tensor):
            x = (x, x)
        if self.project and hasattr(self, "lin"):
            x = (self.lin(x[0]).relu(), x[1])
        num_dst = x[1].size(0) if x[1] is not None else (size[1] if size is not None else x[0].size(0))
        graph = _plain_graph(edge_index, x[0].size(0), num_dst)
        out = Fn.aggregate(graph, x[0], self.aggr)          # sage_conv.py:134 propagate
        out = self.lin_l(out)
        if self.root_weight and x[1] is not None:
            out = out + self.lin_r(x[1])
        if self.normalize:
            out = F.normalize(out, p=2.0, dim=-1)
        return out

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, aggr={self.aggr})"


class GINConv(torch.nn.Module):
    def __init__(self, nn: torch.nn.Module, eps: float = 0.0, train_eps: bool = False, **kwargs):
        super().__init__()
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.full((1, ), float(eps)))     # shape [1] as gin_conv.py:63-65
        else:
            self.register_buffer("eps", torch.full((1, ), float(eps)))

    def forward(self, x, edge_index: Adj, size=None) -> Tensor:
        if isinstance(x, Tensor):
            x = (x, x)
        graph = _plain_graph(edge_index, x[0].size(0), x[1].size(0) if x[1] is not None else x[0].size(0))
        out = Fn.aggregate(graph, x[0], "sum")              # gin_conv.py:88
        if x[1] is not None:
            out = out + (1 + self.eps) * x[1]               # gin_conv.py:90-92
        return self.nn(out)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(nn={self.nn})"


class RGCNConv(torch.nn.Module):
    """Relational GCN with the per-relation semantics of the reference's loop path
    (rgcn_conv.py:257-280): out_i = sum_r aggr_{j in N_r(i)} x_j W_r + x_i root + bias.

    B200 mapping: edges are keyed by the virtual destination `dst * R + r`, so ONE gather-reduce
    sweep produces H [N, R*F_in] (per-relation mean/sum for every node) and the R small GEMMs of
    the reference collapse into ONE GEMM with K = R * F_in against weight.view(R*F_in, F_out).
    """

    def __init__(self, in_channels: int, out_channels: int, num_relations: int, aggr: str = "mean",
                 root_weight: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        if kwargs.get("num_bases") is not None or kwargs.get("num_blocks") is not None:
            raise NotImplementedError("basis / block-diagonal decomposition is not on the fused path")
        if aggr not in ("mean", "sum", "add", "max", "min"):
            raise ValueError(f"aggr='{aggr}' is not on the fused path")
        self.in_channels, self.out_channels, self.num_relations, self.aggr = in_channels, out_channels, num_relations, aggr
        self.weight = torch.nn.Parameter(torch.empty(num_relations, in_channels, out_channels))
        self.root = torch.nn.Parameter(torch.empty(in_channels, out_channels)) if root_weight else None
        self.bias = torch.nn.Parameter(torch.zeros(out_channels)) if bias else None
        glorot_(self.weight)
        if self.root is not None:
            glorot_(self.root)
        self._cached_graph = None

    def relation_graph(self, edge_index: Tensor, edge_type: Tensor, num_nodes: int) -> CSRGraph:
        R = self.num_relations
        vdst = edge_index[1].to(torch.int64) * R + edge_type.to(torch.int64)
        return CSRGraph(edge_index[0], vdst, num_nodes, num_nodes * R)

    def forward(self, x: Tensor, edge_index: Adj, edge_type: Optional[Tensor] = None) -> Tensor:
        N, R = x.size(0), self.num_relations
        if isinstance(edge_index, CSRGraph):
            graph = edge_index
        else:
            assert edge_type is not None
            graph = self.relation_graph(edge_index, edge_type, N)
        h = Fn.aggregate(graph, x, self.aggr)                              # [N*R, F_in]
        out = h.view(N, R * self.in_channels) @ self.weight.view(R * self.in_channels, self.out_channels).to(x.dtype)
        if self.root is not None:
            out = out + x @ self.root.to(x.dtype)
        if self.bias is not None:
            out = out + self.bias.to(x.dtype)
        return out

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, "
                f"num_relations={self.num_relations})")


class GATConv(torch.nn.Module):
    """Mirror of torch_geometric.nn.GATConv (nn/conv/gat_conv.py:27-413) for the homogeneous,
    edge_attr-free case; attention + aggregation run in the fused kernel (csrc/gat.cu).
    Attention dropout (training-time, gat_conv.py:405) is not fused: dropout must be 0."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                 bias: bool = True, residual: bool = False, **kwargs):
        super().__init__()
        if kwargs.get("edge_dim") is not None:
            raise NotImplementedError("edge_dim is not on the fused path")
        if dropout != 0.0:
            raise NotImplementedError("attention dropout is not fused; use dropout=0")
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negative_slope, self.dropout, self.add_self_loops, self.residual = negative_slope, dropout, add_self_loops, residual
        self.lin = _Lin(in_channels, heads * out_channels, bias=False)
        self.att_src = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        glorot_(self.att_src)
        glorot_(self.att_dst)
        total = heads * out_channels if concat else out_channels
        self.res = _Lin(in_channels, total, bias=False) if residual else None
        self.bias = torch.nn.Parameter(torch.zeros(total)) if bias else None

    def graph_for(self, edge_index: Adj, num_nodes: int) -> CSRGraph:
        if isinstance(edge_index, CSRGraph):
            return edge_index
        if self.add_self_loops:
            edge_index = U.remove_then_add_self_loops(edge_index, num_nodes)     # gat_conv.py:342-346
        return CSRGraph(edge_index[0], edge_index[1], num_nodes, num_nodes)

    def forward(self, x: Tensor, edge_index: Adj, return_attention_weights: Optional[bool] = None):
        H, C = self.heads, self.out_channels
        assert x.dim() == 2, "Static graphs not supported in 'GATConv'"
        graph = self.graph_for(edge_index, x.size(0))
        xh = self.lin(x)                                                          # [N, H*C]
        x3 = xh.view(-1, H, C)
        a_src = (x3 * self.att_src.to(xh.dtype)).sum(dim=-1)                      # gat_conv.py:330-331
        a_dst = (x3 * self.att_dst.to(xh.dtype)).sum(dim=-1)
        want = return_attention_weights is not None
        res = Fn.gat_attention(graph, xh, a_src.float(), a_dst.float(), H, C, self.negative_slope, want)
        out, alpha = res if want else (res, None)
        if not self.concat:
            out = out.view(-1, H, C).mean(dim=1)
        if self.res is not None:
            out = out + self.res(x)
        if self.bias is not None:
            out = out + self.bias.to(out.dtype)
        if want:
            # alpha is in the engine's CSR order; hand back the matching edge list
            ei = torch.stack([graph.col.long(), graph.dst_csr.long()])
            return out, (ei, alpha)
        return out

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"

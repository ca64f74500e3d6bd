"""Subclasses of the reference's OWN layer classes with the fused path inside (`class B200GCNConv(torch_geometric.nn.GCNConv)`).

Everything a user of the reference relies on is inherited unchanged -- constructor, parameters, `reset_parameters`,
`state_dict`, hooks, `explain`, `decomposed_layers`, jittable / TorchScript plumbing, CPU execution -- because the
object IS the reference layer.  Only `forward` is overridden: when the inputs are CUDA float32 / bfloat16 tensors and
the layer is in a mode the fused kernels implement, it runs the engine's functional core (`nn/conv.py`); otherwise it
calls `super().forward(...)`, i.e. the reference code (which, with `plugin.install()`, still lands in the engine
through the routed `scatter` / `softmax` / lazy gather).  `message_and_aggregate` is overridden too, with
`SUPPORTS_FUSED_EDGE_INDEX = True` (nn/conv/message_passing.py:108,475-497), so that `propagate` on a destination-sorted
`EdgeIndex` takes the fused branch with the EdgeIndex's cached CSR.

A fall-through happens for: CPU / other dtypes, `explain=True`, `decomposed_layers > 1`, any registered
propagate / message / aggregate hook (message_passing.py:776-922), `SparseTensor` / `torch.sparse` adjacencies,
and layer options the kernels do not cover.  Attention dropout (training mode) runs inside the fused sweep.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch_geometric.nn as tgnn
from torch import Tensor
from torch_geometric import EdgeIndex
from torch_geometric.typing import (Adj, NoneType, OptPairTensor, OptTensor, PairTensor, Size,  # noqa: F401
                                    SparseTensor)  # (names the inherited `# propagate_type:` annotations are evaluated with)

from .. import dense
from .. import utils as U
from ..graph import CSRGraph, cached_graph
from ..nn import conv as C
from . import graphs, routing
from ._util import plain

# reference class name -> subclass defined below.  The subclasses carry a `B200` prefix on purpose: the reference's
# Inspector caches class sources by `cls.__name__` (torch_geometric/inspector.py:323-334), so a subclass that reused
# its parent's name would hide the parent's `# propagate_type:` annotation and get a `propagate` without arguments.
LAYERS = {n: "B200" + n for n in ("GCNConv", "SAGEConv", "GraphConv", "GINConv", "GATConv", "GATv2Conv", "TransformerConv",
                                  "RGCNConv", "FastRGCNConv")}


def _has_hooks(self) -> bool:
    for name in ("_propagate_forward_pre_hooks", "_propagate_forward_hooks", "_message_forward_pre_hooks",
                 "_message_forward_hooks", "_aggregate_forward_pre_hooks", "_aggregate_forward_hooks",
                 "_message_and_aggregate_forward_pre_hooks", "_message_and_aggregate_forward_hooks",
                 "_edge_update_forward_pre_hooks", "_edge_update_forward_hooks"):
        if len(getattr(self, name, ())) > 0:
            return True
    return False


def _fast(self, *tensors) -> bool:
    if self.explain or self.decomposed_layers > 1 or _has_hooks(self) or routing._compiling() or torch.jit.is_scripting():
        return False
    return all(t is None or routing.engine_ok(t) for t in tensors)


_plain = plain


def _graph(edge_index, num_src: int, num_dst: int, flow: str, **kw) -> Optional[CSRGraph]:
    """CSRGraph for a [2, E] tensor (cached by identity), an EdgeIndex (its own cached CSR when sorted the right way)
    or a CSRGraph; None for SparseTensor / torch.sparse inputs (those stay with the reference)."""
    if isinstance(edge_index, CSRGraph):
        return edge_index
    if not isinstance(edge_index, Tensor) or edge_index.layout != torch.strided or edge_index.dim() != 2:
        return None
    if isinstance(edge_index, EdgeIndex) and not kw:
        if flow == "source_to_target" and edge_index.is_sorted_by_col:
            return graphs.graph_from_edge_index(edge_index, transpose=True)
        if flow == "target_to_source" and edge_index.is_sorted_by_row:
            return graphs.graph_from_edge_index(edge_index, transpose=False)
    return cached_graph(_plain(edge_index), num_src, num_dst, flow=flow, **kw)


def _pair(x):
    return (x, x) if isinstance(x, Tensor) else (x[0], x[1])


def _ndst(x, size):
    if x[1] is not None:
        return x[1].size(0)
    return size[1] if size is not None and size[1] is not None else x[0].size(0)


class _FusedEdgeIndexMixin:
    """`propagate(EdgeIndex sorted by destination, x=...)` -> `message_and_aggregate` -> the CSR kernel with the
    EdgeIndex's own cached structure (graph_conv.py:104-110 is the only reference layer that does this correctly)."""
    SUPPORTS_FUSED_EDGE_INDEX = True


class B200GCNConv(tgnn.GCNConv):
    def forward(self, x, edge_index, edge_weight: Optional[Tensor] = None) -> Tensor:
        if (isinstance(x, Tensor) and x.dim() == 2 and _fast(self, x) and (edge_weight is None or not edge_weight.requires_grad)):
            g = self._b200_graph(edge_index, edge_weight, x.size(0))
            if g is not None:
                return C.gcn_conv(x, g, self.lin.weight, self.bias)
        return super().forward(x, edge_index, edge_weight)

    def _b200_graph(self, edge_index, edge_weight, num_nodes: int) -> Optional[CSRGraph]:
        if isinstance(edge_index, CSRGraph):
            return edge_index
        if not isinstance(edge_index, Tensor) or edge_index.layout != torch.strided:
            return None
        cache = self.__dict__.get("_b200_cached_graph")
        if self.cached and cache is not None:
            return cache
        ei = _plain(edge_index)
        if self.normalize:
            g = U.gcn_norm_graph(ei, edge_weight, num_nodes, self.improved, self.add_self_loops, self.flow)
        else:
            g = cached_graph(ei, num_nodes, num_nodes, flow=self.flow)
            if edge_weight is not None:
                g = g.with_values(g.to_csr_order(edge_weight.detach().float()))
        if self.cached:
            self.__dict__["_b200_cached_graph"] = g
        return g

    def reset_parameters(self):
        super().reset_parameters()
        self.__dict__.pop("_b200_cached_graph", None)


class B200SAGEConv(_FusedEdgeIndexMixin, tgnn.SAGEConv):
    def forward(self, x, edge_index, size=None) -> Tensor:
        xs = _pair(x)
        if (isinstance(self.aggr, str) and self.aggr in ("mean", "sum", "add", "max", "min") and xs[0].dim() == 2
                and _fast(self, xs[0], xs[1])):
            g = _graph(edge_index, xs[0].size(0), _ndst(xs, size), self.flow)
            if g is not None:
                if self.project and hasattr(self, "lin"):
                    xs = (dense.linear(xs[0], self.lin.weight, self.lin.bias, relu=True), xs[1])
                return C.sage_conv(xs[0], xs[1], g, self.aggr, self.lin_l.weight, self.lin_l.bias,
                                   self.lin_r.weight if self.root_weight else None, self.normalize)
        return super().forward(x, edge_index, size)

    def message_and_aggregate(self, adj_t, x) -> Tensor:
        if isinstance(adj_t, EdgeIndex):
            return adj_t.matmul(other=x[0], reduce=self.aggr, transpose=True)
        return super().message_and_aggregate(adj_t, x)


class B200GraphConv(tgnn.GraphConv):
    def forward(self, x, edge_index, edge_weight: Optional[Tensor] = None, size=None) -> Tensor:
        xs = _pair(x)
        if (isinstance(self.aggr, str) and self.aggr in ("mean", "sum", "add", "max", "min") and xs[0].dim() == 2
                and _fast(self, xs[0], xs[1])):
            g = _graph(edge_index, xs[0].size(0), _ndst(xs, size), self.flow)
            if g is not None:
                return C.graph_conv(xs[0], xs[1], g, self.aggr, self.lin_rel.weight, self.lin_rel.bias, self.lin_root.weight,
                                    edge_weight)
        return super().forward(x, edge_index, edge_weight, size)


class B200GINConv(_FusedEdgeIndexMixin, tgnn.GINConv):
    def forward(self, x, edge_index, size=None) -> Tensor:
        xs = _pair(x)
        if xs[0].dim() == 2 and _fast(self, xs[0], xs[1]):
            g = _graph(edge_index, xs[0].size(0), _ndst(xs, size), self.flow)
            if g is not None:
                return self.nn(C.gin_aggregate(xs[0], xs[1], g, self.eps))
        return super().forward(x, edge_index, size)

    def message_and_aggregate(self, adj_t, x) -> Tensor:
        if isinstance(adj_t, EdgeIndex):
            return adj_t.matmul(other=x[0], reduce=self.aggr, transpose=True)
        return super().message_and_aggregate(adj_t, x)


def _attn_fast(self, *tensors) -> bool:
    return _fast(self, *tensors)


def _drop(self) -> float:
    """Attention dropout runs inside the fused sweep (same distribution as F.dropout, its own random stream)."""
    return float(self.dropout) if self.training else 0.0


class B200GATConv(tgnn.GATConv):
    def forward(self, x, edge_index, edge_attr=None, size=None, return_attention_weights=None):
        xs = _pair(x)
        if (xs[0].dim() == 2 and _attn_fast(self, xs[0], xs[1], edge_attr) and return_attention_weights is None
                and isinstance(edge_index, Tensor) and edge_index.layout == torch.strided and size is None):
            H, Cc = self.heads, self.out_channels
            lin_s, lin_d = (self.lin, self.lin) if self.lin is not None else (self.lin_src, self.lin_dst)
            same = isinstance(x, Tensor)
            res = None
            if self.res is not None and xs[1] is not None:
                res = dense.linear(xs[1], self.res.weight)
            xh_src = dense.linear(xs[0], lin_s.weight)
            xh_dst = None if (same and self.lin is not None) else (None if xs[1] is None else dense.linear(xs[1], lin_d.weight))
            n_src, n_dst = xs[0].size(0), (xs[1].size(0) if xs[1] is not None else xs[0].size(0))
            ei = _plain(edge_index)
            g = cached_graph(ei, n_src, n_dst, flow=self.flow, loops="gat" if self.add_self_loops else None,
                             loop_nodes=min(n_src, n_dst))
            s_edge = None
            if edge_attr is not None and self.lin_edge is not None:
                ea = edge_attr if not self.add_self_loops else C.edge_attr_with_loops(ei, edge_attr, min(n_src, n_dst),
                                                                                      self.fill_value, self.flow)
                if ea.dim() == 1:
                    ea = ea.view(-1, 1)
                s_edge = C._head_dot(dense.linear(ea, self.lin_edge.weight), self.att_edge, H, Cc)
            att_dst = self.att_dst if (xs[1] is not None) else None
            return C.gat_conv(xh_src, xh_dst, g, self.att_src, att_dst, H, Cc, self.negative_slope, self.concat, res,
                              self.bias, s_edge, False, _drop(self))
        return super().forward(x, edge_index, edge_attr, size, return_attention_weights)


class B200GATv2Conv(tgnn.GATv2Conv):
    def forward(self, x, edge_index, edge_attr=None, return_attention_weights=None):
        xs = _pair(x)
        if (xs[0].dim() == 2 and xs[1] is not None and _attn_fast(self, xs[0], xs[1], edge_attr)
                and (edge_attr is None) == (self.lin_edge is None)
                and return_attention_weights is None and isinstance(edge_index, Tensor) and edge_index.layout == torch.strided):
            H, Cc = self.heads, self.out_channels
            res = dense.linear(xs[1], self.res.weight) if self.res is not None else None
            x_l = dense.linear(xs[0], self.lin_l.weight, self.lin_l.bias)
            if self.share_weights and isinstance(x, Tensor):
                x_r = x_l
            else:
                x_r = dense.linear(xs[1], self.lin_r.weight, self.lin_r.bias)
            ei = _plain(edge_index)
            g = cached_graph(ei, x_l.size(0), x_r.size(0), flow=self.flow,
                             loops="gat" if self.add_self_loops else None, loop_nodes=min(x_l.size(0), x_r.size(0)))
            e_feat = None
            if edge_attr is not None:                                           # gatv2_conv.py:318-325, 358-360
                ea = edge_attr if not self.add_self_loops else C.edge_attr_with_loops(
                    ei, edge_attr, min(x_l.size(0), x_r.size(0)), self.fill_value, self.flow)
                e_feat = dense.linear(ea.view(-1, 1) if ea.dim() == 1 else ea, self.lin_edge.weight)
            return C.gatv2_conv(x_l, x_r, g, self.att, H, Cc, self.negative_slope, self.concat, res, self.bias, False,
                                _drop(self), e_feat)
        return super().forward(x, edge_index, edge_attr, return_attention_weights)


class B200TransformerConv(tgnn.TransformerConv):
    def forward(self, x, edge_index, edge_attr=None, return_attention_weights=None):
        xs = _pair(x)
        if (xs[0].dim() == 2 and xs[1] is not None and _attn_fast(self, xs[0], xs[1], edge_attr)
                and (edge_attr is None) == (self.lin_edge is None) and return_attention_weights is None
                and isinstance(edge_index, Tensor) and edge_index.layout == torch.strided):
            H, Cc = self.heads, self.out_channels
            query = dense.linear(xs[1], self.lin_query.weight, self.lin_query.bias)
            w_kv = torch.cat([self.lin_key.weight, self.lin_value.weight], dim=0)
            b_kv = None if self.lin_key.bias is None else torch.cat([self.lin_key.bias, self.lin_value.bias], dim=0)
            kv = dense.linear(xs[0], w_kv, b_kv)
            g = _graph(edge_index, xs[0].size(0), xs[1].size(0), self.flow)
            x_skip = dense.linear(xs[1], self.lin_skip.weight, self.lin_skip.bias) if self.root_weight else None
            w_beta = self.lin_beta.weight if self.lin_beta is not None else None
            e_feat = None
            if edge_attr is not None:                                           # transformer_conv.py:258-261
                e_feat = dense.linear(edge_attr.view(-1, 1) if edge_attr.dim() == 1 else edge_attr, self.lin_edge.weight)
            return C.transformer_conv(query, kv, g, H, Cc, self.concat, x_skip, w_beta, False, _drop(self), e_feat)
        return super().forward(x, edge_index, edge_attr, return_attention_weights)


class _RGCNMixin:
    def forward(self, x, edge_index, edge_type=None) -> Tensor:
        if (isinstance(x, Tensor) and x.dim() == 2 and x.is_floating_point() and _fast(self, x) and edge_type is not None
                and isinstance(edge_index, Tensor) and edge_index.layout == torch.strided
                and isinstance(self.aggr, str) and self.aggr in ("mean", "sum", "add", "max", "min")):
            g = cached_graph(_plain(edge_index), x.size(0), x.size(0) * self.num_relations, flow=self.flow,
                             edge_type=edge_type, num_relations=self.num_relations)
            w = C.rgcn_weight(self.weight, getattr(self, "comp", None) if self.num_bases is not None else None,
                              self.num_relations, self.in_channels_l, self.out_channels, self.num_blocks)
            return C.rgcn_conv(x, g, w, self.root, self.bias, self.aggr)
        return super().forward(x, edge_index, edge_type)


class B200RGCNConv(_RGCNMixin, tgnn.RGCNConv):
    pass


class B200FastRGCNConv(_RGCNMixin, tgnn.FastRGCNConv):
    pass

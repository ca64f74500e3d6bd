"""The mini-batch side of the path: what decides whether a sampled subgraph reaches
the CSR kernels or the atomic COO fallback is the FORMAT it arrives in.

`NeighborLoader` / `neighbor_sample` emit the sampled edges hop by hop in BFS order, each hop grouped by destination
(`loader/utils.py:108-133`: "we expect (row, col) to be sorted by col (CSC layout)"; the `EdgeIndex` integration is a
TODO there) -- i.e. the edge list of a sampled subgraph is ALREADY destination-sorted.  So:

  sampled_graph(edge_index, num_src, num_dst)   adopts it as a CSR with one `index2ptr` pass: no sort, no host sync
                                                (`bounded_degree=True`: fan-out bounded rows need no long-row plan)
  trim_to_layer(layer, nodes_per_hop, edges_per_hop, x, graph | edge_index, edge_attr)
                                                mirror of utils/_trim_to_layer.py:20-217; for a CSRGraph the trimmed
                                                layer graph is a pair of prefix VIEWS (`CSRGraph.trim`)
  coalesce(edge_index, edge_attr, num_nodes, reduce, is_sorted, sort_by_row)
                                                mirror of utils/_coalesce.py:131-175 on the engine's stable radix sort
                                                and segmented reduce (duplicates merged by a CSR sweep, not by atomics)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch
from torch import Tensor

from . import _debug
from . import functional as Fn
from . import ops
from .graph import CSRGraph


def sampled_graph(edge_index: Tensor, num_src: int, num_dst: int, edge_weight: Optional[Tensor] = None,
                  bounded_degree: bool = True) -> CSRGraph:
    """CSRGraph of a neighbour-sampled subgraph whose edges are grouped by destination (edge_index[1] non-decreasing),
    adopted without sorting.  In debug mode (`pytorch_geometric_b200.debug()`) the sortedness is verified."""
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must have shape [2, E]")
    dst = edge_index[1].contiguous()
    if _debug.enabled() and dst.numel():
        mn, mx, srt = ops.index_stats(dst)
        if not srt:
            raise ValueError("sampled_graph: edge_index[1] is not sorted; use CSRGraph(...) (one stable sort) instead")
        if mn < 0 or mx >= num_dst:
            raise IndexError(f"Found indices in 'edge_index' outside [0, {num_dst})")
    ptr = ops.index2ptr(dst, num_dst)
    return CSRGraph.from_csr(ptr, edge_index[0], num_src, edge_weight, bounded_degree=bounded_degree)


def _trim_feat(x: Tensor, layer: int, per_hop: List[int]) -> Tensor:
    return x if layer <= 0 else x.narrow(0, 0, x.size(0) - per_hop[-layer])


def _trim_adj(adj, layer: int, src_per_hop: List[int], dst_per_hop: List[int], edges_per_hop: List[int]):
    if layer <= 0:
        return adj
    if isinstance(adj, CSRGraph):
        return adj.trim(adj.num_dst - dst_per_hop[-layer], adj.num_src - src_per_hop[-layer],
                        adj.num_edges - edges_per_hop[-layer])
    if isinstance(adj, Tensor):
        return adj.narrow(1, 0, adj.size(1) - edges_per_hop[-layer])
    raise ValueError(f"Unsupported 'edge_index' type '{type(adj)}'")


def trim_to_layer(layer: int, num_sampled_nodes_per_hop: Union[List[int], Dict], num_sampled_edges_per_hop: Union[List[int], Dict],
                  x, edge_index, edge_attr=None):
    """Mirror of torch_geometric.utils.trim_to_layer (homogeneous lists or heterogeneous dicts); `edge_index` entries
    may be `[2, E]` tensors or `CSRGraph`s (then the result is a view-trimmed CSRGraph)."""
    if layer <= 0:
        return x, edge_index, edge_attr
    if isinstance(num_sampled_edges_per_hop, dict):
        assert isinstance(num_sampled_nodes_per_hop, dict) and isinstance(x, dict) and isinstance(edge_index, dict)
        x = {k: _trim_feat(v, layer, num_sampled_nodes_per_hop[k]) for k, v in x.items()}
        edge_index = {k: _trim_adj(v, layer, num_sampled_nodes_per_hop[k[0]], num_sampled_nodes_per_hop[k[-1]],
                                   num_sampled_edges_per_hop[k]) for k, v in edge_index.items()}
        if edge_attr is not None:
            edge_attr = {k: _trim_feat(v, layer, num_sampled_edges_per_hop[k]) for k, v in edge_attr.items()}
        return x, edge_index, edge_attr
    x = _trim_feat(x, layer, num_sampled_nodes_per_hop)
    edge_index = _trim_adj(edge_index, layer, num_sampled_nodes_per_hop, num_sampled_nodes_per_hop, num_sampled_edges_per_hop)
    if edge_attr is not None:
        edge_attr = _trim_feat(edge_attr, layer, num_sampled_edges_per_hop)
    return x, edge_index, edge_attr


def coalesce(edge_index: Tensor, edge_attr: Optional[Tensor] = None, num_nodes: Optional[int] = None, reduce: str = "sum",
             is_sorted: bool = False, sort_by_row: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
    """Mirror of torch_geometric.utils.coalesce: row-wise (or column-wise) sorts the edges and merges duplicates,
    reducing their attributes.  Stable radix sort by the composite key + a segmented reduce over the runs of equal
    keys (deterministic; the reference scatters with atomics).  Returns (edge_index, edge_attr)."""
    if not edge_index.is_cuda:
        raise RuntimeError("pytorch_geometric_b200 ops run on CUDA tensors only (no CPU fallback)")
    if num_nodes is None:
        num_nodes = (ops.index_stats(edge_index.reshape(-1))[1] + 1) if edge_index.numel() else 0
    if num_nodes * num_nodes > 2**63 - 1:
        raise ValueError("'coalesce' will result in an overflow")
    E = edge_index.size(1)
    major, minor = (edge_index[0], edge_index[1]) if sort_by_row else (edge_index[1], edge_index[0])
    key = major.to(torch.int64) * num_nodes + minor.to(torch.int64)
    if not is_sorted and E:
        key, perm, _ = ops.sort_by_key(key, num_nodes * num_nodes, want_sorted=True, want_ptr=False)
        perm = perm.long()
        edge_index = edge_index[:, perm]
        if edge_attr is not None:
            edge_attr = edge_attr[perm]
    if E == 0:
        return edge_index, edge_attr
    first = torch.ones(E, dtype=torch.bool, device=key.device)
    first[1:] = key[1:] != key[:-1]
    n_unique = int(first.sum())                                 # the reference's `mask.all()` is the same host read
    if n_unique == E:
        return edge_index, edge_attr
    out_index = edge_index[:, first]
    if edge_attr is None:
        return out_index, None
    starts = torch.nonzero(first).view(-1)
    ptr = torch.cat([starts, starts.new_tensor([E])])
    if reduce in ("sum", "add", "mean", "min", "max") and edge_attr.dtype in (torch.float32, torch.bfloat16):
        merged = Fn.segment(edge_attr, ptr, "sum" if reduce == "add" else reduce)
    else:
        seg = torch.cumsum(first.long(), 0) - 1
        merged = torch.zeros((n_unique, ) + tuple(edge_attr.shape[1:]), dtype=edge_attr.dtype, device=edge_attr.device)
        merged = merged.scatter_reduce(0, seg.view((-1, ) + (1, ) * (edge_attr.dim() - 1)).expand_as(edge_attr), edge_attr,
                                       {"sum": "sum", "add": "sum", "mean": "mean", "min": "amin", "max": "amax", "mul": "prod"}[reduce],
                                       include_self=False)
    return out_index, merged

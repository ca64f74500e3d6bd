"""CSRGraph handles for the reference's own graph containers, WITHOUT re-sorting what they already cache.

  EdgeIndex  (torch_geometric/edge_index.py:237-246, 589-696)  sort_order + `_indptr`, `_T_perm`, `_T_index`, `_T_indptr`
  Index      (torch_geometric/index.py:244-299)                `is_sorted` + `_indptr` (what `edge_index[i]` returns)
  torch.sparse_csr tensors                                      crow_indices / col_indices / values
  plain (src ids, dst ids) pairs                                one stable sort, cached by tensor identity

The engine keeps its own int32 copy of the index arrays (12 B/edge instead of 24) -- a streaming conversion, not a
sort -- and hangs the handle on the EdgeIndex object (`_b200_graphs`), so every later call is free.
"""
from __future__ import annotations

import weakref
from collections import OrderedDict
from typing import Optional

import torch
from torch import Tensor

from ..graph import CSRGraph
from ._util import plain

_PAIR_CACHE: "OrderedDict" = OrderedDict()     # (ptr_j, ptr_i, E, n_src, n_dst) -> (weak j, weak i, versions, graph)
_PAIR_CACHE_EDGES = 600_000_000                # evict least-recently-used graphs above this many cached edges


def graph_from_edge_index(ei, transpose: bool) -> CSRGraph:
    """The CSR that `EdgeIndex.matmul(other, transpose=...)` multiplies with (edge_index.py:1925-1970):
    transpose=False: rows = ei[0] (needs sort_order 'row'), out[r] = sum A[r,c] other[c];
    transpose=True:  rows = ei[1] (needs sort_order 'col'), out[c] = sum A[r,c] other[r] -- the message-passing case
    (flow source_to_target aggregates at edge_index[1])."""
    cache = ei.__dict__.setdefault("_b200_graphs", {})
    g = cache.get(transpose)
    if g is not None:
        return g
    if not transpose:
        (rowptr, col), _ = ei.get_csr()          # sorted by row: perm is None
        num_src = ei.get_sparse_size(1)
        t_cached = ei._T_perm is not None
        t = ei.get_csc() if t_cached else None   # ((colptr, row), perm): the transposed structure, if already cached
    else:
        (rowptr, col), _ = ei.get_csc()
        num_src = ei.get_sparse_size(0)
        t_cached = ei._T_perm is not None
        t = ei.get_csr() if t_cached else None
    transposed = None
    if t is not None and t[1] is not None:
        (ptr_t, idx_t), perm_t = t
        transposed = (ptr_t, idx_t, perm_t)
    g = CSRGraph.from_csr(plain(rowptr), plain(col), num_src,
                          transposed=None if transposed is None else tuple(plain(t) for t in transposed))
    cache[transpose] = g
    return g


def graph_from_sparse_csr(adj: Tensor) -> CSRGraph:
    """torch.sparse_csr adjacency whose rows are destinations (the `adj_t` convention of utils/_spmm.py:12-136)."""
    crow, col = adj.crow_indices(), adj.col_indices()
    key = ("csr", crow.data_ptr(), col.data_ptr(), col.numel(), adj.size(0), adj.size(1))
    hit = _PAIR_CACHE.get(key)
    if hit is not None:
        _PAIR_CACHE.move_to_end(key)
        return hit[3]
    g = CSRGraph.from_csr(crow, col, adj.size(1))
    _remember(key, (crow, col, None, g))
    return g


def _remember(key, entry) -> None:
    _PAIR_CACHE[key] = entry
    total = sum(e[3].num_edges for e in _PAIR_CACHE.values())
    while total > _PAIR_CACHE_EDGES and len(_PAIR_CACHE) > 1:
        _, old = _PAIR_CACHE.popitem(last=False)
        total -= old[3].num_edges


def graph_from_pair(index_j: Tensor, index_i: Tensor, num_src: int, num_dst: int, ptr: Optional[Tensor] = None) -> CSRGraph:
    """CSR for messages j -> i given the two aligned index vectors a MessagePassing layer collects
    (`edge_index_j`, `edge_index_i`, collect.jinja:67-75).  With `ptr` (or a sorted `Index` carrying its indptr) the
    pair already IS a CSR and nothing is sorted; otherwise one stable sort by destination.  Cached by the identity
    of the two tensors (storage pointer, length, version), the strong references living in the cache entry."""
    if ptr is None and getattr(index_i, "is_sorted", False) and getattr(index_i, "_indptr", None) is not None:
        ptr = index_i._indptr
    index_j, index_i, ptr = plain(index_j), plain(index_i), plain(ptr)
    key = ("pair", index_j.data_ptr(), index_i.data_ptr(), index_i.numel(), int(num_src), int(num_dst), ptr is not None)
    hit = _PAIR_CACHE.get(key)
    if hit is not None and hit[2] == (index_j._version, index_i._version):
        _PAIR_CACHE.move_to_end(key)
        return hit[3]
    if ptr is not None and ptr.numel() == num_dst + 1:
        g = CSRGraph.from_csr(ptr, index_j, num_src)
    else:
        g = CSRGraph(index_j, index_i, num_src, num_dst)
    _remember(key, (index_j, index_i, (index_j._version, index_i._version), g))
    return g


def clear_cache() -> None:
    _PAIR_CACHE.clear()

"""Call-time routing of the reference's functional seams to the engine.

Every routed function keeps the reference's signature, validation and error messages.  Routing rule: the engine takes
a call iff the feature tensor is a CUDA tensor of dtype float32 / bfloat16 and the reduction is one it implements with
the reference's semantics; everything else -- CPU tensors, float64 / half / integer data, reduce='any' (whose CUDA result is unspecified in the reference itself) -- falls through to the UNTOUCHED reference function (`__wrapped__`), which is also how the parity oracle
keeps working next to the engine.  Nothing is ever silently computed on the CPU by the engine.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from .. import functional as Fn
from .. import ops
from .. import utils as U
from ..graph import CSRGraph
from . import graphs
from ._util import plain
from .lazy import LazyRows

_ENGINE_DTYPES = (torch.float32, torch.bfloat16)
_REDUCE = {"sum": "sum", "add": "sum", "mean": "mean", "min": "min", "amin": "min", "max": "max", "amax": "max",
           "mul": "mul"}


def engine_ok(t) -> bool:
    return isinstance(t, Tensor) and t.is_cuda and t.dtype in _ENGINE_DTYPES


def _compiling() -> bool:
    try:
        return torch.compiler.is_compiling()
    except Exception:
        return False


_plain = plain


def _sorted_ptr(index, dim_size: Optional[int]):
    """indptr of a sorted `Index` (index.py:244-299), or None: the only sortedness evidence the reference carries."""
    if not getattr(index, "is_sorted", False):
        return None
    try:
        ptr = index.get_indptr()
    except Exception:
        return None
    if dim_size is not None and ptr.numel() != dim_size + 1:
        return None
    return _plain(ptr)


# ------------------------------------------------------------------------------------------------ fused gather + reduce
def fused_lazy_reduce(lazy: LazyRows, index: Tensor, ptr: Optional[Tensor], dim_size: Optional[int], reduce: str):
    """aggregate(LazyRows(x, index_j[, w]), index_i) == one CSR gather-reduce; None when not applicable."""
    r = _REDUCE.get(reduce)
    src = lazy._src
    if r is None or r == "mul" or not engine_ok(src) or index is None:
        return None
    if lazy._scale is not None and r in ("min", "max") and lazy._scale.requires_grad:
        return None
    if dim_size is None:
        dim_size = getattr(index, "dim_size", None)
        if dim_size is None:
            return None
    g = graphs.graph_from_pair(lazy._index, index, src.size(0), int(dim_size), ptr=ptr)
    x2 = src if src.dim() == 2 else src.reshape(src.size(0), -1)
    w = lazy._scale
    if w is not None and w.dtype != torch.float32:
        w = w.float()
    out = Fn.aggregate(g, x2, r, w)
    return out if src.dim() == 2 else out.view((int(dim_size), ) + tuple(src.shape[1:]))


# ------------------------------------------------------------------------------------------------ utils.scatter
def make_scatter(theirs):
    def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None, reduce: str = "sum") -> Tensor:
        if isinstance(src, LazyRows):
            d = dim + src.dim() if dim < 0 else dim
            if d == 0 and not _compiling():
                out = fused_lazy_reduce(src, index, None, dim_size, reduce)
                if out is not None:
                    return out
            src = src.materialise()
        r = _REDUCE.get(reduce)
        if (not engine_ok(src) or _compiling() or r is None
                or not isinstance(index, Tensor) or index.dim() != 1):
            return theirs(src, index, dim, dim_size, reduce)           # incl. every argument error of the reference
        d = src.dim() + dim if dim < 0 else dim
        if d < 0 or d >= src.dim():
            return theirs(src, index, dim, dim_size, reduce)
        ptr = _sorted_ptr(index, dim_size) if r != "mul" else None
        if ptr is not None:                                            # sorted Index: deterministic CSR kernel, no atomics
            x = src if d == 0 else src.movedim(d, 0).contiguous()
            out = Fn.segment(x, ptr, r)
            return out if d == 0 else out.movedim(0, d)
        return U.scatter(src, _plain(index), d, dim_size, r)
    scatter.__wrapped__ = theirs
    scatter.__name__, scatter.__doc__ = "scatter", theirs.__doc__
    return scatter


def make_segment(theirs):
    def segment(src: Tensor, ptr: Tensor, reduce: str = "sum") -> Tensor:
        if isinstance(src, LazyRows):
            src = src.materialise()
        if not engine_ok(src) or _compiling() or ptr.dim() != 1 or reduce not in ("sum", "mean", "min", "max"):
            return theirs(src, ptr, reduce)
        return Fn.segment(src, _plain(ptr), reduce)
    segment.__wrapped__ = theirs
    segment.__name__, segment.__doc__ = "segment", theirs.__doc__
    return segment


def make_softmax(theirs):
    def softmax(src: Tensor, index: Optional[Tensor] = None, ptr: Optional[Tensor] = None,
                num_nodes: Optional[int] = None, dim: int = 0) -> Tensor:
        if isinstance(src, LazyRows):
            src = src.materialise()
        if not engine_ok(src) or _compiling() or (index is None and ptr is None) or (ptr is not None and ptr.dim() != 1):
            return theirs(src, index, ptr, num_nodes, dim)
        if ptr is None:
            ptr = _sorted_ptr(index, num_nodes)
        out = U.softmax(src, None if ptr is not None else _plain(index), None if ptr is None else _plain(ptr), num_nodes, dim)
        return out.to(src.dtype)                                       # the reference returns src's dtype
    softmax.__wrapped__ = theirs
    softmax.__name__, softmax.__doc__ = "softmax", theirs.__doc__
    return softmax


# ------------------------------------------------------------------------------------------------ spmm / EdgeIndex.matmul
def make_spmm(theirs):
    def spmm(src, other: Tensor, reduce: str = "sum") -> Tensor:
        r = "sum" if reduce == "add" else reduce
        if isinstance(src, CSRGraph):
            return U.spmm(src, other, reduce)
        if (engine_ok(other) and not _compiling() and isinstance(src, Tensor) and src.layout == torch.sparse_csr
                and r in ("sum", "mean", "min", "max") and other.dim() == 2 and src.dim() == 2):
            val = src.values()
            if val.dim() == 1 and val.dtype in _ENGINE_DTYPES + (torch.float64, ) and not val.requires_grad:
                g = graphs.graph_from_sparse_csr(src)
                return Fn.aggregate(g, other, r, val.float())          # differentiable wrt `other` (training included)
        return theirs(src, other, reduce)                              # EdgeIndex inputs reach edge_index._spmm below
    spmm.__wrapped__ = theirs
    spmm.__name__, spmm.__doc__ = "spmm", theirs.__doc__
    return spmm


def make_edge_index_spmm(theirs):
    """edge_index._spmm (edge_index.py:1925-1970): CUDA operands go to the CSR kernel with the EdgeIndex's own cached
    structure (no re-sort), forward and backward, all four reductions, value gradients included."""
    def _spmm(input, other: Tensor, value: Optional[Tensor] = None, reduce: str = "sum", transpose: bool = False) -> Tensor:
        r = "sum" if reduce == "add" else reduce
        if (not engine_ok(other) or _compiling() or other.dim() != 2 or r not in ("sum", "mean", "min", "max")
                or (value is not None and (value.dim() != 1 or not value.is_floating_point()))
                or (value is not None and value.requires_grad and r not in ("sum", "mean"))):
            return theirs(input, other, value, reduce, transpose)
        if (not transpose and not input.is_sorted_by_row) or (transpose and not input.is_sorted_by_col):
            return theirs(input, other, value, reduce, transpose)      # raises the reference's ValueError
        g = graphs.graph_from_edge_index(input, transpose)
        return Fn.aggregate(g, other, r, value)
    _spmm.__wrapped__ = theirs
    return _spmm


# ------------------------------------------------------------------------------------------------ Aggregation.reduce
def make_aggr_reduce(theirs):
    """nn/aggr/base.py:173-185.  The reference ignores `ptr` unless deterministic mode is on; the engine uses it
    whenever it is there (same result, test/nn/aggr/test_basic.py:63) because the CSR kernel is the deterministic,
    atomics-free path -- and this is where a LazyRows message meets its destination index."""
    def reduce(self, x: Tensor, index: Optional[Tensor] = None, ptr: Optional[Tensor] = None,
               dim_size: Optional[int] = None, dim: int = -2, reduce: str = "sum") -> Tensor:
        d = dim + x.dim() if dim < 0 else dim
        if isinstance(x, LazyRows):
            if d == 0 and not _compiling() and index is not None:
                out = fused_lazy_reduce(x, index, None if ptr is None else _plain(ptr), dim_size, reduce)
                if out is not None:
                    return out
            x = x.materialise()
        if engine_ok(x) and not _compiling() and ptr is not None and ptr.dim() == 1 and reduce in ("sum", "mean", "min", "max"):
            xm = x if d == 0 else x.movedim(d, 0).contiguous()
            out = Fn.segment(xm, _plain(ptr), reduce)
            return out if d == 0 else out.movedim(0, d)
        return theirs(self, x, index, ptr, dim_size, dim, reduce)
    reduce.__wrapped__ = theirs
    return reduce


# ------------------------------------------------------------------------------------------------ MessagePassing._index_select
def make_index_select(theirs):
    """nn/conv/message_passing.py:263-267: the gather of `_collect` / `_lift` becomes lazy (see lazy.py)."""
    def _index_select(self, src: Tensor, index) -> Tensor:
        if (engine_ok(src) and not _compiling() and not torch.jit.is_scripting() and isinstance(index, Tensor)
                and index.dim() == 1 and src.dim() >= 2 and (self.node_dim == 0 or self.node_dim == -src.dim())
                and not isinstance(src, LazyRows) and not getattr(self, "explain", False)):
            return LazyRows(src, index)
        return theirs(self, src, index)
    _index_select.__wrapped__ = theirs
    return _index_select

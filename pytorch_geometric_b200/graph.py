"""CSRGraph -- the engine's cached graph structure for one edge set.

Plays the role of the reference's `EdgeIndex` caches (`_indptr`, `_T_perm`, `_T_index`,
`_T_indptr`, `_value`; torch_geometric/edge_index.py:237-246, get_csr/get_csc :626-696) and of a
`torch_sparse.SparseTensor` `adj_t` (rowptr/col/value + csr2csc): a destination-sorted CSR used
by the forward gather-reduce and, built lazily, the source-sorted CSR used by the backward.

HBM layout: int32 `rowptr`/`col` whenever #nodes and #edges < 2^31 (the
reference keeps int64), fp32 edge values in CSR order, an int32 `perm` (CSR slot -> original edge
id) so per-edge tensors given in the caller's order can be permuted once.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops

DEFAULT_CHUNK = 512  # edges per long-row chunk (rows above this are split, see csr_reduce.cuh)
_INT32_MAX = 2**31 - 1


class CSRGraph:
    def __init__(self, src: Tensor, dst: Tensor, num_src: int, num_dst: int,
                 edge_weight: Optional[Tensor] = None, chunk: int = DEFAULT_CHUNK,
                 idx_dtype: Optional[torch.dtype] = None):
        if not src.is_cuda:
            raise RuntimeError("CSRGraph lives on a CUDA device (no CPU fallback)")
        E = src.numel()
        if idx_dtype is None:
            idx_dtype = torch.int32 if max(num_src, num_dst, E) < _INT32_MAX else torch.int64
        self.idx_dtype = idx_dtype
        self.num_src, self.num_dst, self.num_edges = int(num_src), int(num_dst), int(E)
        self.chunk = int(chunk)
        self.device = src.device
        self._src = ops.convert_index(src.contiguous(), idx_dtype)   # original edge order
        self._dst = ops.convert_index(dst.contiguous(), idx_dtype)
        # forward structure: stable sort by destination
        _, self.perm, self.rowptr = ops.sort_by_key(self._dst, self.num_dst, want_sorted=False)
        self.col = ops.permute(self._src, self.perm)
        self.val = None if edge_weight is None else ops.permute(edge_weight.detach().float().contiguous(), self.perm)
        self.plan = ops.LongRowPlan(self.rowptr, self.chunk)
        # backward structure (lazy)
        self._t_built = False
        self.perm_t = self.rowptr_t = self.col_t = self.val_t = self.plan_t = None
        self._mean_val_t = None
        self._inv_perm = None
        self._inv_perm_t = None
        self._dst_csr = None
        self._t2csr = None

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_edge_index(cls, edge_index: Tensor, num_nodes: Optional[int] = None,
                        edge_weight: Optional[Tensor] = None, num_src: Optional[int] = None,
                        num_dst: Optional[int] = None, **kw) -> "CSRGraph":
        """edge_index[0] = source (j), edge_index[1] = destination (i): flow source_to_target
        (message_passing.py:31 / collect.jinja:67-68)."""
        if edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must have shape [2, E]")
        if num_nodes is None and (num_src is None or num_dst is None):
            _, mx, _ = ops.index_stats(edge_index.reshape(-1)) if edge_index.numel() else (0, -1, True)
            num_nodes = mx + 1   # maybe_num_nodes (utils/num_nodes.py:12-40): a D2H sync, as in the reference
        return cls(edge_index[0], edge_index[1], num_src if num_src is not None else num_nodes,
                   num_dst if num_dst is not None else num_nodes, edge_weight, **kw)

    # ------------------------------------------------------------------ transposed structure
    def build_transpose(self) -> None:
        if self._t_built:
            return
        _, self.perm_t, self.rowptr_t = ops.sort_by_key(self._src, self.num_src, want_sorted=False)
        self.col_t = ops.permute(self._dst, self.perm_t)
        self.plan_t = ops.LongRowPlan(self.rowptr_t, self.chunk)
        self._t_built = True
        if self.val is not None:
            self.val_t = self.to_csc_order(self.from_csr_order(self.val))

    # per-edge tensor permutations (1-D fp32 / int tensors)
    def to_csr_order(self, per_edge: Tensor) -> Tensor:
        return ops.permute(per_edge.contiguous(), self.perm)

    def to_csc_order(self, per_edge: Tensor) -> Tensor:
        self.build_transpose()
        return ops.permute(per_edge.contiguous(), self.perm_t)

    def _inverse(self, perm: Tensor) -> Tensor:
        inv = torch.empty_like(perm)
        inv[perm.long()] = torch.arange(perm.numel(), device=perm.device, dtype=perm.dtype)
        return inv

    def from_csr_order(self, per_edge_csr: Tensor) -> Tensor:
        if self._inv_perm is None:
            self._inv_perm = self._inverse(self.perm)
        return ops.permute(per_edge_csr.contiguous(), self._inv_perm)

    def from_csc_order(self, per_edge_csc: Tensor) -> Tensor:
        self.build_transpose()
        if self._inv_perm_t is None:
            self._inv_perm_t = self._inverse(self.perm_t)
        return ops.permute(per_edge_csc.contiguous(), self._inv_perm_t)

    def to_csr_order_rows(self, per_edge: Tensor) -> Tensor:
        """[E, k] per-edge rows (caller's edge order) -> CSR order."""
        return ops.gather_rows(per_edge.contiguous(), self.perm)

    def from_csr_order_rows(self, per_edge_csr: Tensor) -> Tensor:
        if self._inv_perm is None:
            self._inv_perm = self._inverse(self.perm)
        return ops.gather_rows(per_edge_csr.contiguous(), self._inv_perm)

    @property
    def t2csr(self) -> Tensor:
        """CSR slot of every transposed-CSR slot (the reference's csr2csc / _T_perm role)."""
        if getattr(self, "_t2csr", None) is None:
            self.build_transpose()
            if self._inv_perm is None:
                self._inv_perm = self._inverse(self.perm)
            self._t2csr = ops.permute(self._inv_perm, self.perm_t)
        return self._t2csr

    @property
    def dst_csr(self) -> Tensor:
        """destination of every CSR slot (ptr2index of rowptr)."""
        if self._dst_csr is None:
            self._dst_csr = ops.ptr2index(self.rowptr, self.num_edges)
        return self._dst_csr

    def in_degree(self) -> Tensor:
        return self.rowptr[1:] - self.rowptr[:-1]

    def mean_val_t(self) -> Tensor:
        """1 / max(in_degree(dst), 1) per transposed-CSR slot: backward weights of 'mean'."""
        if self._mean_val_t is None:
            self.build_transpose()
            inv = 1.0 / self.in_degree().clamp(min=1).to(torch.float32)
            self._mean_val_t = ops.gather_rows(inv.view(-1, 1), self.col_t).view(-1)
        return self._mean_val_t

    def with_values(self, val_csr: Optional[Tensor]) -> "CSRGraph":
        """Shallow copy sharing the structure but carrying different (static) CSR-ordered values."""
        g = object.__new__(CSRGraph)
        g.__dict__.update(self.__dict__)
        g.val = val_csr
        g.val_t = None
        if val_csr is not None and self._t_built:
            g.val_t = g.to_csc_order(g.from_csr_order(val_csr))
        return g

    def nbytes(self) -> int:
        n = 0
        for t in (self.rowptr, self.col, self.perm, self.val, self.rowptr_t, self.col_t, self.perm_t, self.val_t,
                  self._src, self._dst):
            if t is not None:
                n += t.numel() * t.element_size()
        return n

    def __repr__(self) -> str:
        return (f"CSRGraph(num_src={self.num_src}, num_dst={self.num_dst}, num_edges={self.num_edges}, "
                f"idx={self.idx_dtype}, long_rows={self.plan.n_long}, chunks={self.plan.n_chunks})")

"""LazyRows -- the hook behind `MessagePassing._collect / _lift / _index_select`
(nn/conv/message_passing.py:263-333, collect.jinja:118-139).

The reference materialises `x_j = x.index_select(node_dim, edge_index_j)` ([E, F]: 113 GB at the headline shape),
runs `message` on it and scatters the result.  With the plug-in installed, `_index_select` on a CUDA fp32 / bf16
feature matrix returns a `LazyRows`: a tensor subclass that only REMEMBERS (matrix, index[, per-edge scale]).

  * `message` returning `x_j` or `edge_weight.view(-1, 1) * x_j` (GCNConv, SAGEConv, GINConv, GraphConv, ... --
    gcn_conv.py:270-271, graph_conv.py:100-101) keeps it lazy: the multiplication is folded into the scale;
  * `aggregate` -> `Aggregation.reduce` -> `scatter` / `segment` (nn/aggr/base.py:173-185) sees the LazyRows and runs
    ONE fused gather-reduce over a CSR (`b200mp_spmm_csr`) -- adopted from the sorted `Index`/`ptr` the layer
    collected, or built by one cached stable sort -- instead of index_select + atomics;
  * anything else a layer does with `x_j` (concatenation, an MLP, attention logits, ...) materialises it through
    `__torch_function__` with the very `index_select` the reference would have run, so behaviour is unchanged.

explain mode and `decomposed_layers > 1` keep working: they call the same `_index_select` / `aggregate`.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from ._util import plain

_META = {"size", "dim", "numel", "stride", "is_floating_point", "is_complex", "is_contiguous", "element_size", "nelement",
         "ndimension", "type", "__len__", "is_cuda", "dtype", "device", "shape", "requires_grad", "ndim", "layout", "names",
         "is_sparse", "is_quantized", "is_meta", "grad_fn", "is_leaf", "data_ptr", "_version", "__get__", "__repr__",
         "__format__", "__class__", "__hash__", "__reduce_ex__", "untyped_storage", "storage_offset"}


class LazyRows(Tensor):
    """rows `index` of `src` along dim 0 (times `scale[e]` per row when set), not yet gathered."""

    @staticmethod
    def __new__(cls, src: Tensor, index: Tensor, scale: Optional[Tensor] = None):
        shape = (index.numel(), ) + tuple(src.shape[1:])
        r = Tensor._make_wrapper_subclass(cls, shape, dtype=src.dtype, device=src.device, requires_grad=False)
        r._src, r._index, r._scale = src, index, scale
        return r

    def __repr__(self):                                            # noqa: D105
        return f"LazyRows(rows={self._index.numel()}, of={tuple(self._src.shape)}, scaled={self._scale is not None})"

    def materialise(self) -> Tensor:
        """What the reference computes: src.index_select(0, index) (* scale)."""
        with torch._C.DisableTorchFunctionSubclass():
            idx = plain(self._index)
            out = self._src.index_select(0, idx)
            if self._scale is not None:
                s = self._scale
                out = s.view((-1, ) + (1, ) * (out.dim() - 1)) * out
        return out

    def _scaled_by(self, w: Tensor) -> Optional["LazyRows"]:
        """self * w for a per-edge weight w of shape [E] / [E, 1, ...]; None when w is anything else."""
        E = self._index.numel()
        if not isinstance(w, Tensor) or isinstance(w, LazyRows) or w.numel() != E or w.dim() == 0:
            return None
        if w.dim() > 1 and tuple(w.shape) != (E, ) + (1, ) * (w.dim() - 1):
            return None
        if not w.is_floating_point() or w.device != self.device:
            return None
        w1 = w.reshape(-1)
        return LazyRows(self._src, self._index, w1 if self._scale is None else self._scale * w1)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name in _META:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if name in ("mul", "__mul__", "__rmul__", "multiply") and len(args) == 2 and not kwargs:
            a, b = args
            lazy, other = (a, b) if isinstance(a, LazyRows) else (b, a)
            if isinstance(lazy, LazyRows):
                r = lazy._scaled_by(other)
                if r is not None:
                    return r
        # anything else: gather now (exactly the reference's index_select) and carry on with a plain tensor

        def mat(v):
            if isinstance(v, LazyRows):
                return v.materialise()
            if isinstance(v, (list, tuple)):
                return type(v)(mat(u) for u in v)
            return v
        with torch._C.DisableTorchFunctionSubclass():
            return func(*mat(args), **{k: mat(v) for k, v in kwargs.items()})

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # reached only by operations that bypassed __torch_function__ (e.g. autograd internals): materialise
        def mat(v):
            if isinstance(v, LazyRows):
                return v.materialise()
            if isinstance(v, (list, tuple)):
                return type(v)(mat(u) for u in v)
            return v
        return func(*mat(args), **{k: mat(v) for k, v in (kwargs or {}).items()})

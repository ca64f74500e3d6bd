"""Aggregation modules: mirrors of torch_geometric.nn.aggr.{Sum,Mean,Max,Min,Softmax}Aggregation
(nn/aggr/base.py:102-185, basic.py:12-50,142-218) on the sm_100a kernels.

`__call__(x, index=None, ptr=None, dim_size=None, dim=-2)` has the reference's meaning and error
behaviour.  Unlike the reference (base.py:177-180 only uses `ptr` in deterministic mode), the CSR
kernel is used whenever `ptr` is given: the result is the same (test/nn/aggr/test_basic.py:63)
and it is the deterministic, atomics-free path.  Without `ptr` the index is treated as unsorted
(atomic COO kernel) unless `index_sorted=True` is passed.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from .. import functional as Fn
from .. import ops
from .. import utils as U


class Aggregation(torch.nn.Module):
    reduce_op = "sum"

    def __call__(self, x: Tensor, index: Optional[Tensor] = None, ptr: Optional[Tensor] = None,
                 dim_size: Optional[int] = None, dim: int = -2, index_sorted: bool = False, **kwargs) -> Tensor:
        if dim >= x.dim() or dim < -x.dim():
            raise ValueError(f"Encountered invalid dimension '{dim}' of source tensor with {x.dim()} dimensions")
        if index is None and ptr is None:
            index = x.new_zeros(x.size(dim), dtype=torch.long)
        if ptr is not None:
            if dim_size is None:
                dim_size = ptr.numel() - 1
            elif dim_size != ptr.numel() - 1:
                raise ValueError(f"Encountered invalid 'dim_size' (got '{dim_size}' but expected "
                                 f"'{ptr.numel() - 1}')")
        if index is not None and dim_size is None:
            dim_size = (ops.index_stats(index)[1] + 1) if index.numel() > 0 else 0
        if index is not None and index.numel() > 0 and ptr is None:
            mx = ops.index_stats(index)[1]
            if mx >= dim_size:
                raise ValueError(f"Encountered invalid 'dim_size' (got '{dim_size}' but expected >= '{mx + 1}')")
        return super().__call__(x, index=index, ptr=ptr, dim_size=dim_size, dim=dim, index_sorted=index_sorted,
                                **kwargs)

    def reduce(self, x: Tensor, index: Optional[Tensor], ptr: Optional[Tensor], dim_size: Optional[int], dim: int,
               reduce: str, index_sorted: bool = False) -> Tensor:
        d = dim + x.dim() if dim < 0 else dim
        if ptr is not None:
            xm = x if d == 0 else x.movedim(d, 0).contiguous()
            out = Fn.segment(xm, ptr, reduce)
            return out if d == 0 else out.movedim(0, d)
        return U.scatter(x, index, d, dim_size, reduce, sorted=index_sorted)

    def forward(self, x, index=None, ptr=None, dim_size=None, dim=-2, index_sorted=False):
        return self.reduce(x, index, ptr, dim_size, dim, self.reduce_op, index_sorted)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"


class SumAggregation(Aggregation):
    reduce_op = "sum"


class MeanAggregation(Aggregation):
    reduce_op = "mean"


class MaxAggregation(Aggregation):
    reduce_op = "max"


class MinAggregation(Aggregation):
    reduce_op = "min"


class SoftmaxAggregation(Aggregation):
    """alpha = softmax(t * x) per group; out = sum(alpha * x)   (nn/aggr/basic.py:142-218)."""

    def __init__(self, t: float = 1.0, learn: bool = False, semi_grad: bool = False, channels: int = 1):
        super().__init__()
        if learn and semi_grad:
            raise ValueError("Cannot enable 'semi_grad' if 't' is learnable")
        self._init_t = t
        self.learn, self.semi_grad, self.channels = learn, semi_grad, channels
        self.t = torch.nn.Parameter(torch.full((channels, ), float(t))) if learn else t

    def forward(self, x, index=None, ptr=None, dim_size=None, dim=-2, index_sorted=False):
        t = self.t
        if self.channels != 1:
            shape = [1] * x.dim()
            shape[-1] = -1
            t = t.view(shape)
        alpha = x
        if not isinstance(t, (int, float)) or t != 1:
            alpha = x * t
        d = dim + x.dim() if dim < 0 else dim
        if not self.learn and self.semi_grad:
            with torch.no_grad():
                alpha = U.softmax(alpha, index, ptr, dim_size, d)
        else:
            alpha = U.softmax(alpha, index, ptr, dim_size, d)
        return self.reduce(x * alpha, index, ptr, dim_size, dim, "sum", index_sorted)


def aggregation_resolver(name: str) -> Aggregation:
    table = {"sum": SumAggregation, "add": SumAggregation, "mean": MeanAggregation, "max": MaxAggregation,
             "min": MinAggregation, "softmax": SoftmaxAggregation}
    if name not in table:
        raise ValueError(f"Could not resolve '{name}' among the aggregations on the hot path {sorted(table)}")
    return table[name]()

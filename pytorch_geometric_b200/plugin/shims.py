"""The optional-extension operator signatures the reference binds to, served by the
engine: flipping `torch_geometric.typing.WITH_*` to True with these modules bound makes every `torch_scatter.*`,
`pyg_lib.ops.*` and `torch.ops.torch_sparse.*` call site of the reference land in the sm_100a kernels.

  torch_scatter.scatter(src, index, dim, out=None, dim_size=None, reduce=...)      utils/_scatter.py:115,135
  torch_scatter.scatter_max / scatter_min(...) -> (out, arg)                       utils/_scatter.py:156
  torch_scatter.segment_csr(src, indptr, out=None, reduce=...)                     utils/_segment.py:34
  torch.ops.torch_sparse.spmm_sum(row, rowptr, col, value, colptr, csr2csc, mat)   edge_index.py:1798-1800
  torch.ops.torch_sparse.spmm_mean(row, rowptr, col, value, rowcount, colptr, csr2csc, mat)          :1802-1805
  torch.ops.torch_sparse.spmm_min / spmm_max(rowptr, col, value, mat) -> (out, arg)                  :1807-1810
  pyg_lib.ops.softmax_csr(src, ptr, dim)                                           utils/_softmax.py:58
  pyg_lib.ops.index_sort(inputs, max_value) -> (values, perm)                      utils/_index_sort.py:32
  pyg_lib.ops.segment_matmul(inputs, ptr, other) / grouped_matmul(inputs, others, biases)   nn/dense/linear.py:255,304-330

CUDA fp32 / bf16 operands run in the engine.  CPU operands fall through to the reference's own ATen branch (the
shim calls the reference function with the extension flag switched off for the duration of the call) -- the engine
itself never computes on the CPU.  torch.ops.torch_sparse.* are `torch.library` operators with a CUDA implementation,
a Meta (shape) implementation so tracing stays legal, and autograd formulas that use the transposed structure the
reference passes (colptr, csr2csc), exactly like torch_sparse's own.

Semantics note: real torch_scatter / torch_sparse route the min / max gradient to ONE arg element; the reference
without extensions (ATen `scatter_reduce`) splits it evenly among ties.  `scatter(..., reduce='max')` here keeps the
ATen rule (the pinned oracle); the explicit `(out, arg)` operators return the first extremum like the originals.
"""
from __future__ import annotations

import contextlib
import types
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from .. import dense
from .. import functional as Fn
from .. import ops
from .. import utils as U

_ENGINE = (torch.float32, torch.bfloat16)


def _ok(t) -> bool:
    return isinstance(t, Tensor) and t.is_cuda and t.dtype in _ENGINE


@contextlib.contextmanager
def _flags_off(*names):
    import torch_geometric.typing as T
    old = {n: getattr(T, n) for n in names}
    try:
        for n in names:
            setattr(T, n, False)
        yield
    finally:
        for n, v in old.items():
            setattr(T, n, v)


# ================================================================================================ torch_scatter
def _ts_scatter(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None, dim_size: Optional[int] = None,
                reduce: str = "sum") -> Tensor:
    if out is not None:
        raise NotImplementedError("torch_scatter.scatter(out=...) is not used by the reference and not provided")
    d = dim + src.dim() if dim < 0 else dim
    if index.dim() != 1:                                    # torch_scatter broadcasts the index; the reference passes 1-D
        index = index.movedim(d, 0).reshape(index.size(d), -1)[:, 0] if index.dim() == src.dim() else index.reshape(-1)
    if _ok(src):
        return U.scatter(src, index, d, dim_size, reduce)
    from torch_geometric.utils import _scatter as S
    with _flags_off("WITH_TORCH_SCATTER"):
        return getattr(S.scatter, "__wrapped__", S.scatter)(src, index, d, dim_size, reduce)


def _ts_scatter_arg(which: str):
    def fn(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
           dim_size: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        if out is not None:
            raise NotImplementedError("out= is not provided")
        d = dim + src.dim() if dim < 0 else dim
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() else 0
        if not (isinstance(src, Tensor) and src.is_cuda):
            raise NotImplementedError(f"torch_scatter.scatter_{which} shim: CUDA tensors only (the reference itself only "
                                      "calls it when the real extension is installed)")
        x = src if d == 0 else src.movedim(d, 0)
        x32 = x.contiguous().float()
        o = ops.scatter_coo(x32, index.contiguous(), dim_size, which)
        arg = ops.scatter_arg(x32, index.contiguous(), o)
        o = o.to(src.dtype)
        return (o, arg) if d == 0 else (o.movedim(0, d), arg.movedim(0, d))
    fn.__name__ = f"scatter_{which}"
    return fn


def _ts_segment_csr(src: Tensor, indptr: Tensor, out: Optional[Tensor] = None, reduce: str = "sum") -> Tensor:
    if out is not None:
        raise NotImplementedError("out= is not provided")
    if indptr.dim() != 1:                                   # expand_left'ed ptr of Aggregation.reduce: [1, ..., N+1]
        d = indptr.dim() - 1
        x = src if d == 0 else src.movedim(d, 0)
        res = _ts_segment_csr(x.contiguous(), indptr.reshape(-1), None, reduce)
        return res if d == 0 else res.movedim(0, d)
    if _ok(src):
        return U.segment(src, indptr, reduce)
    from torch_geometric.utils import _segment as S
    with _flags_off("WITH_TORCH_SCATTER"):
        return getattr(S.segment, "__wrapped__", S.segment)(src, indptr, reduce)


def torch_scatter_module() -> types.ModuleType:
    m = types.ModuleType("torch_scatter")
    m.__doc__ = "pytorch_geometric_b200 shim of the torch_scatter operators the reference calls"
    m.scatter = _ts_scatter
    for r in ("sum", "add", "mean", "mul", "min", "max"):
        if r in ("min", "max"):
            setattr(m, f"scatter_{r}", _ts_scatter_arg(r))
        else:
            setattr(m, f"scatter_{r}", (lambda rr: lambda src, index, dim=-1, out=None, dim_size=None:
                                        _ts_scatter(src, index, dim, out, dim_size, rr))(r))
    m.segment_csr = _ts_segment_csr
    m.__version__ = "b200mp-shim"
    return m


# ================================================================================================ torch.ops.torch_sparse
_SPARSE_LIB = None


def _mean_rowcount(rowptr: Tensor) -> Tensor:
    return (rowptr[1:] - rowptr[:-1]).clamp(min=1).to(torch.float32)


def register_torch_sparse_ops() -> bool:
    """Defines torch.ops.torch_sparse.spmm_{sum,mean,min,max} with torch_sparse's schemas (CUDA + Meta + autograd).
    Returns False when the namespace already has them (the real extension is installed)."""
    global _SPARSE_LIB
    if _SPARSE_LIB is not None:
        return True
    try:
        torch.ops.torch_sparse.spmm_sum                                            # noqa: B018
        return False
    except (AttributeError, RuntimeError):
        pass
    lib = torch.library.Library("torch_sparse", "DEF")
    lib.define("spmm_sum(Tensor? row, Tensor rowptr, Tensor col, Tensor? value, Tensor? colptr, Tensor? csr2csc, Tensor mat) -> Tensor")
    lib.define("spmm_mean(Tensor? row, Tensor rowptr, Tensor col, Tensor? value, Tensor? rowcount, Tensor? colptr, Tensor? csr2csc, Tensor mat) -> Tensor")
    lib.define("spmm_min(Tensor rowptr, Tensor col, Tensor? value, Tensor mat) -> (Tensor, Tensor)")
    lib.define("spmm_max(Tensor rowptr, Tensor col, Tensor? value, Tensor mat) -> (Tensor, Tensor)")

    def fwd(rowptr, col, value, mat, reduce):
        if mat.dtype not in _ENGINE:
            raise TypeError("torch_sparse.spmm_* (b200mp): float32 / bfloat16 only")
        v = None if value is None else value.detach().float()
        return ops.spmm_csr(rowptr, col, v, mat.detach(), rowptr.numel() - 1, reduce)

    def sum_cuda(row, rowptr, col, value, colptr, csr2csc, mat):
        return fwd(rowptr, col, value, mat, "sum")

    def mean_cuda(row, rowptr, col, value, rowcount, colptr, csr2csc, mat):
        return fwd(rowptr, col, value, mat, "mean")

    def arg_cuda(reduce):
        def f(rowptr, col, value, mat):
            m32 = mat.detach().float().contiguous()
            v = None if value is None else value.detach().float()
            out = ops.spmm_csr(rowptr, col, v, m32, rowptr.numel() - 1, reduce)
            arg = ops.spmm_csr_arg(rowptr, col, v, m32, out)
            return out.to(mat.dtype), arg
        return f

    def meta1(*args):
        rowptr, mat = (args[1], args[-1])
        return mat.new_empty((rowptr.numel() - 1, mat.size(1)))

    def meta2(rowptr, col, value, mat):
        n = rowptr.numel() - 1
        return mat.new_empty((n, mat.size(1))), mat.new_empty((n, mat.size(1)), dtype=torch.long)

    lib.impl("spmm_sum", sum_cuda, "CUDA")
    lib.impl("spmm_mean", mean_cuda, "CUDA")
    lib.impl("spmm_min", arg_cuda("min"), "CUDA")
    lib.impl("spmm_max", arg_cuda("max"), "CUDA")
    lib.impl("spmm_sum", meta1, "Meta")
    lib.impl("spmm_mean", meta1, "Meta")
    lib.impl("spmm_min", meta2, "Meta")
    lib.impl("spmm_max", meta2, "Meta")

    # ---- autograd (torch_sparse/csrc/spmm.cpp: grad_mat = A^T grad through (colptr, csr2csc); grad_value = SDDMM)
    def make_backward(is_mean):
        def setup(ctx, inputs, output):
            if is_mean:
                row, rowptr, col, value, rowcount, colptr, csr2csc, mat = inputs
            else:
                row, rowptr, col, value, colptr, csr2csc, mat = inputs
            ctx.save_for_backward(row, rowptr, col, value, colptr, csr2csc, mat)

        def backward(ctx, grad):
            row, rowptr, col, value, colptr, csr2csc, mat = ctx.saved_tensors
            grad = grad.contiguous()
            need_mat = ctx.needs_input_grad[-1]
            need_val = value is not None and ctx.needs_input_grad[3]
            g_mat = g_val = None
            scale = None
            if is_mean:
                scale = 1.0 / _mean_rowcount(rowptr)                         # per destination row
            if need_mat:
                if colptr is None or csr2csc is None or row is None:
                    raise RuntimeError("spmm backward needs row / colptr / csr2csc (the reference passes them when "
                                       "`other.requires_grad`, edge_index.py:1789-1796)")
                row_t = row.index_select(0, csr2csc)
                w = None if value is None else value.detach().float().index_select(0, csr2csc)
                if scale is not None:
                    s_t = scale.index_select(0, row_t)
                    w = s_t if w is None else w * s_t
                g_mat = ops.spmm_csr(colptr, row_t, w, grad, mat.size(0), "sum")
            if need_val:
                dot = ops.sddmm_csr(rowptr, col, grad, mat.detach())         # <grad[row(e)], mat[col[e]]> in CSR order
                if scale is not None:
                    dot = dot * scale.index_select(0, ops.ptr2index(rowptr, col.numel()))
                g_val = dot.to(value.dtype)
            n_in = 8 if is_mean else 7
            res = [None] * n_in
            res[3], res[-1] = g_val, g_mat
            return tuple(res)
        return setup, backward

    for name, is_mean in (("spmm_sum", False), ("spmm_mean", True)):
        setup, backward = make_backward(is_mean)
        torch.library.register_autograd(f"torch_sparse::{name}", backward, setup_context=setup, lib=lib)

    def arg_setup(ctx, inputs, output):
        rowptr, col, value, mat = inputs
        ctx.save_for_backward(col, value, output[1], mat)

    def arg_backward(ctx, grad, _grad_arg):
        col, value, arg, mat = ctx.saved_tensors
        nnz = col.numel()
        if nnz == 0:
            return None, None, (None if value is None else torch.zeros_like(value)), torch.zeros_like(mat)
        valid = arg < nnz
        a = arg.clamp(max=nnz - 1)
        g = torch.where(valid, grad, torch.zeros_like(grad))
        src_row = col.long()[a]                                              # [n_rows, F]: the producing source row
        g_mat = g_val = None
        # one producer per output element (torch_sparse semantics): element-wise scatters, done by ATen
        if ctx.needs_input_grad[3]:
            gm = g if value is None else g * value.detach().to(g.dtype)[a]
            g_mat = torch.zeros_like(mat).scatter_add_(0, src_row, gm)
        if value is not None and ctx.needs_input_grad[2]:
            contrib = g * mat.detach().gather(0, src_row)
            g_val = torch.zeros(nnz, dtype=contrib.dtype, device=contrib.device).scatter_add_(0, a.reshape(-1), contrib.reshape(-1))
            g_val = g_val.to(value.dtype)
        return None, None, g_val, g_mat

    for name in ("spmm_min", "spmm_max"):
        torch.library.register_autograd(f"torch_sparse::{name}", arg_backward, setup_context=arg_setup, lib=lib)
    _SPARSE_LIB = lib
    return True


# ================================================================================================ pyg_lib.ops
def _pl_softmax_csr(src: Tensor, ptr: Tensor, dim: int = 0) -> Tensor:
    if _ok(src):
        return U.softmax(src, None, ptr, None, dim).to(src.dtype)
    from torch_geometric.utils import _softmax as S
    with _flags_off("WITH_SOFTMAX", "WITH_TORCH_SCATTER"):
        return getattr(S.softmax, "__wrapped__", S.softmax)(src, None, ptr, None, dim)


def _pl_index_sort(inputs: Tensor, max_value: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    if isinstance(inputs, Tensor) and inputs.is_cuda and inputs.dtype in (torch.int32, torch.int64) and inputs.dim() == 1:
        return U.index_sort(inputs, max_value)
    return inputs.sort(stable=True)


def _pl_segment_matmul(inputs: Tensor, ptr: Tensor, other: Tensor) -> Tensor:
    """out[ptr[r]:ptr[r+1]] = inputs[ptr[r]:ptr[r+1]] @ other[r]   (nn/dense/linear.py:248-255, rgcn_conv.py:288)."""
    return dense.segment_matmul(inputs, ptr, other)


def _pl_grouped_matmul(inputs: List[Tensor], others: List[Tensor], biases: Optional[List[Tensor]] = None) -> List[Tensor]:
    return dense.grouped_matmul(inputs, others, biases)


def pyg_lib_module() -> types.ModuleType:
    m = types.ModuleType("pyg_lib")
    m.__doc__ = "pytorch_geometric_b200 shim of the pyg_lib.ops operators on the aggregation path"
    m.ops = types.ModuleType("pyg_lib.ops")
    m.ops.softmax_csr = _pl_softmax_csr
    m.ops.index_sort = _pl_index_sort
    m.ops.segment_matmul = _pl_segment_matmul
    m.ops.grouped_matmul = _pl_grouped_matmul
    m.__version__ = "b200mp-shim"
    return m

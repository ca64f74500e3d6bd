"""The layer's dense transform `x W^T (+ b)` (reference: nn/dense/linear.py:121-127, F.linear).

Two back ends, same fp32-level accuracy:
  * "tf32x3" (default on shapes it supports): the hand-written tcgen05/TMEM/TMA 3xTF32 GEMMs of
    csrc/gemm_tf32x3.cu -- forward, grad-input and the split-K grad-weight product all read x, g
    and W exactly as they lie in HBM;
  * "cublas": torch.nn.functional.linear in strict fp32 (what the reference runs) -- used for
    shapes outside the kernel's limits (reduction dim % 32, output width in {64,128,256k}) and for
    non-fp32 inputs.  A plain library GEMM, not a fallback of the aggregation path.
`set_backend("cublas")` forces the library path (A/B measurements in bench.py).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from . import ops
from ._lib import check, lib

_BACKEND = "tf32x3"
DEFAULT_GEMM_MODE = 1   # library default of b200mp_set_option("gemm_mode"): 0 = SS, 1 = TS (A operand in TMEM)
DEFAULT_GEMM_PREFETCH = 0   # library default of b200mp_set_option("gemm_prefetch") (k-blocks ahead, 0 = off)
_B_SPLIT = False            # True: pass W unsplit and let the kernel split its B tiles (w_lo == NULL in the C ABI)


def set_b_split(on: bool) -> None:
    global _B_SPLIT
    _B_SPLIT = bool(on)


def get_b_split() -> bool:
    return _B_SPLIT


def set_backend(name: str) -> None:
    global _BACKEND
    if name not in ("tf32x3", "cublas"):
        raise ValueError("backend must be 'tf32x3' or 'cublas'")
    _BACKEND = name


def get_backend() -> str:
    return _BACKEND


def _width_ok(n: int) -> bool:
    return n in (64, 128) or (n > 0 and n % 256 == 0)


def supported(x: Tensor, weight: Tensor) -> bool:
    if not (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.dim() == 2):
        return False
    n, k = weight.shape
    # forward: K % 32, width N;  grad_input: reduction N % 32, width K;  grad_weight: N % 128, width K
    return k % 32 == 0 and n % 128 == 0 and _width_ok(n) and _width_ok(k) and x.size(0) < 2**31


def split_tf32(w: Tensor):
    w = w.detach().contiguous()
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    ops._timed("split_tf32", 1, lib().b200mp_split_tf32, w.data_ptr(), hi.data_ptr(), lo.data_ptr(), w.numel(), ops._stream())
    return hi, lo


def prepare_weight(weight: Tensor):
    """(w_hi, w_lo) for the kernels; (w, None) when the kernel splits B tiles itself."""
    n, k = weight.shape
    if _B_SPLIT and n % 128 == 0 and k % 128 == 0:
        return weight.detach().contiguous(), None
    return split_tf32(weight)


def linear_forward(x: Tensor, w_hi: Tensor, w_lo: Optional[Tensor], out: Tensor = None) -> Tensor:
    m, k = x.shape
    n = w_hi.size(0)
    if out is None:
        y = torch.empty((m, n), dtype=torch.float32, device=x.device)
    else:
        if out.shape != (m, n) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous fp32 [M, N] tensor")
        y = out
    ops._timed("linear_tf32x3", 1, lib().b200mp_linear_tf32x3, x.data_ptr(), w_hi.data_ptr(), ops._p(w_lo),
               y.data_ptr(), m, n, k, ops._stream())
    return y


def linear_grad_input(g: Tensor, w_hi: Tensor, w_lo: Optional[Tensor]) -> Tensor:
    m, n = g.shape
    k = w_hi.size(1)
    gx = torch.empty((m, k), dtype=torch.float32, device=g.device)
    ops._timed("linear_grad_input_tf32x3", 1, lib().b200mp_linear_grad_input_tf32x3, g.data_ptr(), w_hi.data_ptr(),
               ops._p(w_lo), gx.data_ptr(), m, n, k, ops._stream())
    return gx


_WS = {}


def linear_grad_weight(g: Tensor, x: Tensor) -> Tensor:
    m, n = g.shape
    k = x.size(1)
    gw = torch.empty((n, k), dtype=torch.float32, device=g.device)
    nbytes = lib().b200mp_linear_grad_weight_workspace_bytes(m, n, k)
    key = (g.device, nbytes)
    ws = _WS.get(key)
    if ws is None:
        _WS.clear()
        ws = _WS[key] = torch.empty(nbytes, dtype=torch.uint8, device=g.device)
    ops._timed("linear_grad_weight_tf32x3", 2, lib().b200mp_linear_grad_weight_tf32x3, g.data_ptr(), x.data_ptr(),
               gw.data_ptr(), m, n, k, ws.data_ptr(), ws.numel(), ops._stream())
    return gw


class _LinearTF32x3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor):
        x = x.contiguous()
        w_hi, w_lo = prepare_weight(weight)
        ctx.save_for_backward(x, w_hi, w_lo)
        return linear_forward(x, w_hi, w_lo)

    @staticmethod
    def backward(ctx, g: Tensor):
        x, w_hi, w_lo = ctx.saved_tensors
        g = g.contiguous()
        gx = linear_grad_input(g, w_hi, w_lo) if ctx.needs_input_grad[0] else None
        gw = linear_grad_weight(g, x) if ctx.needs_input_grad[1] else None
        return gx, gw


def linear(x: Tensor, weight: Tensor, bias=None) -> Tensor:
    """x W^T + b with fp32 accuracy; tensor cores where the shape allows."""
    if _BACKEND == "tf32x3" and supported(x, weight):
        y = _LinearTF32x3.apply(x, weight)
        return y if bias is None else y + bias
    return F.linear(x, weight.to(x.dtype), None if bias is None else bias.to(x.dtype))

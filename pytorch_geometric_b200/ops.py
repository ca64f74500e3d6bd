"""Tensor-level wrappers over the C ABI (include/b200mp.h).  No autograd here (see functional.py).

PyTorch is used for device memory (the caching allocator owns every buffer) and the current
stream; all arithmetic happens in libb200mp.so.  Every function refuses non-CUDA tensors: there is
no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from ._lib import check, lib

F32, BF16 = 0, 1
I32, I64 = 0, 1
REDUCE = {"sum": 0, "add": 0, "mean": 1, "min": 2, "amin": 2, "max": 3, "amax": 3, "mul": 4}


class _Launches:
    """Number of engine kernels launched so far (bench.py reports the delta as gpu_launches)."""
    count = 0


LAUNCHES = _Launches()


class _Profile:
    """Optional per-op CUDA-event timing on the launching stream (bench.py's roofline numbers).
    Disabled by default: zero overhead on the product path."""

    def __init__(self):
        self.enabled = False
        self.events = {}

    def reset(self, enabled: bool = False):
        self.enabled = enabled
        self.events = {}

    def begin(self, name: str):
        if not self.enabled:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        self.events.setdefault(name, []).append((e0, e1))
        return e1

    def summary(self) -> dict:
        torch.cuda.synchronize()
        out = {}
        for k, v in self.events.items():
            each = [a.elapsed_time(b) for a, b in v]
            out[k] = {"ms_total": sum(each), "calls": len(each), "ms_each": each}
        return out


PROFILE = _Profile()


def _timed(name: str, n_kernels: int, fn, *args):
    LAUNCHES.count += n_kernels
    end = PROFILE.begin(name)
    rc = fn(*args)
    if end is not None:
        end.record()
    check(rc, name)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pytorch_geometric_b200 ops run on CUDA tensors only (no CPU fallback); "
                               f"got a tensor on {t.device}")


def _idt(t: Tensor) -> int:
    if t.dtype == torch.int32:
        return I32
    if t.dtype == torch.int64:
        return I64
    raise TypeError(f"index tensors must be int32 or int64, got {t.dtype}")


def _vdt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"feature tensors must be float32 or bfloat16, got {t.dtype}")


def _same_idx(*ts: Optional[Tensor]) -> int:
    ds = {t.dtype for t in ts if t is not None}
    if len(ds) != 1:
        raise TypeError(f"index tensors of one call must share a dtype, got {ds}")
    return _idt(next(t for t in ts if t is not None))


def _ws(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------ structure
def degree(index: Tensor, num_nodes: int) -> Tensor:
    _cuda(index)
    index = index.contiguous()
    deg = torch.empty(num_nodes, dtype=index.dtype, device=index.device)
    check(lib().b200mp_degree(_p(index), index.numel(), num_nodes, _p(deg), _idt(index), _stream()), "degree")
    return deg


def index2ptr(index: Tensor, size: int) -> Tensor:
    _cuda(index)
    index = index.contiguous()
    ptr = torch.empty(size + 1, dtype=index.dtype, device=index.device)
    check(lib().b200mp_index2ptr(_p(index), index.numel(), size, _p(ptr), _idt(index), _stream()), "index2ptr")
    return ptr


def ptr2index(ptr: Tensor, output_size: Optional[int] = None) -> Tensor:
    _cuda(ptr)
    ptr = ptr.contiguous()
    n = int(ptr[-1]) if output_size is None else int(output_size)
    index = torch.empty(n, dtype=ptr.dtype, device=ptr.device)
    check(lib().b200mp_ptr2index(_p(ptr), ptr.numel() - 1, n, _p(index), _idt(ptr), _stream()), "ptr2index")
    return index


def index_stats(index: Tensor) -> Tuple[int, int, bool]:
    """(min, max, is_sorted) of an index tensor -- one kernel + one D2H read."""
    _cuda(index)
    index = index.contiguous()
    st = torch.empty(3, dtype=torch.int64, device=index.device)
    check(lib().b200mp_index_stats(_p(index), index.numel(), _p(st), _idt(index), _stream()), "index_stats")
    mn, mx, srt = st.tolist()
    return mn, mx, bool(srt)


def sort_by_key(keys: Tensor, num_nodes: int, want_sorted: bool = True,
                want_ptr: bool = True) -> Tuple[Optional[Tensor], Tensor, Optional[Tensor]]:
    """Stable sort of keys in [0, num_nodes): (keys_sorted, perm, ptr)."""
    _cuda(keys)
    keys = keys.contiguous()
    n = keys.numel()
    it = _idt(keys)
    ws = _ws(lib().b200mp_sort_workspace_bytes(n, num_nodes, it), keys.device)
    ks = torch.empty_like(keys) if want_sorted else None
    perm = torch.empty_like(keys)
    ptr = torch.empty(num_nodes + 1, dtype=keys.dtype, device=keys.device) if want_ptr else None
    check(lib().b200mp_sort_by_key(_p(keys), n, num_nodes, _p(ks), _p(perm), _p(ptr), _p(ws), ws.numel(), it,
                                   _stream()), "sort_by_key")
    return ks, perm, ptr


def permute(src: Tensor, perm: Tensor) -> Tensor:
    """out[i] = src[perm[i]] for 1-D tensors of 4- or 8-byte elements."""
    _cuda(src, perm)
    src, perm = src.contiguous(), perm.contiguous()
    out = torch.empty(perm.numel(), dtype=src.dtype, device=src.device)
    check(lib().b200mp_permute(_p(src), _p(perm), _p(out), perm.numel(), src.element_size(), _idt(perm),
                               _stream()), "permute")
    return out


def convert_index(index: Tensor, dtype: torch.dtype) -> Tensor:
    _cuda(index)
    if index.dtype == dtype:
        return index
    index = index.contiguous()
    out = torch.empty(index.shape, dtype=dtype, device=index.device)
    check(lib().b200mp_convert_index(_p(index), _idt(index), _p(out), _idt(out), index.numel(), _stream()),
          "convert_index")
    return out


def self_loops(row: Tensor, col: Tensor, weight: Optional[Tensor], num_nodes: int, fill_value: float = 1.0,
               mode: int = 0) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
    """mode 0: add_remaining_self_loops; mode 1: remove_self_loops + add_self_loops."""
    _cuda(row, col, weight)
    row, col = row.contiguous(), col.contiguous()
    it = _same_idx(row, col)
    E = row.numel()
    if weight is not None:
        weight = weight.contiguous().float()
    ws = _ws(lib().b200mp_self_loops_workspace_bytes(E, num_nodes, it), row.device)
    r2 = torch.empty(E + num_nodes, dtype=row.dtype, device=row.device)
    c2 = torch.empty_like(r2)
    w2 = torch.empty(E + num_nodes, dtype=torch.float32, device=row.device) if weight is not None else None
    n_out = torch.empty(1, dtype=torch.int64, device=row.device)
    check(lib().b200mp_self_loops(_p(row), _p(col), _p(weight), E, num_nodes, float(fill_value), mode, _p(r2),
                                  _p(c2), _p(w2), _p(n_out), _p(ws), ws.numel(), it, _stream()), "self_loops")
    n = int(n_out.item())  # one D2H sync, as the reference's boolean-mask indexing has
    return r2[:n], c2[:n], (w2[:n] if w2 is not None else None)


def gcn_norm_csr(rowptr: Tensor, src: Tensor, weight: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """(deg_inv_sqrt [N], normalised weights [E]) on a destination-sorted CSR."""
    _cuda(rowptr, src, weight)
    it = _same_idx(rowptr, src)
    N, E = rowptr.numel() - 1, src.numel()
    dinv = torch.empty(N, dtype=torch.float32, device=rowptr.device)
    w_out = torch.empty(E, dtype=torch.float32, device=rowptr.device)
    check(lib().b200mp_gcn_norm_csr(_p(rowptr), _p(src), _p(weight), N, E, _p(dinv), _p(w_out), it, _stream()),
          "gcn_norm_csr")
    return dinv, w_out


class LongRowPlan:
    """Rows with more than `chunk` edges, cut into chunks (b200mp_csr_plan_*)."""

    __slots__ = ("long_rows", "chunk_ptr", "n_long", "n_chunks", "chunk", "_partials")

    def __init__(self, rowptr: Tensor, chunk: int):
        _cuda(rowptr)
        it = _idt(rowptr)
        n_rows = rowptr.numel() - 1
        counts = torch.empty(2, dtype=torch.int64, device=rowptr.device)
        check(lib().b200mp_csr_plan_count(_p(rowptr), n_rows, chunk, _p(counts), it, _stream()), "csr_plan_count")
        self.n_long, self.n_chunks = (int(v) for v in counts.tolist())
        self.chunk = int(chunk)
        self.long_rows = self.chunk_ptr = None
        self._partials = None
        if self.n_long:
            self.long_rows = torch.empty(self.n_long, dtype=torch.int64, device=rowptr.device)
            self.chunk_ptr = torch.empty(self.n_long + 1, dtype=torch.int64, device=rowptr.device)
            ws = _ws(lib().b200mp_csr_plan_workspace_bytes(n_rows, self.n_long, it), rowptr.device)
            check(lib().b200mp_csr_plan_fill(_p(rowptr), n_rows, chunk, self.n_long, _p(self.long_rows),
                                             _p(self.chunk_ptr), _p(ws), ws.numel(), it, _stream()),
                  "csr_plan_fill")

    @classmethod
    def empty(cls, chunk: int = 512) -> "LongRowPlan":
        """A plan without long rows, built WITHOUT the device->host count: always correct (a row above `chunk` edges
        is then simply walked by one lane group), meant for graphs whose degrees are bounded by construction
        (neighbour-sampled mini-batches: degree <= fan-out)."""
        p = object.__new__(cls)
        p.n_long = p.n_chunks = 0
        p.chunk = int(chunk)
        p.long_rows = p.chunk_ptr = p._partials = None
        return p

    def partials(self, feat: int, device) -> Optional[Tensor]:
        if not self.n_long:
            return None
        need = self.n_chunks * feat
        if self._partials is None or self._partials.numel() < need:
            self._partials = torch.empty(need, dtype=torch.float32, device=device)
        return self._partials


# ------------------------------------------------------------------ the hot path
def spmm_csr(rowptr: Tensor, col: Tensor, val: Optional[Tensor], x: Tensor, n_rows: int, reduce: str = "sum",
             plan: Optional[LongRowPlan] = None, out: Optional[Tensor] = None,
             bias: Optional[Tensor] = None, x_halo: Optional[Tensor] = None, accumulate: bool = False,
             peer_ptrs: Optional[int] = None, peer_rows: int = 0, relu_mask: Optional[Tensor] = None) -> Tensor:
    """out[i,:] = REDUCE_{e in row i} val[e] * x[col[e],:] (+ bias)  (x: [n_cols, F] contiguous)."""
    _cuda(rowptr, col, val, x)
    if x.dim() != 2:
        raise ValueError("spmm_csr expects a 2-D feature matrix")
    x = x.contiguous()
    it = _same_idx(rowptr, col)
    if val is not None and (val.dtype != torch.float32 or not val.is_contiguous()):
        val = val.contiguous().float()
    F = x.size(1)
    if out is None:
        out = torch.empty(n_rows, F, dtype=x.dtype, device=x.device)
    if plan is not None and plan.n_long:
        part = plan.partials(F, x.device)
        args = (_p(plan.long_rows), _p(plan.chunk_ptr), plan.n_long, plan.n_chunks, plan.chunk, _p(part))
    else:
        args = (None, None, 0, 0, 0, None)
    if bias is not None:
        _cuda(bias)
        if bias.dtype != torch.float32 or not bias.is_contiguous():
            bias = bias.detach().float().contiguous()
        if bias.numel() != F:
            raise ValueError("bias must have one entry per feature")
    if relu_mask is not None:
        _cuda(relu_mask)
        if not accumulate or relu_mask.dtype != x.dtype or tuple(relu_mask.shape) != (n_rows, F) or not relu_mask.is_contiguous():
            raise ValueError("relu_mask needs accumulate=True and a contiguous [n_rows, F] tensor of x's dtype")
    n_cols, n_local = x.size(0), 0
    if x_halo is not None:
        _cuda(x_halo)
        if x_halo.dtype != x.dtype or x_halo.dim() != 2 or x_halo.size(1) != F or not x_halo.is_contiguous():
            raise ValueError("x_halo must be a contiguous [n_halo, F] tensor of x's dtype")
        n_local, n_cols = x.size(0), x.size(0) + x_halo.size(0)
    _timed("spmm_csr", 2 if args[2] else 1, lib().b200mp_spmm_csr, _p(rowptr), _p(col), _p(val), _p(x), _p(out),
           n_rows, n_cols, F, REDUCE[reduce], *args, _p(bias), _p(x_halo), n_local, int(bool(accumulate)),
           peer_ptrs, int(peer_rows), _p(relu_mask), it, _vdt(x), _stream())
    return out


SEGMENT_PLAN_MIN_ROWS = 1 << 16   # below this many source rows a hub segment cannot matter: skip the plan (and its sync)
SEGMENT_CHUNK = 512               # rows per chunk of a long segment (same as graph.DEFAULT_CHUNK)


_PLAN_CACHE: "dict" = {}            # id(ptr) -> (weakref(ptr), version, plan or None)
_PLAN_CACHE_MAX = 32


def segment_plan(ptr: Tensor, n_src: int) -> Optional["LongRowPlan"]:
    """Long-segment plan for segment_csr / multi_aggr_csr (None when the input is too small to need one).
    Counting the long segments needs one device->host read, so the plan is cached per `ptr` tensor OBJECT
    (weak reference + version counter): a layer that hands over the same ptr every step -- an EdgeIndex's
    cached indptr, a loader's batch ptr -- pays that read once and enqueues asynchronously afterwards."""
    if n_src < SEGMENT_PLAN_MIN_ROWS:
        return None
    import weakref
    key = id(ptr)
    hit = _PLAN_CACHE.get(key)
    if hit is not None and hit[0]() is ptr and hit[1] == ptr._version:
        return hit[2]
    plan = LongRowPlan(ptr, SEGMENT_CHUNK)
    plan = plan if plan.n_long else None
    if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
        for k in [k for k, v in _PLAN_CACHE.items() if v[0]() is None] or list(_PLAN_CACHE)[:1]:
            _PLAN_CACHE.pop(k, None)
    try:
        _PLAN_CACHE[key] = (weakref.ref(ptr), ptr._version, plan)
    except TypeError:
        pass
    return plan


def segment_csr(src: Tensor, ptr: Tensor, reduce: str = "sum", plan: Optional["LongRowPlan"] = None) -> Tensor:
    _cuda(src, ptr)
    src = src.contiguous()
    n_rows = ptr.numel() - 1
    flat = src.view(src.size(0), -1)
    out = torch.empty((n_rows, flat.size(1)), dtype=src.dtype, device=src.device)
    pargs, _ = _plan_args(plan, flat.size(1), src.device)
    _timed("segment_csr", 2 if pargs[2] else 1, lib().b200mp_segment_csr, _p(ptr), _p(flat), _p(out), n_rows,
           flat.size(0), flat.size(1), REDUCE[reduce], *pargs, _idt(ptr), _vdt(src), _stream())
    return out.view((n_rows, ) + tuple(src.shape[1:]))


def minmax_ties(rowptr: Tensor, col: Tensor, val: Optional[Tensor], x: Tensor, out: Tensor,
                count_self_zero: bool = True) -> Tensor:
    _cuda(rowptr, col, val, x, out)
    it = _same_idx(rowptr, col)
    ties = torch.empty(out.shape, dtype=torch.float32, device=out.device)
    check(lib().b200mp_minmax_ties(_p(rowptr), _p(col), _p(val), _p(x), _p(out), _p(ties), out.size(0),
                                   out.size(1), int(count_self_zero), it, _vdt(x), _stream()), "minmax_ties")
    return ties


def minmax_backward(rowptr_t: Tensor, col_t: Tensor, val_t: Optional[Tensor], x: Tensor, out: Tensor,
                    grad_out: Tensor, ties: Tensor) -> Tensor:
    _cuda(rowptr_t, col_t, val_t, x, out, grad_out, ties)
    it = _same_idx(rowptr_t, col_t)
    gx = torch.empty_like(x)
    check(lib().b200mp_minmax_backward(_p(rowptr_t), _p(col_t), _p(val_t), _p(x), _p(out), _p(grad_out),
                                       _p(ties), _p(gx), x.size(0), x.size(1), it, _vdt(x), _stream()),
          "minmax_backward")
    return gx


def sddmm_csr(rowptr: Tensor, col: Tensor, a: Tensor, b: Tensor) -> Tensor:
    """dot[e] = <a[row(e),:], b[col[e],:]> in CSR edge order."""
    _cuda(rowptr, col, a, b)
    it = _same_idx(rowptr, col)
    a, b = a.contiguous(), b.contiguous()
    dot = torch.empty(col.numel(), dtype=torch.float32, device=a.device)
    check(lib().b200mp_sddmm_csr(_p(rowptr), _p(col), _p(a), _p(b), _p(dot), rowptr.numel() - 1, a.size(1), it,
                                 _vdt(a), _stream()), "sddmm_csr")
    return dot


def scatter_coo(src: Tensor, index: Tensor, n_rows: int, reduce: str = "sum") -> Tensor:
    """Atomic COO fallback for an unsorted index; fp32, src: [E, F]."""
    _cuda(src, index)
    if src.dtype != torch.float32:
        raise TypeError("scatter_coo is fp32 only (use the CSR path for bf16)")
    src, index = src.contiguous(), index.contiguous()
    flat = src.view(src.size(0), -1)
    out = torch.empty((n_rows, flat.size(1)), dtype=torch.float32, device=src.device)
    count = torch.empty(n_rows, dtype=torch.float32, device=src.device) if REDUCE[reduce] in (1, 2, 3) else None
    check(lib().b200mp_scatter_coo(_p(flat), _p(index), _p(out), _p(count), flat.size(0), n_rows, flat.size(1),
                                   REDUCE[reduce], _idt(index), _stream()), "scatter_coo")
    return out.view((n_rows, ) + tuple(src.shape[1:]))


def scatter_arg(src: Tensor, index: Tensor, out: Tensor) -> Tensor:
    """arg[i,f] = smallest e with index[e] == i and src[e,f] == out[i,f] (src.size(0) for empty groups): the second
    output of torch_scatter.scatter_max / scatter_min for an `out` computed by scatter_coo(min / max)."""
    _cuda(src, index, out)
    if src.dtype != torch.float32 or out.dtype != torch.float32:
        raise TypeError("scatter_arg is fp32 only")
    src, index, out = src.contiguous(), index.contiguous(), out.contiguous()
    flat, oflat = src.view(src.size(0), -1), out.view(out.size(0), -1)
    arg = torch.empty(oflat.shape, dtype=torch.int64, device=src.device)
    _timed("scatter_arg", 2, lib().b200mp_scatter_arg, _p(flat), _p(index), _p(oflat), _p(arg), flat.size(0), oflat.size(0),
           flat.size(1), _idt(index), _stream())
    return arg.view(out.shape)


def spmm_csr_arg(rowptr: Tensor, col: Tensor, val: Optional[Tensor], x: Tensor, out: Tensor) -> Tensor:
    """arg[i,f] = first CSR slot of row i whose (weighted) value equals out[i,f] (nnz for empty rows): the second
    output of torch.ops.torch_sparse.spmm_min / spmm_max."""
    _cuda(rowptr, col, val, x, out)
    if x.dtype != torch.float32 or out.dtype != torch.float32:
        raise TypeError("spmm_csr_arg is fp32 only")
    it = _same_idx(rowptr, col)
    x, out = x.contiguous(), out.contiguous()
    if val is not None:
        val = val.contiguous().float()
    arg = torch.empty(out.shape, dtype=torch.int64, device=x.device)
    _timed("spmm_csr_arg", 1, lib().b200mp_spmm_csr_arg, _p(rowptr), _p(col), _p(val), _p(x), _p(out), _p(arg), out.size(0),
           out.size(1), col.numel(), it, _stream())
    return arg


def index_add_rows(out: Tensor, index: Tensor, src: Tensor) -> Tensor:
    """out[index[e], :] += src[e, :] in place (fp32, atomics)."""
    _cuda(out, index, src)
    if out.dtype != torch.float32 or src.dtype != torch.float32 or not out.is_contiguous():
        raise TypeError("index_add_rows works on contiguous fp32 tensors")
    src, index = src.contiguous(), index.contiguous()
    check(lib().b200mp_index_add_rows(_p(src), _p(index), _p(out), index.numel(), out.size(1), _idt(index),
                                      _stream()), "index_add_rows")
    return out


def gather_rows(x: Tensor, index: Tensor, scale: Optional[Tensor] = None) -> Tensor:
    _cuda(x, index, scale)
    x, index = x.contiguous(), index.contiguous()
    flat = x.view(x.size(0), -1)
    out = torch.empty((index.numel(), flat.size(1)), dtype=x.dtype, device=x.device)
    check(lib().b200mp_gather_rows(_p(flat), _p(index), _p(scale), _p(out), index.numel(), flat.size(1),
                                   _idt(index), _vdt(x), _stream()), "gather_rows")
    return out.view((index.numel(), ) + tuple(x.shape[1:]))


def _softmax_edge_op(op: int, a: Tensor, b: Optional[Tensor], row: Optional[Tensor], dst: Optional[Tensor]) -> Tensor:
    out = torch.empty_like(a)
    _timed("softmax_edge_op", 1, lib().b200mp_softmax_edge_op, op, _p(a), _p(b), _p(row), _p(dst), _p(out), a.size(0),
           a.size(1), _idt(dst) if dst is not None else I64, _stream())
    return out


def softmax_csr(src: Tensor, ptr: Tensor, plan: Optional["LongRowPlan"] = None,
                index: Optional[Tensor] = None) -> Tensor:
    """Per-group softmax over ptr ranges.  With a long-row plan (hub groups) the reference's own sequence
    -- segment max, exp(x - max), segment sum, divide (_softmax.py:82-88) -- runs on the chunked segmented
    reduce and edge-parallel kernels; otherwise one fused three-pass kernel per group."""
    _cuda(src, ptr)
    src = src.contiguous().float()
    flat = src.view(src.size(0), -1)
    if plan is not None and plan.n_long:
        if index is None:
            index = ptr2index(ptr, flat.size(0))
        mx = segment_csr(flat, ptr, "max", plan)
        ex = _softmax_edge_op(0, flat, None, mx, index)
        den = segment_csr(ex, ptr, "sum", plan)
        return _softmax_edge_op(1, ex, None, den, index).view(src.shape)
    out = torch.empty_like(flat)
    _timed("softmax_csr", 1, lib().b200mp_softmax_csr, _p(ptr), _p(flat), _p(out), ptr.numel() - 1, flat.size(0),
           flat.size(1), _idt(ptr), _stream())
    return out.view(src.shape)


def softmax_csr_backward(out: Tensor, grad_out: Tensor, ptr: Tensor, plan: Optional["LongRowPlan"] = None,
                         index: Optional[Tensor] = None) -> Tensor:
    _cuda(out, grad_out, ptr)
    out, grad_out = out.contiguous(), grad_out.contiguous().float()
    flat = out.view(out.size(0), -1)
    gflat = grad_out.view(flat.shape)
    if plan is not None and plan.n_long:
        if index is None:
            index = ptr2index(ptr, flat.size(0))
        dot = segment_csr(_softmax_edge_op(2, flat, gflat, None, None), ptr, "sum", plan)
        return _softmax_edge_op(3, flat, gflat, dot, index).view(out.shape)
    g = torch.empty_like(flat)
    _timed("softmax_csr_backward", 1, lib().b200mp_softmax_csr_backward, _p(ptr), _p(flat), _p(gflat), _p(g),
           ptr.numel() - 1, flat.size(0), flat.size(1), _idt(ptr), _stream())
    return g.view(out.shape)


def _plan_args(plan, feat: int, device):
    """(long_rows, chunk_ptr, n_long, n_chunks, chunk, partials[n_chunks*feat]) for the C ABI."""
    if plan is None or not plan.n_long:
        return (None, None, 0, 0, 0, None), None
    part = plan.partials(feat, device)
    return (_p(plan.long_rows), _p(plan.chunk_ptr), plan.n_long, plan.n_chunks, plan.chunk, _p(part)), part


def gat_fused_csr(rowptr: Tensor, col: Tensor, xh: Tensor, a_src: Tensor, a_dst: Tensor, heads: int, chan: int,
                  slope: float, want_alpha: bool = False, plan: Optional["LongRowPlan"] = None,
                  dst_of_edge: Optional[Tensor] = None):
    """Fused GAT forward: returns (out [n_rows, H*C], row_max, row_den, alpha or None)."""
    _cuda(rowptr, col, xh, a_src, a_dst)
    it = _same_idx(rowptr, col)
    xh = xh.contiguous()
    a_src, a_dst = a_src.contiguous().float(), a_dst.contiguous().float()
    n_rows = rowptr.numel() - 1
    out = torch.empty((n_rows, heads * chan), dtype=xh.dtype, device=xh.device)
    row_max = torch.empty((n_rows, heads), dtype=torch.float32, device=xh.device)
    row_den = torch.empty_like(row_max)
    alpha = torch.empty((col.numel(), heads), dtype=torch.float32, device=xh.device) if want_alpha else None
    if want_alpha and dst_of_edge is None:
        dst_of_edge = ptr2index(rowptr, col.numel())
    pargs, _part = _plan_args(plan, heads * chan, xh.device)
    part_ms = torch.empty(plan.n_chunks * heads * 2, dtype=torch.float32, device=xh.device) if pargs[2] else None
    _timed("gat_fused_csr", 2 if pargs[2] else 1, lib().b200mp_gat_fused_csr, _p(rowptr), _p(col), _p(dst_of_edge), _p(xh),
           _p(a_src), _p(a_dst), _p(out), _p(row_max), _p(row_den), _p(alpha), n_rows, col.numel(), heads, chan,
           float(slope), *pargs, _p(part_ms), it, _vdt(xh), _stream())
    return out, row_max, row_den, alpha


def gat_fused_csr_backward(rowptr, col, dst_of_edge, rowptr_t, col_t, t2csr, xh, a_src, a_dst, row_max, row_den, out,
                           grad_out, heads: int, chan: int, slope: float, plan: Optional["LongRowPlan"] = None):
    """Returns (grad_xh [n_src, H*C], grad_a_src [n_src, H], grad_a_dst [n_rows, H])."""
    _cuda(rowptr, col, dst_of_edge, rowptr_t, col_t, t2csr, xh, grad_out)
    it = _same_idx(rowptr, col, dst_of_edge, rowptr_t, col_t, t2csr)
    grad_out = grad_out.contiguous()
    n_rows, n_src = rowptr.numel() - 1, rowptr_t.numel() - 1
    grad_pre = torch.empty((col.numel(), heads), dtype=torch.float32, device=xh.device)
    rowdot = torch.empty((n_rows, heads), dtype=torch.float32, device=xh.device)
    gxh = torch.empty_like(xh)
    gas = torch.empty((n_src, heads), dtype=torch.float32, device=xh.device)
    gad = torch.empty((n_rows, heads), dtype=torch.float32, device=xh.device)
    pargs, _part = _plan_args(plan, heads, xh.device)
    _timed("gat_fused_csr_backward", 5, lib().b200mp_gat_fused_csr_backward, _p(rowptr), _p(col), _p(dst_of_edge),
           _p(rowptr_t), _p(col_t), _p(t2csr), _p(xh), _p(a_src), _p(a_dst), _p(row_max), _p(row_den), _p(out),
           _p(grad_out), _p(grad_pre), _p(rowdot), _p(gxh), _p(gas), _p(gad), n_rows, n_src, col.numel(), heads, chan,
           float(slope), *pargs, it, _vdt(xh), _stream())
    return gxh, gas, gad


ATTN_MODES = {"gat": 0, "gatv2": 1, "dot": 2}


def attn_supported(heads: int, chan: int, dtype: torch.dtype) -> bool:
    """True when [*, heads*chan] rows of `dtype` are on the vector path of csrc/attention.cu."""
    if dtype not in (torch.float32, torch.bfloat16):
        return False
    return bool(lib().b200mp_attn_supported(heads, chan, BF16 if dtype == torch.bfloat16 else F32))


def _row_stride(t: Optional[Tensor], hc: int) -> int:
    """Row stride in elements of a [n, heads*chan] operand that may be a column slice of a wider matrix."""
    if t is None:
        return 0
    if t.dim() != 2 or t.size(1) != hc or t.stride(1) != 1:
        raise ValueError("attention operands must be [n, heads*chan] with unit inner stride")
    return t.stride(0)


def attn_forward(mode: str, rowptr: Tensor, col: Tensor, v: Tensor, heads: int, chan: int, *, k: Optional[Tensor] = None,
                 q: Optional[Tensor] = None, s_src: Optional[Tensor] = None, s_dst: Optional[Tensor] = None,
                 att: Optional[Tensor] = None, s_edge: Optional[Tensor] = None, slope: float = 0.2, scale: float = 1.0,
                 want_alpha: bool = False, plan: Optional["LongRowPlan"] = None, dropout_p: float = 0.0, dropout_seed: int = 0,
                 edge_feat: Optional[Tensor] = None):
    """Fused edge-softmax attention + aggregation (b200mp_attn_csr_forward).  v / k: [n_src, H*C] (column slices of
    a wider matrix are fine), q: [n_rows, H*C].  Returns (out, row_max, row_den, alpha or None).
    dropout_p / dropout_seed: attention dropout fused into the sweep (pass the same pair to attn_backward).
    edge_feat: [E, H*C] per-edge feature rows in CSR order (gatv2 / dot modes, `edge_dim` layers)."""
    _cuda(rowptr, col, v, k, q, s_src, s_dst, att, s_edge, edge_feat)
    if edge_feat is not None:
        if mode == "gat" or edge_feat.shape != (col.numel(), heads * chan) or edge_feat.dtype != v.dtype:
            raise ValueError("edge_feat must be [E, heads*chan] of the value dtype (gatv2 / dot modes)")
        edge_feat = edge_feat.contiguous()
    it = _same_idx(rowptr, col)
    hc = heads * chan
    n_rows = rowptr.numel() - 1
    out = torch.empty((n_rows, hc), dtype=v.dtype, device=v.device)
    row_max = torch.empty((n_rows, heads), dtype=torch.float32, device=v.device)
    row_den = torch.empty_like(row_max)
    alpha = torch.empty((col.numel(), heads), dtype=torch.float32, device=v.device) if want_alpha else None
    pargs, _part = _plan_args(plan, hc, v.device)
    part_ms = torch.empty(plan.n_chunks * heads * 2, dtype=torch.float32, device=v.device) if pargs[2] else None
    launches = 1 + (1 if pargs[2] else 0) + (1 if want_alpha else 0)
    _timed("attn_forward", launches, lib().b200mp_attn_csr_forward, ATTN_MODES[mode], _p(rowptr), _p(col), _p(v), _p(k), _p(q),
           _p(s_src), _p(s_dst), _p(att), _p(s_edge), _row_stride(v, hc), _row_stride(k, hc), _row_stride(q, hc), _p(out),
           _p(row_max), _p(row_den), _p(alpha), n_rows, col.numel(), heads, chan, float(slope), float(scale), *pargs,
           _p(part_ms), float(dropout_p), int(dropout_seed) & 0xFFFFFFFFFFFFFFFF, _p(edge_feat), it, _vdt(v), _stream())
    return out, row_max, row_den, alpha


def attn_backward(mode: str, rowptr, col, rowptr_t, col_t, t2csr, v: Tensor, heads: int, chan: int, row_max, row_den, out,
                  grad_out, *, k=None, q=None, s_src=None, s_dst=None, att=None, s_edge=None, slope: float = 0.2,
                  scale: float = 1.0, plan=None, plan_t=None, grad_v: Optional[Tensor] = None,
                  grad_k: Optional[Tensor] = None, dropout_p: float = 0.0, dropout_seed: int = 0,
                  edge_feat: Optional[Tensor] = None):
    """Backward of attn_forward.  Returns a dict with grad_v, and per mode grad_k / grad_q / grad_s_src /
    grad_s_dst / grad_att / grad_s_edge.  grad_v / grad_k may be preallocated (column slices of one matrix)."""
    _cuda(rowptr, col, rowptr_t, col_t, t2csr, v, grad_out, edge_feat)
    it = _same_idx(rowptr, col, rowptr_t, col_t, t2csr)
    grad_ef = None
    if edge_feat is not None:
        edge_feat = edge_feat.contiguous()
        grad_ef = torch.empty_like(edge_feat)
    m = ATTN_MODES[mode]
    hc = heads * chan
    dev = v.device
    n_rows, n_src, n_edges = rowptr.numel() - 1, rowptr_t.numel() - 1, col.numel()
    grad_out = grad_out.contiguous()
    out = out.contiguous()
    pair = torch.empty((n_edges, heads, 2), dtype=torch.float32, device=dev)
    if grad_v is None:
        grad_v = torch.empty((n_src, hc), dtype=v.dtype, device=dev)
    if m == 2 and grad_k is None:
        grad_k = torch.empty((n_src, hc), dtype=v.dtype, device=dev)
    if _row_stride(grad_v, hc) != _row_stride(v, hc) or (m == 2 and _row_stride(grad_k, hc) != _row_stride(k, hc)):
        raise ValueError("grad_v / grad_k must have the row stride of v / k")
    grad_q = torch.empty((n_rows, hc), dtype=v.dtype, device=dev) if m != 0 else None
    if q is not None and _row_stride(q, hc) != hc:
        q = q.contiguous()
    gss = torch.empty((n_src, heads), dtype=torch.float32, device=dev) if m == 0 else None
    gsd = torch.empty((n_rows, heads), dtype=torch.float32, device=dev) if m == 0 else None
    gatt = gatt_part = None
    if m == 1:
        gatt = torch.empty(hc, dtype=torch.float32, device=dev)
        gatt_part = torch.empty(int(lib().b200mp_attn_gatt_rows()) * hc, dtype=torch.float32, device=dev)
    w = int(lib().b200mp_attn_backward_partial_width(m, heads, chan, 0))
    wt = int(lib().b200mp_attn_backward_partial_width(m, heads, chan, 1))
    if plan is not None and plan_t is plan:   # one plan object cannot describe both the CSR and its transpose
        raise ValueError("pass distinct plans for the CSR and the transposed CSR")
    pa, _keep = _plan_args(plan, w, dev)
    pt, _keep_t = _plan_args(plan_t, wt, dev)
    _timed("attn_backward", 2 + (1 if pa[2] else 0) + (1 if pt[2] else 0) + (1 if m == 1 else 0), lib().b200mp_attn_csr_backward, m,
           _p(rowptr), _p(col), _p(rowptr_t), _p(col_t), _p(t2csr), _p(v), _p(k), _p(q), _p(s_src), _p(s_dst), _p(att),
           _p(s_edge), _row_stride(v, hc), _row_stride(k, hc), _row_stride(q, hc), _p(row_max), _p(row_den), _p(out),
           _p(grad_out), _p(pair), _p(grad_v), _p(grad_k), _p(grad_q), _p(gss), _p(gsd), _p(gatt), _p(gatt_part), n_rows,
           n_src, n_edges, heads, chan, float(slope), float(scale), pa[0], pa[1], pa[2], pa[3], pa[4], pa[5], pt[0],
           pt[1], pt[2], pt[3], pt[5], float(dropout_p), int(dropout_seed) & 0xFFFFFFFFFFFFFFFF, _p(edge_feat), _p(grad_ef), it,
           _vdt(v), _stream())
    res = {"grad_v": grad_v, "grad_k": grad_k, "grad_q": grad_q, "grad_s_src": gss, "grad_s_dst": gsd, "grad_att": gatt,
           "grad_edge_feat": grad_ef}
    if s_edge is not None:
        res["grad_s_edge"] = pair[:, :, 1]
    return res


def column_sum(x: Tensor) -> Tensor:
    """sum over rows of a [n, F] matrix in fp32 (the bias gradient); deterministic, two launches."""
    _cuda(x)
    if x.dim() != 2:
        raise ValueError("column_sum expects a 2-D matrix")
    x = x.contiguous()
    n, F = x.shape
    out = torch.empty(F, dtype=torch.float32, device=x.device)
    parts = int(lib().b200mp_column_sum_parts(n))
    ws = torch.empty(parts * max(F, 1), dtype=torch.float32, device=x.device)
    _timed("column_sum", 2, lib().b200mp_column_sum, _p(x), _p(out), _p(ws), parts, n, F, _vdt(x), _stream())
    return out


def head_dot_supported(x: Tensor, heads: int, chan: int) -> bool:
    return (x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and x.dim() == 2 and x.size(1) == heads * chan
            and bool(lib().b200mp_head_dot_supported(heads, chan, _vdt(x))))


def head_dot(x: Tensor, att_a: Tensor, att_b: Optional[Tensor], heads: int, chan: int):
    """(s_a, s_b) with s_*[n, h] = sum_c x[n, h, c] * att_*[h, c] in fp32, one read of x (b200mp_head_dot)."""
    _cuda(x, att_a, att_b)
    x = x.contiguous()
    att_a = att_a.reshape(-1).float().contiguous()
    att_b = None if att_b is None else att_b.reshape(-1).float().contiguous()
    n = x.size(0)
    s_a = torch.empty(n, heads, dtype=torch.float32, device=x.device)
    s_b = None if att_b is None else torch.empty(n, heads, dtype=torch.float32, device=x.device)
    _timed("head_dot", 1, lib().b200mp_head_dot, _p(x), _p(att_a), _p(att_b), _p(s_a), _p(s_b), n, heads, chan, _vdt(x),
           _stream())
    return s_a, s_b


def head_dot_backward(x: Tensor, att_a: Tensor, att_b: Optional[Tensor], g_a: Tensor, g_b: Optional[Tensor],
                      add: Optional[Tensor], heads: int, chan: int, want_grad_x: bool = True):
    """(grad_x, grad_att_a, grad_att_b): grad_x = g_a (x) att_a + g_b (x) att_b (+ add), grad_att_* fp32 [heads*chan]."""
    _cuda(x, att_a, att_b, g_a, g_b, add)
    x = x.contiguous()
    att_a = att_a.reshape(-1).float().contiguous()
    att_b = None if att_b is None else att_b.reshape(-1).float().contiguous()
    g_a = g_a.float().contiguous()
    g_b = None if g_b is None else g_b.float().contiguous()
    if add is not None:
        add = add.to(x.dtype).contiguous()
    n, F = x.shape
    gx = torch.empty_like(x) if want_grad_x else None
    parts = int(lib().b200mp_head_dot_parts(n, heads, chan, _vdt(x)))
    if parts == 0:
        z = torch.zeros(F, dtype=torch.float32, device=x.device)
        return gx, z, (None if att_b is None else z.clone())
    pa = torch.empty(parts, F, dtype=torch.float32, device=x.device)
    pb = None if att_b is None else torch.empty(parts, F, dtype=torch.float32, device=x.device)
    _timed("head_dot_backward", 1, lib().b200mp_head_dot_backward, _p(x), _p(att_a), _p(att_b), _p(g_a), _p(g_b), _p(add),
           _p(gx), _p(pa), _p(pb), parts, n, heads, chan, _vdt(x), _stream())
    return gx, column_sum(pa), (None if pb is None else column_sum(pb))


MULTI_AGGRS = ("sum", "mean", "min", "max", "var", "std")
MULTI_HIT_MASK = True       # emit the forward's hit bits for the backward (A/B switch for benchmarks)


def multi_aggr_csr(rowptr: Tensor, col: Optional[Tensor], x: Tensor, n_rows: int, want, plan=None,
                   with_ties: bool = False, count_self_zero: bool = True) -> dict:
    """Every aggregation named in `want` (subset of MULTI_AGGRS) from one sweep over the CSR rows.
    col=None: segment mode (x is the destination-sorted [E, F] message matrix).  Returns a dict
    name -> [n_rows, F]; with_ties adds fp32 'ties_min' / 'ties_max' when min / max are wanted."""
    _cuda(rowptr, col, x)
    if x.dim() != 2:
        raise ValueError("multi_aggr_csr expects a 2-D feature matrix")
    x = x.contiguous()
    it = _same_idx(rowptr, col) if col is not None else _idt(rowptr)
    F = x.size(1)
    res = {}
    for name in want:
        if name not in MULTI_AGGRS:
            raise ValueError(f"cannot fuse aggregation '{name}' (supported: {MULTI_AGGRS})")
        res[name] = torch.empty(n_rows, F, dtype=x.dtype, device=x.device)
    if with_ties:
        for name in ("min", "max"):
            if name in res:
                res["ties_" + name] = torch.empty(n_rows, F, dtype=torch.float32, device=x.device)
    # hit bits for the backward (one byte per edge and 16-byte vector): only where its masked kernel exists
    if (MULTI_HIT_MASK and with_ties and col is not None and ("ties_min" in res or "ties_max" in res)
            and lib().b200mp_multi_aggr_mask_supported(F, _vdt(x), 0)):
        res["hit_mask"] = torch.empty(col.numel() * (F // 4), dtype=torch.uint8, device=x.device)
    pargs, _ = _plan_args(plan, 6 * F, x.device)
    _timed("multi_aggr_csr", (3 if "hit_mask" in res else 2) if pargs[2] else 1, lib().b200mp_multi_aggr_csr, _p(rowptr),
           _p(col), _p(x), *[_p(res.get(n)) for n in MULTI_AGGRS], _p(res.get("ties_min")), _p(res.get("ties_max")),
           _p(res.get("hit_mask")), n_rows, x.size(0), F, int(bool(count_self_zero)), *pargs, it, _vdt(x), _stream())
    return res


def multi_aggr_prepare_backward(rowptr: Tensor, grads: dict, mean: Optional[Tensor], std: Optional[Tensor],
                                ties_min: Optional[Tensor], ties_max: Optional[Tensor], semi_grad: bool):
    """(term_a, term_b, g_min / ties_min, g_max / ties_max) for multi_aggr_backward from the output
    gradients `grads` (name -> [n_rows, F] or None), in one elementwise kernel."""
    ref = next(g for g in grads.values() if g is not None)
    n_rows, F = ref.shape
    g = {k: (None if v is None else v.contiguous()) for k, v in grads.items()}
    _cuda(rowptr, *g.values(), mean, std, ties_min, ties_max)

    def new():
        return torch.empty(n_rows, F, dtype=torch.float32, device=ref.device)

    need_a = any(g.get(k) is not None for k in ("sum", "mean", "var", "std"))
    need_b = any(g.get(k) is not None for k in ("var", "std"))
    term_a = new() if need_a else None
    term_b = new() if need_b else None
    gmin = new() if g.get("min") is not None else None
    gmax = new() if g.get("max") is not None else None
    _timed("multi_aggr_prepare_backward", 1, lib().b200mp_multi_aggr_prepare_backward, _p(rowptr), _p(g.get("sum")),
           _p(g.get("mean")), _p(g.get("var")), _p(g.get("std")), _p(g.get("min")), _p(g.get("max")), _p(mean),
           _p(std), _p(ties_min), _p(ties_max), _p(term_a), _p(term_b), _p(gmin), _p(gmax), n_rows, F,
           int(bool(semi_grad)), _idt(rowptr), _vdt(ref), _stream())
    return term_a, (None if semi_grad else term_b), gmin, gmax


def multi_aggr_backward(ptr: Optional[Tensor], idx: Tensor, x: Tensor, term_a: Optional[Tensor],
                        term_b: Optional[Tensor], out_min: Optional[Tensor], g_min: Optional[Tensor],
                        out_max: Optional[Tensor], g_max: Optional[Tensor], segment_mode: bool,
                        hit_mask: Optional[Tensor] = None, t2csr: Optional[Tensor] = None) -> Tensor:
    """grad wrt the message values (see b200mp_multi_aggr_backward).  segment_mode: idx = destination of
    every message; otherwise (ptr, idx) is the transposed CSR and x the [n_src, F] source matrix.
    hit_mask / t2csr: the forward's hit bits (multi_aggr_csr's 'hit_mask') and the CSR slot of every transposed slot."""
    _cuda(ptr, idx, x, term_a, term_b, out_min, g_min, out_max, g_max, hit_mask, t2csr)
    if hit_mask is not None:
        if t2csr is None or segment_mode or t2csr.dtype != idx.dtype:
            raise ValueError("hit_mask needs gather mode and t2csr of the index dtype")
        t2csr = t2csr.contiguous()
    x = x.contiguous()

    def f32(t):
        return None if t is None else t.contiguous().float()

    term_a, term_b, g_min, g_max = f32(term_a), f32(term_b), f32(g_min), f32(g_max)
    out_min = None if out_min is None else out_min.contiguous()
    out_max = None if out_max is None else out_max.contiguous()
    gx = torch.empty_like(x)
    it = _idt(idx) if segment_mode else _same_idx(ptr, idx)
    _timed("multi_aggr_backward", 1, lib().b200mp_multi_aggr_backward, _p(ptr), _p(idx), _p(x), _p(term_a),
           _p(term_b), _p(out_min), _p(g_min), _p(out_max), _p(g_max), _p(hit_mask), _p(t2csr), _p(gx), x.size(0), x.size(1),
           int(bool(segment_mode)), it, _vdt(x), _stream())
    return gx


def device_info() -> dict:
    import ctypes
    sm, ma, mi, l2 = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
    check(lib().b200mp_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(l2)))
    return {"sm_count": sm.value, "cc": (ma.value, mi.value), "l2_bytes": l2.value}


def set_option(name: str, value: int) -> None:
    """Runtime switches of the library (b200mp_set_option), e.g. set_option("spmm_impl", 1)."""
    check(lib().b200mp_set_option(name.encode(), int(value)), "set_option")

"""`torch.library` registration of the engine's operators (namespace `b200mp`), each with a CUDA implementation (the
C-ABI kernel through ops.py) and a Meta implementation (shape / dtype inference), so FakeTensor propagation,
`torch.export` and `torch.compile` tracing stay legal -- the role `torch.ops.torch_scatter.*` / `torch.ops.pyg.*`
schemas play for the reference's optional extensions.  Forward-only operators: the
differentiable front ends are the Python autograd functions in functional.py (their backward needs the cached
transposed structure of a CSRGraph, which a flat operator schema cannot carry).

    torch.ops.b200mp.spmm_csr(rowptr, col, value?, x, n_rows, reduce)  -> Tensor [n_rows, F]
    torch.ops.b200mp.segment_csr(src, ptr, reduce)                      -> Tensor [len(ptr) - 1, ...]
    torch.ops.b200mp.scatter_coo(src, index, dim_size, reduce)          -> Tensor [dim_size, ...]
    torch.ops.b200mp.softmax_csr(src, ptr)                              -> Tensor like src
    torch.ops.b200mp.index_sort(keys, max_value)                        -> (Tensor sorted, Tensor perm)
"""
from __future__ import annotations

import torch

from .. import ops

_LIB = None


def register() -> bool:
    global _LIB
    if _LIB is not None:
        return True
    lib = torch.library.Library("b200mp", "DEF")
    lib.define("spmm_csr(Tensor rowptr, Tensor col, Tensor? value, Tensor x, int n_rows, str reduce) -> Tensor")
    lib.define("segment_csr(Tensor src, Tensor ptr, str reduce) -> Tensor")
    lib.define("scatter_coo(Tensor src, Tensor index, int dim_size, str reduce) -> Tensor")
    lib.define("softmax_csr(Tensor src, Tensor ptr) -> Tensor")
    lib.define("index_sort(Tensor keys, int max_value) -> (Tensor, Tensor)")

    lib.impl("spmm_csr", lambda rowptr, col, value, x, n_rows, reduce: ops.spmm_csr(rowptr, col, value, x, n_rows, reduce), "CUDA")
    lib.impl("segment_csr", lambda src, ptr, reduce: ops.segment_csr(src, ptr, reduce, ops.segment_plan(ptr, src.size(0))), "CUDA")
    lib.impl("scatter_coo", lambda src, index, dim_size, reduce: ops.scatter_coo(src, index, dim_size, reduce), "CUDA")
    lib.impl("softmax_csr", lambda src, ptr: ops.softmax_csr(src, ptr, ops.segment_plan(ptr, src.size(0))), "CUDA")

    def index_sort_cuda(keys, max_value):
        ks, perm, _ = ops.sort_by_key(keys, int(max_value) + 1, want_sorted=True, want_ptr=False)
        return ks, perm.to(torch.int64)

    lib.impl("index_sort", index_sort_cuda, "CUDA")
    lib.impl("spmm_csr", lambda rowptr, col, value, x, n_rows, reduce: x.new_empty((n_rows, x.size(1))), "Meta")
    lib.impl("segment_csr", lambda src, ptr, reduce: src.new_empty((ptr.numel() - 1, ) + tuple(src.shape[1:])), "Meta")
    lib.impl("scatter_coo", lambda src, index, dim_size, reduce: src.new_empty((dim_size, ) + tuple(src.shape[1:])), "Meta")
    lib.impl("softmax_csr", lambda src, ptr: torch.empty_like(src), "Meta")
    lib.impl("index_sort", lambda keys, max_value: (torch.empty_like(keys), keys.new_empty(keys.shape, dtype=torch.int64)), "Meta")
    _LIB = lib
    return True

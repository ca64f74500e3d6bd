"""ctypes/numpy front end of the CPU oracle (oracle/mp_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of mp_oracle.c.  The product package
(pytorch_geometric_b200/) never imports this module; tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg do, as the checker.  Parity status: PINNED against golden vectors
generated from the reference (tests/golden/make_golden.py) and the reference's own
known-answer tests.

All functions take / return numpy arrays: float32 features, int64 indices.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmp_oracle.so")

REDUCE = {"sum": 0, "add": 0, "mean": 1, "min": 2, "max": 3, "mul": 4}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        for name in ("oracle_add_remaining_self_loops", "oracle_remove_then_add_self_loops",
                     "oracle_gcn_norm", "oracle_gat_attention"):
            getattr(_lib, name).restype = ctypes.c_int64
    return _lib


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


_c = ctypes.c_int64


def degree(index, N):
    index = _i(index)
    out = np.zeros(N, np.int64)
    lib().oracle_degree(_p(index), _c(index.size), _c(N), _p(out))
    return out


def index2ptr(index, N):
    index = _i(index)
    out = np.zeros(N + 1, np.int64)
    lib().oracle_index2ptr(_p(index), _c(index.size), _c(N), _p(out))
    return out


def ptr2index(ptr):
    ptr = _i(ptr)
    N = ptr.size - 1
    out = np.zeros(int(ptr[-1]), np.int64)
    lib().oracle_ptr2index(_p(ptr), _c(N), _p(out))
    return out


def stable_sort_by_key(keys, N):
    keys = _i(keys)
    perm = np.zeros(keys.size, np.int64)
    ptr = np.zeros(N + 1, np.int64)
    lib().oracle_stable_sort_by_key(_p(keys), _c(keys.size), _c(N), _p(perm), _p(ptr))
    return perm, ptr


def add_remaining_self_loops(row, col, w, N, fill_value=1.0):
    row, col, w = _i(row), _i(col), _f(w)
    E = row.size
    r2 = np.zeros(E + N, np.int64)
    c2 = np.zeros(E + N, np.int64)
    w2 = np.zeros(E + N, np.float32) if w is not None else None
    Ep = lib().oracle_add_remaining_self_loops(_p(row), _p(col), _p(w), _c(E), _c(N),
                                               ctypes.c_float(fill_value), _p(r2), _p(c2), _p(w2))
    return r2[:Ep], c2[:Ep], (w2[:Ep] if w is not None else None)


def scatter(src, index, N, reduce="sum"):
    src, index = _f(src), _i(index)
    flat = src.reshape(src.shape[0], -1)
    E, F = flat.shape
    out = np.zeros((N, F), np.float32)
    lib().oracle_scatter(_p(flat), _p(index), _c(E), _c(F), _c(N), REDUCE[reduce], _p(out))
    return out.reshape((N, ) + src.shape[1:])


def scatter_backward(grad_out, src, out, index, reduce="sum"):
    grad_out, src, out, index = _f(grad_out), _f(src), _f(out), _i(index)
    flat = src.reshape(src.shape[0], -1)
    E, F = flat.shape
    N = out.shape[0]
    g = np.zeros((E, F), np.float32)
    lib().oracle_scatter_backward(_p(grad_out), _p(flat), _p(out), _p(index), _c(E), _c(F), _c(N),
                                  REDUCE[reduce], _p(g))
    return g.reshape(src.shape)


def segment(src, ptr, reduce="sum"):
    src, ptr = _f(src), _i(ptr)
    flat = src.reshape(src.shape[0], -1)
    F = flat.shape[1]
    N = ptr.size - 1
    out = np.zeros((N, F), np.float32)
    lib().oracle_segment(_p(flat), _p(ptr), _c(N), _c(F), REDUCE[reduce], _p(out))
    return out.reshape((N, ) + src.shape[1:])


def softmax(src, index, N):
    src, index = _f(src), _i(index)
    flat = src.reshape(src.shape[0], -1)
    E, H = flat.shape
    out = np.zeros((E, H), np.float32)
    lib().oracle_softmax(_p(flat), _p(index), _c(E), _c(H), _c(N), _p(out))
    return out.reshape(src.shape)


def softmax_backward(grad_out, out, index, N):
    grad_out, out, index = _f(grad_out), _f(out), _i(index)
    flat = out.reshape(out.shape[0], -1)
    E, H = flat.shape
    g = np.zeros((E, H), np.float32)
    lib().oracle_softmax_backward(_p(grad_out), _p(flat), _p(index), _c(E), _c(H), _c(N), _p(g))
    return g.reshape(out.shape)


def gather_scatter(x, src, dst, w, N_dst, reduce="sum"):
    x, src, dst, w = _f(x), _i(src), _i(dst), _f(w)
    F = x.shape[1]
    out = np.zeros((N_dst, F), np.float32)
    lib().oracle_gather_scatter(_p(x), _p(src), _p(dst), _p(w), _c(src.size), _c(F), _c(N_dst),
                                REDUCE[reduce], _p(out))
    return out


def gather_scatter_backward(grad_out, x, out, src, dst, w, reduce="sum", need_grad_w=False):
    grad_out, x, out, src, dst, w = _f(grad_out), _f(x), _f(out), _i(src), _i(dst), _f(w)
    F = x.shape[1]
    gx = np.zeros_like(x)
    gw = np.zeros(src.size, np.float32) if need_grad_w else None
    lib().oracle_gather_scatter_backward(_p(grad_out), _p(x), _p(out), _p(src), _p(dst), _p(w),
                                         _c(src.size), _c(F), _c(x.shape[0]), _c(out.shape[0]),
                                         REDUCE[reduce], _p(gx), _p(gw))
    return gx, gw


def spmm_csr(rowptr, col, val, x, reduce="sum"):
    rowptr, col, val, x = _i(rowptr), _i(col), _f(val), _f(x)
    N = rowptr.size - 1
    F = x.shape[1]
    out = np.zeros((N, F), np.float32)
    lib().oracle_spmm_csr(_p(rowptr), _p(col), _p(val), _p(x), _c(N), _c(F), REDUCE[reduce], _p(out))
    return out


def gcn_norm(row, col, w, N, improved=False, add_self_loops=True):
    row, col, w = _i(row), _i(col), _f(w)
    E = row.size
    r2 = np.zeros(E + N, np.int64)
    c2 = np.zeros(E + N, np.int64)
    w2 = np.zeros(E + N, np.float32)
    Ep = lib().oracle_gcn_norm(_p(row), _p(col), _p(w), _c(E), _c(N), int(improved),
                               int(add_self_loops), _p(r2), _p(c2), _p(w2))
    return r2[:Ep], c2[:Ep], w2[:Ep]


def linear(a, weight, bias=None):
    a, weight, bias = _f(a), _f(weight), _f(bias)
    M, K = a.shape
    Nn = weight.shape[0]
    out = np.zeros((M, Nn), np.float32)
    lib().oracle_linear(_p(a), _p(weight), _p(bias), _c(M), _c(K), _c(Nn), _p(out))
    return out


def gcn_conv(x, row, col, w, weight, bias, improved=False, add_self_loops=True):
    x, row, col, w, weight, bias = _f(x), _i(row), _i(col), _f(w), _f(weight), _f(bias)
    N, Fin = x.shape
    Fout = weight.shape[0]
    out = np.zeros((N, Fout), np.float32)
    lib().oracle_gcn_conv(_p(x), _p(row), _p(col), _p(w), _p(weight), _p(bias), _c(N), _c(row.size),
                          _c(Fin), _c(Fout), int(improved), int(add_self_loops), _p(out))
    return out


def gcn_conv_backward(grad_out, x, row, col, w, weight, improved=False, add_self_loops=True):
    grad_out, x, row, col, w, weight = _f(grad_out), _f(x), _i(row), _i(col), _f(w), _f(weight)
    N, Fin = x.shape
    Fout = weight.shape[0]
    gx = np.zeros_like(x)
    gw = np.zeros_like(weight)
    gb = np.zeros(Fout, np.float32)
    lib().oracle_gcn_conv_backward(_p(grad_out), _p(x), _p(row), _p(col), _p(w), _p(weight), _c(N),
                                   _c(row.size), _c(Fin), _c(Fout), int(improved),
                                   int(add_self_loops), _p(gx), _p(gw), _p(gb))
    return gx, gw, gb


def sage_conv(x, row, col, w_l, b_l, w_r, reduce="mean"):
    x, row, col, w_l, b_l, w_r = _f(x), _i(row), _i(col), _f(w_l), _f(b_l), _f(w_r)
    N, Fin = x.shape
    Fout = w_l.shape[0]
    out = np.zeros((N, Fout), np.float32)
    lib().oracle_sage_conv(_p(x), _p(row), _p(col), _p(w_l), _p(b_l), _p(w_r), _c(N), _c(row.size),
                           _c(Fin), _c(Fout), REDUCE[reduce], _p(out))
    return out


def gin_aggregate(x, row, col, eps=0.0):
    x, row, col = _f(x), _i(row), _i(col)
    N, F = x.shape
    out = np.zeros((N, F), np.float32)
    lib().oracle_gin_aggregate(_p(x), _p(row), _p(col), _c(N), _c(row.size), _c(F),
                               ctypes.c_float(eps), _p(out))
    return out


def gat_attention(xh, att_src, att_dst, row, col, slope=0.2, add_self_loops=True):
    """xh: [N,H,C].  Returns (out [N,H*C], alpha [E',H], row', col')."""
    xh, att_src, att_dst, row, col = _f(xh), _f(att_src), _f(att_dst), _i(row), _i(col)
    N, H, C = xh.shape
    E = row.size
    r2 = np.zeros(E + N, np.int64)
    c2 = np.zeros(E + N, np.int64)
    alpha = np.zeros((E + N, H), np.float32)
    out = np.zeros((N, H * C), np.float32)
    Ep = lib().oracle_gat_attention(_p(xh), _p(att_src), _p(att_dst), _p(row), _p(col), _c(N), _c(E),
                                    _c(H), _c(C), ctypes.c_float(slope), int(add_self_loops),
                                    _p(r2), _p(c2), _p(alpha), _p(out))
    return out, alpha[:Ep], r2[:Ep], c2[:Ep]


def rgcn_conv(x, row, col, edge_type, weight, root, bias, reduce="mean"):
    x, row, col, edge_type = _f(x), _i(row), _i(col), _i(edge_type)
    weight, root, bias = _f(weight), _f(root), _f(bias)
    N, Fin = x.shape
    R, _, Fout = weight.shape
    out = np.zeros((N, Fout), np.float32)
    lib().oracle_rgcn_conv(_p(x), _p(row), _p(col), _p(edge_type), _p(weight), _p(root), _p(bias),
                           _c(N), _c(row.size), _c(R), _c(Fin), _c(Fout), REDUCE[reduce], _p(out))
    return out


FUSABLE = ("sum", "mean", "min", "max", "var", "std")


def fused_aggregation(x, index, N, aggrs):
    """torch_geometric/nn/aggr/fused.py:191-336 -- FusedAggregation.forward: one scatter per base reduction
    (sum, x*x sum, min, max) plus the shared clamp(count, 1); mean (:243-256), var = pow_sum/count - mean*mean
    (:258-283), std = sqrt(clamp(var, 1e-5)) with values <= sqrt(1e-5) set to 0 (:319-323).  fp32 throughout,
    every intermediate rounded like the reference's separate ATen ops."""
    x, index = _f(x), _i(index)
    if x.shape[0] == 0:                      # test/nn/aggr/test_fused.py:46-55
        return [np.zeros((N, x.shape[1]), np.float32) for _ in aggrs]
    cnt = np.maximum(degree(index, N).astype(np.float32), np.float32(1)).reshape(-1, 1)
    out = {}
    s = scatter(x, index, N, "sum")
    mean = (s / cnt).astype(np.float32)
    if "var" in aggrs or "std" in aggrs:
        pow_sum = scatter((x * x).astype(np.float32), index, N, "sum")
        var = ((pow_sum / cnt).astype(np.float32) - (mean * mean).astype(np.float32)).astype(np.float32)
        sd = np.sqrt(np.maximum(var, np.float32(1e-5))).astype(np.float32)
        sd = np.where(sd <= np.float32(np.sqrt(1e-5)), np.float32(0), sd).astype(np.float32)
        out["var"], out["std"] = var, sd
    out["sum"], out["mean"] = s, mean
    if "min" in aggrs:
        out["min"] = scatter(x, index, N, "min")
    if "max" in aggrs:
        out["max"] = scatter(x, index, N, "max")
    return [out[a] for a in aggrs]


def fused_aggregation_backward(grads, x, index, N, aggrs, semi_grad=False):
    """Gradient of fused_aggregation wrt x, term by term as autograd derives it from fused.py:
    sum/mean -> gather (scatter_add_ backward), min/max -> ATen scatter_reduce rule
    (oracle_scatter_backward), var -> 2 x g/cnt (dropped under semi_grad, basic.py:106-110) - 2 mean g/cnt,
    std -> g / (2 std) where var >= 1e-5 survived the mask."""
    x, index = _f(x), _i(index)
    cnt = np.maximum(degree(index, N).astype(np.float64), 1.0).reshape(-1, 1)
    outs = dict(zip(aggrs, fused_aggregation(x, index, N, aggrs)))
    mean = scatter(x, index, N, "sum").astype(np.float64) / cnt
    gx = np.zeros(x.shape, np.float64)
    gvar = np.zeros((N, x.shape[1]), np.float64)
    for a, g in zip(aggrs, grads):
        g = _f(g).astype(np.float64)
        if a == "sum":
            gx += g[index]
        elif a == "mean":
            gx += (g / cnt)[index]
        elif a in ("min", "max"):
            gx += scatter_backward(g.astype(np.float32), x, outs[a], index, a).astype(np.float64)
        elif a == "var":
            gvar += g
        elif a == "std":
            sd = outs["std"].astype(np.float64)
            gvar += np.where(sd > 0, g * 0.5 / np.where(sd > 0, sd, 1.0), 0.0)
    if "var" in aggrs or "std" in aggrs:
        gx += (-2.0 * gvar * mean / cnt)[index]
        if not semi_grad:
            gx += 2.0 * x.astype(np.float64) * (gvar / cnt)[index]
    return gx.astype(np.float32)


def remove_then_add_self_loops(row, col, N):
    """torch_geometric/utils/loop.py:71-131 + 382-492 as GATConv/GATv2Conv call them (gat_conv.py:342-346,
    gatv2_conv.py:310-316): every existing loop is dropped, then one loop per node is appended in node order."""
    row, col = _i(row), _i(col)
    keep = row != col
    loops = np.arange(N, dtype=np.int64)
    return np.concatenate([row[keep], loops]), np.concatenate([col[keep], loops])


def gatv2_attention(x_l, x_r, att, row, col, slope=0.2, add_self_loops=True):
    """torch_geometric/nn/conv/gatv2_conv.py:310-331,356-378 -- GATv2Conv after the two linear maps:
    e = (leaky_relu(x_r[i] + x_l[j]) * att).sum(-1) per head, alpha = softmax over the in-edges of i
    (utils/_softmax.py:82-88), out[i] = sum_j alpha * x_l[j].  x_l, x_r: [N, H, C]; att: [H, C].
    Groundwork for SURVEY section 8(f) rank 2 (no CUDA path yet).  Returns (out [N, H*C], alpha [E', H], row', col')."""
    x_l, x_r, att = _f(x_l), _f(x_r), _f(att)
    N, H, C = x_l.shape
    row, col = _i(row), _i(col)
    if add_self_loops:
        row, col = remove_then_add_self_loops(row, col, N)
    s = (x_r[col] + x_l[row]).astype(np.float32)
    s = np.where(s > 0, s, np.float32(slope) * s).astype(np.float32)
    e = (s * att[None]).astype(np.float32).sum(-1, dtype=np.float32)
    alpha = softmax(e, col, N)
    msg = (x_l[row] * alpha[:, :, None]).astype(np.float32).reshape(row.size, H * C)
    return scatter(msg, col, N, "sum"), alpha, row, col

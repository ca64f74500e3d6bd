"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs, against the committed golden fixtures, and -- at large sizes -- through size-independent
properties.  Bar: bit-exact for integer/index work; <= 1e-5 relative for fp32 aggregation
(stated per test); bf16 within bf16 rounding (1e-2)."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import oracle as O

pytestmark = pytest.mark.gpu

import pytorch_geometric_b200 as pgb  # noqa: E402
from pytorch_geometric_b200 import ops, utils as U  # noqa: E402
from pytorch_geometric_b200.graph import CSRGraph  # noqa: E402
from pytorch_geometric_b200.nn import (GCNConv, GINConv, MaxAggregation, MeanAggregation, MinAggregation,  # noqa: E402
                                       RGCNConv, SAGEConv, SoftmaxAggregation, SumAggregation)

DEV = "cuda"
RTOL = 1e-5  # BASELINE.json north_star: "within 1e-5 relative for fp32 aggregation"


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def npy(t):
    return t.detach().float().cpu().numpy()


def power_law_graph(rng, N, E, hubs=True):
    """dst from a truncated power law (hubs), src uniform; includes duplicates, self loops and
    isolated nodes."""
    if hubs:
        ranks = rng.zipf(2.1, size=E).astype(np.int64)
        dst = (ranks - 1) % max(N - 2, 1)          # node N-1 (and maybe more) stays isolated
        dst = rng.permutation(N)[dst]
    else:
        dst = rng.integers(0, max(N - 1, 1), size=E)
    src = rng.integers(0, N, size=E)
    return src.astype(np.int64), dst.astype(np.int64)


def sum_scale(x, src, dst, w, N):
    """sum over |terms| per output row: the natural scale for a relative error of a sum."""
    return O.gather_scatter(np.abs(x), src, dst, None if w is None else np.abs(w), N, "sum")


# ------------------------------------------------------------------ integer / index work (bit-exact)
@pytest.mark.parametrize("idt", [torch.int32, torch.int64])
def test_structure_bit_exact(idt):
    rng = np.random.default_rng(0)
    N, E = 1000, 20000
    src, dst = power_law_graph(rng, N, E)
    d = cu(dst, idt)
    assert np.array_equal(npy(ops.degree(d, N)).astype(np.int64), O.degree(dst, N))
    ks, perm, ptr = ops.sort_by_key(d, N)
    operm, optr = O.stable_sort_by_key(dst, N)
    assert np.array_equal(perm.cpu().numpy().astype(np.int64), operm)
    assert np.array_equal(ptr.cpu().numpy().astype(np.int64), optr)
    assert np.array_equal(ks.cpu().numpy().astype(np.int64), dst[operm])
    assert np.array_equal(ops.index2ptr(ks, N).cpu().numpy().astype(np.int64), O.index2ptr(dst[operm], N))
    assert np.array_equal(ops.ptr2index(ptr, E).cpu().numpy().astype(np.int64), O.ptr2index(optr))
    mn, mx, srt = ops.index_stats(d)
    assert (mn, mx, srt) == (int(dst.min()), int(dst.max()), False)
    assert ops.index_stats(ks)[2] is True
    assert np.array_equal(ops.permute(cu(src, idt), perm).cpu().numpy().astype(np.int64), src[operm])


def test_structure_golden():
    g = load_golden("structure")
    ptr = ops.index2ptr(cu(g["idx_sorted"]), 11)
    assert np.array_equal(ptr.cpu().numpy(), g["ptr"])
    assert np.array_equal(ops.ptr2index(ptr).cpu().numpy(), g["ptr2idx"])
    _, perm, _ = ops.sort_by_key(cu(g["ei"][1]), 11)
    assert np.array_equal(perm.cpu().numpy(), g["stable_perm"])
    assert np.array_equal(ops.degree(cu(g["deg_index"]), 3).cpu().numpy(), [3, 1, 1])
    ei, w = U.add_remaining_self_loops(cu(np.stack([g["row"], g["col"]])), cu(g["w"]), 1.0, 3)
    assert np.array_equal(ei.cpu().numpy(), g["asl_ei"])
    assert np.array_equal(npy(w), g["asl_w"])
    ei, w = U.add_remaining_self_loops(cu(g["ei"]), cu(g["wr"]), 2.0, 11)
    assert np.array_equal(ei.cpu().numpy(), g["asl_eiR"])
    assert np.array_equal(npy(w), g["asl_wR"])
    ei, _ = U.add_remaining_self_loops(cu(np.stack([g["row"], g["col"]])), None, None, 3)
    assert np.array_equal(ei.cpu().numpy(), g["asl_ei_now"])


def test_self_loops_vs_oracle_random():
    rng = np.random.default_rng(1)
    N, E = 500, 6000
    src, dst = power_law_graph(rng, N, E)
    dst[:200] = src[:200]                      # many (duplicated) self loops
    w = rng.random(E).astype(np.float32)
    r, c, ww = ops.self_loops(cu(src), cu(dst), cu(w), N, 2.0, 0)
    orr, oc, ow = O.add_remaining_self_loops(src, dst, w, N, 2.0)
    assert np.array_equal(r.cpu().numpy(), orr) and np.array_equal(c.cpu().numpy(), oc)
    assert np.array_equal(npy(ww), ow)         # bit-exact incl. the "last duplicate wins" rule
    r, c, _ = ops.self_loops(cu(src), cu(dst), None, N, 1.0, 1)
    assert np.array_equal(r.cpu().numpy()[-N:], np.arange(N))
    assert (r != c).sum().item() == (src != dst).sum()


@pytest.mark.parametrize("tag,use_w,improved,asl", [("a", False, False, True), ("b", True, False, True),
                                                     ("c", True, True, True), ("d", True, False, False)])
def test_gcn_norm_golden(tag, use_w, improved, asl):
    g = load_golden("gcn_norm")
    ei, w = U.gcn_norm(cu(g["ei"]), cu(g["wr"]) if use_w else None, int(g["N"]), improved, asl)
    assert np.array_equal(ei.cpu().numpy(), g["ei_" + tag])          # integer work: exact
    assert_close(npy(w), g["w_" + tag], rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------ the hot path vs the oracle
FEATS = [1, 3, 4, 8, 16, 20, 32, 64, 128, 256, 512, 640]


@pytest.mark.parametrize("feat", FEATS)
@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
def test_aggregate_vs_oracle(feat, reduce):
    rng = np.random.default_rng(feat * 7 + len(reduce))
    N, E = 300, 4000
    src, dst = power_law_graph(rng, N, E)
    x = rng.standard_normal((N, feat)).astype(np.float32)
    w = (rng.random(E) + 0.5).astype(np.float32)
    for weights in (None, w):
        for idt, chunk in ((torch.int32, 512), (torch.int64, 16)):     # chunk=16 forces the long-row path
            g = CSRGraph(cu(src), cu(dst), N, N, None if weights is None else cu(weights), chunk=chunk, idx_dtype=idt)
            if chunk == 16:
                assert g.plan.n_long > 0
            out = npy(pgb.aggregate(g, cu(x), reduce))
            ref = O.gather_scatter(x, src, dst, weights, N, reduce)
            scale = sum_scale(x, src, dst, weights, N) if reduce in ("sum", "mean") else np.abs(ref)
            err = np.abs(out - ref)
            assert (err <= RTOL * scale + 1e-30).all(), f"max rel err {np.max(err / (scale + 1e-30)):.2e}"
            if reduce in ("min", "max"):
                assert np.array_equal(out, ref)
            elif feat % 4 == 0:
                # rows walked by a single lane group (deg <= chunk): CSR order == the reference's edge
                # order and products are rounded before the add => bit-identical to the CPU reference
                short = O.degree(dst, N) <= chunk
                assert np.array_equal(out[short], ref[short])


@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
@pytest.mark.parametrize("feat", [5, 64, 256])
def test_aggregate_backward_vs_oracle(reduce, feat):
    rng = np.random.default_rng(11 + feat)
    N, E = 200, 3000
    src, dst = power_law_graph(rng, N, E)
    x = (np.round(rng.standard_normal((N, feat)) * 2) / 2).astype(np.float32)   # coarse grid: real ties, exact 0s
    gout = rng.standard_normal((N, feat)).astype(np.float32)
    g = CSRGraph(cu(src), cu(dst), N, N, chunk=32)
    xt = cu(x).requires_grad_()
    out = pgb.aggregate(g, xt, reduce)
    out.backward(cu(gout))
    oout = O.gather_scatter(x, src, dst, None, N, reduce)
    gx, _ = O.gather_scatter_backward(gout, x, oout, src, dst, None, reduce)
    scale = O.gather_scatter(np.abs(gout), dst, src, None, N, "sum") + 1e-30
    assert (np.abs(npy(xt.grad) - gx) <= RTOL * scale).all()


def test_aggregate_weight_grad_vs_oracle():
    rng = np.random.default_rng(5)
    N, E, feat = 150, 2000, 48
    src, dst = power_law_graph(rng, N, E)
    x = rng.standard_normal((N, feat)).astype(np.float32)
    w = (rng.random(E) + 0.5).astype(np.float32)
    gout = rng.standard_normal((N, feat)).astype(np.float32)
    g = CSRGraph(cu(src), cu(dst), N, N)
    xt, wt = cu(x).requires_grad_(), cu(w).requires_grad_()
    out = pgb.aggregate(g, xt, "sum", edge_weight=wt)
    out.backward(cu(gout))
    oout = O.gather_scatter(x, src, dst, w, N, "sum")
    gx, gw = O.gather_scatter_backward(gout, x, oout, src, dst, w, "sum", need_grad_w=True)
    # relative to the sum of |terms| (hub rows are summed chunk-wise, i.e. in a different order)
    assert (np.abs(npy(out) - oout) <= RTOL * sum_scale(x, src, dst, w, N) + 1e-30).all()
    gscale = O.gather_scatter(np.abs(gout), dst, src, np.abs(w), N, "sum") + 1e-30
    assert (np.abs(npy(xt.grad) - gx) <= RTOL * gscale).all()
    wscale = (np.abs(gout[dst]) * np.abs(x[src])).sum(1) + 1e-30
    assert (np.abs(npy(wt.grad) - gw) <= 1e-5 * wscale).all()


@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
def test_spmm_golden(reduce):
    g = load_golden("spmm")
    N = int(g["N"])
    ei = g["ei_sorted"]               # EdgeIndex.matmul: out[row] = reduce x[col]  => dst=row, src=col
    graph = CSRGraph(cu(ei[1]), cu(ei[0]), N, N)
    x = cu(g["x"]).requires_grad_()
    out = U.spmm(graph, x, reduce)
    assert_close(npy(out), g["out_" + reduce], rtol=RTOL, atol=1e-6)
    out.backward(cu(g["gout_" + reduce]))
    assert_close(npy(x.grad), g["gx_" + reduce], rtol=RTOL, atol=1e-6)


def test_spmm_weighted_golden():
    g = load_golden("spmm")
    N = int(g["N"])
    ei = g["ei_sorted"]
    graph = CSRGraph(cu(ei[1]), cu(ei[0]), N, N)
    x, v = cu(g["x"]).requires_grad_(), cu(g["val_sorted"]).requires_grad_()
    out = pgb.aggregate(graph, x, "sum", edge_weight=v)
    assert_close(npy(out), g["out_wsum"], rtol=RTOL, atol=1e-6)
    out.backward(cu(g["gout_wsum"]))
    assert_close(npy(x.grad), g["gx_wsum"], rtol=RTOL, atol=1e-6)
    assert_close(npy(v.grad), g["gval_wsum"], rtol=RTOL, atol=1e-5)
    csr = torch.sparse_csr_tensor(ops.index2ptr(cu(ei[0]), N), cu(ei[1]), cu(g["val_sorted"]), (N, N))
    assert_close(npy(U.spmm(csr, cu(g["x"]), "sum")), g["out_spmm_wsum"], rtol=RTOL, atol=1e-6)
    assert_close(npy(U.spmm(csr, cu(g["x"]), "mean")), g["out_spmm_wmean"], rtol=RTOL, atol=1e-6)


def test_aggregate_bf16():
    rng = np.random.default_rng(3)
    N, E, feat = 400, 6000, 128
    src, dst = power_law_graph(rng, N, E)
    xb = torch.from_numpy(rng.standard_normal((N, feat)).astype(np.float32)).bfloat16()
    g = CSRGraph(cu(src), cu(dst), N, N)
    for reduce in ("sum", "mean", "max"):
        out = pgb.aggregate(g, xb.to(DEV), reduce).float().cpu().numpy()
        ref = O.gather_scatter(xb.float().numpy(), src, dst, None, N, reduce)   # fp32 oracle on bf16-rounded inputs
        scale = sum_scale(xb.float().numpy(), src, dst, None, N) if reduce != "max" else np.abs(ref)
        assert (np.abs(out - ref) <= 1e-2 * scale + 1e-30).all()


def test_empty_and_ragged_inputs():
    x = torch.randn(5, 8, device=DEV)
    empty = torch.empty(0, dtype=torch.int64, device=DEV)
    g = CSRGraph(empty, empty, 5, 5)
    for reduce in ("sum", "mean", "min", "max"):
        assert torch.equal(pgb.aggregate(g, x, reduce), torch.zeros(5, 8, device=DEV))
    g = CSRGraph(torch.tensor([0, 1], device=DEV), torch.tensor([4, 4], device=DEV), 5, 5)
    out = pgb.aggregate(g, x, "max")
    assert torch.equal(out[:4], torch.zeros(4, 8, device=DEV))
    assert torch.equal(out[4], torch.maximum(x[0], x[1]))
    with pytest.raises(RuntimeError):
        pgb.aggregate(g, x.cpu(), "sum")                  # no CPU fallback
    with pytest.raises(ValueError):
        pgb.aggregate(g, x, "prod")


# ------------------------------------------------------------------ scatter / segment / softmax
@pytest.mark.parametrize("F", [1, 5])
@pytest.mark.parametrize("red", ["sum", "mean", "min", "max", "mul"])
def test_scatter_golden(F, red):
    g = load_golden(f"scatter_F{F}")
    N = int(g["N"])
    src = cu(g["src"]).requires_grad_(red != "mul")
    out = U.scatter(src, cu(g["index"]), 0, N, red)
    assert_close(npy(out), g["out_" + red], rtol=RTOL, atol=1e-6)
    if red != "mul":
        out.backward(cu(g["gout_" + red]))
        assert_close(npy(src.grad), g["gsrc_" + red], rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("feat", [1, 4, 7, 64])
@pytest.mark.parametrize("red", ["sum", "mean", "min", "max", "mul"])
def test_scatter_coo_and_sorted_vs_oracle(feat, red):
    rng = np.random.default_rng(feat)
    N, E = 120, 2500
    index = rng.integers(0, N - 3, size=E)
    src = rng.standard_normal((E, feat)).astype(np.float32)
    if red == "mul":
        src = (1.0 + 0.01 * src).astype(np.float32)
    ref = O.scatter(src, index, N, red)
    out = npy(U.scatter(cu(src), cu(index), 0, N, red))
    tol = 1e-4 if red == "mul" else RTOL
    scale = O.scatter(np.abs(src), index, N, "sum") if red in ("sum", "mean") else np.abs(ref)
    assert (np.abs(out - ref) <= tol * scale + 1e-30).all()
    if red != "mul":
        order = np.argsort(index, kind="stable")
        out_s = npy(U.scatter(cu(src[order]), cu(index[order]), 0, N, red, sorted=True))
        assert (np.abs(out_s - ref) <= tol * scale + 1e-30).all()
    # dim != 0 and error behaviour (test/utils/test_scatter.py:13-24)
    out_t = npy(U.scatter(cu(src.T.copy()), cu(index), 1, N, red))
    assert (np.abs(out_t.T - ref) <= tol * scale + 1e-30).all()
    with pytest.raises(ValueError, match="must be one-dimensional"):
        U.scatter(cu(src), cu(index).view(-1, 1), 0, N)
    with pytest.raises(ValueError, match="must lay between"):
        U.scatter(cu(src), cu(index), 2, N)
    with pytest.raises(ValueError, match="invalid `reduce` argument"):
        U.scatter(cu(src), cu(index), 0, N, "std")


def test_scatter_mul_backward_and_any_follow_the_aten_rules():
    """The two reduce modes that used to be gaps: `mul` = new_ones().scatter_reduce_('prod', include_self=True) with ATen's
    zero-aware backward, `any` = new_zeros().scatter_() (utils/_scatter.py:75-77,130-135).  Reference = those ATen calls on
    the CPU (where `any` deterministically keeps the last member in index order)."""
    rng = np.random.default_rng(4)
    N, E, F = 60, 900, 5
    index = torch.from_numpy(rng.integers(0, N - 4, size=E))
    src = torch.from_numpy((1.0 + 0.3 * rng.standard_normal((E, F))).astype(np.float32))
    src[rng.random((E, F)) < 0.02] = 0.0                                   # groups with one zero, a few with several
    src[index == 7] = src[index == 7].abs() + 0.1
    src[torch.nonzero(index == 7)[0], 2] = 0.0                            # exactly one zero in group 7, feature 2
    gout = torch.from_numpy(rng.standard_normal((N, F)).astype(np.float32))
    # ---- mul
    a = src.clone().requires_grad_()
    want = a.new_ones(N, F).scatter_reduce_(0, index.view(-1, 1).expand(-1, F), a, "prod", include_self=True)
    want.backward(gout)
    b = src.clone().to(DEV).requires_grad_()
    got = U.scatter(b, index.to(DEV), 0, N, "mul")
    got.backward(gout.to(DEV))
    assert_close(npy(got), want.detach().numpy(), rtol=2e-5, atol=1e-6)
    scale = float(a.grad.abs().max())
    assert_close(npy(b.grad), a.grad.numpy(), rtol=1e-4, atol=1e-5 * scale)
    assert (npy(b.grad)[a.grad.numpy() == 0] == 0).all()                   # groups with >= 2 zeros: exactly zero
    # ---- any
    a = src.clone().requires_grad_()
    want = a.new_zeros(N, F).scatter_(0, index.view(-1, 1).expand(-1, F), a)
    want.backward(gout)
    b = src.clone().to(DEV).requires_grad_()
    got = U.scatter(b, index.to(DEV), 0, N, "any")
    got.backward(gout.to(DEV))
    assert np.array_equal(npy(got), want.detach().numpy())
    assert np.array_equal(npy(b.grad), a.grad.numpy())
    got_t = U.scatter(src.T.contiguous().to(DEV), index.to(DEV), 1, N, "any")        # dim != 0
    assert np.array_equal(npy(got_t).T, want.detach().numpy())


@pytest.mark.parametrize("red", ["sum", "mean", "min", "max"])
def test_segment_golden_and_dense(red):
    g = load_golden("segment")
    assert_close(npy(U.segment(cu(g["src"]), cu(g["ptr"]), red)), g["out_" + red], rtol=RTOL, atol=1e-6)
    rng = np.random.default_rng(0)
    src = rng.standard_normal((2000, 16)).astype(np.float32)
    ptr = np.array([0, 0, 500, 1000, 1500, 2000])
    out = npy(U.segment(cu(src), cu(ptr), red))
    assert_close(out, O.segment(src, ptr, red), rtol=RTOL, atol=1e-4)
    st = cu(src).requires_grad_()
    U.segment(st, cu(ptr), red).sum().backward()
    assert st.grad.shape == st.shape


def test_softmax_golden_and_oracle():
    g = load_golden("softmax")
    assert_close(npy(U.softmax(cu(g["src1"]), cu(g["index1"]))), [0.5, 0.5, 1, 1])
    assert_close(npy(U.softmax(cu(g["src1"]), None, cu(g["ptr1"]))), [0.5, 0.5, 1, 1])
    N = int(g["N"])
    src = cu(g["src"]).requires_grad_()
    out = U.softmax(src, None, cu(g["ptr"]))
    assert_close(npy(out), g["out"], rtol=RTOL, atol=1e-7)
    out.backward(cu(g["gout"]))
    assert_close(npy(src.grad), g["gsrc"], rtol=1e-4, atol=1e-6)
    # index path on an unsorted index, vs the oracle
    rng = np.random.default_rng(2)
    E, H, Nn = 5000, 8, 300
    index = rng.integers(0, Nn, size=E)
    s = rng.standard_normal((E, H)).astype(np.float32) * 3
    assert_close(npy(U.softmax(cu(s), cu(index), num_nodes=Nn)), O.softmax(s, index, Nn), rtol=RTOL, atol=1e-7)


def test_aggr_modules_golden():
    g = load_golden("aggr")
    N = int(g["N"])
    x, index, ptr = cu(g["x"]), cu(g["index"]), cu(g["ptr"])
    for name, mod in (("sum", SumAggregation()), ("mean", MeanAggregation()), ("max", MaxAggregation()),
                      ("min", MinAggregation()), ("softmax", SoftmaxAggregation(t=1.0))):
        assert_close(npy(mod(x, index, dim_size=N)), g["out_" + name], rtol=RTOL, atol=1e-6, msg=name)
        assert_close(npy(mod(x, ptr=ptr)), g["out_" + name], rtol=RTOL, atol=1e-6, msg=name + " ptr")
        assert_close(npy(mod(x, index, dim_size=N, index_sorted=True)), g["out_" + name], rtol=RTOL, atol=1e-6)
    # a too-small dim_size: the reference finds it by catching the backend's exception (aggr/base.py:130-139); the
    # engine validates it in debug mode only (no per-call device->host read otherwise) and never writes out of range
    with pgb.debug():
        with pytest.raises(ValueError, match="invalid 'dim_size'"):
            SumAggregation()(x, index, dim_size=2)
    small = SumAggregation()(x, index, dim_size=2)                       # rows >= 2 are dropped, rows < 2 are exact
    assert_close(npy(small), g["out_sum"][:2], rtol=RTOL, atol=1e-6)
    with pytest.raises(ValueError, match="invalid dimension"):
        SumAggregation()(x, index, dim=5)


# ------------------------------------------------------------------ layers vs golden (the reference itself)
def test_gcn_conv_cora_two_layers_golden():
    """BASELINE.json configs[0]: 2-layer GCNConv, Cora-shaped (2708 / 10556 / h=16), fwd + bwd."""
    g = load_golden("gcn_cora")
    c1, c2 = GCNConv(16, 16).to(DEV), GCNConv(16, 16).to(DEV)
    with torch.no_grad():
        c1.lin.weight.copy_(cu(g["w1"])); c1.bias.copy_(cu(g["b1"]))
        c2.lin.weight.copy_(cu(g["w2"])); c2.bias.copy_(cu(g["b2"]))
    x = cu(g["x"]).requires_grad_()
    ei = cu(g["ei"])
    h1 = c1(x, ei)
    out = c2(h1.relu(), ei)
    assert_close(npy(h1), g["h1"], rtol=1e-5, atol=1e-5)
    assert_close(npy(out), g["out"], rtol=1e-5, atol=1e-5)
    out.backward(cu(g["gout"]))
    assert_close(npy(x.grad), g["gx"], rtol=1e-4, atol=1e-5)
    for got, key in ((c1.lin.weight.grad, "gw1"), (c1.bias.grad, "gb1"), (c2.lin.weight.grad, "gw2"), (c2.bias.grad, "gb2")):
        assert_close(npy(got), g[key], rtol=1e-4, atol=1e-4, msg=key)


def test_gcn_conv_weighted_improved_golden():
    g = load_golden("gcn_small")
    conv = GCNConv(5, 7, improved=True).to(DEV)
    with torch.no_grad():
        conv.lin.weight.copy_(cu(g["weight"])); conv.bias.copy_(cu(g["bias"]))
    x = cu(g["x"]).requires_grad_()
    out = conv(x, cu(g["ei"]), cu(g["w"]))
    assert_close(npy(out), g["out"], rtol=1e-5, atol=1e-6)
    out.backward(cu(g["gout"]))
    assert_close(npy(x.grad), g["gx"], rtol=1e-5, atol=1e-6)
    assert_close(npy(conv.lin.weight.grad), g["gweight"], rtol=1e-5, atol=1e-5)
    assert_close(npy(conv.bias.grad), g["gbias"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("aggr", ["mean", "max", "sum"])
def test_sage_conv_golden(aggr):
    g = load_golden("sage_gin")
    conv = SAGEConv(6, 5, aggr=aggr).to(DEV)
    with torch.no_grad():
        conv.lin_l.weight.copy_(cu(g["wl_" + aggr])); conv.lin_l.bias.copy_(cu(g["bl_" + aggr]))
        conv.lin_r.weight.copy_(cu(g["wr_" + aggr]))
    x = cu(g["x"]).requires_grad_()
    out = conv(x, cu(g["ei"]))
    assert_close(npy(out), g["out_" + aggr], rtol=1e-5, atol=1e-6)
    out.backward(cu(g["gout_" + aggr]))
    assert_close(npy(x.grad), g["gx_" + aggr], rtol=1e-5, atol=1e-6)
    assert_close(npy(conv.lin_l.weight.grad), g["gwl_" + aggr], rtol=1e-5, atol=1e-5)
    assert_close(npy(conv.lin_r.weight.grad), g["gwr_" + aggr], rtol=1e-5, atol=1e-5)


def test_gin_conv_golden():
    g = load_golden("sage_gin")
    conv = GINConv(torch.nn.Identity(), eps=0.25).to(DEV)
    assert_close(npy(conv(cu(g["x"]), cu(g["ei"]))), g["gin_out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("aggr", ["mean", "sum"])
def test_rgcn_conv_golden(aggr):
    g = load_golden("rgcn")
    conv = RGCNConv(5, 4, int(g["R"]), aggr=aggr).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(cu(g["weight_" + aggr])); conv.root.copy_(cu(g["root_" + aggr])); conv.bias.copy_(cu(g["bias_" + aggr]))
    x = cu(g["x"]).requires_grad_()
    out = conv(x, cu(g["ei"]), cu(g["et"]))
    assert_close(npy(out), g["out_" + aggr], rtol=1e-5, atol=1e-6)
    out.backward(cu(g["gout_" + aggr]))
    assert_close(npy(x.grad), g["gx_" + aggr], rtol=1e-5, atol=1e-6)
    assert_close(npy(conv.weight.grad), g["gweight_" + aggr], rtol=1e-5, atol=1e-5)
    assert_close(npy(conv.root.grad), g["groot_" + aggr], rtol=1e-5, atol=1e-5)


def test_gcn_conv_vs_oracle_medium():
    rng = np.random.default_rng(9)
    N, E, Fi, Fo = 3000, 40000, 64, 128
    src, dst = power_law_graph(rng, N, E)
    x = rng.standard_normal((N, Fi)).astype(np.float32)
    conv = GCNConv(Fi, Fo).to(DEV)
    with torch.no_grad():
        conv.bias.copy_(torch.randn(Fo) * 0.1)
    W, b = npy(conv.lin.weight), npy(conv.bias)
    xt = cu(x).requires_grad_()
    out = conv(xt, cu(np.stack([src, dst])))
    ref = O.gcn_conv(x, src, dst, None, W, b)
    assert_close(npy(out), ref, rtol=1e-4, atol=1e-4)
    gout = rng.standard_normal((N, Fo)).astype(np.float32)
    out.backward(cu(gout))
    gx, gw, gb = O.gcn_conv_backward(gout, x, src, dst, None, W)
    assert_close(npy(xt.grad), gx, rtol=1e-4, atol=1e-4)
    assert_close(npy(conv.lin.weight.grad), gw, rtol=1e-4, atol=1e-3)
    assert_close(npy(conv.bias.grad), gb, rtol=1e-4, atol=1e-3)


# ------------------------------------------------------------------ size-independent properties at scale
def test_large_graph_properties():
    """A graph too large for the oracle in seconds (4M nodes / 40M edges / F=256, several GB of
    traffic): (1) aggregating ones gives the in-degree exactly; (2) linearity
    A(ax + by) == aA(x) + bA(y); (3) the checksum of the output equals the weighted checksum of
    the gathered rows computed edge-wise by a different kernel (SDDMM against a ones vector);
    (4) forward on the transposed graph is the adjoint: <A x, y> == <x, A^T y>."""
    torch.manual_seed(0)
    N, E, Fdim = 4_000_000, 40_000_000, 256
    gen = torch.Generator(device=DEV).manual_seed(1)
    u = torch.rand(E, device=DEV, generator=gen)
    dst = ((u ** 4) * (N - 1)).long()                       # skewed destinations (hubs near id 0)
    src = torch.randint(0, N, (E, ), device=DEV, generator=gen)
    g = CSRGraph(src, dst, N, N)
    assert g.plan.n_long > 0
    deg = g.in_degree().float()
    ones = torch.ones(N, 4, device=DEV)
    assert torch.equal(pgb.aggregate(g, ones, "sum"), deg.view(-1, 1).expand(N, 4))          # (1) exact
    x = torch.randn(N, Fdim, device=DEV, generator=gen)
    y = torch.randn(N, Fdim, device=DEV, generator=gen)
    ax, ay = pgb.aggregate(g, x, "sum"), pgb.aggregate(g, y, "sum")
    lin = pgb.aggregate(g, 2.0 * x - 0.5 * y, "sum")
    scale = pgb.aggregate(g, 2.0 * x.abs() + 0.5 * y.abs(), "sum") + 1e-20
    assert ((lin - (2.0 * ax - 0.5 * ay)).abs() <= 1e-5 * scale).all()                      # (2)
    dots = ops.sddmm_csr(g.rowptr, g.col, torch.ones(N, Fdim, device=DEV), x)               # sum_f x[src_e, f]
    assert abs(ax.double().sum().item() - dots.double().sum().item()) <= 1e-6 * x.abs().double().sum().item() * 10  # (3)
    g.build_transpose()
    aty = ops.spmm_csr(g.rowptr_t, g.col_t, None, y, N, "sum", g.plan_t)
    lhs, rhs = (ax.double() * y.double()).sum().item(), (x.double() * aty.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * (ax.abs().double() * y.abs().double()).sum().item()      # (4)
    mean = pgb.aggregate(g, x, "mean")
    assert ((mean * deg.clamp(min=1).view(-1, 1) - ax).abs() <= 1e-5 * scale + 1e-4).all()
    mx = pgb.aggregate(g, x, "max")
    assert (mx[deg > 0] >= mean[deg > 0] - 1e-5).all() and (mx[deg == 0] == 0).all()


@pytest.mark.parametrize("n,F,dtype", [(0, 8, torch.float32), (1, 1, torch.float32), (777, 5, torch.float32),
                                       (100003, 256, torch.float32), (50000, 64, torch.bfloat16), (4099, 300, torch.float32)])
def test_column_sum_bias_gradient(n, F, dtype):
    """b200mp_column_sum (the layer's bias gradient) vs an fp64 column sum; deterministic."""
    from pytorch_geometric_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(n + F)
    x = torch.randn(n, F, device=DEV, generator=g).to(dtype)
    got = ops.column_sum(x)
    ref = x.double().sum(0)
    scale = x.double().abs().sum(0) + 1e-30
    assert got.dtype == torch.float32 and got.shape == (F, )
    assert ((got.double() - ref).abs() <= 1e-6 * scale + 1e-30).all()
    assert torch.equal(got, ops.column_sum(x))

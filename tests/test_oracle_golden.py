"""Pins the CPU oracle (oracle/mp_oracle.c) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py) and the reference's own known-answer tests."""
import numpy as np
import pytest

from conftest import assert_close, load_golden
from oracle import oracle as O


@pytest.mark.parametrize("F", [1, 5])
@pytest.mark.parametrize("red", ["sum", "mean", "min", "max", "mul"])
def test_scatter(F, red):
    g = load_golden(f"scatter_F{F}")
    N = int(g["N"])
    out = O.scatter(g["src"], g["index"], N, red)
    assert_close(out, g["out_" + red], msg=red)
    if red != "mul":
        gs = O.scatter_backward(g["gout_" + red], g["src"], out, g["index"], red)
        assert_close(gs, g["gsrc_" + red], msg="bwd " + red)


def test_scatter_max_ties_split_evenly():
    # ATen scatter_reduce backward splits evenly among tied maxima.
    src = np.array([[1., -2.], [1., -3.]], np.float32)
    index = np.array([0, 0])
    out = O.scatter(src, index, 1, "max")
    gs = O.scatter_backward(np.ones((1, 2), np.float32), src, out, index, "max")
    assert_close(gs, [[0.5, 1.0], [0.5, 0.0]])


@pytest.mark.parametrize("red", ["sum", "mean", "min", "max"])
def test_segment(red):
    g = load_golden("segment")
    assert_close(O.segment(g["src"], g["ptr"], red), g["out_" + red], msg=red)


def test_segment_matches_dense_reductions():
    # test/utils/test_segment.py:13-31: empty first segment -> 0
    rng = np.random.default_rng(0)
    src = rng.standard_normal((20, 16)).astype(np.float32)
    ptr = np.array([0, 0, 5, 10, 15, 20])
    out = O.segment(src, ptr, "sum")
    assert_close(out[0], np.zeros(16))
    assert_close(out[1:], src.reshape(4, 5, 16).sum(1), rtol=1e-5, atol=1e-5)
    assert_close(O.segment(src, ptr, "max")[1:], src.reshape(4, 5, 16).max(1))
    assert_close(O.segment(src, ptr, "min")[0], np.zeros(16))


def test_softmax():
    g = load_golden("softmax")
    assert_close(O.softmax(g["src1"], g["index1"], 3), [0.5, 0.5, 1, 1])  # test_softmax.py:12-27
    assert_close(O.softmax(g["src1"], g["index1"], 3), g["out1_ptr"])
    N = int(g["N"])
    out = O.softmax(g["src"], g["index"], N)
    assert_close(out, g["out"])
    assert_close(out, g["out_ptr"])
    assert_close(O.softmax_backward(g["gout"], out, g["index"], N), g["gsrc"], rtol=1e-4, atol=1e-6)


def test_structure():
    g = load_golden("structure")
    assert np.array_equal(O.degree(g["deg_index"], 3), [3, 1, 1])  # test/utils/test_degree.py
    assert np.array_equal(O.degree(g["deg_index"], 3), g["deg"])
    assert np.array_equal(O.index2ptr(g["idx_sorted"], 11), g["ptr"])
    assert np.array_equal(O.ptr2index(g["ptr"]), g["ptr2idx"])
    perm, ptr = O.stable_sort_by_key(g["ei"][1], 11)
    assert np.array_equal(perm, g["stable_perm"])
    assert np.array_equal(ptr, g["ptr"])
    # test/utils/test_loop.py:220-291 shape of the answer; SURVEY section 9 exact values
    r, c, w = O.add_remaining_self_loops(g["row"], g["col"], g["w"], 3, 1.0)
    assert np.array_equal(np.stack([r, c]), g["asl_ei"])
    assert np.array_equal(np.stack([r, c]), [[0, 1, 2, 0, 1, 2], [1, 0, 1, 0, 1, 2]])
    assert_close(w, g["asl_w"])
    assert_close(w, [3, 4, 6, 2, 1, 5])
    r, c, _ = O.add_remaining_self_loops(g["row"], g["col"], None, 3)
    assert np.array_equal(np.stack([r, c]), g["asl_ei_now"])
    r, c, w = O.add_remaining_self_loops(g["ei"][0], g["ei"][1], g["wr"], 11, 2.0)
    assert np.array_equal(np.stack([r, c]), g["asl_eiR"])
    assert_close(w, g["asl_wR"])


@pytest.mark.parametrize("tag,use_w,improved,asl", [("a", False, False, True), ("b", True, False, True),
                                                     ("c", True, True, True), ("d", True, False, False)])
def test_gcn_norm(tag, use_w, improved, asl):
    g = load_golden("gcn_norm")
    r, c, w = O.gcn_norm(g["ei"][0], g["ei"][1], g["wr"] if use_w else None, int(g["N"]), improved, asl)
    assert np.array_equal(np.stack([r, c]), g["ei_" + tag])
    assert_close(w, g["w_" + tag], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("red", ["sum", "mean", "min", "max"])
def test_spmm_and_matmul(red):
    g = load_golden("spmm")
    N = int(g["N"])
    ei, x = g["ei_sorted"], g["x"]
    # EdgeIndex.matmul: out[row] = reduce over col of x[col]  (edge_index.py:1903-1922)
    out = O.gather_scatter(x, ei[1], ei[0], None, N, red)
    assert_close(out, g["out_" + red], msg=red)
    ptr = O.index2ptr(ei[0], N)
    assert_close(O.spmm_csr(ptr, ei[1], None, x, red), g["out_" + red], msg="csr " + red)
    gx, _ = O.gather_scatter_backward(g["gout_" + red], x, out, ei[1], ei[0], None, red)
    assert_close(gx, g["gx_" + red], rtol=1e-5, atol=1e-6, msg="gx " + red)


def test_spmm_weighted():
    g = load_golden("spmm")
    N = int(g["N"])
    ei, x, val = g["ei_sorted"], g["x"], g["val_sorted"]
    out = O.gather_scatter(x, ei[1], ei[0], val, N, "sum")
    assert_close(out, g["out_wsum"])
    ptr = O.index2ptr(ei[0], N)
    assert_close(O.spmm_csr(ptr, ei[1], val, x, "sum"), g["out_spmm_wsum"])
    assert_close(O.spmm_csr(ptr, ei[1], val, x, "mean"), g["out_spmm_wmean"])
    gx, gw = O.gather_scatter_backward(g["gout_wsum"], x, out, ei[1], ei[0], val, "sum", True)
    assert_close(gx, g["gx_wsum"], rtol=1e-5, atol=1e-6)
    assert_close(gw, g["gval_wsum"], rtol=1e-5, atol=1e-6)


def test_aggr_modules():
    g = load_golden("aggr")
    N = int(g["N"])
    for red in ("sum", "mean", "max", "min"):
        assert_close(O.scatter(g["x"], g["index"], N, red), g["out_" + red], msg=red)
        assert_close(O.segment(g["x"], g["ptr"], red), g["out_" + red], msg="ptr " + red)
    # SoftmaxAggregation (aggr/basic.py:196-215): alpha = softmax(x*t); out = sum(x*alpha)
    alpha = O.softmax(g["x"], g["index"], N)
    assert_close(O.scatter(g["x"] * alpha, g["index"], N, "sum"), g["out_softmax"])


def test_gcn_conv_cora_two_layers():
    g = load_golden("gcn_cora")
    ei, x = g["ei"], g["x"]
    h1 = O.gcn_conv(x, ei[0], ei[1], None, g["w1"], g["b1"])
    assert_close(h1, g["h1"], rtol=1e-4, atol=1e-5)
    a1 = np.maximum(h1, 0)
    out = O.gcn_conv(a1, ei[0], ei[1], None, g["w2"], g["b2"])
    assert_close(out, g["out"], rtol=1e-4, atol=1e-5)
    ga1, gw2, gb2 = O.gcn_conv_backward(g["gout"], a1, ei[0], ei[1], None, g["w2"])
    assert_close(gw2, g["gw2"], rtol=1e-4, atol=1e-4)
    assert_close(gb2, g["gb2"], rtol=1e-4, atol=1e-4)
    gh1 = ga1 * (h1 > 0)
    gx, gw1, gb1 = O.gcn_conv_backward(gh1, x, ei[0], ei[1], None, g["w1"])
    assert_close(gx, g["gx"], rtol=1e-4, atol=1e-5)
    assert_close(gw1, g["gw1"], rtol=1e-4, atol=1e-4)
    assert_close(gb1, g["gb1"], rtol=1e-4, atol=1e-4)


def test_gcn_conv_weighted_improved():
    g = load_golden("gcn_small")
    ei = g["ei"]
    out = O.gcn_conv(g["x"], ei[0], ei[1], g["w"], g["weight"], g["bias"], improved=True)
    assert_close(out, g["out"], rtol=1e-5, atol=1e-6)
    gx, gw, gb = O.gcn_conv_backward(g["gout"], g["x"], ei[0], ei[1], g["w"], g["weight"], improved=True)
    assert_close(gx, g["gx"], rtol=1e-5, atol=1e-6)
    assert_close(gw, g["gweight"], rtol=1e-5, atol=1e-5)
    assert_close(gb, g["gbias"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("aggr", ["mean", "max", "sum"])
def test_sage_conv(aggr):
    g = load_golden("sage_gin")
    ei = g["ei"]
    out = O.sage_conv(g["x"], ei[0], ei[1], g["wl_" + aggr], g["bl_" + aggr], g["wr_" + aggr], aggr)
    assert_close(out, g["out_" + aggr], rtol=1e-5, atol=1e-6)


def test_gin_aggregate():
    g = load_golden("sage_gin")
    assert_close(O.gin_aggregate(g["x"], g["ei"][0], g["ei"][1], 0.25), g["gin_out"], rtol=1e-5, atol=1e-6)


def test_gat_attention():
    g = load_golden("gat")
    H, C = 4, 3
    xh = O.linear(g["x"], g["lin"]).reshape(-1, H, C)
    out, alpha, r2, c2 = O.gat_attention(xh, g["att_src"].reshape(H, C), g["att_dst"].reshape(H, C),
                                         g["ei"][0], g["ei"][1], float(g["slope"]))
    assert np.array_equal(np.stack([r2, c2]), g["ei2"])  # test_gat_conv.py:55-58 edge order
    assert_close(alpha, g["alpha"], rtol=1e-5, atol=1e-6)
    assert_close(out + g["bias"], g["out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("aggr", ["mean", "sum"])
def test_rgcn_conv(aggr):
    g = load_golden("rgcn")
    out = O.rgcn_conv(g["x"], g["ei"][0], g["ei"][1], g["et"], g["weight_" + aggr], g["root_" + aggr],
                      g["bias_" + aggr], aggr)
    assert_close(out, g["out_" + aggr], rtol=1e-5, atol=1e-6)


FUSED_CASES = {"all": ["sum", "mean", "min", "max", "var", "std"], "pna": ["mean", "min", "max", "std"],
               "sumstd": ["sum", "std"], "var": ["var"], "minmax": ["min", "max"]}


@pytest.mark.parametrize("case", sorted(FUSED_CASES))
def test_fused_aggregation(case):
    # FusedAggregation of the reference (tests/golden/make_golden_fused.py): empty groups, ties at 0,
    # a constant group (std mask)
    g = load_golden("fused_aggr")
    aggrs, N = FUSED_CASES[case], int(g["N"])
    outs = O.fused_aggregation(g["x"], g["index"], N, aggrs)
    for a, o in zip(aggrs, outs):
        if a == "std":      # ATen's vectorised CPU sqrt is not correctly rounded (1 ulp off near ties)
            assert_close(o, g[f"{case}_out_{a}"], rtol=2.5e-7, atol=0, msg=f"{case}/std")
            assert np.array_equal(o == 0, g[f"{case}_out_{a}"] == 0)
        else:
            assert np.array_equal(o, g[f"{case}_out_{a}"]), f"{case}/{a} not bit-exact"
    gx = O.fused_aggregation_backward([g[f"{case}_gout_{a}"] for a in aggrs], g["x"], g["index"], N, aggrs)
    assert_close(gx, g[f"{case}_gx"], rtol=1e-5, atol=1e-6, msg=case)


def test_fused_aggregation_empty_input():
    # test/nn/aggr/test_fused.py:46-55: no messages at all -> zeros of shape [dim_size, F] per aggregation
    outs = O.fused_aggregation(np.zeros((0, 6), np.float32), np.zeros(0, np.int64), 5, ["mean", "var", "std"])
    assert all(o.shape == (5, 6) and float(np.abs(o).sum()) == 0.0 for o in outs)


# ---------------------------------------------------------------- more of the reference's own known-answer tests
def test_softmax_matches_dense_softmax_forward_and_backward():
    # test/utils/test_softmax.py:30-45: groups of two rows == dense softmax over dim 1, same for the gradient of mean()
    rng = np.random.default_rng(3)
    src = rng.random((4, 8)).astype(np.float32)
    index = np.array([0, 0, 1, 1])
    out = O.softmax(src, index, 2)
    d = src.reshape(2, 2, 8).astype(np.float64)
    e = np.exp(d - d.max(1, keepdims=True))
    dense = e / e.sum(1, keepdims=True)
    assert_close(out, dense.reshape(4, 8), rtol=1e-6, atol=1e-7)
    g = np.full((4, 8), 1.0 / 32, np.float32)                       # d mean() / d out
    dense_grad = dense * (g.reshape(2, 2, 8) - (g.reshape(2, 2, 8) * dense).sum(1, keepdims=True))
    assert_close(O.softmax_backward(g, out, index, 2), dense_grad.reshape(4, 8), rtol=1e-4, atol=1e-8)


def test_softmax_single_group_is_plain_softmax():
    # test/utils/test_softmax.py:48-62 (dim = 0 cases)
    rng = np.random.default_rng(4)
    for shape in ((4, ), (4, 16)):
        src = rng.standard_normal(shape).astype(np.float32)
        d = src.astype(np.float64)
        e = np.exp(d - d.max(0, keepdims=True))
        assert_close(O.softmax(src, np.zeros(4, np.int64), 1).reshape(shape), e / e.sum(0, keepdims=True), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("red", ["sum", "mean"])
def test_spmm_basic_equals_dense_matmul(red):
    # test/utils/test_spmm.py:18-38, including the isolated-node case for `mean`
    rng = np.random.default_rng(5)
    for zero_row in (False, True):
        a = rng.standard_normal((5, 4)).astype(np.float32)
        if zero_row:
            a[0] = 0.0
        other = rng.standard_normal((4, 8)).astype(np.float32)
        rows, cols = np.nonzero(a)
        rowptr = O.index2ptr(rows, 5)
        out = O.spmm_csr(rowptr, cols, a[rows, cols], other, red)
        want = a.astype(np.float64) @ other.astype(np.float64)
        if red == "mean":
            # to_sparse_csr drops explicit zeros, so the mean divides by the stored entries per row (4, or 1 when empty)
            want = want / np.maximum(np.diff(rowptr), 1)[:, None]
        assert out.shape == (5, 8)
        assert_close(out, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("aggr", ["mean", "sum", "max", "min", "var", "std"])
def test_basic_aggregation_index_equals_ptr(aggr):
    # test/nn/aggr/test_basic.py:35-63 (index and ptr give the same answer) and :66-75 (var against its definition)
    rng = np.random.default_rng(6)
    x = rng.standard_normal((6, 16)).astype(np.float32)
    index = np.array([0, 0, 1, 1, 1, 2])
    ptr = np.array([0, 2, 5, 6])
    if aggr in ("var", "std"):
        out = O.fused_aggregation(x, index, 3, [aggr])[0]
        mean = O.scatter(x, index, 3, "mean")
        var = O.scatter(((x - mean[index]) ** 2).astype(np.float32), index, 3, "mean")
        want = var if aggr == "var" else np.where(np.sqrt(np.maximum(var, 1e-5)) <= np.sqrt(1e-5), 0, np.sqrt(np.maximum(var, 1e-5)))
        # E[x^2] - mean^2 cancels: the reference's own test allows 1e-6 on var, which is ~1e-5 on a small std
        assert_close(out, want, rtol=1e-4, atol=1e-6 if aggr == "var" else 2e-5)
    else:
        out = O.scatter(x, index, 3, aggr)
        assert out.shape == (3, 16)
        assert_close(out, O.segment(x, ptr, aggr), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("tag", ["sep", "shared"])
def test_gatv2_attention_groundwork(tag):
    """SURVEY section 8(f) rank 2 (next row, no CUDA path yet): the oracle's GATv2 restatement is already pinned
    to the reference's GATv2Conv (tests/golden/make_golden_gatv2.py), forward incl. the attention weights."""
    g = load_golden("gatv2")
    H, C = int(g["H"]), int(g["C"])
    x = g["x"]
    x_l = (x @ g[f"{tag}_lin_l_w"].T + g[f"{tag}_lin_l_b"]).astype(np.float32).reshape(-1, H, C)
    x_r = (x @ g[f"{tag}_lin_r_w"].T + g[f"{tag}_lin_r_b"]).astype(np.float32).reshape(-1, H, C)
    out, alpha, row, col = O.gatv2_attention(x_l, x_r, g[f"{tag}_att"], g["ei"][0], g["ei"][1])
    assert np.array_equal(np.stack([row, col]), g[f"{tag}_ei2"])          # remove + add self loops, same edge order
    assert_close(alpha, g[f"{tag}_alpha"], rtol=1e-5, atol=1e-7)
    assert_close(out + g[f"{tag}_bias"], g[f"{tag}_out"], rtol=1e-5, atol=1e-6)
    sums = np.zeros((x.shape[0], H)); np.add.at(sums, col, alpha)
    assert_close(sums, np.ones_like(sums), rtol=1e-5, atol=0)            # every node has its self loop

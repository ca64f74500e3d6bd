"""GPU parity of the one-sweep multi-aggregation (csrc/multi_aggr.cu) against the golden run of the
reference's FusedAggregation / MultiAggregation and against the oracle restatement of fused.py:191-336."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import oracle as O

pytestmark = pytest.mark.gpu

import pytorch_geometric_b200 as pgb  # noqa: E402,F401
from pytorch_geometric_b200 import functional as Fn  # noqa: E402
from pytorch_geometric_b200 import ops  # noqa: E402
from pytorch_geometric_b200.graph import CSRGraph  # noqa: E402
from pytorch_geometric_b200.nn import FusedAggregation, MultiAggregation, StdAggregation, VarAggregation  # noqa: E402

DEV = "cuda"
CASES = {"all": ["sum", "mean", "min", "max", "var", "std"], "pna": ["mean", "min", "max", "std"],
         "sumstd": ["sum", "std"], "var": ["var"], "minmax": ["min", "max"]}
STD_RTOL = 2.5e-7     # the reference's CPU sqrt is 1 ulp off near ties; everything else is bit-exact


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().float().cpu().numpy()


def check_outputs(aggrs, outs, ref, exact=True, rtol=1e-5, atol=1e-6):
    for a, o, r in zip(aggrs, outs, ref):
        if a == "std" and exact:
            assert_close(npy(o), r, rtol=STD_RTOL, atol=0, msg="std")
            assert np.array_equal(npy(o) == 0, r == 0)
        elif exact:
            assert np.array_equal(npy(o), r), f"{a} not bit-exact (max err {np.abs(npy(o) - r).max():.3e})"
        else:
            assert_close(npy(o), r, rtol=rtol, atol=atol, msg=a)


@pytest.mark.parametrize("mode", ["unsorted", "sorted", "ptr"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_fused_aggregation_golden(case, mode):
    g = load_golden("fused_aggr")
    aggrs, N = CASES[case], int(g["N"])
    x, index = g["x"], g["index"]
    order = np.arange(index.size)
    if mode != "unsorted":
        order = np.argsort(index, kind="stable")
    xs = cu(x[order]).requires_grad_()
    idx = cu(index[order])
    aggr = FusedAggregation(aggrs)
    if mode == "ptr":
        outs = aggr(xs, idx, ptr=ops.index2ptr(idx, N), dim_size=N)
    else:
        outs = aggr(xs, idx, dim_size=N, index_sorted=(mode == "sorted"))
    check_outputs(aggrs, outs, [g[f"{case}_out_{a}"] for a in aggrs])
    torch.autograd.backward(outs, [cu(g[f"{case}_gout_{a}"]) for a in aggrs])
    assert_close(npy(xs.grad), g[f"{case}_gx"][order], rtol=1e-5, atol=1e-6, msg=f"{case} grad")


def test_multi_aggregation_cat_golden():
    g = load_golden("fused_aggr")
    N = int(g["N"])
    x = cu(g["x"]).requires_grad_()
    m = MultiAggregation(CASES["pna"], mode="cat")
    out = m(x, cu(g["index"]), dim_size=N)
    assert m.get_out_channels(5) == 20
    assert_close(npy(out), g["multi_cat_out"], rtol=STD_RTOL, atol=0)
    out.backward(cu(g["multi_cat_gout"]))
    assert_close(npy(x.grad), g["multi_cat_gx"], rtol=1e-5, atol=1e-6)


def test_fused_matches_the_separate_aggregations_reference_test():
    # test/nn/aggr/test_fused.py:19-43: fused outputs == the individual aggregations, same for gradients
    torch.manual_seed(1)
    x = torch.randn(6, 1, device=DEV)
    y = x.clone()
    index = torch.tensor([0, 0, 1, 1, 1, 3], device=DEV)
    x.requires_grad_(True)
    y.requires_grad_(True)
    aggrs = ["sum", "mean", "min", "max", "var", "std"]
    out = torch.cat(FusedAggregation(aggrs)(x, index), dim=-1)
    from pytorch_geometric_b200.nn import aggregation_resolver
    expected = torch.cat([aggregation_resolver(a)(y, index, index_sorted=True) for a in aggrs], dim=-1)
    assert torch.allclose(out, expected, atol=1e-5)
    out.mean().backward()
    expected.mean().backward()
    assert torch.allclose(x.grad, y.grad, atol=1e-5)


def test_empty_fused_std_aggregation():
    # test/nn/aggr/test_fused.py:46-55
    aggr = FusedAggregation(["mean", "var", "std"])
    x = torch.empty(0, 6, device=DEV)
    index = torch.empty(0, dtype=torch.long, device=DEV)
    out = torch.cat(aggr(x, index, dim_size=5), dim=-1)
    assert out.size() == (5, 18)
    assert float(out.abs().sum()) == 0.0


def test_fused_errors():
    with pytest.raises(ValueError, match="should be a list or tuple"):
        FusedAggregation("sum")
    with pytest.raises(ValueError, match="should not be empty"):
        FusedAggregation([])
    with pytest.raises(ValueError, match="not fusable"):
        FusedAggregation(["sum", "softmax"])


@pytest.mark.parametrize("F,dtype,chunk", [(1, torch.float32, 512), (7, torch.float32, 512), (64, torch.float32, 16),
                                           (256, torch.float32, 16), (132, torch.float32, 512),
                                           (96, torch.float32, 512), (128, torch.float32, 16),
                                           (64, torch.bfloat16, 16), (8, torch.bfloat16, 512)])
def test_gather_mode_vs_oracle(F, dtype, chunk):
    """Graph form: out_k[i] = aggr_k over x[src] of the in-edges of i; hub rows chunked when chunk = 16."""
    rng = np.random.default_rng(F + chunk)
    N, E = 700, 20000
    src = rng.integers(0, N, size=E)
    src[rng.random(E) < 0.02] = 3                                   # one hub SOURCE (~400 out-edges): the backward's index batches
    dst = ((rng.random(E) ** 3) * (N - 5)).astype(np.int64)        # skewed, last rows empty
    x = rng.standard_normal((N, F)).astype(np.float32)
    x[rng.random((N, F)) < 0.2] = 0.0                               # post-ReLU zeros -> ties at 0
    x = torch.from_numpy(x).to(dtype).float().numpy()
    aggrs = ["sum", "mean", "min", "max", "var", "std"]
    order = np.argsort(dst, kind="stable")
    ref = O.fused_aggregation(x[src[order]], dst[order], N, aggrs)
    g = CSRGraph(cu(src), cu(dst), N, N, chunk=chunk)
    if chunk == 16:
        assert g.plan.n_long > 0
    xt = cu(x).to(dtype).requires_grad_()
    outs = Fn.multi_aggregate(g, xt, aggrs)
    with torch.no_grad():                                           # inference sweeps (no tie counts, no hit mask): same values
        for sub in (aggrs, ["min", "max"], ["sum", "std"]):
            for name, o in zip(sub, Fn.multi_aggregate(g, xt, sub)):
                assert torch.equal(o, outs[aggrs.index(name)]), name
    deg = np.bincount(dst, minlength=N)
    short = deg <= chunk
    if dtype == torch.float32:
        check_outputs(aggrs, [o[cu(short)] for o in outs], [r[short] for r in ref])       # CSR order == oracle order
        scale = max(1.0, float(np.abs(ref[0]).max()))
        check_outputs(aggrs, outs, ref, exact=False, rtol=1e-4, atol=1e-5 * scale)       # hub rows: chunk-order sums
    else:
        for a, o, r in zip(aggrs, outs, ref):
            assert np.abs(npy(o) - r).max() <= 1e-2 * max(1.0, np.abs(r).max()), a
    if dtype != torch.float32:
        return
    gouts = [rng.standard_normal((N, F)).astype(np.float32) for _ in aggrs]
    torch.autograd.backward(outs, [cu(go) for go in gouts])
    gmsg = O.fused_aggregation_backward(gouts, x[src[order]], dst[order], N, aggrs)        # per message
    gx = np.zeros((N, F), np.float64)
    np.add.at(gx, src[order], gmsg.astype(np.float64))
    assert_close(npy(xt.grad), gx, rtol=2e-4, atol=2e-4 * max(1.0, np.abs(gx).max() / 10), msg="grad x")


def test_semi_grad_and_single_modules():
    rng = np.random.default_rng(5)
    N, E, F = 50, 900, 12
    index = np.sort(rng.integers(0, N, size=E))
    x = rng.standard_normal((E, F)).astype(np.float32)
    for cls, name in ((VarAggregation, "var"), (StdAggregation, "std")):
        for semi in (False, True):
            xt = cu(x).requires_grad_()
            out = cls(semi_grad=semi)(xt, cu(index), dim_size=N, index_sorted=True)
            ref = O.fused_aggregation(x, index, N, [name])[0]
            assert_close(npy(out), ref, rtol=STD_RTOL if name == "std" else 0, atol=0, msg=name)
            go = rng.standard_normal((N, F)).astype(np.float32)
            out.backward(cu(go))
            gref = O.fused_aggregation_backward([go], x, index, N, [name], semi_grad=semi)
            assert_close(npy(xt.grad), gref, rtol=1e-4, atol=1e-5, msg=f"{name} semi={semi}")


def test_multi_aggregation_modes():
    rng = np.random.default_rng(9)
    N, E, F = 40, 600, 8
    index = cu(rng.integers(0, N, size=E))
    x = cu(rng.standard_normal((E, F)).astype(np.float32))
    outs = FusedAggregation(["sum", "max", "std"])(x, index, dim_size=N)
    for mode in ("sum", "mean", "max", "min", "logsumexp", "std", "var"):
        got = MultiAggregation(["sum", "max", "std"], mode=mode)(x, index, dim_size=N)
        want = getattr(torch, mode)(torch.stack(outs, 0), dim=0)
        want = want if isinstance(want, torch.Tensor) else want[0]
        assert torch.equal(got, want), mode
    proj = MultiAggregation(["sum", "softmax"], mode="proj", mode_kwargs=dict(in_channels=F, out_channels=3)).to(DEV)
    assert proj(x, index, dim_size=N).shape == (N, 3)


@pytest.mark.parametrize("F", [8, 64, 132])
def test_segment_mode_backward_runs_of_messages(F):
    """Segment mode (destination-sorted [E, F] messages, what MultiAggregation / PNAConv pass): the backward walks runs of 8
    consecutive messages per lane group (E not a multiple of 8, runs that straddle several destinations)."""
    rng = np.random.default_rng(F)
    N, E = 900, 20011
    index = np.sort(((rng.random(E) ** 2) * (N - 3)).astype(np.int64))
    x = rng.standard_normal((E, F)).astype(np.float32)
    x[rng.random((E, F)) < 0.2] = 0.0
    aggrs = ["sum", "mean", "min", "max", "var", "std"]
    xt = cu(x).requires_grad_()
    idx = cu(index)
    outs = Fn.multi_aggregate((ops.index2ptr(idx, N), idx, None), xt, aggrs)
    ref = O.fused_aggregation(x, index, N, aggrs)
    check_outputs(aggrs, outs, ref)
    gouts = [rng.standard_normal((N, F)).astype(np.float32) for _ in aggrs]
    torch.autograd.backward(outs, [cu(go) for go in gouts])
    gmsg = O.fused_aggregation_backward(gouts, x, index, N, aggrs)
    assert_close(npy(xt.grad), gmsg, rtol=2e-5, atol=2e-5 * max(1.0, np.abs(gmsg).max() / 10), msg="grad messages")

"""GPU parity of the round-2 dense kernels against fp64 products:
  * pair GEMM (b200mp_gemm_pair_tf32x3): two A streams into one accumulator, two outputs, bias / ReLU epilogue
    (dense.linear with epilogue, linear_pair, matmul_pair) forward and backward;
  * grouped GEMM (b200mp_segment_matmul_tf32x3 == pyg_lib.ops.segment_matmul, nn/dense/linear.py:248-255): ragged
    segments incl. empty ones and segments that are not multiples of the 128-row tile, forward + both gradients;
  * HeteroLinear against the reference's naive per-type loop."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from pytorch_geometric_b200 import dense, ops  # noqa: E402
from pytorch_geometric_b200.nn import HeteroLinear  # noqa: E402

DEV = "cuda"


def _close(a, b, tol=1e-5, what=""):   # 3xTF32: ~2^-21 of sum|terms|; sum|terms| is a few times max|result| here
    a, b = a.detach().double(), b.detach().double()
    err = (a - b).abs().max().item()
    assert err <= tol * max(b.abs().max().item(), 1e-6), f"{what}: max err {err:.3e} (scale {b.abs().max().item():.3e})"


@pytest.mark.parametrize("m", [1000, 129, 4096])
@pytest.mark.parametrize("relu", [False, True])
def test_linear_pair_and_epilogue(m, relu):
    g = torch.Generator(device=DEV).manual_seed(m)
    ka, kb, n = 256, 128, 256
    a = torch.randn(m, ka, device=DEV, generator=g).requires_grad_()
    b = torch.randn(m, kb, device=DEV, generator=g).requires_grad_()
    wa = (torch.randn(n, ka, device=DEV, generator=g) / 16).requires_grad_()
    wb = (torch.randn(n, kb, device=DEV, generator=g) / 11).requires_grad_()
    bias = torch.randn(n, device=DEV, generator=g).requires_grad_()
    gout = torch.randn(m, n, device=DEV, generator=g)
    n0 = ops.LAUNCHES.count
    y = dense.linear_pair(a, wa, b, wb, bias, relu=relu)
    y.backward(gout)
    assert ops.LAUNCHES.count > n0
    ref_in = [t.detach().double().requires_grad_() for t in (a, wa, b, wb, bias)]
    yr = ref_in[0] @ ref_in[1].t() + ref_in[2] @ ref_in[3].t() + ref_in[4]
    if relu:
        # the mask of OUR output: an element with |y| ~ 1e-7 may round to the other side of 0 in fp32, which flips a
        # whole gradient term -- legitimate at a discontinuity, so the reference uses the same active set
        yr = yr * (y.detach() > 0)
    yr.backward(gout.double())
    _close(y, yr, what="y")
    for got, want, name in zip((a, wa, b, wb, bias), ref_in, ("ga", "gwa", "gb", "gwb", "gbias")):
        _close(got.grad, want.grad, tol=2e-5, what=name)
    # single-stream linear with the bias / ReLU epilogue
    a2 = a.detach().requires_grad_()
    y2 = dense.linear(a2, wa.detach(), bias.detach(), relu=relu)
    y2r = a.detach().double() @ wa.detach().double().t() + bias.detach().double()
    _close(y2, y2r.relu() if relu else y2r, what="linear+epilogue")


def test_matmul_pair_row_major_weights():
    g = torch.Generator(device=DEV).manual_seed(3)
    m, ka, kb, n = 2000, 512, 128, 128
    a = torch.randn(m, ka, device=DEV, generator=g).requires_grad_()
    b = torch.randn(m, kb, device=DEV, generator=g).requires_grad_()
    wa = (torch.randn(ka, n, device=DEV, generator=g) / 20).requires_grad_()
    wb = (torch.randn(kb, n, device=DEV, generator=g) / 11).requires_grad_()
    bias = torch.randn(n, device=DEV, generator=g).requires_grad_()
    gout = torch.randn(m, n, device=DEV, generator=g)
    y = dense.matmul_pair(a, wa, b, wb, bias)
    y.backward(gout)
    ref_in = [t.detach().double().requires_grad_() for t in (a, wa, b, wb, bias)]
    yr = ref_in[0] @ ref_in[1] + ref_in[2] @ ref_in[3] + ref_in[4]
    yr.backward(gout.double())
    _close(y, yr, what="y")
    for got, want, name in zip((a, wa, b, wb, bias), ref_in, ("ga", "gwa", "gb", "gwb", "gbias")):
        _close(got.grad, want.grad, tol=2e-5, what=name)


@pytest.mark.parametrize("sizes", [[300, 0, 128, 1, 1000, 77], [128, 128], [5, 3, 0, 0, 9], [4000]])
def test_segment_matmul_grouped_kernel(sizes):
    g = torch.Generator(device=DEV).manual_seed(len(sizes))
    R, K, N = len(sizes), 128, 256
    ptr = torch.tensor([0] + sizes).cumsum(0).to(DEV)
    m = int(ptr[-1])
    x = torch.randn(m, K, device=DEV, generator=g).requires_grad_()
    w = (torch.randn(R, K, N, device=DEV, generator=g) / 11).requires_grad_()
    gout = torch.randn(m, N, device=DEV, generator=g)
    prof = ops.PROFILE
    prof.reset(enabled=True)
    y = dense.segment_matmul(x, ptr, w)
    calls = {k: v["calls"] for k, v in prof.summary().items()}
    prof.reset(enabled=False)
    assert calls.get("segment_matmul_tf32x3", 0) == 1, calls            # ONE launch for all segments
    y.backward(gout)
    xr, wr = x.detach().double().requires_grad_(), w.detach().double().requires_grad_()
    yr = torch.empty(m, N, dtype=torch.float64, device=DEV)
    outs = []
    for r in range(R):
        s, e = int(ptr[r]), int(ptr[r + 1])
        outs.append(xr[s:e] @ wr[r])
    yr = torch.cat(outs)
    yr.backward(gout.double())
    _close(y, yr, what="segment_matmul")
    _close(x.grad, xr.grad, tol=2e-5, what="grad_inputs")
    _close(w.grad, wr.grad, tol=2e-5, what="grad_other")


@pytest.mark.parametrize("is_sorted", [True, False])
def test_hetero_linear_vs_naive_loop(is_sorted):
    g = torch.Generator(device=DEV).manual_seed(9)
    T, K, N, m = 5, 64, 128, 3000
    tv = torch.randint(0, T, (m, ), device=DEV, generator=g)
    if is_sorted:
        tv = tv.sort()[0]
    lin = HeteroLinear(K, N, T, is_sorted=is_sorted).to(DEV)
    x = torch.randn(m, K, device=DEV, generator=g)
    out = lin(x, tv)
    want = torch.empty(m, N, dtype=torch.float64, device=DEV)
    for k in range(T):
        msk = tv == k
        want[msk] = x[msk].double() @ lin.weight[k].double() + lin.bias[k].double()
    _close(out, want, what="HeteroLinear")

"""Parity of the layers behind BASELINE configs 2, 3 and 5 at >= 1 M edges on a power-law graph (round-1 covered them
only on toy goldens and 300-3000-node graphs): SAGEConv(mean), GATConv (fp32 forward vs the C oracle; bf16 forward AND
backward vs the unfused fp64 formula) and RGCNConv against the C oracle / fp64, on the same generator bench.py uses."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import synth_graph  # noqa: E402
from pytorch_geometric_b200.graph import CSRGraph, cached_graph  # noqa: E402
from pytorch_geometric_b200.nn import conv as C  # noqa: E402
from test_gpu_attention import ref_attention  # noqa: E402

DEV = "cuda"
N, E = 120_000, 1_200_000


def _rel_to_terms(got, want, scale, tol):
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bad = err > tol * np.maximum(scale, 1e-30)
    assert not bad.any(), f"max err/scale {(err / np.maximum(scale, 1e-30)).max():.3e} (tol {tol})"


def test_sage_mean_1m_edges_vs_oracle():
    ei = synth_graph(N, E, 11, DEV)
    g = torch.Generator(device=DEV).manual_seed(1)
    F = 64
    x = torch.randn(N, F, device=DEV, generator=g)
    wl = torch.randn(F, F, device=DEV, generator=g) / 8
    wr = torch.randn(F, F, device=DEV, generator=g) / 8
    bl = torch.randn(F, device=DEV, generator=g) * 0.1
    graph = CSRGraph(ei[0], ei[1], N, N)
    assert graph.plan.n_long > 0                                   # hub rows exercise the chunked path
    out = C.sage_conv(x, x, graph, "mean", wl, bl, wr)
    src, dst = ei[0].cpu().numpy(), ei[1].cpu().numpy()
    want = O.sage_conv(x.cpu().numpy(), src, dst, wl.cpu().numpy(), bl.cpu().numpy(), wr.cpu().numpy(), "mean")
    scale = O.sage_conv(np.abs(x.cpu().numpy()), src, dst, np.abs(wl.cpu().numpy()), np.abs(bl.cpu().numpy()),
                        np.abs(wr.cpu().numpy()), "mean")
    _rel_to_terms(out.cpu().numpy(), want, scale, 1e-5)


def test_rgcn_mean_1m_edges_vs_oracle():
    ei = synth_graph(N, E, 12, DEV)
    g = torch.Generator(device=DEV).manual_seed(2)
    F, R = 32, 4
    et = torch.randint(0, R, (E, ), device=DEV, generator=g)
    x = torch.randn(N, F, device=DEV, generator=g)
    W = torch.randn(R, F, F, device=DEV, generator=g) / 6
    root = torch.randn(F, F, device=DEV, generator=g) / 6
    b = torch.randn(F, device=DEV, generator=g) * 0.1
    graph = cached_graph(ei, N, N * R, edge_type=et, num_relations=R)
    out = C.rgcn_conv(x, graph, W, root, b, "mean")
    a = lambda t: t.cpu().numpy()                                    # noqa: E731
    want = O.rgcn_conv(a(x), a(ei[0]), a(ei[1]), a(et), a(W), a(root), a(b), "mean")
    scale = O.rgcn_conv(np.abs(a(x)), a(ei[0]), a(ei[1]), a(et), np.abs(a(W)), np.abs(a(root)), np.abs(a(b)), "mean")
    _rel_to_terms(out.cpu().numpy(), want, scale, 1e-5)


def test_gat_1m_edges_fp32_vs_oracle_and_bf16_forward_backward_vs_fp64():
    ei = synth_graph(N, E, 13, DEV)
    g = torch.Generator(device=DEV).manual_seed(3)
    H, Cc = 8, 16
    xh = torch.randn(N, H * Cc, device=DEV, generator=g)
    att_s = torch.randn(1, H, Cc, device=DEV, generator=g) * 0.3
    att_d = torch.randn(1, H, Cc, device=DEV, generator=g) * 0.3
    graph = cached_graph(ei, N, N, loops="gat", loop_nodes=N)
    out = C.gat_conv(xh, None, graph, att_s, att_d, H, Cc)
    a = lambda t: t.cpu().numpy()                                    # noqa: E731
    want, _, _, _ = O.gat_attention(a(xh).reshape(N, H, Cc), a(att_s).reshape(H, Cc), a(att_d).reshape(H, Cc), a(ei[0]), a(ei[1]),
                                    0.2, add_self_loops=True)
    assert np.abs(out.cpu().numpy() - want).max() <= 2e-5 * np.abs(want).max()
    # ---- bf16 storage, forward + backward (config 3's dtype), against the unfused formula in fp64
    bf = torch.bfloat16
    xb = xh.to(bf).requires_grad_()
    a_s = (xb.detach().float().view(N, H, Cc) * att_s).sum(-1).requires_grad_()
    a_d = (xb.detach().float().view(N, H, Cc) * att_d).sum(-1).requires_grad_()
    from pytorch_geometric_b200 import functional as Fn
    ob = Fn.attention("gat", graph, H, Cc, v=xb, s_src=a_s, s_dst=a_d)
    gout = torch.randn(N, H * Cc, device=DEV, generator=g).to(bf)
    ob.backward(gout)
    src2, dst2 = graph.col.long(), graph.dst_csr.long()             # the edges incl. the inserted loops, CSR order
    r_in = {"v": xb.detach().double().requires_grad_(), "s_src": a_s.detach().double().requires_grad_(),
            "s_dst": a_d.detach().double().requires_grad_()}
    ro, _ = ref_attention("gat", src2, dst2, N, H, Cc, **r_in)
    ro.backward(gout.double())
    def close(x_, y_, t, what):
        e = (x_.double() - y_).abs().max().item()
        assert e <= t * y_.abs().max().item(), f"{what}: {e:.3e} vs {y_.abs().max().item():.3e}"
    close(ob, ro, 1.2e-2, "out bf16")
    close(xb.grad, r_in["v"].grad, 2e-2, "grad_v bf16")
    close(a_s.grad, r_in["s_src"].grad, 2e-2, "grad_s_src")
    close(a_d.grad, r_in["s_dst"].grad, 2e-2, "grad_s_dst")


@pytest.mark.parametrize("aggr,relu", [("mean", True), ("sum", False)])
def test_fused_sage_layer_forward_backward_vs_fp64(aggr, relu):
    """The one-node SAGE layer (`_SageFused`: pair GEMMs + accumulate epilogue) against the unfused formula in fp64."""
    n, e, F = 50_000, 600_000, 128
    ei = synth_graph(n, e, 17, DEV)
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(n, F, device=DEV, generator=g).requires_grad_()
    wl = (torch.randn(F, F, device=DEV, generator=g) / 11).requires_grad_()
    wr = (torch.randn(F, F, device=DEV, generator=g) / 11).requires_grad_()
    bl = (torch.randn(F, device=DEV, generator=g) * 0.1).requires_grad_()
    gout = torch.randn(n, F, device=DEV, generator=g)
    graph = CSRGraph(ei[0], ei[1], n, n)
    assert C._sage_fusable(x, x, graph, aggr, wl, wr)
    y = C.sage_conv(x, x, graph, aggr, wl, bl, wr, relu=relu)
    y.backward(gout)
    xr, wlr, wrr, blr = (t.detach().double().requires_grad_() for t in (x, wl, wr, bl))
    agg = torch.zeros(n, F, dtype=torch.float64, device=DEV).index_add_(0, ei[1], xr[ei[0]])
    if aggr == "mean":
        agg = agg / torch.bincount(ei[1], minlength=n).clamp(min=1).double().view(-1, 1)
    yr = agg @ wlr.t() + xr @ wrr.t() + blr
    if relu:
        yr = yr * (y.detach() > 0)          # same active set as the fp32 result (see test_gpu_dense_ext.py)
    yr.backward(gout.double())
    for got, want, name, tol in ((y, yr, "y", 1e-5), (x.grad, xr.grad, "gx", 2e-5), (wl.grad, wlr.grad, "gWl", 2e-5),
                                 (wr.grad, wrr.grad, "gWr", 2e-5), (bl.grad, blr.grad, "gb", 2e-5)):
        err = (got.double() - want).abs().max().item()
        assert err <= tol * max(want.abs().max().item(), 1e-6) * 4, f"{name}: {err:.3e} vs {want.abs().max().item():.3e}"


def test_fused_sage_stack_with_relu_mask_in_the_sweep_epilogue():
    """Two stacked fused SAGE layers with the ReLU backward of layer 1 applied by layer 2's accumulate epilogue
    (input_is_relu / grad_masked_by_consumer): gradients equal the plain stacking of the same layers."""
    n, e, F = 30_000, 400_000, 128
    ei = synth_graph(n, e, 19, DEV)
    g = torch.Generator(device=DEV).manual_seed(6)
    graph = CSRGraph(ei[0], ei[1], n, n)
    P = [(torch.randn(F, F, device=DEV, generator=g) / 11) for _ in range(4)]
    bs = [torch.randn(F, device=DEV, generator=g) * 0.1 for _ in range(2)]
    x0 = torch.randn(n, F, device=DEV, generator=g)
    gout = torch.randn(n, F, device=DEV, generator=g)
    res = []
    for hints in (False, True):
        x = x0.clone().requires_grad_()
        ps = [p.clone().requires_grad_() for p in P]
        h = C.sage_conv(x, x, graph, "mean", ps[0], bs[0], ps[1], relu=True, grad_masked_by_consumer=hints)
        y = C.sage_conv(h, h, graph, "mean", ps[2], bs[1], ps[3], input_is_relu=hints)
        y.backward(gout)
        res.append([y.detach(), x.grad] + [p.grad for p in ps])
    for a, b in zip(*res):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5 * b.abs().max().item())

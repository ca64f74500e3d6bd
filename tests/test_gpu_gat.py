"""GPU parity of the fused GAT path (csrc/gat.cu) against the golden GATConv run of the reference
and against the oracle's restatement of gat_conv.py:330-409."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import oracle as O

pytestmark = pytest.mark.gpu

import pytorch_geometric_b200 as pgb  # noqa: E402
from pytorch_geometric_b200 import functional as Fn  # noqa: E402
from pytorch_geometric_b200.graph import CSRGraph  # noqa: E402
from pytorch_geometric_b200.nn import GATConv  # noqa: E402

DEV = "cuda"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().float().cpu().numpy()


def test_gat_conv_golden_forward_backward():
    g = load_golden("gat")
    conv = GATConv(6, 3, heads=4).to(DEV)
    with torch.no_grad():
        conv.lin.weight.copy_(cu(g["lin"]))
        conv.att_src.copy_(cu(g["att_src"]))
        conv.att_dst.copy_(cu(g["att_dst"]))
        conv.bias.copy_(cu(g["bias"]))
    x = cu(g["x"]).requires_grad_()
    out, (ei, alpha) = conv(x, cu(g["ei"]), return_attention_weights=True)
    assert_close(npy(out), g["out"], rtol=1e-5, atol=1e-6)
    # alpha comes back in CSR order: compare as a multiset keyed by (src, dst, value)
    def canon(ei_, a_):
        rows = np.concatenate([ei_.T.astype(np.float64), np.round(a_.astype(np.float64), 5)], axis=1)
        return rows[np.lexsort(rows.T[::-1])]
    assert_close(canon(ei.cpu().numpy(), npy(alpha)), canon(g["ei2"], g["alpha"]), rtol=1e-4, atol=2e-5)
    out.backward(cu(g["gout"]))
    assert_close(npy(x.grad), g["gx"], rtol=1e-4, atol=1e-5)
    assert_close(npy(conv.lin.weight.grad), g["glin"], rtol=1e-4, atol=1e-5)
    assert_close(npy(conv.att_src.grad), g["gatt_src"], rtol=1e-4, atol=1e-5)
    assert_close(npy(conv.att_dst.grad), g["gatt_dst"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("chunk", [512, 16])   # 16: hub rows go through the chunked online-softmax + merge path
@pytest.mark.parametrize("H,C,dtype", [(8, 16, torch.float32), (8, 16, torch.bfloat16), (4, 3, torch.float32),
                                       (1, 64, torch.float32), (2, 128, torch.float32), (8, 4, torch.bfloat16)])
def test_gat_attention_vs_oracle(H, C, dtype, chunk):
    rng = np.random.default_rng(H * 100 + C)
    N, E = 500, 8000
    src = rng.integers(0, N, size=E)
    dst = ((rng.random(E) ** 2) * (N - 1)).astype(np.int64)
    xh = rng.standard_normal((N, H, C)).astype(np.float32)
    xh = torch.from_numpy(xh).to(dtype).float().numpy()            # inputs rounded to the test dtype
    att_s = rng.standard_normal((H, C)).astype(np.float32) * 0.5
    att_d = rng.standard_normal((H, C)).astype(np.float32) * 0.5
    ref_out, ref_alpha, r2, c2 = O.gat_attention(xh, att_s, att_d, src, dst, 0.2, add_self_loops=True)
    g = CSRGraph(cu(r2), cu(c2), N, N, chunk=chunk)
    if chunk == 16:
        assert g.plan.n_long > 0
    a_src = (xh * att_s).sum(-1).astype(np.float32)
    a_dst = (xh * att_d).sum(-1).astype(np.float32)
    out, alpha = Fn.gat_attention(g, cu(xh.reshape(N, H * C)).to(dtype), cu(a_src), cu(a_dst), H, C, 0.2, True)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    scale = np.abs(ref_out).max()
    assert np.abs(npy(out) - ref_out).max() <= max(tol * scale, 2e-5 if dtype == torch.float32 else 0)
    # alpha in CSR order == oracle alpha permuted by the stable sort on dst
    perm = g.perm.cpu().numpy().astype(np.int64)
    assert_close(npy(alpha), ref_alpha[perm], rtol=1e-4, atol=1e-6)


def test_gat_backward_matches_autograd_of_unfused_formula():
    """Gradient check against torch autograd applied to the reference's unfused formula
    (gather, leaky_relu, scatter-softmax, weighted scatter-add) written with plain torch ops."""
    torch.manual_seed(0)
    N, E, H, C = 300, 4000, 8, 16
    src = torch.randint(0, N, (E, ), device=DEV)
    dst = (torch.rand(E, device=DEV) ** 2 * (N - 1)).long()
    g = CSRGraph(src, dst, N, N, chunk=8)          # most rows are "hubs": chunked forward + chunked grad_a_dst
    assert g.plan.n_long > 0
    xh = torch.randn(N, H * C, device=DEV, requires_grad=True)
    a_s = torch.randn(N, H, device=DEV, requires_grad=True)
    a_d = torch.randn(N, H, device=DEV, requires_grad=True)
    gout = torch.randn(N, H * C, device=DEV)
    out = Fn.gat_attention(g, xh, a_s, a_d, H, C, 0.2)
    out.backward(gout)
    got = [t.grad.clone() for t in (xh, a_s, a_d)]
    for t in (xh, a_s, a_d):
        t.grad = None
    logit = torch.nn.functional.leaky_relu(a_s[src] + a_d[dst], 0.2)
    mx = torch.full((N, H), -float("inf"), device=DEV).scatter_reduce(0, dst.view(-1, 1).expand(-1, H), logit, "amax")
    ex = (logit - mx[dst]).exp()
    den = torch.zeros(N, H, device=DEV).index_add_(0, dst, ex) + 1e-16
    alpha = ex / den[dst]
    ref = torch.zeros(N, H, C, device=DEV).index_add_(0, dst, alpha.unsqueeze(-1) * xh.view(N, H, C)[src]).view(N, H * C)
    assert_close(npy(out), npy(ref), rtol=1e-4, atol=1e-5)
    ref.backward(gout)
    for a, b, name in zip(got, (xh.grad, a_s.grad, a_d.grad), ("xh", "a_src", "a_dst")):
        assert_close(npy(a), npy(b), rtol=1e-3, atol=1e-4, msg=name)

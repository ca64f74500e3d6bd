"""GPU parity of the fused attention family (csrc/attention.cu: GAT / GATv2 / dot-product scores, edge softmax,
weighted aggregation, and the two-sweep backward) against
  * the reference's unfused formula written with plain torch ops in fp64 (gather, score, scatter-softmax of
    utils/_softmax.py:82-88, weighted scatter-add) and differentiated by torch autograd,
  * the oracle's restatement of GATv2Conv and the golden run of the reference's GATv2Conv (tests/golden/gatv2.npz).
Shapes cover one vector per head, several lanes per head, two vectors per lane, hub rows cut into chunks (chunk=16),
int64 indices, bf16 storage, keys|values fused in one matrix, and a per-edge additive score (edge_dim)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from oracle import oracle as O

pytestmark = pytest.mark.gpu

from pytorch_geometric_b200 import functional as Fn  # noqa: E402
from pytorch_geometric_b200.graph import CSRGraph  # noqa: E402

DEV = "cuda"


def ref_attention(mode, src, dst, n_dst, H, C, v, k=None, q=None, s_src=None, s_dst=None, att=None, s_edge=None,
                  slope=0.2, scale=1.0, keep=None, e_feat=None):
    vj = v[src].view(-1, H, C)
    ef = 0.0 if e_feat is None else e_feat.view(-1, H, C)          # lin_edge(edge_attr) of the edge_dim layers
    if mode == "gat":
        pre = s_src[src] + s_dst[dst]
        if s_edge is not None:
            pre = pre + s_edge
        s = F.leaky_relu(pre, slope)
    elif mode == "gatv2":
        s = (F.leaky_relu(vj + q[dst].view(-1, H, C) + ef, slope) * att.view(1, H, C)).sum(-1)      # gatv2_conv.py:358-362
    else:
        s = (q[dst].view(-1, H, C) * (k[src].view(-1, H, C) + ef)).sum(-1) * scale                    # transformer_conv.py:258-263
        vj = vj + ef                                                                                # transformer_conv.py:270-272
    idx = dst.view(-1, 1).expand(-1, H)
    smax = torch.full((n_dst, H), -math.inf, dtype=s.dtype, device=s.device).scatter_reduce(0, idx, s.detach(), "amax")
    ex = (s - smax[dst]).exp()
    den = torch.zeros(n_dst, H, dtype=s.dtype, device=s.device).index_add(0, dst, ex) + 1e-16
    alpha = ex / den[dst]
    if keep is not None:                                   # F.dropout(alpha): kept coefficients scaled by 1 / (1 - p)
        alpha = alpha * keep
    out = torch.zeros(n_dst, H, C, dtype=s.dtype, device=s.device).index_add(0, dst, alpha.unsqueeze(-1) * vj)
    return out.view(n_dst, H * C), alpha


def _problem(mode, H, C, dtype, seed, n_src=400, n_dst=300, E=6000, with_edge=False):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n_src, (E, ), generator=g)
    dst = (torch.rand(E, generator=g) ** 2.5 * (n_dst - 1)).long()           # skewed: a few hub destinations
    dst[:20] = n_dst - 1 - torch.arange(20) % 7                               # some rows with few edges; others empty
    rnd = lambda *s: torch.randn(*s, generator=g).to(dtype).double()          # noqa: E731  (inputs rounded to the test dtype)
    p = {"src": src, "dst": dst, "v": rnd(n_src, H * C)}
    if mode == "gat":
        p.update(s_src=torch.randn(n_src, H, generator=g).double(), s_dst=torch.randn(n_dst, H, generator=g).double())
        if with_edge:
            p["s_edge"] = torch.randn(E, H, generator=g).double() * 0.5
    elif mode == "gatv2":
        p.update(q=rnd(n_dst, H * C), att=torch.randn(H * C, generator=g).double() * 0.5)
    else:
        p.update(q=rnd(n_dst, H * C), k=rnd(n_src, H * C))
    p["gout"] = rnd(n_dst, H * C)
    return p


CASES = [(8, 16, torch.float32), (8, 16, torch.bfloat16), (1, 64, torch.float32), (4, 8, torch.float32),
         (2, 128, torch.float32), (4, 32, torch.bfloat16), (16, 8, torch.bfloat16), (3, 4, torch.float32)]


@pytest.mark.parametrize("chunk", [512, 16])
@pytest.mark.parametrize("mode", ["gat", "gatv2", "dot"])
@pytest.mark.parametrize("H,C,dtype", CASES)
def test_attention_forward_backward_vs_unfused_fp64(mode, H, C, dtype, chunk):
    p = _problem(mode, H, C, dtype, seed=H * 131 + C + len(mode), with_edge=(mode == "gat" and C == 16))
    n_src, n_dst = p["v"].size(0), p["gout"].size(0)
    names = [n for n in ("v", "k", "q", "s_src", "s_dst", "att", "s_edge") if n in p]
    scale = 1.0 / math.sqrt(C)
    # ---- fp64 reference with autograd
    ref_in = {n: p[n].clone().to(DEV).requires_grad_() for n in names}
    ref_out, ref_alpha = ref_attention(mode, p["src"].to(DEV), p["dst"].to(DEV), n_dst, H, C, slope=0.2, scale=scale, **ref_in)
    ref_out.backward(p["gout"].to(DEV))
    # ---- engine
    idx_dtype = torch.int64 if (H, C) == (4, 8) else None
    graph = CSRGraph(p["src"].to(DEV), p["dst"].to(DEV), n_src, n_dst, chunk=chunk, idx_dtype=idx_dtype)
    if chunk == 16:
        assert graph.plan.n_long > 0
    feat = ("v", "k", "q")
    ours = {n: (p[n].to(dtype) if n in feat else p[n].float()).to(DEV).requires_grad_() for n in names}
    out, alpha = Fn.attention(mode, graph, H, C, negative_slope=0.2, scale=scale, return_alpha=True, **ours)
    out.backward(p["gout"].to(dtype).to(DEV))
    fp32 = dtype == torch.float32
    tol = 2e-5 if fp32 else 1.5e-2

    def close(a, b, what, t=tol):
        a, b = a.detach().double(), b.detach().double()
        err = (a - b).abs().max().item()
        assert err <= t * max(b.abs().max().item(), 1e-3), f"{what}: max err {err:.3e} (scale {b.abs().max().item():.3e})"

    close(out, ref_out, "out")
    perm = graph.perm.long()
    close(alpha, ref_alpha[perm], "alpha", 1e-4 if fp32 else 1.5e-2)
    for n in names:
        close(ours[n].grad, ref_in[n].grad, "grad_" + n, (2e-4 if fp32 else 3e-2))


def test_dot_attention_with_fused_key_value_matrix():
    """keys | values as the two halves of one [N, 2HC] product: strided operands, one gradient matrix."""
    H, C = 8, 16
    p = _problem("dot", H, C, torch.float32, seed=5)
    graph = CSRGraph(p["src"].to(DEV), p["dst"].to(DEV), p["v"].size(0), p["gout"].size(0))
    kv = torch.cat([p["k"], p["v"]], dim=1).float().to(DEV).requires_grad_()
    q = p["q"].float().to(DEV).requires_grad_()
    out = Fn.attention("dot", graph, H, C, q=q, kv=kv, scale=0.25)
    out.backward(p["gout"].float().to(DEV))
    k2 = p["k"].float().to(DEV).requires_grad_()
    v2 = p["v"].float().to(DEV).requires_grad_()
    q2 = p["q"].float().to(DEV).requires_grad_()
    out2 = Fn.attention("dot", graph, H, C, q=q2, k=k2, v=v2, scale=0.25)
    out2.backward(p["gout"].float().to(DEV))
    assert torch.equal(out, out2) and torch.equal(q.grad, q2.grad)
    assert torch.equal(kv.grad[:, :H * C], k2.grad) and torch.equal(kv.grad[:, H * C:], v2.grad)


def test_gatv2_attention_vs_oracle_and_golden_layer():
    """(1) the fused attention on the golden's projected inputs against the oracle's restatement of
    gatv2_conv.py:310-378 (H = 4, C = 3: off the vector path -> zero-padded heads); (2) the whole GATv2Conv layer,
    forward and backward, against what the reference's GATv2Conv produced (tests/golden/gatv2.npz)."""
    from pytorch_geometric_b200.nn import GATv2Conv
    g = load_golden("gatv2")
    H, C = int(g["H"]), int(g["C"])
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)          # noqa: E731
    for tag, share in (("sep", False), ("shared", True)):
        x = g["x"]
        N = x.shape[0]
        x_l = x @ g[f"{tag}_lin_l_w"].T + g[f"{tag}_lin_l_b"]
        x_r = x @ g[f"{tag}_lin_r_w"].T + g[f"{tag}_lin_r_b"]
        out_ref, alpha_ref, r2, c2 = O.gatv2_attention(x_l.reshape(N, H, C), x_r.reshape(N, H, C), g[f"{tag}_att"],
                                                       g["ei"][0], g["ei"][1], 0.2, add_self_loops=True)
        graph = CSRGraph(cu(r2), cu(c2), N, N)
        out, alpha = Fn.attention("gatv2", graph, H, C, v=cu(x_l.astype(np.float32)), q=cu(x_r.astype(np.float32)),
                                  att=cu(g[f"{tag}_att"].reshape(-1)), negative_slope=0.2, return_alpha=True)
        np.testing.assert_allclose(out.cpu().numpy(), out_ref.reshape(N, H * C), rtol=2e-5, atol=2e-6)
        perm = graph.perm.cpu().numpy().astype(np.int64)
        np.testing.assert_allclose(alpha.cpu().numpy(), alpha_ref[perm], rtol=2e-5, atol=1e-6)
        # ---- the layer against the reference's own run
        conv = GATv2Conv(x.shape[1], C, heads=H, share_weights=share).to(DEV)
        with torch.no_grad():
            conv.lin_l.weight.copy_(cu(g[f"{tag}_lin_l_w"]))
            conv.lin_l.bias.copy_(cu(g[f"{tag}_lin_l_b"]))
            if not share:
                conv.lin_r.weight.copy_(cu(g[f"{tag}_lin_r_w"]))
                conv.lin_r.bias.copy_(cu(g[f"{tag}_lin_r_b"]))
            conv.att.copy_(cu(g[f"{tag}_att"]).view(1, H, C))
            conv.bias.copy_(cu(g[f"{tag}_bias"]))
        xt = cu(x).requires_grad_()
        y = conv(xt, cu(g["ei"]))
        np.testing.assert_allclose(y.detach().cpu().numpy(), g[f"{tag}_out"], rtol=2e-5, atol=2e-6)
        y.backward(cu(g[f"{tag}_gout"]))
        np.testing.assert_allclose(xt.grad.cpu().numpy(), g[f"{tag}_gx"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(conv.att.grad.view(H, C).cpu().numpy(), g[f"{tag}_g_att"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(conv.lin_l.weight.grad.cpu().numpy(), g[f"{tag}_g_lin_l_w"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("H,C,dtype,pair", [(8, 16, torch.bfloat16, True), (8, 16, torch.float32, True),
                                           (3, 4, torch.float32, True), (1, 128, torch.float32, False),
                                           (2, 8, torch.bfloat16, True), (5, 32, torch.float32, False),
                                           (4, 6, torch.float32, True)])
def test_head_dot_terms_and_their_backward(H, C, dtype, pair):
    """alpha_src / alpha_dst = (x * att).sum(-1) (gat_conv.py:330-331) from one read of x (csrc/head_dot.cu), and the
    one-pass backward, against the formula in fp64; (4, 6) is a shape the kernel does not take (torch fallback)."""
    from pytorch_geometric_b200 import ops
    from pytorch_geometric_b200.nn.conv import _head_dot
    g = torch.Generator(device=DEV).manual_seed(H * 100 + C)
    N = 3001
    x = torch.randn(N, H * C, device=DEV, generator=g).to(dtype).requires_grad_()
    att_a = torch.randn(1, H, C, device=DEV, generator=g).requires_grad_()
    att_b = torch.randn(1, H, C, device=DEV, generator=g).requires_grad_() if pair else None
    assert ops.head_dot_supported(x, H, C) == ((C * x.element_size()) % 16 == 0)
    r = _head_dot(x, att_a, H, C, att_b)
    s_a, s_b = r if pair else (r, None)
    x64 = x.detach().double().requires_grad_()
    a64 = att_a.detach().double().requires_grad_()
    b64 = att_b.detach().double().requires_grad_() if pair else None
    ra = (x64.view(N, H, C) * a64).sum(-1)
    rb = (x64.view(N, H, C) * b64).sum(-1) if pair else None
    tol = 1e-5 * float((x64.view(N, H, C).abs() * a64.abs()).sum(-1).max())
    assert float((s_a.double() - ra).abs().max()) <= tol
    ga = torch.randn(N, H, device=DEV, generator=g)
    gb = torch.randn(N, H, device=DEV, generator=g) if pair else None
    if pair:
        assert float((s_b.double() - rb).abs().max()) <= 1e-5 * float((x64.view(N, H, C).abs() * b64.abs()).sum(-1).max())
        torch.autograd.backward([s_a, s_b], [ga, gb])
        torch.autograd.backward([ra, rb], [ga.double(), gb.double()])
    else:
        s_a.backward(ga)
        ra.backward(ga.double())
    gx_tol = (2.0 ** -8 if dtype == torch.bfloat16 else 1e-5) * float(x64.grad.abs().max() + 1)
    assert float((x.grad.double() - x64.grad).abs().max()) <= gx_tol
    for got, ref in ((att_a.grad, a64.grad), ) + (((att_b.grad, b64.grad), ) if pair else ()):
        assert got.shape == ref.shape
        assert float((got.double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max() + N ** 0.5)


@pytest.mark.parametrize("chunk", [512, 16])
@pytest.mark.parametrize("mode,H,C,dtype", [("gat", 8, 16, torch.float32), ("gat", 8, 16, torch.bfloat16), ("gatv2", 4, 8, torch.float32),
                                            ("dot", 2, 128, torch.float32), ("gat", 3, 4, torch.float32)])
def test_attention_dropout_inside_the_sweep(mode, H, C, dtype, chunk):
    """F.dropout(alpha, p, training) (gat_conv.py:404, gatv2_conv.py:376, transformer_conv.py:268) fused into the sweeps.
    torch's random stream cannot be reproduced inside a kernel, so: (1) the dropped coefficients the engine returns are
    alpha / (1 - p) or 0, with a keep rate of 1 - p, the same for the same seed and different for another seed;
    (2) out and every gradient equal the unfused fp64 formula evaluated with THAT mask."""
    p = _problem(mode, H, C, dtype, seed=H * 17 + C + len(mode))
    n_src, n_dst = p["v"].size(0), p["gout"].size(0)
    names = [n for n in ("v", "k", "q", "s_src", "s_dst", "att") if n in p]
    scale = 1.0 / math.sqrt(C)
    graph = CSRGraph(p["src"].to(DEV), p["dst"].to(DEV), n_src, n_dst, chunk=chunk)
    perm = graph.perm.long()
    feat = ("v", "k", "q")
    ours = {n: (p[n].to(dtype) if n in feat else p[n].float()).to(DEV).requires_grad_() for n in names}
    pd = 0.4
    kw = dict(negative_slope=0.2, scale=scale, return_alpha=True)
    with torch.no_grad():
        _, alpha0 = Fn.attention(mode, graph, H, C, **kw, **ours)
        _, alpha_b = Fn.attention(mode, graph, H, C, dropout_p=pd, dropout_seed=99, **kw, **ours)
        _, alpha_c = Fn.attention(mode, graph, H, C, dropout_p=pd, dropout_seed=7, **kw, **ours)
    out, alpha_d = Fn.attention(mode, graph, H, C, dropout_p=pd, dropout_seed=7, **kw, **ours)
    assert torch.equal(alpha_c, alpha_d) and not torch.equal(alpha_b, alpha_d)
    kept = alpha_d != 0
    live = alpha0 > 1e-30
    rate = (kept & live).sum().item() / live.sum().item()
    assert abs(rate - (1 - pd)) < 0.02, rate
    assert torch.allclose(alpha_d[kept], alpha0[kept] / (1 - pd), rtol=1e-6, atol=0)
    per_head = (kept & live).float().sum(0) / live.float().sum(0)
    assert float((per_head - (1 - pd)).abs().max()) < 0.05                    # heads draw independent masks
    out.backward(p["gout"].to(dtype).to(DEV))
    # the unfused formula with this mask (CSR order -> the caller's edge order)
    keep = torch.empty_like(alpha_d, dtype=torch.float64)
    keep[perm] = kept.double() / (1 - pd)
    ref_in = {n: p[n].clone().to(DEV).requires_grad_() for n in names}
    ref_out, _ = ref_attention(mode, p["src"].to(DEV), p["dst"].to(DEV), n_dst, H, C, slope=0.2, scale=scale, keep=keep, **ref_in)
    ref_out.backward(p["gout"].to(DEV))
    fp32 = dtype == torch.float32

    def close(a, b, what, t):
        a, b = a.detach().double(), b.detach().double()
        err = (a - b).abs().max().item()
        assert err <= t * max(b.abs().max().item(), 1e-3), f"{what}: max err {err:.3e} (scale {b.abs().max().item():.3e})"

    close(out, ref_out, "out", 2e-5 if fp32 else 1.5e-2)
    for n in names:
        close(ours[n].grad, ref_in[n].grad, "grad_" + n, 2e-4 if fp32 else 3e-2)


def test_attention_layers_train_with_dropout():
    """The layers no longer refuse dropout > 0 in training mode; eval mode is deterministic and equals dropout = 0."""
    from pytorch_geometric_b200.nn import GATConv, GATv2Conv, TransformerConv
    torch.manual_seed(3)
    N, E = 500, 6000
    ei = torch.randint(0, N, (2, E), device=DEV)
    x = torch.randn(N, 32, device=DEV)
    for cls in (GATConv, GATv2Conv, TransformerConv):
        torch.manual_seed(11)
        a = cls(32, 16, heads=4, dropout=0.5).to(DEV)
        torch.manual_seed(11)
        b = cls(32, 16, heads=4, dropout=0.0).to(DEV)
        b.load_state_dict(a.state_dict())
        a.train()
        torch.manual_seed(5)
        y1 = a(x, ei)
        torch.manual_seed(5)
        y2 = a(x, ei)
        y3 = a(x, ei)
        assert torch.equal(y1, y2) and not torch.equal(y1, y3)                  # seeded from torch's generator
        y1.sum().backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in a.parameters())
        a.eval(), b.eval()
        assert torch.equal(a(x, ei), b(x, ei))
        # E[dropout output] = no-dropout output: the mean over draws approaches the deterministic layer
        a.train()
        with torch.no_grad():
            g = a.graph_for(ei, N) if hasattr(a, "graph_for") else ei          # (GATConv: build the CSR once)
            mean = torch.stack([a(x, g) for _ in range(100)]).mean(0)
            ref = b(x, ei)
        assert float((mean - ref).abs().mean() / ref.abs().mean()) < 0.1, cls.__name__


@pytest.mark.parametrize("chunk", [512, 16])
@pytest.mark.parametrize("mode,H,C,dtype", [("gatv2", 8, 16, torch.float32), ("gatv2", 8, 16, torch.bfloat16), ("gatv2", 2, 128, torch.float32),
                                            ("gatv2", 3, 4, torch.float32), ("dot", 8, 16, torch.float32), ("dot", 4, 32, torch.bfloat16),
                                            ("dot", 2, 128, torch.float32), ("dot", 1, 4, torch.float32)])
def test_attention_with_per_edge_feature_rows(mode, H, C, dtype, chunk):
    """`edge_dim`: lin_edge(edge_attr) [E, H*C] inside GATv2's leaky_relu (gatv2_conv.py:358-360) / added to the key and the
    value of TransformerConv (transformer_conv.py:258-272) -- forward, returned coefficients, and every gradient incl. the
    per-edge one, vs the unfused fp64 formula; with dropout on top for one case per mode."""
    p = _problem(mode, H, C, dtype, seed=H * 7 + C + 3 * len(mode))
    n_src, n_dst, E = p["v"].size(0), p["gout"].size(0), p["src"].numel()
    p["e_feat"] = (torch.randn(E, H * C, generator=torch.Generator().manual_seed(C)) * 0.7).to(dtype).double()
    names = [n for n in ("v", "k", "q", "att", "e_feat") if n in p]
    scale = 1.0 / math.sqrt(C)
    graph = CSRGraph(p["src"].to(DEV), p["dst"].to(DEV), n_src, n_dst, chunk=chunk)
    perm = graph.perm.long()
    feat = ("v", "k", "q", "e_feat")
    ours = {n: (p[n].to(dtype) if n in feat else p[n].float()).to(DEV).requires_grad_() for n in names}
    pd = 0.3 if (H, C) == (8, 16) and dtype == torch.float32 else 0.0
    out, alpha = Fn.attention(mode, graph, H, C, negative_slope=0.2, scale=scale, return_alpha=True, dropout_p=pd, dropout_seed=5,
                              **ours)
    out.backward(p["gout"].to(dtype).to(DEV))
    keep = None
    if pd > 0:
        keep = torch.empty_like(alpha, dtype=torch.float64)
        keep[perm] = (alpha != 0).double() / (1 - pd)
    ref_in = {n: p[n].clone().to(DEV).requires_grad_() for n in names}
    ref_out, ref_alpha = ref_attention(mode, p["src"].to(DEV), p["dst"].to(DEV), n_dst, H, C, slope=0.2, scale=scale, keep=keep, **ref_in)
    ref_out.backward(p["gout"].to(DEV))
    fp32 = dtype == torch.float32

    def close(a, b, what, t):
        a, b = a.detach().double(), b.detach().double()
        err = (a - b).abs().max().item()
        assert err <= t * max(b.abs().max().item(), 1e-3), f"{what}: max err {err:.3e} (scale {b.abs().max().item():.3e})"

    close(out, ref_out, "out", 2e-5 if fp32 else 1.5e-2)
    close(alpha, ref_alpha[perm], "alpha", 1e-4 if fp32 else 1.5e-2)
    for n in names:
        close(ours[n].grad, ref_in[n].grad, "grad_" + n, 2e-4 if fp32 else 3e-2)

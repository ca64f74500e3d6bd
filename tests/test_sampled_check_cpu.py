"""The sampled-row checker of bench.py (oracle/sampled.py) checked on CPU: fed with the full-graph oracle's own
GCNConv results it must report a tiny error, and it must catch a corrupted row / gradient."""
import numpy as np
import torch

from oracle import oracle as O
from oracle import sampled


def _problem(seed=0, N=3000, E=40000, Fi=32, Fo=48):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, N, size=E)
    dst = ((rng.random(E) ** 4) * (N - 1)).astype(np.int64)          # skewed: a few rows above 512 edges
    dst[: N // 3] = rng.integers(0, N // 2, size=N // 3)
    x = rng.standard_normal((N, Fi)).astype(np.float32)
    W = (rng.standard_normal((Fo, Fi)) / np.sqrt(Fi)).astype(np.float32)
    b = rng.standard_normal(Fo).astype(np.float32)
    gout = rng.standard_normal((N, Fo)).astype(np.float32)
    out = O.gcn_conv(x, src, dst, None, W, b)
    gx, gw, gb = O.gcn_conv_backward(gout, x, src, dst, None, W)
    t = torch.from_numpy
    return dict(ei=t(np.stack([src, dst])), x=t(x), weight=t(W), bias=t(b), gout=t(gout), out=t(out), gx=t(gx),
                gw=t(gw), gb=t(gb), N=N)


def _run(p, **over):
    q = dict(p)
    q.update(over)
    return sampled.gcn_check(q["ei"], 0, q["N"], q["x"], q["weight"], q["bias"], q["gout"], q["out"], q["gx"], q["gw"],
                             q["gb"], n_rows=512, max_edges=20000, seed=3)


def test_sampled_check_accepts_the_oracle_itself():
    p = _problem()
    r = _run(p)
    assert r["ok"], r
    assert r["dst_rows"] >= 256 and r["src_rows"] >= 128 and r["max_rel"] < 2e-6


def test_sampled_check_catches_wrong_rows_and_gradients():
    p = _problem(1)
    base = _run(p)
    assert base["ok"]
    bad = p["out"].clone()
    bad += 1e-3 * bad.abs().mean()                                    # every row slightly off
    assert not _run(p, out=bad)["ok"]
    bad = p["gx"].clone() * (1 + 1e-3)
    assert not _run(p, gx=bad)["ok"]
    bad = p["gw"].clone()
    bad[3, 5] += 0.05 * bad.abs().max()
    assert not _run(p, gw=bad)["ok"]
    assert not _run(p, gb=p["gb"] * 1.001)["ok"]

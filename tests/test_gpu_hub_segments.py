"""Hub destinations in the ptr/sorted-index entry points (utils.segment / scatter(sorted) / softmax):
groups far longer than the long-row chunk go through the chunk plan instead of being walked by a single
lane group.  Checked against the oracle at the default thresholds (one 40 k-edge hub) and with the
thresholds lowered so that most groups are chunked."""
import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import oracle as O

pytestmark = pytest.mark.gpu

import pytorch_geometric_b200 as pgb  # noqa: E402,F401
from pytorch_geometric_b200 import ops  # noqa: E402
from pytorch_geometric_b200 import utils as U  # noqa: E402

DEV = "cuda"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().float().cpu().numpy()


def hub_index(rng, N, E, hub_edges):
    idx = np.concatenate([np.full(hub_edges, 7), rng.integers(0, N - 3, size=E - hub_edges)])
    return np.sort(idx)


@pytest.fixture(params=["default", "lowered"])
def thresholds(request, monkeypatch):
    if request.param == "lowered":
        monkeypatch.setattr(ops, "SEGMENT_PLAN_MIN_ROWS", 0)
        monkeypatch.setattr(ops, "SEGMENT_CHUNK", 16)
    return request.param


@pytest.mark.parametrize("red", ["sum", "mean", "min", "max"])
def test_segment_with_hub(red, thresholds):
    rng = np.random.default_rng(11)
    N, E, F = 300, (70000 if thresholds == "default" else 6000), 8
    index = hub_index(rng, N, E, E * 4 // 7)
    src = rng.standard_normal((E, F)).astype(np.float32)
    src[rng.random((E, F)) < 0.1] = 0.0
    ptr = O.index2ptr(index, N)
    assert ops.segment_plan(cu(ptr), E) is not None
    x = cu(src).requires_grad_()
    out = U.segment(x, cu(ptr), red)
    ref = O.segment(src, ptr, red)
    assert_close(npy(out), ref, rtol=1e-4, atol=5e-3 if red == "sum" else 1e-6, msg=red)
    go = rng.standard_normal((N, F)).astype(np.float32)
    out.backward(cu(go))
    if red in ("sum", "mean"):
        gref = O.scatter_backward(go, src, ref, index, red)
        assert_close(npy(x.grad), gref, rtol=1e-5, atol=1e-6, msg="bwd " + red)
    else:   # _segment_reduce rule: even split among ties, no zero-initialised self
        eq = (src == ref[index])
        ties = np.zeros((N, F)); np.add.at(ties, index, eq)
        assert_close(npy(x.grad), eq * (go / np.maximum(ties, 1))[index], rtol=1e-5, atol=1e-6, msg="bwd " + red)
    # the sorted-index entry of scatter takes the same path
    assert torch.equal(U.scatter(cu(src), cu(index), 0, N, red, sorted=True), out.detach())


def test_softmax_with_hub(thresholds):
    rng = np.random.default_rng(12)
    N, E, H = 300, (70000 if thresholds == "default" else 6000), 8
    index = hub_index(rng, N, E, E * 4 // 7)
    src = (rng.standard_normal((E, H)) * 3).astype(np.float32)
    ptr = O.index2ptr(index, N)
    x = cu(src).requires_grad_()
    out = U.softmax(x, None, cu(ptr))
    ref = O.softmax(src, index, N)
    assert_close(npy(out), ref, rtol=1e-4, atol=1e-9)
    sums = np.zeros((N, H)); np.add.at(sums, index, npy(out))
    assert_close(sums[np.unique(index)], np.ones((np.unique(index).size, H)), rtol=1e-4, atol=0)
    go = rng.standard_normal((E, H)).astype(np.float32)
    out.backward(cu(go))
    assert_close(npy(x.grad), O.softmax_backward(go, ref, index, N), rtol=1e-3, atol=1e-7)
    # unsorted index: sort, chunked softmax, un-permute
    perm = rng.permutation(E)
    out2 = U.softmax(cu(src[perm]), cu(index[perm]), num_nodes=N)
    assert_close(npy(out2), ref[perm], rtol=1e-4, atol=1e-9)

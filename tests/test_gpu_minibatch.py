"""GPU parity of the mini-batch side (pytorch_geometric_b200/minibatch.py) against the golden run of the reference's
`trim_to_layer` + 3 x SAGEConv on a BFS-ordered sampled subgraph, and of `coalesce` (tests/golden/minibatch.npz)."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu

import pytorch_geometric_b200 as pgb  # noqa: E402
from pytorch_geometric_b200 import minibatch as MB, ops  # noqa: E402
from pytorch_geometric_b200.nn import SAGEConv  # noqa: E402

DEV = "cuda"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_trim_to_layer_on_an_adopted_sampled_graph_matches_the_reference_model():
    g = load_golden("minibatch")
    nodes_per_hop, edges_per_hop = g["nodes_per_hop"].tolist(), g["edges_per_hop"].tolist()
    ei, x = cu(g["ei"]), cu(g["x"]).requires_grad_()
    n = x.size(0)
    with pgb.debug():                                   # verifies that the loader-order edge list is destination-sorted
        graph = MB.sampled_graph(ei, n, n)
    assert graph.perm is None and graph.plan.n_long == 0
    convs = []
    for i in range(3):
        c = SAGEConv(x.size(1), x.size(1)).to(DEV)
        with torch.no_grad():
            c.lin_l.weight.copy_(cu(g[f"conv{i}_lin_l_w"]))
            c.lin_l.bias.copy_(cu(g[f"conv{i}_lin_l_b"]))
            c.lin_r.weight.copy_(cu(g[f"conv{i}_lin_r_w"]))
        convs.append(c)
    h, adj = x, graph
    sorts0 = ops.LAUNCHES.count
    for i, conv in enumerate(convs):
        h, adj, _ = MB.trim_to_layer(i, nodes_per_hop, edges_per_hop, h, adj)
        assert h.size(0) == int(g[f"trim{i}_n"]) and adj.num_edges == g[f"trim{i}_ei"].shape[1]
        assert adj.num_dst == adj.num_src == h.size(0)
        # the trimmed graph is a VIEW of the adopted CSR
        assert adj.col.data_ptr() == graph.col.data_ptr()
        h = conv(h, adj)
        if i < 2:
            h = h.relu()
        assert_close(h.detach().cpu().numpy(), g[f"h{i}"], rtol=1e-5, atol=1e-6, msg=f"layer {i}")
    out = h[:nodes_per_hop[0]]
    out.backward(cu(g["gout"]))
    assert_close(x.grad.cpu().numpy(), g["gx"], rtol=1e-4, atol=1e-6, msg="grad_x")
    del sorts0
    # the tensor form of trim_to_layer is the reference's narrow()
    _, e2, _ = MB.trim_to_layer(2, nodes_per_hop, edges_per_hop, x.detach(), ei)
    assert torch.equal(e2.cpu(), ei[:, :ei.size(1) - edges_per_hop[-2]].cpu())


def test_sampled_graph_rejects_unsorted_edges_in_debug_mode():
    ei = torch.tensor([[0, 1, 2], [2, 0, 1]], device=DEV)
    with pgb.debug():
        with pytest.raises(ValueError, match="not sorted"):
            MB.sampled_graph(ei, 3, 3)


@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
@pytest.mark.parametrize("by_row", [True, False])
def test_coalesce_golden(reduce, by_row):
    g = load_golden("minibatch")
    oe, oa = MB.coalesce(cu(g["c_ei"]), cu(g["c_attr"]), num_nodes=7, reduce=reduce, sort_by_row=by_row)
    assert np.array_equal(oe.cpu().numpy(), g[f"c_{reduce}_{int(by_row)}_ei"])          # integer work: bit-exact
    assert_close(oa.cpu().numpy(), g[f"c_{reduce}_{int(by_row)}_attr"], rtol=1e-6, atol=1e-7)
    # already-coalesced input comes back unchanged
    oe2, oa2 = MB.coalesce(oe, oa, num_nodes=7, reduce=reduce, sort_by_row=by_row)
    assert torch.equal(oe2, oe) and torch.equal(oa2, oa)

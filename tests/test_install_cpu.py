"""The reference-side binding (pytorch_geometric_b200/install.py) against the unmodified reference package
(fixture `tg`: baseline/_ref, else /root/reference in the build container; skipped where neither exists)."""
import os
import sys

import pytest
import torch


def test_install_rebinds_consumers_and_cpu_falls_through(tg):
    import torch_geometric.nn.aggr.base as aggr_base
    import torch_geometric.nn.conv.gcn_conv as gcn_conv
    from torch_geometric.utils import _scatter

    import pytorch_geometric_b200.install as b200

    orig = _scatter.scatter
    src = torch.randn(10, 4)
    index = torch.tensor([0, 1, 0, 1, 2, 2, 3, 3, 3, 0])
    before = orig(src, index, 0, 5, "sum")
    counts = b200.install(layers=False)
    try:
        assert counts["scatter"] >= 10 and counts["softmax"] >= 1 and counts["spmm"] >= 1
        assert _scatter.scatter is not orig and aggr_base.scatter is _scatter.scatter
        assert gcn_conv.scatter is _scatter.scatter
        # CPU tensors still run the untouched reference: bit-identical results
        assert torch.equal(_scatter.scatter(src, index, 0, 5, "sum"), before)
        conv = tg.nn.GCNConv(4, 3)
        ei = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 0]])
        x = torch.randn(4, 4)
        out1 = conv(x, ei)
    finally:
        b200.uninstall()
    assert _scatter.scatter is orig and aggr_base.scatter is orig
    assert torch.equal(conv(x, ei), out1)


def test_install_layers_swaps_conv_classes(tg):
    import pytorch_geometric_b200.install as b200
    from pytorch_geometric_b200 import nn as ours

    orig = tg.nn.GCNConv
    b200.install(layers=True)
    try:
        assert tg.nn.GCNConv is ours.GCNConv and tg.nn.GATConv is ours.GATConv
    finally:
        b200.uninstall()
    assert tg.nn.GCNConv is orig


def test_install_routes_fused_aggregation_and_cpu_falls_through(tg):
    from torch_geometric.nn.aggr.fused import FusedAggregation

    import pytorch_geometric_b200.install as b200

    orig = FusedAggregation.forward
    aggr = FusedAggregation(["sum", "mean", "max", "std"])
    x = torch.randn(12, 3)
    index = torch.tensor([0, 0, 1, 1, 1, 3, 3, 3, 3, 4, 4, 0])
    before = aggr(x, index, dim_size=6)
    counts = b200.install()
    try:
        assert counts["fused_aggregation"] == 1 and FusedAggregation.forward is not orig
        after = aggr(x, index, dim_size=6)             # CPU tensors: untouched reference
        assert all(torch.equal(a, b) for a, b in zip(before, after))
        # a list with 'mul' is not fusable by the engine and must fall through as well
        assert len(FusedAggregation(["sum", "mul"])(x, index, dim_size=6)) == 2
    finally:
        b200.uninstall()
    assert FusedAggregation.forward is orig

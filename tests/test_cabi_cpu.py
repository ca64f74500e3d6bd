"""CPU-side checks of the boundary: the C-ABI library loads without a GPU, exports every symbol
include/b200mp.h declares, and the host-side mirror refuses to run anywhere but on CUDA."""
import numpy as np
import pytest
import torch

import pytorch_geometric_b200 as pgb
from pytorch_geometric_b200 import _lib, ops, utils as U
from pytorch_geometric_b200.nn import GCNConv, SumAggregation


def test_library_loads_and_exports_every_declared_symbol():
    lib = pgb.lib()
    syms = pgb.header_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.b200mp_version() == b"0.1.0"
    # every symbol the Python binding declares is in the header, and vice versa
    assert sorted(_lib._SIGS) == syms


def test_no_cpu_fallback():
    x = torch.randn(4, 8)
    idx = torch.tensor([0, 1, 1, 3])
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.scatter_coo(x, idx, 4, "sum")
    with pytest.raises(RuntimeError, match="CUDA"):
        U.scatter(x, idx, 0, 4, "sum")
    with pytest.raises(RuntimeError, match="CUDA"):
        pgb.CSRGraph(idx, idx, 4, 4)
    with pytest.raises(RuntimeError):
        GCNConv(8, 8)(x, torch.stack([idx, idx]))


def test_argument_errors_match_reference_messages():
    x = torch.randn(4, 8)
    idx = torch.tensor([0, 1, 1, 3])
    with pytest.raises(ValueError, match="must be one-dimensional"):      # _scatter.py:37-39
        U.scatter(x, idx.view(2, 2), 0, 4)
    with pytest.raises(ValueError, match="must lay between"):            # _scatter.py:43-45
        U.scatter(x, idx, 3, 4)
    with pytest.raises(ValueError, match="invalid `reduce` argument"):   # _scatter.py:138
        U.scatter(x, idx, 0, 4, "foo")
    with pytest.raises(ValueError, match="invalid dimension"):           # aggr/base.py:105-107
        SumAggregation()(x, idx, dim=4)


def test_state_dict_layout_matches_reference_layers():
    sd = GCNConv(8, 16).state_dict()
    assert set(sd) == {"bias", "lin.weight"} and sd["lin.weight"].shape == (16, 8)
    from pytorch_geometric_b200.nn import RGCNConv, SAGEConv
    assert set(SAGEConv(8, 16).state_dict()) == {"lin_l.weight", "lin_l.bias", "lin_r.weight"}
    sd = RGCNConv(8, 16, 3).state_dict()
    assert sd["weight"].shape == (3, 8, 16) and sd["root"].shape == (8, 16) and sd["bias"].shape == (16, )

"""Plugs the engine into an installed `torch_geometric` (the reference-side binding of
INTEGRATION.md).  The reference dispatches at call time on Python-level seams
(SURVEY.md section 8(b)); `install()` rebinds exactly those:

  torch_geometric.utils._scatter.scatter      -> utils.scatter      (CUDA tensors only)
  torch_geometric.utils._segment.segment      -> utils.segment
  torch_geometric.utils._softmax.softmax      -> utils.softmax
  torch_geometric.utils._spmm.spmm            -> utils.spmm  (CSRGraph / torch.sparse_csr on CUDA)
  torch_geometric.nn.aggr.fused.FusedAggregation.forward -> the one-sweep multi-aggregation
        (so MultiAggregation / PNA-style `aggr=[...]` lists on CUDA tensors take one pass over the messages)
  torch_geometric.nn.{GCNConv,SAGEConv,GINConv,GATConv,RGCNConv} -> the fused layers (optional)

plus every module that imported those names (`from torch_geometric.utils import scatter` binds
early: 64 modules).  CPU tensors fall through to the untouched reference implementation, which is
also how the parity oracle keeps working next to the engine.  `uninstall()` restores everything.
"""
from __future__ import annotations

import sys
from typing import Dict, List, Tuple

_PATCHED: List[Tuple[object, str, object]] = []


def _route(ours, theirs, probe):
    def dispatch(*args, **kwargs):
        t = probe(args, kwargs)
        if t is not None and getattr(t, "is_cuda", False):
            return ours(*args, **kwargs)
        return theirs(*args, **kwargs)
    dispatch.__name__ = getattr(theirs, "__name__", "dispatch")
    dispatch.__doc__ = getattr(theirs, "__doc__", None)
    dispatch.__wrapped__ = theirs
    return dispatch


def _first_tensor(args, kwargs):
    import torch
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            return a
    return None


def install(layers: bool = False) -> Dict[str, int]:
    """Rebinds the reference's seams to the engine.  Returns {name: number of rebinds}."""
    import torch_geometric  # noqa: F401  (must be importable: this is the reference-side binding)
    from torch_geometric.utils import _scatter, _segment, _softmax, _spmm

    from . import utils as U

    if _PATCHED:
        return {}
    table = {
        "scatter": (_scatter, U.scatter),
        "segment": (_segment, U.segment),
        "softmax": (_softmax, U.softmax),
    }
    counts: Dict[str, int] = {}
    for name, (home, ours) in table.items():
        theirs = getattr(home, name)
        routed = _route(ours, theirs, _first_tensor)
        n = 0
        for mod in list(sys.modules.values()):
            if mod is None or not getattr(mod, "__name__", "").startswith("torch_geometric"):
                continue
            if getattr(mod, name, None) is theirs:
                _PATCHED.append((mod, name, theirs))
                setattr(mod, name, routed)
                n += 1
        counts[name] = n

    # spmm: only our own adjacency handle / CUDA sparse CSR goes to the engine
    from .graph import CSRGraph
    theirs_spmm = _spmm.spmm

    def spmm(src, other, reduce: str = "sum"):
        if isinstance(src, CSRGraph) or (getattr(other, "is_cuda", False) and getattr(src, "layout", None) is not None
                                         and str(src.layout) == "torch.sparse_csr" and not other.requires_grad):
            return U.spmm(src, other, reduce)
        return theirs_spmm(src, other, reduce)

    n = 0
    for mod in list(sys.modules.values()):
        if mod is None or not getattr(mod, "__name__", "").startswith("torch_geometric"):
            continue
        if getattr(mod, "spmm", None) is theirs_spmm:
            _PATCHED.append((mod, "spmm", theirs_spmm))
            setattr(mod, "spmm", spmm)
            n += 1
    counts["spmm"] = n

    # FusedAggregation.forward (nn/aggr/fused.py:191): every fusable list except those containing 'mul'
    from torch_geometric.nn.aggr.fused import FusedAggregation as TheirFused

    from .nn import aggr as our_aggr
    theirs_fwd = TheirFused.forward

    def fused_forward(self, x, index=None, ptr=None, dim_size=None, dim=-2):
        names = [our_aggr.FusedAggregation.NAME.get(n) for n in self.aggr_names]
        if not getattr(x, "is_cuda", False) or None in names or x.dim() != 2 or index is None:
            return theirs_fwd(self, x, index, ptr, dim_size, dim)
        if dim_size is None:
            dim_size = ptr.numel() - 1 if ptr is not None else (int(index.max()) + 1 if index.numel() > 0 else 0)
        uniq = list(dict.fromkeys(names))
        # the reference ignores `ptr` here and scatters by `index`; a given ptr means the index is sorted
        outs = dict(zip(uniq, our_aggr._fused_forward(uniq, self.semi_grad, x, index, ptr, dim_size, dim, False)))
        return [outs[n] for n in names]

    _PATCHED.append((TheirFused, "forward", theirs_fwd))
    TheirFused.forward = fused_forward
    counts["fused_aggregation"] = 1

    if layers:
        import torch_geometric.nn as tgnn
        import torch_geometric.nn.conv as tgconv

        from . import nn as ours_nn
        for cls in ("GCNConv", "SAGEConv", "GINConv", "GATConv", "RGCNConv"):
            for mod in (tgnn, tgconv):
                if hasattr(mod, cls):
                    _PATCHED.append((mod, cls, getattr(mod, cls)))
                    setattr(mod, cls, getattr(ours_nn, cls))
        counts["layers"] = 5
    return counts


def uninstall() -> None:
    while _PATCHED:
        mod, name, orig = _PATCHED.pop()
        setattr(mod, name, orig)

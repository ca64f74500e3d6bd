"""Shared helpers of the plug-in."""
from torch import Tensor


def plain(t):
    """The raw index tensor behind the reference's wrapper subclasses (`Index`, `EdgeIndex`: torch_geometric/index.py:324,
    edge_index.py:721 `as_tensor()`); plain tensors pass through."""
    if t is None or type(t) is Tensor:
        return t
    as_t = getattr(t, "as_tensor", None)
    if callable(as_t):
        return as_t()
    return t.as_subclass(Tensor)

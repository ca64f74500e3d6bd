"""Debug switch, the counterpart of torch_geometric.debug() (torch_geometric/debug.py): with it on, the host-side
mirrors validate index ranges with a device->host read before launching (what the reference learns from the
backend's exception, aggr/base.py:130-139, message_passing.py:269-290); off (default) nothing synchronises."""
_ENABLED = False


def enabled() -> bool:
    return _ENABLED


class debug:
    """`with pytorch_geometric_b200.debug(): ...` or `pytorch_geometric_b200.set_debug(True)`."""

    def __init__(self, on: bool = True):
        self.on, self.prev = bool(on), None

    def __enter__(self):
        global _ENABLED
        self.prev, _ENABLED = _ENABLED, self.on
        return self

    def __exit__(self, *exc):
        global _ENABLED
        _ENABLED = self.prev
        return False


def set_debug(on: bool) -> None:
    global _ENABLED
    _ENABLED = bool(on)

"""Backward-compatible entry point of the reference-side binding: see `pytorch_geometric_b200.plugin`."""
from .plugin import install, installed, uninstall  # noqa: F401

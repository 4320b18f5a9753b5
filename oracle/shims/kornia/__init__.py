"""Shim: see oracle/shims/README.md."""
from . import geometry  # noqa: F401

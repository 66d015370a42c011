"""Namespace overlay for the reference's `torch_utils` package (see ../README.md)."""
import os

_ref = os.environ.get("LAYOUTDETR_REFERENCE", "/root/reference")
_ref_pkg = os.path.join(_ref, "torch_utils")
if os.path.isdir(_ref_pkg) and _ref_pkg not in __path__:
    __path__.append(_ref_pkg)

"""Namespace overlay for the reference's `training` package: hot-path modules come from layoutdetr_b200,
everything else from the unmodified checkout at $LAYOUTDETR_REFERENCE."""
import os

_ref = os.environ.get("LAYOUTDETR_REFERENCE", "/root/reference")
_ref_pkg = os.path.join(_ref, "training")
if os.path.isdir(_ref_pkg) and _ref_pkg not in __path__:
    __path__.append(_ref_pkg)

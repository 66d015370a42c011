"""Namespace overlay for the reference's `metrics` package: the two layout-metric modules `metric_main` dispatches to come from
layoutdetr_b200 (one GPU sweep serves both), everything else from the unmodified checkout at $LAYOUTDETR_REFERENCE."""
import os

_ref = os.environ.get("LAYOUTDETR_REFERENCE", "/root/reference")
_ref_pkg = os.path.join(_ref, "metrics")
if os.path.isdir(_ref_pkg) and _ref_pkg not in __path__:
    __path__.append(_ref_pkg)

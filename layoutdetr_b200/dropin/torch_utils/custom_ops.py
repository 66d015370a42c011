"""Overlay for torch_utils/custom_ops.py (reference :62 `get_plugin` JIT-compiles the CUDA plugins with nvcc at first
use).  The sm_100a ops are prebuilt into liblayoutdetr_sm100.so, so there is nothing to compile at run time; the
module keeps the attribute `train.py` / `training_loop.py` touch (`custom_ops.verbosity`) and fails loudly if
something still asks for a JIT plugin."""
from layoutdetr_b200 import _lib

verbosity = 'brief'


def get_plugin(module_name, sources, headers=None, source_dir=None, **build_kwargs):
    _lib.lib()
    raise RuntimeError("torch_utils.custom_ops.get_plugin(%r): the LayoutDETR hot-path ops are served by "
                       "liblayoutdetr_sm100.so (layoutdetr_b200.torch_utils.ops); no JIT plugin is available" % module_name)

"""layoutdetr_b200 — B200-native (sm_100a) implementation of the LayoutDETR generator /
discriminator forward-backward hot path behind the reference's own Python API.

Layout:
  csrc/            hand-written CUDA kernels + the C-ABI (include/layoutdetr_sm100.h)
  _lib.py          ctypes loader of liblayoutdetr_sm100.so (fails loudly when missing)
  kernels.py       thin tensor->pointer wrappers over the C-ABI
  ...              host-side mirror of the reference interface (training/, torch_utils/)
"""
__version__ = "0.1.0"

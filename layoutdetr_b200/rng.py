"""Dropout generator state shared by the training-mode kernels.

The reference trains G and D in `.train()` (training/training_loop.py:133-134): DETR dropout 0.1
(training/detr_transformer.py:185-194) and BERT hidden / attention-probability dropout 0.1 (configs/med_config.json:5,7;
training/med.py:96,213,240,318) are live, even inside the frozen text encoder.  Here every dropout site draws its mask from a
counter-based Philox4x32-10 stream (csrc/common.cuh: DropoutRng) addressed by

    key     = the 64-bit seed                      (device, `manual_seed`)
    counter = (element group, site, step)          group = 8 consecutive elements

`site` is a host-side integer handed out per dropout call (`next_site`); the backward pass of that call re-uses it, so the
mask is regenerated from coordinates instead of being stored.  `step` lives on the device and is bumped by `advance()` — one
tiny kernel at the start of every training iteration, captured into the iteration's CUDA graph — so a replayed graph (whose
site numbers are baked in) still draws fresh masks every step.
"""
import ctypes

import torch

from ._lib import lib, check

_state = {}          # device index -> uint32[4] tensor {seed_lo, seed_hi, step, 0}
_site = [0]
_seed = [0x5EED5EED12345678]
ENABLED = [True]     # global switch: False forces p = 0 everywhere (deterministic eval semantics even under .train())


def _dev_index(device=None):
    if device is None:
        return torch.cuda.current_device()
    device = torch.device(device)
    return device.index if device.index is not None else torch.cuda.current_device()


def state(device=None):
    """Device tensor holding the generator state (created on first use)."""
    idx = _dev_index(device)
    st = _state.get(idx)
    if st is None:
        s = _seed[0]
        host = torch.tensor([s & 0xFFFFFFFF, (s >> 32) & 0xFFFFFFFF, 0, 0], dtype=torch.int64).to(torch.int32)
        st = host.to(torch.device("cuda", idx))
        _state[idx] = st
    return st


def manual_seed(seed, device=None):
    """Re-seed (and reset the step of) the dropout stream; in place, so captured graphs keep reading the same buffer."""
    _seed[0] = int(seed) & 0xFFFFFFFFFFFFFFFF
    s = _seed[0]
    vals = [s & 0xFFFFFFFF, (s >> 32) & 0xFFFFFFFF, 0, 0]
    vals = [v - (1 << 32) if v >= (1 << 31) else v for v in vals]
    for idx, st in _state.items():
        if device is None or idx == _dev_index(device):
            st.copy_(torch.tensor(vals, dtype=torch.int32))
    _site[0] = 0


def next_site():
    _site[0] = (_site[0] + 1) & 0xFFFFFFFF
    return _site[0]


def advance(device=None):
    """step += 1 on the device (stream-ordered, graph-capturable)."""
    st = state(device)
    check(lib().ld_rng_advance(ctypes.c_void_p(st.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
          "ld_rng_advance")


def p_of(module_training, p):
    """Effective dropout probability of a site."""
    return float(p) if (module_training and ENABLED[0] and p > 0.0) else 0.0

"""Host-side engine state: bf16 weight shadows, direct gradient accumulation, GEMM heuristics.

The master parameters stay fp32 `nn.Parameter`s (the training loop reads / writes `.grad`, the
optimizer updates them, `state_dict` keys match the reference).  Tensor cores read bf16 shadows that
are refreshed lazily whenever a parameter's version counter changes (optimizer step, load_state_dict,
EMA copy).
"""
import weakref

import torch

from . import kernels as K

_shadow = {}          # (id(param), tag) -> (version, data_ptr, generation, tensor, params, make_fn)
_managed = {}         # id(param) -> bf16 [N, K] view kept fresh by the fused optimizer kernel (flat storage)
_generation = {}      # id(param) -> int, bumped when a kernel rewrites the parameter through a raw pointer
_SM_COUNT = None


def sm_count():
    global _SM_COUNT
    if _SM_COUNT is None:
        _SM_COUNT = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return _SM_COUNT


def clear_cache():
    _shadow.clear()
    _managed.clear()
    _generation.clear()


def register_managed(param, shadow):
    """`shadow` (bf16 [N, K] view of the flat bf16 buffer) is rewritten together with `param` by ld_adam_flat /
    ld_ema_flat, so it is always current."""
    _managed[id(param)] = (param.data_ptr(), shadow)


def bump_generation(params):
    """Invalidate derived shadows (concatenations, conv re-layouts, padded copies) after an in-kernel update."""
    for p in params:
        _generation[id(p)] = _generation.get(id(p), 0) + 1


def _lookup(key, params):
    ent = _shadow.get(key)
    if ent is None:
        return None
    ver = tuple(p._version for p in params)
    ptr = tuple(p.data_ptr() for p in params)
    gen = tuple(_generation.get(id(p), 0) for p in params)
    if ent[0] == ver and ent[1] == ptr and ent[2] == gen:
        return ent[3]
    return None


def _store(key, params, value, fn=None):
    # Refresh IN PLACE when a previous shadow of the same shape exists: the device address of a derived shadow must stay
    # stable so that a captured CUDA graph (which bakes pointers) keeps reading the buffer a later refresh writes.
    old = _shadow.get(key)
    if old is not None and torch.is_tensor(old[3]) and old[3].shape == value.shape and old[3].dtype == value.dtype \
            and old[3].device == value.device:
        old[3].copy_(value)
        value = old[3]
    if fn is None and old is not None:
        fn = old[5]
    _shadow[key] = (tuple(p._version for p in params), tuple(p.data_ptr() for p in params),
                    tuple(_generation.get(id(p), 0) for p in params), value, tuple(weakref.ref(p) for p in params), fn)
    return value


def refresh_stale(only_ids=None):
    """Recompute (in place) every cached shadow whose source parameters changed since it was made, on the CURRENT stream.
    The lane scheduler (lanes.py) calls this before it forks: a shadow refreshed lazily inside one lane would race with
    readers in another lane.  Plain weight shadows first — concatenations are built from them.  `only_ids`: restrict to
    shadows derived from parameters / buffers with these `id()`s (the modules the caller is about to run)."""
    stale = []
    for key, ent in list(_shadow.items()):
        params = tuple(r() for r in ent[4])
        if any(p is None for p in params):          # the module is gone
            del _shadow[key]
            continue
        if ent[5] is None or (only_ids is not None and not all(id(p) in only_ids for p in params)):
            continue
        if ent[0] != tuple(p._version for p in params) or ent[1] != tuple(p.data_ptr() for p in params) or \
                ent[2] != tuple(_generation.get(id(p), 0) for p in params):
            stale.append((0 if key[1] == "w" else 1, key, params, ent[5]))
    stale.sort(key=lambda t: t[0])
    for _, key, params, fn in stale:
        with torch.no_grad():
            v = fn()
        _store(key, params, v, fn)
    return len(stale)


def pad8(n):
    return (n + 7) // 8 * 8


def w_bf16(param):
    """bf16 shadow of a 2-D weight [N, K] with K zero-padded to a multiple of 8 (TMA stride rule)."""
    man = _managed.get(id(param))
    if man is not None and man[0] == param.data_ptr():
        return man[1]
    key = (id(param), "w")
    v = _lookup(key, (param,))
    if v is None:
        def make():
            w = param.detach()
            w2 = w.reshape(w.shape[0], -1)
            return K.cast_pad(w2, torch.bfloat16, pad8(w2.shape[1]))
        with torch.no_grad():
            v = make()
        v = _store(key, (param,), v, make)
    return v


def w_cat_bf16(params):
    """bf16 shadow of several [Ni, K] weights concatenated along N (fused QKV projections)."""
    key = (tuple(id(p) for p in params), "cat")
    v = _lookup(key, params)
    if v is None:
        make = lambda: torch.cat([w_bf16(p) for p in params], dim=0)
        with torch.no_grad():
            v = make()
        v = _store(key, params, v, make)
    return v


def b_cat_f32(params):
    key = (tuple(id(p) for p in params), "bcat")
    v = _lookup(key, params)
    if v is None:
        make = lambda: torch.cat([p.detach().float() for p in params], dim=0).contiguous()
        with torch.no_grad():
            v = make()
        v = _store(key, params, v, make)
    return v


def derived(param_tuple, tag, fn):
    """Cache an arbitrary tensor derived from parameters / buffers (e.g. folded FrozenBN scale+bias,
    conv weights re-laid out as [Cout, kh*kw*Cin])."""
    key = (tuple(id(p) for p in param_tuple), tag)
    v = _lookup(key, param_tuple)
    if v is None:
        with torch.no_grad():
            v = fn()
        v = _store(key, param_tuple, v, fn)
    return v


def grad_buffer(param):
    """fp32 gradient accumulator of a parameter (created zeroed on first use).  Backward kernels add
    straight into it (GEMM epilogue accumulate / atomics) instead of materialising a temporary that
    autograd would add afterwards."""
    if param.grad is None:
        param.grad = torch.zeros_like(param, dtype=torch.float32, memory_format=torch.contiguous_format)
    return param.grad


def wgrad_split_k(M_out, N_out, K_red):
    """Split-K factor for weight-gradient GEMMs (few output tiles, long reduction)."""
    tiles = ((M_out + 127) // 128) * ((N_out + 127) // 128)
    kb = (K_red + 63) // 64
    sms = sm_count()
    if tiles >= sms or kb <= 8:
        return 1
    return int(max(1, min(kb // 4, (2 * sms) // tiles)))

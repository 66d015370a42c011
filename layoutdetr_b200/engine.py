"""Host-side engine state: bf16 weight shadows, direct gradient accumulation, GEMM heuristics.

The master parameters stay fp32 `nn.Parameter`s (the training loop reads / writes `.grad`, the
optimizer updates them, `state_dict` keys match the reference).  Tensor cores read bf16 shadows that
are refreshed lazily whenever a parameter's version counter changes (optimizer step, load_state_dict,
EMA copy).
"""
import weakref

import torch

from . import kernels as K

_shadow = {}          # (id(param)..., tag) -> (versions, data_ptrs, generations, tensor, weakrefs of the params, make_fn(*params))
_managed = {}         # id(param) -> [weakref(param), data_ptr, version, bf16 [N, K] view kept fresh by the fused optimizer kernel]
_generation = {}      # id(param) -> int, bumped when a kernel rewrites the parameter through a raw pointer
_SM_COUNT = None

# Ownership rules (a long run deep-copies G_ema for every snapshot / evaluation sweep and drops the copies again):
#   * nothing in these tables holds a strong reference to a parameter — make functions receive the live tensors as arguments;
#   * an entry is only ever returned for the very object it was built from (`weakref() is param`, so a recycled id() cannot hit);
#   * entries die with their parameters (weakref.finalize), taking the bf16 shadows with them.


def sm_count():
    global _SM_COUNT
    if _SM_COUNT is None:
        _SM_COUNT = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return _SM_COUNT


def clear_cache():
    _shadow.clear()
    _managed.clear()
    _generation.clear()


def _evict(key):
    _shadow.pop(key, None)


def _forget_param(pid):
    _managed.pop(pid, None)
    _generation.pop(pid, None)


def register_managed(param, shadow):
    """`shadow` (bf16 [N, K] view of the flat bf16 buffer) is rewritten together with `param` by ld_adam_flat /
    ld_ema_flat.  A torch-level in-place write to the parameter (load_state_dict, copy_params_and_buffers, a broadcast from
    rank 0) bumps its version counter instead; w_bf16 / refresh_stale then re-cast the shadow before anything reads it."""
    _managed[id(param)] = [weakref.ref(param), param.data_ptr(), param._version, shadow]
    weakref.finalize(param, _forget_param, id(param))


def _managed_shadow(param):
    man = _managed.get(id(param))
    if man is None or man[0]() is not param or man[1] != param.data_ptr():
        return None
    if man[2] != param._version:                       # written behind the optimizer kernel's back: re-cast in place
        with torch.no_grad():
            man[3].copy_(param.detach().reshape(man[3].shape))
        man[2] = param._version
    return man[3]


def bump_generation(params):
    """Invalidate derived shadows (concatenations, conv re-layouts, padded copies) after an in-kernel update."""
    for p in params:
        _generation[id(p)] = _generation.get(id(p), 0) + 1


def _stamp(params):
    return (tuple(p._version for p in params), tuple(p.data_ptr() for p in params),
            tuple(_generation.get(id(p), 0) for p in params))


def _lookup(key, params):
    ent = _shadow.get(key)
    if ent is None:
        return None
    if any(r() is not p for r, p in zip(ent[4], params)):       # same id(), different object: a stale entry of a dead module
        del _shadow[key]
        return None
    if (ent[0], ent[1], ent[2]) == _stamp(params):
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
    if old is None:
        for p in params:
            weakref.finalize(p, _evict, key)
    ver, ptr, gen = _stamp(params)
    _shadow[key] = (ver, ptr, gen, value, tuple(weakref.ref(p) for p in params), fn)
    return value


def refresh_stale(only_ids=None):
    """Recompute (in place) every cached shadow whose source parameters changed since it was made, on the CURRENT stream.
    The lane scheduler (lanes.py) calls this before it forks: a shadow refreshed lazily inside one lane would race with
    readers in another lane.  Plain weight shadows first — concatenations are built from them.  `only_ids`: restrict to
    shadows derived from parameters / buffers with these `id()`s (the modules the caller is about to run)."""
    n = 0
    for pid, man in list(_managed.items()):
        p = man[0]()
        if p is None:
            _managed.pop(pid, None)
        elif (only_ids is None or pid in only_ids) and man[2] != p._version:
            _managed_shadow(p)
            n += 1
    stale = []
    for key, ent in list(_shadow.items()):
        params = tuple(r() for r in ent[4])
        if any(p is None for p in params):          # the module is gone
            del _shadow[key]
            continue
        if ent[5] is None or (only_ids is not None and not all(id(p) in only_ids for p in params)):
            continue
        if (ent[0], ent[1], ent[2]) != _stamp(params):
            stale.append((0 if key[1] == "w" else 1, key, params, ent[5]))
    stale.sort(key=lambda t: t[0])
    for _, key, params, fn in stale:
        with torch.no_grad():
            v = fn(*params)
        _store(key, params, v, fn)
    return n + len(stale)


def pad8(n):
    return (n + 7) // 8 * 8


def _make_w(param):
    w = param.detach()
    w2 = w.reshape(w.shape[0], -1)
    return K.cast_pad(w2, torch.bfloat16, pad8(w2.shape[1]))


def w_bf16(param):
    """bf16 shadow of a 2-D weight [N, K] with K zero-padded to a multiple of 8 (TMA stride rule)."""
    man = _managed_shadow(param)
    if man is not None:
        return man
    key = (id(param), "w")
    v = _lookup(key, (param,))
    if v is None:
        with torch.no_grad():
            v = _make_w(param)
        v = _store(key, (param,), v, _make_w)
    return v


def _make_wcat(*params):
    return torch.cat([w_bf16(p) for p in params], dim=0)


def _make_bcat(*params):
    return torch.cat([p.detach().float() for p in params], dim=0).contiguous()


def w_cat_bf16(params):
    """bf16 shadow of several [Ni, K] weights concatenated along N (fused QKV projections)."""
    params = tuple(params)
    key = (tuple(id(p) for p in params), "cat")
    v = _lookup(key, params)
    if v is None:
        with torch.no_grad():
            v = _make_wcat(*params)
        v = _store(key, params, v, _make_wcat)
    return v


def b_cat_f32(params):
    params = tuple(params)
    key = (tuple(id(p) for p in params), "bcat")
    v = _lookup(key, params)
    if v is None:
        with torch.no_grad():
            v = _make_bcat(*params)
        v = _store(key, params, v, _make_bcat)
    return v


def derived(param_tuple, tag, fn):
    """Cache an arbitrary tensor derived from parameters / buffers (e.g. folded FrozenBN scale+bias, conv weights re-laid out
    as [Cout, kh*kw*Cin]).  `fn(*param_tuple)` builds it and must not capture the tensors (it is kept for in-place refreshes)."""
    param_tuple = tuple(param_tuple)
    key = (tuple(id(p) for p in param_tuple), tag)
    v = _lookup(key, param_tuple)
    if v is None:
        with torch.no_grad():
            v = fn(*param_tuple)
        v = _store(key, param_tuple, v, fn)
    return v


_grad_watch = {}        # id(param) -> (GradExchange, bucket): parameters whose gradient writes are reported (data-parallel runs only)
_grad_pending = []      # writes handed out by grad_buffer since the last grad_writes_done()


def grad_buffer(param):
    """fp32 gradient accumulator of a parameter (created zeroed on first use).  Backward kernels add
    straight into it (GEMM epilogue accumulate / atomics) instead of materialising a temporary that
    autograd would add afterwards."""
    if param.grad is None:
        param.grad = torch.zeros_like(param, dtype=torch.float32, memory_format=torch.contiguous_format)
    if _grad_watch:
        ent = _grad_watch.get(id(param))
        if ent is not None:
            _grad_pending.append(ent)
    return param.grad


def watch_grad(param, exchange, bucket):
    _grad_watch[id(param)] = (exchange, bucket)


def unwatch_grads(exchange):
    for k in [k for k, v in _grad_watch.items() if v[0] is exchange]:
        del _grad_watch[k]


def grad_writes_done():
    """Called at the end of every hand-written backward node (functional._Fn): the kernels that write the buffers handed out
    by grad_buffer since the last call have been issued on the current stream (exchange.GradExchange)."""
    if _grad_pending:
        pend = list(_grad_pending)
        _grad_pending.clear()
        for ex, b in pend:
            ex.written(b)


def wgrad_split_k(M_out, N_out, K_red):
    """Split-K factor for weight-gradient GEMMs (few output tiles, long reduction; partial sums meet through fp32 atomics).

    Cost model fitted to the B200 sweep in profiles/r2_splitk_probe.txt (both operands MN-major): a CTA spends ~0.43 us per
    64-deep K block on a 128 x 128 tile (a CTA pair ~0.40 us on its 256 x 256 tile), a launch costs ~5 us, and the atomic adds
    of all splits drain at ~160 G fp32 atomics / s — the best factor balances the K loop against sk * M * N atomics.  (The
    round-1 rule "fill 2 x the SMs" over-split small outputs: 256 x 64 x 65536 ran 148 splits in 40 us, 32 splits take 20 us;
    512 x 4608 x 1024 ran 2 splits in 34 us, one takes 15 us.)"""
    kb = (K_red + 63) // 64
    sms = sm_count()
    tiles = ((M_out + 127) // 128) * ((N_out + 127) // 128)
    pairs = ((M_out + 255) // 256) * ((N_out + 255) // 256)
    pair_ok = M_out >= 256 and N_out > 128
    best, best_cost = 1, None
    for sk in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128):
        if sk > max(1, kb // 2):
            break
        per_split = -(-kb // sk)
        if pair_ok and pairs * sk >= (sms * 3 // 8 if per_split >= 32 else sms // 2):      # gemm_sm100.cu takes the CTA-pair kernel
            loop = (-(-(pairs * sk) // (sms // 2))) * per_split * 0.40
        else:
            loop = (-(-(tiles * sk) // sms)) * per_split * 0.43
        atomics = 0.0 if sk == 1 else sk * M_out * N_out / 160000.0
        cost = 5.0 + loop + atomics
        if best_cost is None or cost < best_cost * 0.97:        # prefer fewer splits on near ties
            best, best_cost = sk, cost
    return best

"""The pieces of the reference's torch_utils/misc.py that the training loop of the hot path uses (reference lines in each
docstring), written against the product modules: name-matched weight copies, replica-consistency check, module summary
through forward hooks, the data-parallel sampler.  Everything else of that file (assert_shape, profiled_function, ...)
is not on the path."""
import re

import torch

from ..training.sampler import InfiniteSampler  # noqa: F401  (reference :114-148)


def params_and_buffers(module):
    """Parameters then buffers, in registration order (reference :154-156)."""
    assert isinstance(module, torch.nn.Module)
    return list(module.parameters()) + list(module.buffers())


def named_params_and_buffers(module):
    assert isinstance(module, torch.nn.Module)
    return list(module.named_parameters()) + list(module.named_buffers())


def copy_params_and_buffers(src_module, dst_module, require_all=False):
    """dst tensors <- src tensors of the same name, in place (reference :162-169).  In-place matters here: with flat storage
    (flat.FlatParams) the destination tensors are views of one buffer, and the bf16 shadows notice the version bump."""
    src = dict(named_params_and_buffers(src_module))
    with torch.no_grad():
        for name, t in named_params_and_buffers(dst_module):
            if name not in src:
                if require_all:
                    raise KeyError("%s missing from the source module" % name)
                continue
            t.copy_(src[name].detach().to(t.device, t.dtype))


def check_ddp_consistency(module, ignore_regex=None, group=None):
    """Every rank holds the same values as rank 0 (reference :183-194): broadcast rank 0's copy and compare bit for bit
    (NaNs compare equal).  Raises on the first mismatching tensor."""
    full = type(module).__name__
    for name, t in named_params_and_buffers(module):
        if ignore_regex is not None and re.fullmatch(ignore_regex, full + "." + name):
            continue
        t = t.detach()
        if t.is_floating_point():
            t = torch.nan_to_num(t)
        other = t.clone()
        torch.distributed.broadcast(tensor=other, src=0, group=group)
        if not bool((t == other).all()):
            raise AssertionError("replicas diverged at %s.%s" % (full, name))


def weight_checksum(module):
    """One fp64 number per module that is identical on identical replicas: sum over tensors of (i + 1) * sum(t) in fp64."""
    acc = None
    for i, (_, t) in enumerate(named_params_and_buffers(module)):
        if not t.is_floating_point():
            continue
        v = t.detach().double().sum() * float(i + 1)
        acc = v if acc is None else acc + v
    return acc


def print_module_summary(module, inputs, max_nesting=3, skip_redundant=True, file=None):
    """Run `module(*inputs)` once with a forward hook on every sub-module and print one row per sub-module: parameter count,
    buffer count, output shape, output dtype (reference :199-266).  Returns the module's outputs.  The hooks only look at
    outputs that are tensors (the product modules pass bf16 2-D activations between sub-modules; holders whose forward is
    never called simply do not show up)."""
    assert isinstance(module, torch.nn.Module) and not isinstance(module, torch.jit.ScriptModule)
    rows_raw = []
    depth = [0]

    def pre(_m, _in):
        depth[0] += 1

    def post(m, _in, out):
        depth[0] -= 1
        if depth[0] <= max_nesting:
            outs = out if isinstance(out, (tuple, list)) else [out]
            rows_raw.append((m, [t for t in outs if isinstance(t, torch.Tensor)]))

    hooks = []
    for m in module.modules():
        hooks.append(m.register_forward_pre_hook(pre))
        hooks.append(m.register_forward_hook(post))
    try:
        outputs = module(*inputs)
    finally:
        for h in hooks:
            h.remove()

    names = {m: n for n, m in module.named_modules()}
    seen = set()
    table = [[type(module).__name__, "Parameters", "Buffers", "Output shape", "Datatype"], ["---"] * 5]
    tot_p = tot_b = 0
    for m, outs in rows_raw:
        own_p = [p for p in m.parameters() if id(p) not in seen]
        own_b = [b for b in m.buffers() if id(b) not in seen]
        seen.update(id(t) for t in own_p + own_b)
        if skip_redundant and not own_p and not own_b and not outs:
            continue
        np_, nb_ = sum(p.numel() for p in own_p), sum(b.numel() for b in own_b)
        tot_p += np_
        tot_b += nb_
        name = "<top-level>" if m is module else names.get(m, "?")
        shape = str(list(outs[0].shape)) if outs else "-"
        dtype = str(outs[0].dtype).split(".")[-1] if outs else "-"
        table.append([name, str(np_) if np_ else "-", str(nb_) if nb_ else "-", shape, dtype])
    table += [["---"] * 5, ["Total", str(tot_p), str(tot_b), "-", "-"]]
    widths = [max(len(r[c]) for r in table) for c in range(5)]
    print(file=file)
    for r in table:
        print("  ".join(cell + " " * (w - len(cell)) for cell, w in zip(r, widths)), file=file)
    print(file=file)
    return outputs

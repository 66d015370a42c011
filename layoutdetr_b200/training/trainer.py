"""One data-parallel training iteration of LayoutDETR (the body of the reference's hot loop,
training/training_loop.py:274-328): phases Gmain / Greg / Dmain / Dreg, gradient all-reduce, Adam, G_ema.

Differences from the reference are mechanical, not numerical:
  * parameters, gradients and Adam state live in flat buffers (flat.FlatParams): the all-reduce runs in place
    on the gradient buffer (no torch.cat / split) and nan_to_num + Adam + bf16 refresh is one kernel;
  * the EMA skips parameters that cannot change (the frozen text encoder) — lerp(p, p, beta) == p;
  * Greg / Dreg are no-ops at the reference defaults (pl_weight = r1_gamma = 0: zero_grad + an optimizer step
    over parameters without gradients), so they launch nothing;
  * dropout (G / D in `.train()` as in training_loop.py:133-134) is drawn in-kernel from a device-resident Philox state that is
    advanced once per iteration inside the captured graph (rng.py); modules in `.eval()` give the deterministic iteration;
  * independent parts of the iteration run on parallel CUDA streams (lanes.py): the frozen text-encoder calls, the
    branches of a forward pass that do not feed each other, and the real-sample discriminator pass (which does not
    depend on G) next to Gmain.  Same kernels, same operands; `LD_LANES=0` gives the single-stream schedule.
"""
import copy

import torch

from .. import engine as E
from .. import rng as RNG
from ..exchange import GradExchange
from ..flat import FlatParams
from ..lanes import LANES
from . import networks_detr as nd
from .loss import StyleGAN2Loss


class Trainer:
    def __init__(self, G, D, device, batch_size, num_gpus=1, lr=1e-5, betas=(0.0, 0.99), eps=1e-8,
                 G_reg_interval=4, D_reg_interval=16, ema_kimg=10.0, ema_rampup=0.05, loss_kwargs=None, process_group=None):
        self.G, self.D = G, D
        self.G_ema = copy.deepcopy(G).eval()
        self.device = device
        self.batch_size = batch_size              # total batch over all ranks
        self.num_gpus = num_gpus
        self.pg = process_group
        self.loss = StyleGAN2Loss(device=device, G=G, D=D, **(loss_kwargs or {}))
        self.flat = {"G": FlatParams(G, beta1=betas[0]), "D": FlatParams(D, beta1=betas[0])}
        self.flat_ema = FlatParams(self.G_ema, beta1=0.0, with_grad=False)
        self.opt = {}
        for name, interval in (("G", G_reg_interval), ("D", D_reg_interval)):
            r = interval / (interval + 1) if interval is not None else 1.0        # lazy-regularisation rescale, :191-195
            self.opt[name] = dict(lr=lr * r, beta1=betas[0] ** r, beta2=betas[1] ** r, eps=eps)
        self.ema_kimg, self.ema_rampup = ema_kimg, ema_rampup
        self.cur_nimg = 0
        for m in (G, D, self.G_ema):
            m.requires_grad_(False)
        # lane scheduler state: the first iteration of a Trainer runs single-stream (fills every lazily built host cache)
        self._warmed = False
        self._iter_open = False
        self._lanes_on = False
        self._real_lane = None
        self._text_lane = None
        self._real_issued = False
        self._suspending = False
        self._param_ids = None
        self.segmented = False               # True: _phase_grads("G") ends a CUDA-graph segment (multi-GPU, NCCL between graphs)
        # data-parallel gradient exchange (exchange.py): per network, buckets of the flat gradient buffer all-reduced on a
        # communication stream while the backward pass is still running.  `_expected` (per-bucket write counts of one backward
        # pass) is only set by GraphedStep around a capture, from the counts traced in its warm-up iterations of the SAME
        # batch key; eager iterations always trace (one all-reduce after the backward pass, the reference's order).
        self.exch = {n: GradExchange(f.g, f.params, f.offsets, group=process_group, world=num_gpus, name=n) for n, f in self.flat.items()}
        self._expected = {}
        self._traced = {}

    def _phase(self, name, batch, gen_z):
        self._phase_grads(name, batch, gen_z)
        self._phase_reduce(name)
        self._phase_step(name)

    def _phase_reduce(self, name):
        if self.num_gpus > 1:
            self._traced[name] = self.exch[name].finish()

    def _phase_step(self, name):
        o = self.opt[name]
        self.flat[name].adam_step(o["lr"], o["beta1"], o["beta2"], o["eps"], grad_scale=1.0 / self.num_gpus)

    def _accumulate(self, phase, batch, gen_z, before_backward=None):
        self.loss.accumulate_gradients(phase=phase, bbox_real=batch["bbox_real"], bbox_class=batch["bbox_class"],
                                       bbox_text=batch["bbox_text"], bbox_patch=batch["bbox_patch"],
                                       padding_mask=batch["padding_mask"], background=batch["background"], real_c=batch["c"],
                                       gen_z=gen_z, gen_c=batch["c"], gain=1, cur_nimg=self.cur_nimg,
                                       before_backward=before_backward)

    def _phase_grads(self, name, batch, gen_z):
        """Forward + backward of phase Gmain / Dmain into the flat gradient buffer (no reduce, no step)."""
        if name == "G" and not self._iter_open:
            self.begin_iteration(batch)
        mod = self.G if name == "G" else self.D
        split = name == "D" and self._real_issued                   # the real-sample half already ran on the R lane
        if not split:
            self.flat[name].zero_grad()
            self.exch[name].begin(self._expected.get(name))
        mod.requires_grad_(True)
        mod.text_encoder.requires_grad_(False)
        if split:
            # both halves of Dmain add into the same gradient buffers: the fake-sample backward starts after the R lane
            self._accumulate("Dgen", batch, gen_z, before_backward=self._join_real_lane)
        else:
            self._accumulate(name + "main", batch, gen_z)
        if self._lanes_on:
            LANES.join_children()            # backward kernels ran on the lanes of their forward
        mod.requires_grad_(False)
        if name == "D":
            self.end_iteration()
        elif self.segmented:
            self.join_lanes()                # a CUDA-graph segment ends here: every lane has to be back

    # ---------------------------------------------------------------------------------------------- lane scheduling
    def begin_iteration(self, batch):
        """Start of an iteration: per-iteration caches, and with lanes on: refresh everything derived from the weights on
        this stream, issue the five text-encoder calls on the T lane and the real-sample discriminator pass on the R lane."""
        nd.new_iteration()
        if self.G.training or self.D.training:
            RNG.advance(self.device)         # fresh dropout masks every iteration, also under CUDA-graph replay (sites are baked in)
        self._iter_open = True
        self._real_issued = False
        self._real_lane = self._text_lane = None
        self._lanes_on = LANES.active(1) and self._warmed
        if not self._warmed:                 # first iteration of this Trainer: single stream, fills every lazy host cache
            LANES._suspended += 1
            self._suspending = True
        if not self._lanes_on:
            return
        G, D = self.G, self.D
        if self._param_ids is None:
            self._param_ids = set(id(t) for m in (G, D) for t in list(m.parameters()) + list(m.buffers()))
        E.refresh_stale(self._param_ids)
        for m in (G, D):
            m._front()(batch["bbox_text"], self.device)
        nd.valid_index(batch["padding_mask"])
        real_lane = LANES.active(3)
        # T lane: one call per forward pass, in the order the passes are issued on the host (per-module FIFO)
        calls = [G, D, D, G, D] if real_lane else [G, D, G, D, D]
        if real_lane and LANES.real_first:           # feed the R lane (the longest chain) before Gmain
            calls = [D, G, D, G, D]
        self._text_lane = nd.prefetch_text(calls, batch["bbox_text"], self.device)
        if real_lane:
            sR = LANES.fork("R", detached=True)
            with torch.cuda.stream(sR):
                D.requires_grad_(True)
                D.text_encoder.requires_grad_(False)
                self.flat["D"].zero_grad()
                self.exch["D"].begin(self._expected.get("D"))
                self._accumulate("Dreal", batch, None)
                D.requires_grad_(False)
                LANES.join_children()
            self._real_lane = sR
            self._real_issued = True

    def _join_real_lane(self):
        if self._real_lane is not None:
            LANES.join(self._real_lane)
            self._real_lane = None

    def join_lanes(self):
        """Bring every lane back into the current stream (end of the iteration / of a CUDA-graph segment)."""
        if not self._lanes_on:
            return
        cur = torch.cuda.current_stream()
        if self._text_lane is not None:
            if self._text_lane.cuda_stream != cur.cuda_stream:
                cur.wait_stream(self._text_lane)
            self._text_lane = None
        self._join_real_lane()
        for m in (self.G, self.D):                   # results still queued are complete now: no event needed any more
            for ent in m.__dict__.get("_te_queue", ()):
                ent[1] = None
        LANES.join_children()
        LANES.forget_children()      # a later CUDA-graph segment must only wait for lanes it forked itself

    def end_iteration(self):
        if self._lanes_on:
            self.join_lanes()
            LANES.forget_children()
            left = nd.pending_text((self.G, self.D))
            if left:
                for m in (self.G, self.D):
                    m.__dict__["_te_queue"].clear()
                raise RuntimeError("lane scheduler: %d prefetched text-encoder results were not consumed" % left)
        if self._suspending:
            LANES._suspended -= 1
            self._suspending = False
        self._iter_open = False
        self._real_issued = False
        self._warmed = True

    def iteration(self, batch, z_g, z_d, update_ema=True):
        self.begin_iteration(batch)
        self._phase("G", batch, z_g)
        self._phase("D", batch, z_d)
        if update_ema:
            self.update_ema()
        return self.loss.last

    def iteration_multi(self, batches, zs_g, zs_d, update_ema=True):
        """One optimizer step per network over SEVERAL micro-batches (batch_gpu < batch_size // num_gpus: the gradient
        accumulation of the reference loop, training_loop.py:286-301).  Eager and single-stream; the per-micro-batch text
        front end makes a captured graph pointless here."""
        if self.G.training or self.D.training:
            RNG.advance(self.device)
        with LANES.suspended():
            for name, zs in (("G", zs_g), ("D", zs_d)):
                mod = self.G if name == "G" else self.D
                self.flat[name].zero_grad()
                self.exch[name].begin(None)
                mod.requires_grad_(True)
                mod.text_encoder.requires_grad_(False)
                for batch, z in zip(batches, zs):
                    nd.new_iteration()
                    self._accumulate(name + "main", batch, z)
                mod.requires_grad_(False)
                self._phase_reduce(name)
                self._phase_step(name)
        self._warmed = True
        if update_ema:
            self.update_ema()
        return self.loss.last

    def update_ema(self):
        ema_nimg = self.ema_kimg * 1000
        if self.ema_rampup is not None:
            ema_nimg = min(ema_nimg, self.cur_nimg * self.ema_rampup)
        ema_beta = 0.5 ** (self.batch_size / max(ema_nimg, 1e-8))
        self.flat_ema.ema_from(self.flat["G"], ema_beta)
        self.cur_nimg += self.batch_size


class GraphedStep:
    """The whole training iteration (Gmain + Dmain + Adam + EMA, ~6.5k kernel launches) captured once into a CUDA graph
    and replayed: the step is otherwise CPU-launch-bound (Python + autograd dispatch ~25 us per kernel vs ~22 us of GPU
    work per kernel on average).  Shapes and host-derived facts (which slots are valid, token padding width) are baked into
    a graph, so graphs are keyed on them; everything data-dependent is read from static device buffers that `run` refreshes
    before each replay: images, boxes, classes, z, token ids / masks / lengths and the LM-loss normalisers."""

    def __init__(self, trainer, max_graphs=None, capture_after=None):
        import collections
        import os
        self.tr = trainer
        self.graphs = collections.OrderedDict()    # key -> entry, least recently used first
        self.stream = LANES.main_stream()          # warm-up + capture stream = the main lane
        # A graph bakes in which slots are valid (and, with text_trim, the token width).  Synthetic / bucketed data has a
        # handful of such keys; real data can have one per batch, and a capture costs 2-3 warm-up iterations plus a private
        # memory pool.  So: at most `max_graphs` graphs are kept (least recently used is dropped), and a key is only captured
        # once it has been seen `capture_after` times — until then (and for one-off keys) the iteration runs eagerly.
        self.max_graphs = int(os.environ.get("LD_MAX_GRAPHS", "8")) if max_graphs is None else max_graphs
        self.capture_after = int(os.environ.get("LD_GRAPH_CAPTURE_AFTER", "1")) if capture_after is None else capture_after
        self._seen = collections.Counter()
        self.eager_steps = 0

    def _key(self, host_batch):
        pm = host_batch["padding_mask"]
        G = self.tr.G
        widths = ()
        if G.text_trim or self.tr.D.text_trim:     # the trimmed token widths are baked into the captured GEMM shapes
            widths = tuple(nd.trimmed_widths(m, host_batch["bbox_text"], pm, self.tr.device) for m in (G, self.tr.D))
        return (tuple(host_batch["background"].shape), pm.numpy().tobytes(), bool(G.text_trim), bool(G.text_dedup), widths,
                bool(G.training), bool(self.tr.D.training),
                LANES.level, LANES.text_ctas, LANES.lm_ctas, LANES.high_priority, LANES.dry, LANES.real_first)

    def _refresh_host_derived(self, st, host_mask):
        """Tokenise (host) into the front-ends' persistent device buffers and refresh the LM-loss normalisers."""
        import numpy as np
        from . import networks_detr as nd
        tr = self.tr
        dev = tr.device
        vidx_cpu = np.flatnonzero(~host_mask.numpy().astype(bool).reshape(-1))
        for m in (tr.G, tr.D):
            text = m._front()(st["bbox_text"], dev)
            n = nd.lm_target_count(text, vidx_cpu, m.tokenizer.pad_token_id)
            buf = m.__dict__.get("_inv_n_valid")
            if buf is None:
                buf = torch.zeros(1, dtype=torch.float32, device=dev)
                m.__dict__["_inv_n_valid"] = buf
            buf.fill_(1.0 / max(1, n))

    def _snapshot(self):
        tr = self.tr
        snap = dict(cur_nimg=tr.cur_nimg, steps={n: f.step for n, f in tr.flat.items()})
        for n, f in list(tr.flat.items()) + [("ema", tr.flat_ema)]:
            snap[n] = (f.p.clone(), f.v.clone() if f.v is not None else None, f.m.clone() if f.m is not None else None)
        return snap

    def _restore(self, snap, st):
        """Undo the warm-up / capture side effects and bring every derived bf16 shadow back in line with the weights."""
        from .. import engine as E
        from .. import kernels as K
        tr = self.tr
        tr.cur_nimg = snap["cur_nimg"]
        for n, f in list(tr.flat.items()) + [("ema", tr.flat_ema)]:
            p, v, m = snap[n]
            f.p.copy_(p)
            if v is not None:
                f.v.copy_(v)
            if m is not None:
                f.m.copy_(m)
            f.p16.copy_(K.to_bf16(f.p))
            if n != "ema":
                f.step = snap["steps"][n]
            E.bump_generation(f.params)
        E.refresh_stale()                                        # shadows recomputed in place (stable pointers)
        with torch.no_grad(), LANES.suspended():
            tr.G(st["z_g"], st["bbox_class"], st["bbox_real"], st["bbox_text"], st["bbox_patch"], st["padding_mask"],
                 st["background"], st["c"], reconst=True)
            tr.D(st["bbox_real"], st["bbox_class"], st["bbox_text"], st["bbox_patch"], st["padding_mask"], st["background"],
                 st["c"], reconst=True)
        torch.cuda.synchronize()

    def _replay(self, ent):
        tr = self.tr
        for name in ("G", "D"):
            o = tr.opt[name]
            tr.flat[name].set_hyper(o["lr"], o["beta1"], o["beta2"])
        gr = ent["graphs"]
        if len(gr) == 1:
            gr[0].replay()
        else:
            gr[0].replay(); tr.exch["G"].reduce_all()
            gr[1].replay(); tr.exch["D"].reduce_all()
            gr[2].replay()
        tr.update_ema()                                          # EMA beta depends on the image counter: kept out of the graph

    def close(self):
        """Drop every captured graph (and its memory pool).  With N > 1 the graphs hold captured NCCL collectives: the communicator
        cannot be destroyed while they are alive (ncclCommDestroy waits for its graph-captured operations to be released), so
        call this before torch.distributed.destroy_process_group()."""
        import gc
        for ent in self.graphs.values():
            ent.clear()
        self.graphs.clear()
        gc.collect()
        torch.cuda.synchronize()

    def run_static(self):
        """Replay the captured iteration on whatever the static input buffers currently hold (inputs resident in HBM)."""
        ent = next(reversed(self.graphs.values()))               # most recently used graph
        self._replay(ent)
        return ent["out"]

    def run(self, host_batch, z_g, z_d):
        """host_batch: dict of (pinned) CPU tensors + `bbox_text`; z_g / z_d: device tensors.  Returns the loss-term dict
        (static device tensors, overwritten by the next call)."""
        tr = self.tr
        key = self._key(host_batch)
        ent = self.graphs.get(key)
        if ent is not None:
            self.graphs.move_to_end(key)
        else:
            self._seen[key] += 1
            if self._seen[key] < self.capture_after:             # not worth a capture yet: plain eager iteration
                self.eager_steps += 1
                batch = {k: (v.to(tr.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host_batch.items()}
                return tr.iteration(batch, z_g, z_d)
            while len(self.graphs) >= max(1, self.max_graphs):   # drop the least recently used graph (and its memory pool)
                _, old = self.graphs.popitem(last=False)
                old.clear()
        if ent is None:
            st = {k: (v.to(tr.device) if torch.is_tensor(v) else v) for k, v in host_batch.items()}
            st["z_g"], st["z_d"] = z_g.clone(), z_d.clone()
            self._refresh_host_derived(st, host_batch["padding_mask"])
            # Warm-up + capture must not train: snapshot weights / Adam state / counters and restore them afterwards.
            snap = self._snapshot()
            # warm up on the capture stream (allocator pools, lazy inits, shadow caches and the lane streams reach their
            # steady state; the first iteration of a Trainer runs single-stream), then capture on the same stream
            side = self.stream
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3 if not tr._warmed else 2):
                    tr.iteration(st, st["z_g"], st["z_d"])
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            from .. import _lib
            n0 = _lib.launch_count()
            overlap = tr.num_gpus > 1 and tr.exch["G"].overlap
            tr.segmented = tr.num_gpus > 1 and not overlap
            if not tr.segmented:
                # one graph for the whole iteration.  With N > 1 the bucketed NCCL all-reduces are captured too, on the
                # communication stream, launched from inside the backward pass (exchange.py); the write counts that tell a
                # bucket it is complete were traced by the warm-up iterations above
                if overlap:
                    tr._expected = {n: list(c) for n, c in tr._traced.items() if c is not None}
                segs = [lambda: tr.iteration(st, st["z_g"], st["z_d"], update_ema=False)]
            else:
                # LD_DP_OVERLAP=0: NCCL collectives stay outside the graphs: [Gmain grads] -AR- [Adam(G) + Dmain grads] -AR- [Adam(D)]
                segs = [lambda: tr._phase_grads("G", st, st["z_g"]),
                        lambda: (tr._phase_step("G"), tr._phase_grads("D", st, st["z_d"])),
                        lambda: tr._phase_step("D")]
            graphs, pool = [], None
            for x in tr.exch.values():
                x.stats = dict(early=0, late=0)
            try:
                for i, seg in enumerate(segs):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool, stream=side, capture_error_mode="thread_local" if tr.num_gpus > 1 else "global"):
                        seg()
                    pool = g.pool()
                    graphs.append(g)
                    if tr.segmented and i < len(segs) - 1:       # keep the eager state consistent between captures
                        torch.cuda.synchronize()
            finally:
                tr._expected = {}
            out = {ph: {k: v for k, v in terms.items()} for ph, terms in tr.loss.last.items()}
            tr.segmented = False
            # host-derived device tensors the captured kernels read (valid-slot indices): owned by the entry, so the module-level
            # caches in networks_detr may be recycled without pulling memory from under a graph
            keep = [nd.valid_index(st["padding_mask"])]
            ent = dict(graphs=graphs, static=st, out=out, launches=_lib.launch_count() - n0, keep=keep,
                       exchange={n: dict(buckets=len(x.bounds), early=x.stats["early"], late=x.stats["late"]) for n, x in tr.exch.items()})
            self.graphs[key] = ent
            self._restore(snap, st)
            self._replay(ent)                                    # the first real step on this batch
            return ent["out"]
        st = ent["static"]
        for k, v in host_batch.items():
            if torch.is_tensor(v):
                st[k].copy_(v, non_blocking=True)
        st["z_g"].copy_(z_g)
        st["z_d"].copy_(z_d)
        if host_batch["bbox_text"] != st["bbox_text"]:
            st["bbox_text"] = host_batch["bbox_text"]
            self._refresh_host_derived(st, host_batch["padding_mask"])
        self._replay(ent)
        return ent["out"]

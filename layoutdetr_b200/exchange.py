"""Bucketed gradient exchange of the data-parallel iteration, overlapped with the backward pass.

The reference concatenates every gradient AFTER the backward pass, all-reduces the temporary and splits it back
(training/training_loop.py:303-313).  Here the gradients already live in one flat fp32 buffer (flat.FlatParams), which is cut
into contiguous buckets; a bucket's SUM all-reduce is launched on a communication stream as soon as the LAST write of the
backward pass into it has been issued, so NCCL moves the early buckets over NVLink while the rest of the backward pass is
still running.  Only the final bucket's exchange is exposed.

How "last write" is known: gradient writes happen in two ways — (a) the hand-written backward kernels add straight into
`param.grad` (functional.py: every such site obtains the buffer through engine.grad_buffer, and `_Fn` reports the end of each
backward node through engine.grad_writes_done), (b) torch's own AccumulateGrad for the few parameters that reach plain torch
ops (post-accumulate-grad hooks).  A first iteration in TRACE mode counts the writes per bucket; armed with those counts the
next iterations launch a bucket when its count is reached.  The number of writes is a property of the code path (shapes,
flags), not of the data; a write that arrives after its bucket was launched raises, a bucket that never fills is exchanged
in finish().

Stream rules: every write records an event on the stream it was issued on (lanes.py runs backward nodes on several streams);
the communication stream waits for the newest event of every stream that wrote into the bucket.  finish() makes the current
stream wait for the communication stream.  All of it is legal inside a CUDA-graph capture (NCCL kernels are captured on the
communication stream), which is what lets the multi-GPU iteration be ONE graph like the single-GPU one.

CPU tensors (gloo, tests) take the same code without streams.
"""
import os

import torch

from . import engine as E


_COMM_STREAMS = {}


class GradExchange:
    def __init__(self, g, params, offsets, group=None, world=1, bucket_mb=None, overlap=None, name=""):
        self.g = g
        self.group = group
        self.world = world
        self.name = name
        self.cuda = g.is_cuda
        if bucket_mb is None:
            bucket_mb = float(os.environ.get("LD_DP_BUCKET_MB", "32"))
        if overlap is None:
            overlap = os.environ.get("LD_DP_OVERLAP", "1") != "0"
        self.overlap = overlap
        lim = max(1, int(bucket_mb * (1 << 20) / 4))
        # contiguous buckets over the flat order, closed at parameter boundaries
        self.bounds = []
        self.bucket_of = {}
        lo = 0
        for i, (p, o) in enumerate(zip(params, offsets)):
            end = offsets[i + 1] if i + 1 < len(offsets) else g.numel()
            self.bucket_of[id(p)] = len(self.bounds)
            if end - lo >= lim or i + 1 == len(params):
                self.bounds.append((lo, end))
                lo = end
        nb = len(self.bounds)
        self.expected = None            # writes per bucket of one backward pass (from a trace iteration)
        self.counts = [0] * nb
        self.launched = [False] * nb
        self.events = [dict() for _ in range(nb)]     # bucket -> {stream id: (stream, newest event)}
        self.comm = None
        self.open = False
        self.stats = dict(early=0, late=0)            # buckets launched from inside / after the backward pass
        self._hooks = []
        if world > 1:
            self._watch(params)

    # ------------------------------------------------------------------------------------------ registration
    def _watch(self, params):
        for p in params:
            E.watch_grad(p, self, self.bucket_of[id(p)])
            if hasattr(p, "register_post_accumulate_grad_hook") and p.is_leaf:
                b = self.bucket_of[id(p)]
                rg = p.requires_grad                  # the Trainer keeps parameters frozen between phases; hooks need the flag on
                p.requires_grad_(True)
                self._hooks.append(p.register_post_accumulate_grad_hook(lambda _p, b=b: self.written(b)))
                p.requires_grad_(rg)

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        E.unwatch_grads(self)

    # ------------------------------------------------------------------------------------------ one backward pass
    def begin(self, expected=None):
        """Start of an accumulation period (right after zero_grad).  `expected`: per-bucket write counts from an earlier
        trace, or None to trace this one (buckets are then exchanged in finish())."""
        if self.world <= 1:
            return
        if expected is not None and len(expected) != len(self.bounds):
            raise ValueError("gradient exchange %s: %d expected counts for %d buckets" % (self.name, len(expected), len(self.bounds)))
        self.expected = list(expected) if (expected is not None and self.overlap) else None
        nb = len(self.bounds)
        self.counts = [0] * nb
        self.launched = [False] * nb
        self.events = [dict() for _ in range(nb)]
        self.open = True
        if self.cuda and self.comm is None:
            # ONE communication stream per device for every exchange (G and D buckets interleave while the real-sample lane
            # runs next to Gmain): collectives of one communicator stay in one stream order on every rank
            key = self.g.device.index
            if key not in _COMM_STREAMS:
                _COMM_STREAMS[key] = torch.cuda.Stream(device=self.g.device, priority=-1)
            self.comm = _COMM_STREAMS[key]

    def written(self, b):
        """One write into bucket `b` has been issued on the current stream."""
        if not self.open:
            return
        if self.launched[b]:
            raise RuntimeError("gradient exchange %s: bucket %d was written after its all-reduce had been launched "
                               "(write counts changed since the trace; set LD_DP_OVERLAP=0)" % (self.name, b))
        self.counts[b] += 1
        if self.expected is None:
            return
        if self.cuda:
            s = torch.cuda.current_stream(self.g.device)
            ev = torch.cuda.Event()
            ev.record(s)
            self.events[b][s.cuda_stream] = ev
        if self.counts[b] == self.expected[b]:
            self._launch(b, early=True)

    def _launch(self, b, early):
        lo, hi = self.bounds[b]
        self.launched[b] = True
        self.stats["early" if early else "late"] += 1
        if not self.cuda:
            torch.distributed.all_reduce(self.g[lo:hi], group=self.group)
            return
        if early:
            for ev in self.events[b].values():
                self.comm.wait_event(ev)
        else:
            self.comm.wait_stream(torch.cuda.current_stream(self.g.device))
        with torch.cuda.stream(self.comm):
            torch.distributed.all_reduce(self.g[lo:hi], group=self.group)

    def finish(self):
        """End of the accumulation period: every lane has been joined into the current stream.  Exchanges what is left and
        orders the current stream after the communication stream.  Returns the per-bucket write counts of this period."""
        if self.world <= 1:
            return None
        if not self.open:
            raise RuntimeError("gradient exchange %s: finish() without begin()" % self.name)
        E.grad_writes_done()
        counts = list(self.counts)
        self.open = False
        left = [b for b in range(len(self.bounds)) if not self.launched[b]]
        if len(left) == len(self.bounds):
            # nothing went early (trace iteration / overlap off): one exchange of the whole buffer on the current stream, the
            # reference's order
            self.reduce_all()
            return counts
        for b in left:                  # fewer writes than traced (cannot corrupt anything): exchanged now
            self._launch(b, early=False)
        if self.cuda and self.comm is not None:
            torch.cuda.current_stream(self.g.device).wait_stream(self.comm)
        return counts

    def reduce_all(self):
        """Plain exchange of the whole buffer on the current stream (also the replay-time call of the segmented schedule)."""
        if self.world <= 1:
            return
        self.open = False
        self.launched = [True] * len(self.bounds)
        self.stats["late"] += len(self.bounds)
        torch.distributed.all_reduce(self.g, group=self.group)

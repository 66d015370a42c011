"""Stream lanes: run the independent parts of one training iteration concurrently on one GPU.

At 16 samples per GPU most of LayoutDETR's iteration is latency-bound: ResNet-50, the DETR encoder / decoder stacks, the
StyleGAN2 background decoder and the small heads are ~5 000 kernels of 5-25 us that each use a handful of the 148 SMs,
while the frozen BERT text encoder (5 calls, 65 % of the FLOPs) and the two LM text decoders are long tensor-bound
GEMM chains.  The reference runs all of it on one stream (training/training_loop.py:274-328, training/loss.py:75-218,
training/networks_detr.py:133-187, :279-361).  Here the data-flow graph of the iteration is put on parallel CUDA
streams ("lanes") and captured, lanes included, into the iteration's CUDA graph:

  level 1  T lane     the frozen text-encoder calls (their only input is the token ids) are issued up front, in the
                      order their consumers need them, with the persistent GEMM grid capped (ld_set_stream_cta_limit)
                      so that SMs stay free for the other lanes; consumers wait on an event per call;
  level 2  branches   inside a forward pass: the causal-LM text decoder (depends only on the ids), the unconditional
                      discriminator branch (depends only on boxes / classes) and the background decoder (depends only
                      on the token output) run next to the conditional DETR chain; autograd replays each backward
                      node on the stream of its forward, so the backward pass is branch-parallel too;
  level 3  R lane     the real-sample discriminator pass of Dmain (training/loss.py:161-218) does not depend on G at
                      all: it runs on its own lane next to Gmain; only its gradient accumulation is ordered before the
                      fake-sample backward of Dmain (both add into the same buffers).

Executed work and numerics are those of the single-stream schedule (same kernels, same operands); only floating-point
accumulation ORDER of the two Dmain backward passes changes at level 3.

Rules that keep this safe with the caching allocator and lazily-filled host caches:
  * a tensor produced in one lane and consumed in another is `record_stream`ed on the consumer (fork / join do it);
  * everything derived from parameters (bf16 shadows, re-laid-out conv weights) is refreshed on the main stream before
    any lane forks (engine.refresh_stale), token ids / index tensors are materialised there too;
  * every lane is joined back into the stream that forked it before that stream's results are used, and all lanes are
    joined at the end of the iteration (a CUDA-graph capture requires it).
"""
import contextlib
import ctypes
import os

import torch

from . import _lib


class Lanes:
    def __init__(self):
        self.level = int(os.environ.get("LD_LANES", "3"))
        # grid caps of the tensor-bound lanes (0 = one CTA per SM).  Measured on B200 (profiles/r1_lane_sweep.txt): capping
        # does not pay — the other lanes' kernels are short and get their SMs between the persistent GEMM's waves
        self.text_ctas = int(os.environ.get("LD_LANE_TEXT_CTAS", "0"))         # text-encoder lane
        self.lm_ctas = int(os.environ.get("LD_LANE_LM_CTAS", "0"))             # text-decoder branches
        self.high_priority = int(os.environ.get("LD_LANE_PRIORITY", "1"))      # latency-bound lanes outrank T / LM lanes
        # dry run: the lane schedule's host-side issue order on ONE stream.  Results of a real multi-stream run must agree
        # with it to fp32-accumulation noise — anything larger is a missing dependency between lanes (tests/test_lanes_gpu.py)
        self.dry = int(os.environ.get("LD_LANES_DRY", "0"))
        self.real_first = int(os.environ.get("LD_LANE_REAL_FIRST", "0"))       # T-lane order: real-sample D pass's call first
        self._streams = {}          # (device index, parent stream id, name) -> Stream
        self._children = {}         # parent stream id -> [child Stream]  (branches forked since the last join)
        self._suspended = 0

    # -------------------------------------------------------------------------------------------- configuration
    def active(self, level=1):
        return self.level >= level and self._suspended == 0 and torch.cuda.is_available()

    @contextlib.contextmanager
    def suspended(self):
        """Single-stream execution inside the block (first iteration of a model: fills every lazy host-side cache)."""
        self._suspended += 1
        try:
            yield
        finally:
            self._suspended -= 1

    def configure(self, level=None, text_ctas=None, lm_ctas=None, high_priority=None, dry=None, real_first=None):
        """Change the schedule (drops the lane streams: their grid caps / priorities are fixed at creation)."""
        if level is not None:
            self.level = int(level)
        if text_ctas is not None:
            self.text_ctas = int(text_ctas)
        if lm_ctas is not None:
            self.lm_ctas = int(lm_ctas)
        if high_priority is not None:
            self.high_priority = int(high_priority)
        if dry is not None:
            self.dry = int(dry)
        if real_first is not None:
            self.real_first = int(real_first)
        for s in self._streams.values():
            _lib.lib().ld_set_stream_cta_limit(ctypes.c_void_p(s.cuda_stream), 0)
        self._streams.clear()
        self._children.clear()

    def main_stream(self):
        """A stream for the iteration's main lane (warm-up and graph capture): high priority when priorities are on, so
        the latency-bound chains are scheduled ahead of the capped tensor-bound lanes."""
        if self.high_priority:
            return torch.cuda.Stream(priority=-1)
        return torch.cuda.Stream()

    # -------------------------------------------------------------------------------------------- streams
    def _stream(self, name, cta_limit=0, bulk=False):
        cur = torch.cuda.current_stream()
        if self.dry:
            return cur
        key = (cur.device.index, cur.cuda_stream, name)
        s = self._streams.get(key)
        if s is None:
            prio = 0 if (bulk or not self.high_priority) else -1
            s = torch.cuda.Stream(device=cur.device, priority=prio)
            self._streams[key] = s
            _lib.check(_lib.lib().ld_set_stream_cta_limit(ctypes.c_void_p(s.cuda_stream), int(cta_limit)), "ld_set_stream_cta_limit")
        return s

    def fork(self, name, *tensors, cta_limit=0, bulk=False, detached=False):
        """Stream of lane `name` under the current stream, ordered after everything enqueued on the current stream so far.
        `tensors` are the values the lane will read (allocator bookkeeping).  A `detached` lane is not waited for by
        join_children: its owner joins it explicitly (T and R lanes)."""
        cur = torch.cuda.current_stream()
        s = self._stream(name, cta_limit, bulk)
        if s is cur or s.cuda_stream == cur.cuda_stream:
            return s
        s.wait_stream(cur)
        for t in tensors:
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(s)
        if not detached:
            ch = self._children.setdefault(cur.cuda_stream, [])
            if s not in ch:
                ch.append(s)
        return s

    def join(self, s, *tensors):
        """The current stream waits for lane `s`; `tensors` are the lane's results the current stream will read.  The lane
        stays registered: the backward pass runs on it again and join_children waits for it then."""
        cur = torch.cuda.current_stream()
        if s.cuda_stream == cur.cuda_stream:
            return
        cur.wait_stream(s)
        for t in tensors:
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(cur)

    def join_children(self, stream=None):
        """Wait (recursively) for every lane forked from `stream` (default: current) that has not been joined yet —
        after a backward pass, whose kernels ran on the lanes of their forward."""
        cur = stream or torch.cuda.current_stream()
        for s in list(self._children.get(cur.cuda_stream, [])):
            self.join_children(s)
            cur.wait_stream(s)

    def forget_children(self, stream=None):
        cur = stream or torch.cuda.current_stream()
        for s in list(self._children.get(cur.cuda_stream, [])):
            self.forget_children(s)
        self._children.pop(cur.cuda_stream, None)



LANES = Lanes()

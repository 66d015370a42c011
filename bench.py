#!/usr/bin/env python
"""bench.py — LayoutDETR training-iteration throughput (layout samples/sec) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 16]

A "step" is one full training iteration of the reference's hot loop (training/training_loop.py:274-328):
Gmain (G fwd(reconst) + D fwd + backward + Adam) + Dmain (G fwd, D fwd x2, backward + Adam) + G_ema update, at
16 samples per GPU, 256x256 backgrounds, 8 valid of 9 slots, text padded to 256 tokens (BASELINE.json configs[1]).
Weak scaling: every rank runs its own 16 samples; gradients are all-reduced over NCCL each phase.

Prints ONE JSON line (see the task contract).  `--impl reference` times the CPU oracle port of the same
iteration on the host cores (the reference itself is PyTorch-on-CPU here; /root/reference does not exist on the
GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("LAYOUTDETR_SYNTHETIC_TOKENIZER", "1")       # no vocabulary / checkpoint files on the GPU box:
os.environ.setdefault("LAYOUTDETR_SYNTHETIC_WEIGHTS", "1")         # synthetic data AND synthetic weights, said so in "data"

GFLOP_PER_SAMPLE = 3186.3          # dense fwd+bwd FLOPs of one training iteration, per sample
METRIC = "layout samples/sec @ bs16 256^2 8-query (full G+D training iteration)"

G_KWARGS = dict(z_dim=4, num_bbox_labels=8, img_channels=3, img_height=1024, img_width=1024, c_dim=0,
                background_size=256, bert_f_dim=768, bert_num_heads=4, bert_num_encoder_layers=12,
                bert_num_decoder_layers=2, im_f_dim=512)
D_KWARGS = dict(num_bbox_labels=8, img_channels=3, img_height=1024, img_width=1024, c_dim=0,
                background_size=256, bert_f_dim=768, bert_num_heads=4, bert_num_encoder_layers=12,
                bert_num_decoder_layers=2, im_f_dim=512)


def bench_loop(args, dev, B, world, rank):
    """Throughput of the drop-in entry point `training.training_loop.training_loop` (reference signature, our body): synthetic
    LayoutDataset-shaped items through a DataLoader, host tokenisation, pinned H2D, one CUDA-graph replay per iteration, EMA,
    tick bookkeeping — everything `train.py` would execute per iteration."""
    import torch
    from layoutdetr_b200.training.training_loop import training_loop
    import tempfile
    common = ("num_bbox_labels", "img_channels", "img_height", "img_width", "c_dim", "background_size")
    net = lambda kw, cls: dict({k: v for k, v in kw.items() if k not in common}, class_name="layoutdetr_b200.training.networks_detr." + cls)
    ds = dict(class_name="layoutdetr_b200.training.synthetic_dataset.SyntheticLayoutDataset", num_items=4096, n_valid=8, seed=rank)
    marks = {}
    W, K = max(5, args.warmup), max(40, args.steps)        # wall-clock window of at least 40 iterations: the first iterations after the capture
    # carry loader start-up and pinned-pool growth (measured: 179 samples/s over 20 iterations, 184 over 60, bare replay 187-188)

    def cb(i):
        if i == W or i == W + K:
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
                torch.cuda.synchronize()
            marks[i] = time.perf_counter()

    with tempfile.TemporaryDirectory() as run_dir:
        training_loop(run_dir=run_dir, training_set_kwargs=ds, validation_set_kwargs=ds,
                      data_loader_kwargs=dict(num_workers=3, prefetch_factor=4, pin_memory=False),
                      G_kwargs=net(G_KWARGS, "Generator"), D_kwargs=net(D_KWARGS, "Discriminator"),
                      G_opt_kwargs=dict(class_name="torch.optim.Adam", lr=1e-5, betas=[0, 0.99], eps=1e-8),
                      D_opt_kwargs=dict(class_name="torch.optim.Adam", lr=1e-5, betas=[0, 0.99], eps=1e-8),
                      loss_kwargs={}, metrics=[], random_seed=0, num_gpus=world, rank=rank, batch_size=B * world, batch_gpu=B,
                      G_reg_interval=4, D_reg_interval=16, total_kimg=10 ** 6, kimg_per_tick=10 ** 6, image_snapshot_ticks=None,
                      network_snapshot_ticks=None, max_iterations=W + K, iteration_callback=cb)
    sec = (marks[W + K] - marks[W]) / K
    return dict(value=B * world / sec, unit="samples/s", ms_per_step=sec * 1e3, steps=K, warmup=W,
                entry="layoutdetr_b200.training.training_loop.training_loop (the reference's training/training_loop.py signature)",
                note="wall clock over K iterations incl. DataLoader, host tokenisation, pinned H2D, graph replay, EMA")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"], src="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, reasons=reasons, samples=len(self.rows))


# --------------------------------------------------------------------------------------------------------
def cpu_iteration_seconds(batch, threads):
    """One Gmain+Dmain fwd+bwd of the ORACLE (CPU port of the reference path) at `batch` samples."""
    import torch
    from layoutdetr_b200.synthetic import SyntheticTokenizer, make_inputs
    from layoutdetr_b200.training import networks_detr as nd
    from oracle import train_step
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    G = nd.Generator(**G_KWARGS)
    D = nd.Discriminator(**D_KWARGS)
    sdG = {k: v.detach() for k, v in G.state_dict().items()}
    sdD = {k: v.detach() for k, v in D.state_dict().items()}
    del G, D
    inp = make_inputs(batch, n_valid=8, seed=1)
    tok = SyntheticTokenizer()
    t0 = time.time()
    train_step.iteration(sdG, sdD, tok, inp)
    return time.time() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = max(1, args.cpu_sample)
    times = []
    for i in range(args.warmup_ref + args.steps_ref):
        t = cpu_iteration_seconds(B, threads)
        if i >= args.warmup_ref:
            times.append(t)
    sec = sum(times) / len(times)
    v = B / sec
    line = dict(metric=METRIC, value=v, unit="samples/s", n_gpus=args.gpus, steps=len(times), warmup=args.warmup_ref,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                impl="reference",
                config=dict(workload="bs16 256x256 synthetic, 8 of 9 slots, G+D fwd/bwd (configs[1]); CPU arm runs a bounded sample", sample_batch=B),
                cpu_baseline=dict(value=v, unit="samples/s", cores=threads, kind="port",
                                  sample="%d sample(s) per step: oracle Gmain+Dmain fwd+bwd, dense T=256, fp32" % B),
                e2e=dict(value=v, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
def run_eval(args):
    """`--workload eval` (BASELINE configs[4]): the evaluation sweep — G_ema forward at 64 layouts per
    batch, LayoutNet features of real and generated layouts, overlap / alignment / layout-wise IoU / DocSim, FID — through
    layoutdetr_b200.metrics.eval_sweep.run_sweep with host batches (H2D inside the timed region, D2H of the metric dict),
    next to the CPU oracle of the same sweep on a bounded sample.  Under torchrun every rank sweeps its own shard (weak scaling)
    and the statistics meet in the sweep's single all-reduce; rank 0 prints one JSON line."""
    import torch
    from layoutdetr_b200 import _lib
    from layoutdetr_b200.synthetic import SyntheticTokenizer, make_inputs
    from layoutdetr_b200.training import networks_detr as nd
    from layoutdetr_b200.training.networks_layoutnet import LayoutNet
    from layoutdetr_b200.metrics import eval_sweep
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)      # the sweep's one exchange: all-reduce of the running statistics
    _lib.lib()
    torch.manual_seed(0)
    G = nd.Generator(**G_KWARGS).to(dev).eval().requires_grad_(False)
    net = LayoutNet(13).to(dev).eval().requires_grad_(False)
    B, nb = args.eval_batch, max(16, args.steps)       # >= 16 batches: the sweep ends with ONE host-side FID (256 x 256 sqrtm), which a 4-batch window over-weights
    host = [make_inputs(B, n_valid=8, seed=50 + 1000 * rank + i) for i in range(nb)]      # every rank sweeps its own shard of the layouts
    for hb in host:
        for k, v in hb.items():
            if torch.is_tensor(v):
                hb[k] = v.pin_memory()
    to_dev = lambda hb: {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in hb.items()}

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(1, args.warmup)):
        eval_sweep.run_sweep(G, net, [to_dev(host[0])])                  # tokeniser cache, weight shadows
    sync_all()
    _lib.launch_count_reset()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    t0 = time.perf_counter()
    res = eval_sweep.run_sweep(G, net, (to_dev(hb) for hb in host))     # one sweep over `steps` batches per rank (+ the final all-reduce)
    sync_all()
    sec = time.perf_counter() - t0
    clocks = sampler.stop()
    if world > 1:
        tsec = torch.tensor([sec], device=dev)
        dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
        sec = float(tsec.item())
        dist.destroy_process_group()
    if rank != 0:
        return
    n = B * nb * world
    h2d = sum(v.numel() * v.element_size() for v in host[0].values() if torch.is_tensor(v))
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import layoutdetr_oracle as O
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sdG = {k: v.detach().cpu() for k, v in G.state_dict().items()}
        sdL = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        inp = make_inputs(args.cpu_sample, n_valid=8, seed=50)
        tok = SyntheticTokenizer()
        t1 = time.perf_counter()
        with torch.no_grad():
            fake = O.generator_forward(sdG, tok, inp["z"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"], inp["background"])
            mask = ~inp["padding_mask"]
            O.layoutnet_extract_features(sdL, inp["bbox_real"], inp["bbox_class"], inp["padding_mask"])
            O.layoutnet_extract_features(sdL, fake, inp["bbox_class"], inp["padding_mask"])
            O.compute_overlap(fake, mask), O.compute_alignment(fake, mask), O.layoutwise_iou_docsim(inp["bbox_real"], fake, mask)
        cpu_sec = time.perf_counter() - t1
        cpu = dict(value=args.cpu_sample / cpu_sec, unit="layouts/s", cores=threads, kind="port",
                   sample="%d layouts, oracle sweep (%.1f s)" % (args.cpu_sample, cpu_sec))
    gf = 425.2                                                           # G.forward(reconst=False) GFLOP per sample
    line = dict(metric="eval-sweep layouts/sec (G_ema fwd + LayoutNet features + overlap/alignment/IoU/DocSim + layout FID)",
                value=n / sec, unit="layouts/s", n_gpus=world, steps=nb, warmup=args.warmup, ms_per_step=sec / nb * 1e3, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload="eval sweep, %d layouts per batch, 256x256 backgrounds, 8 of 9 slots (BASELINE configs[4] shapes)" % B,
                            batch_per_gpu=B, l2="inputs larger than L2 per batch (text activations 680 MB)", timing="host clock around the public API call"),
                clocks=clocks, e2e=dict(value=n / sec, unit="layouts/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=8 * 6),
                gpu_launches=int(_lib.launch_count()),
                roofline=dict(bound="tensor", achieved=n / sec * gf / 1e3 / world, peak=peaks()["sustained"], unit="TFLOP/s (per GPU)",
                              frac=n / sec * gf / 1e3 / world / peaks()["sustained"], traffic=None, gflop_per_layout=gf,
                              note="whole sweep against the algorithmic forward FLOPs (dense reference shapes)"),
                cpu_baseline=cpu, result=res)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
def gemm_roofline(torch, K, pk):
    """Dominant kernel (gemm_bf16_2sm_kernel) on the dominant shape — the BERT FFN GEMM of one text-encoder call at bs16:
    M = 16*9*256 tokens, N = 3072, K = 768 — timed alone with CUDA events on the launching stream, AS THE STEP RUNS IT (bias + GELU
    epilogue, bf16 out: 60 of the 68 BERT-layer executions per iteration); the plain epilogue (no bias, no activation) is reported as
    a sub-field."""
    M, N, Kd = 16 * 9 * 256, 3072, 768
    a = torch.randn((M, Kd), device="cuda").to(torch.bfloat16)
    w = torch.randn((N, Kd), device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    ms = timed(lambda: K.linear(a, w, bias, act=K.ACT_GELU, out=out))
    ms_plain = timed(lambda: K.linear(a, w, out=out))
    flops = 2.0 * M * N * Kd
    ach = flops / (ms * 1e-3) / 1e12
    ach_plain = flops / (ms_plain * 1e-3) / 1e12
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture named in traffic_note
    # (algorithmic: 61.3 MB of operands + 226.5 MB of output; part of the output still sits in L2 when the kernel ends)
    two_sm = os.environ.get("LD_GEMM_2SM", "1") != "0"      # this shape has 1728 pair tiles: the cta_group::2 kernel unless disabled
    traffic, note = 235.7e6, "ncu --set full capture of the single-CTA kernel on this shape (profiles/r1_gemm_ncu_full_epilogue_variants.txt)"
    tp = os.path.join(ROOT, "profiles", "r2_gemm_2sm_ncu_traffic.json")
    if two_sm and os.path.exists(tp):
        with open(tp) as f:
            d = json.load(f)
        traffic, note = d["traffic_bytes"], d["note"]
    return dict(bound="tensor", achieved=ach, peak=pk["burst"], unit="TFLOP/s", frac=ach / pk["burst"], traffic=traffic,
                traffic_note=note, kernel="gemm_bf16_2sm_kernel (cta_group::2)" if two_sm else "gemm_bf16_kernel", shape=[M, N, Kd], ms=ms,
                epilogue="bias + GELU, bf16 out (as the step runs it)",
                plain_epilogue=dict(achieved=ach_plain, frac=ach_plain / pk["burst"], ms=ms_plain, note="same shape, no bias / activation"),
                peak_source=pk["src"] + " burst (kernel timed alone)")


def run_loop(args):
    """`--workload loop`: only the drop-in training_loop entry point (N ranks under torchrun)."""
    import torch
    import torch.distributed as dist
    from layoutdetr_b200 import _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    B = args.batch if args.scaling == "weak" else args.batch // world
    res = bench_loop(args, dev, B, world, rank)
    if rank == 0:
        print(json.dumps(dict(metric=METRIC, value=res["value"], unit="samples/s", n_gpus=world, steps=args.steps, warmup=res["warmup"],
                              ms_per_step=res["ms_per_step"], higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="bf16",
                              data="synthetic", config=dict(workload="training_loop entry point, bs%d per GPU, 256x256 synthetic, 8 of 9 slots" % B,
                                                            entry=res["entry"], note=res["note"]))), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from layoutdetr_b200 import _lib, kernels as K
    from layoutdetr_b200.synthetic import make_inputs
    from layoutdetr_b200.training import networks_detr as nd
    from layoutdetr_b200.training.trainer import Trainer, GraphedStep
    from layoutdetr_b200.lanes import LANES
    LANES.configure(level=args.lanes, text_ctas=args.text_ctas, lm_ctas=args.lm_ctas, high_priority=args.lane_priority)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")      # peer traffic over NVLink / NVSwitch only (north star: "NVLink only")
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()        # fail loudly if the CUDA library is missing
    pk = peaks()
    B = args.batch                                # weak scaling: 16 samples on every GPU
    if args.scaling == "strong":                  # the reference's `--batch=16 --gpus=N`: the 16 samples are split over the ranks
        assert args.batch % world == 0, "strong scaling needs batch % gpus == 0"
        B = args.batch // world

    torch.manual_seed(0)                          # identical initial weights on every rank (reference broadcasts from rank 0)
    G = nd.Generator(**G_KWARGS).to(dev)
    D = nd.Discriminator(**D_KWARGS).to(dev)
    for m in (G, D):
        m.text_trim = args.text_trim
        m.text_dedup = args.text_dedup
        # the reference trains G / D in .train() (training_loop.py:133-134): DETR + BERT dropout 0.1 live, also in the frozen encoder
        m.train(bool(args.dropout))
    trainer = Trainer(G, D, dev, batch_size=B * world, num_gpus=world)

    # ---- inputs: resident (value) and host-pinned rotating batches (e2e)
    nb_host = 4
    host_batches = [make_inputs(B, n_valid=8, seed=100 * rank + 1 + i) for i in range(nb_host)]
    for hb in host_batches:
        for k, v in hb.items():
            if torch.is_tensor(v):
                hb[k] = v.pin_memory()
    resident = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host_batches[0].items()}
    gz = torch.Generator(device=dev).manual_seed(1234 + rank)
    zs = [torch.randn((B, 9, 4), device=dev, generator=gz) for _ in range(2)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > L2 (126 MB)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    use_graph = bool(args.graph)
    gs = GraphedStep(trainer) if use_graph else None
    graphed = gs is not None
    if gs is not None:
        gs.run(host_batches[0], zs[0], zs[1])               # warm-up + capture of the whole iteration

    def step_resident():
        if gs is not None:
            gs.run_static()
        else:
            trainer.iteration(resident, zs[0], zs[1])

    def step_e2e(i):
        hb = host_batches[i % nb_host]
        z1 = torch.randn((B, 9, 4), device=dev, generator=gz)
        z2 = torch.randn((B, 9, 4), device=dev, generator=gz)
        if gs is not None:
            last = gs.run(hb, z1, z2)                       # H2D into the graph's static inputs, host tokenisation, replay
        else:
            batch = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in hb.items()}
            last = trainer.iteration(batch, z1, z2)
        vals = torch.stack([v.float().mean() for v in last["Gmain"].values()] + [v.float().mean() for v in last["Dmain"].values()])
        return vals.cpu()                                   # D2H read of the step's loss terms (synchronises)

    # ---- warm-up (also builds caches / cuBLAS-free: nothing to autotune)
    for i in range(args.warmup):
        step_resident()
    sync_all()

    if args.ncu:                                            # one profiled step: ncu --profile-from-start off
        log_path = os.environ.get("LD_GEMM_LOG")            # also dump (M, N, K, ...) of every GEMM of that step, in launch order
        if log_path:
            K.GEMM_LOG = []
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if log_path:
            with open(log_path, "w") as f:
                json.dump([[e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], e[8].get("act"), e[8].get("d_dtype"), bool(e[8].get("r")),
                            bool(e[8].get("conv"))] for e in K.GEMM_LOG], f)
        return

    # ---- timed: resident inputs
    sampler = ClockSampler(local)
    sampler.start()
    _lib.launch_count_reset()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()                                       # flush L2 between timed iterations
        ev[i][0].record()
        step_resident()
        ev[i][1].record()
    sync_all()
    clocks = sampler.stop()
    launches = _lib.launch_count()
    if gs is not None:                                      # replayed kernels are not counted by the library's host-side counter
        launches = next(iter(gs.graphs.values()))["launches"] * args.steps
    ms_dev = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    t = torch.tensor([ms_dev], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    value = B * world / (ms_step * 1e-3)

    # ---- timed: end to end through the public API with host buffers (H2D of the batch + D2H of the losses)
    h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values() if torch.is_tensor(v))
    for i in range(2):
        step_e2e(i)
    sync_all()
    t0 = time.perf_counter()
    d2h = 0
    for i in range(args.steps):
        out = step_e2e(i + 2)
        d2h = out.numel() * out.element_size()
    sync_all()
    e2e_s = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = B * world / float(te.item())

    # ---- data-parallel replicas must hold bit-identical weights after the run (reference check_ddp_consistency,
    # torch_utils/misc.py:183): a 64-bit wrap-around sum of the fp32 bit patterns of G / D / G_ema, gathered from every rank
    replicas = None
    if world > 1:
        sums = torch.stack([f.p.view(torch.int32).to(torch.int64).sum() for f in (trainer.flat["G"], trainer.flat["D"], trainer.flat_ema)])
        allsums = [torch.zeros_like(sums) for _ in range(world)]
        dist.all_gather(allsums, sums)
        replicas = dict(identical=bool(all(torch.equal(allsums[0], a) for a in allsums)), checksum_G_D_Gema=[int(v) for v in allsums[0].tolist()],
                        check="int64 sum of the fp32 bit patterns of every trainable parameter, all ranks")
    exchange = None
    if world > 1 and graphed:
        exchange = dict(next(reversed(gs.graphs.values())).get("exchange") or {},
                        bucket_mb=float(os.environ.get("LD_DP_BUCKET_MB", "32")), overlap=os.environ.get("LD_DP_OVERLAP", "1") != "0",
                        note="buckets of the flat fp32 gradient buffer all-reduced on a communication stream from inside the backward pass "
                             "(early) / after it (late); NCCL kernels captured into the iteration's CUDA graph")

    # ---- exact work-saving variant (reported separately, never as the headline): frozen text-encoder features computed
    # once per iteration per network instead of 2x / 3x, and all-padding token columns dropped (bit-identical results)
    variant = None
    if args.variants and gs is not None:
        for m in (G, D):
            m.text_trim, m.text_dedup = True, True
        gs.run(host_batches[0], zs[0], zs[1])
        vent = gs.graphs[gs._key(host_batches[0])]
        for _ in range(2):
            gs._replay(vent)
        sync_all()
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.zero_()
            ev2[i][0].record(); gs._replay(vent); ev2[i][1].record()
        sync_all()
        tv = torch.tensor([sum(a.elapsed_time(b) for a, b in ev2) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        variant = dict(name="text_dedup+text_trim (exact)", ms_per_step=float(tv.item()), value=B * world / (float(tv.item()) * 1e-3),
                       unit="samples/s", note="executes fewer FLOPs than the dense reference shapes; not the headline")
        for m in (G, D):
            m.text_trim, m.text_dedup = bool(args.text_trim), bool(args.text_dedup)

    # ---- the same iteration through the drop-in entry point (training_loop with the reference's signature)
    loop = None
    if args.loop_steps and world == 1:
        try:
            gs.graphs.clear()                                    # release the captured graphs' memory pool
            torch.cuda.empty_cache()
            import contextlib
            with contextlib.redirect_stdout(sys.stderr):         # the loop prints the reference's progress lines; stdout carries ONE JSON line
                loop = bench_loop(args, dev, B, world, rank)
        except Exception as e:                                   # never lose the headline line to the secondary measurement
            loop = dict(error="%s: %s" % (type(e).__name__, e))

    if rank == 0:
        roof = gemm_roofline(torch, K, pk)
        step_tflops = value * GFLOP_PER_SAMPLE * 1e9 / 1e12 / world
        roof["step"] = dict(achieved=step_tflops, peak=pk["sustained"], unit="TFLOP/s (algorithmic, dense reference shapes, per GPU)",
                            frac=step_tflops / pk["sustained"], gflop_per_sample=GFLOP_PER_SAMPLE)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            nb = max(1, args.cpu_sample)
            sec = cpu_iteration_seconds(nb, threads)
            cpu = dict(value=nb / sec, unit="samples/s", cores=threads, kind="port",
                       sample="%d samples in one batch: oracle Gmain+Dmain fwd+bwd, dense T=256, fp32 (%.1f s)" % (nb, sec))
        line = dict(metric=METRIC, value=value, unit="samples/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_step, higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="bf16", data="synthetic",
                    config=dict(workload="bs16 256x256 synthetic, 8 of 9 slots, G+D fwd/bwd + Adam + EMA (BASELINE configs[1])",
                                batch_per_gpu=B, global_batch=B * world, text_tokens=256, text_trim=bool(args.text_trim),
                                text_dedup=bool(args.text_dedup), l2="flushed between timed steps (256 MiB write)",
                                dropout=("on (train mode: DETR 0.1, BERT hidden / attention 0.1, in-kernel Philox)" if args.dropout
                                         else "off (modules in .eval(): deterministic kernels)"), parallelism="dp%d" % world,
                                launch="cuda graph replay of the captured iteration" if graphed else "eager (one launch per kernel)",
                                lanes=dict(level=LANES.level, text_ctas=LANES.text_ctas, lm_ctas=LANES.lm_ctas, priority=LANES.high_priority,
                                           note="independent sub-graphs of the iteration on parallel streams (same kernels, same operands)")),
                    clocks=clocks, e2e=dict(value=e2e_value, unit="samples/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=int(launches), roofline=roof, cpu_baseline=cpu, variants=[variant] if variant else [], loop=loop,
                    replicas=replicas, exchange=exchange)
        print(json.dumps(line), flush=True)
    if world > 1:
        if gs is not None:
            gs.close()              # captured NCCL collectives must be released before the communicator goes away
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="samples per GPU (weak scaling) / in total (strong scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch samples on every GPU (default); strong: --batch samples split over the GPUs (reference --batch=16 semantics)")
    ap.add_argument("--cpu-sample", type=int, default=4, help="samples in the bounded CPU-baseline step")
    ap.add_argument("--loop-steps", type=int, default=1, help="1: also time the drop-in training_loop entry point (reported under \"loop\")")
    ap.add_argument("--workload", default="train", choices=["train", "eval", "loop"],
                    help="train: the headline training iteration (default); eval: the evaluation sweep at --eval-batch layouts per batch")
    ap.add_argument("--eval-batch", type=int, default=64)
    ap.add_argument("--text-trim", type=int, default=0, help="1: drop all-padding token columns (exact)")
    ap.add_argument("--text-dedup", type=int, default=0, help="1: reuse frozen text-encoder features across the 5 calls (exact)")
    ap.add_argument("--dropout", type=int, default=1, help="1 (default): G / D in .train() as the reference's loop, dropout live; 0: .eval()")
    ap.add_argument("--graph", type=int, default=1, help="1: capture the iteration into a CUDA graph (single-GPU default)")
    ap.add_argument("--variants", type=int, default=1, help="1: also time the exact work-saving variant (reported separately)")
    ap.add_argument("--lanes", type=int, default=None, help="lane scheduler level 0..3 (layoutdetr_b200/lanes.py); default: LD_LANES or 3")
    ap.add_argument("--text-ctas", type=int, default=None, help="persistent-GEMM grid cap of the text-encoder lane")
    ap.add_argument("--lm-ctas", type=int, default=None, help="persistent-GEMM grid cap of the text-decoder branches")
    ap.add_argument("--lane-priority", type=int, default=None, help="1: latency-bound lanes get a higher stream priority")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu", action="store_true", help="profile exactly one resident step between cudaProfilerStart/Stop and exit")
    ap.add_argument("--steps-ref", type=int, default=1)
    ap.add_argument("--warmup-ref", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps_ref = max(1, min(args.steps, 2))
        run_reference(args)
    elif args.workload == "eval":
        run_eval(args)
    elif args.workload == "loop":
        run_loop(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""`training_loop(rank, **c)` — the reference's training driver (training/training_loop.py:63-467) with the same keyword
arguments, run-directory artefacts (stats.jsonl, network-snapshot-*.pkl with keys G / D / G_ema / augment_pipe /
training_set_kwargs, metric-*.jsonl through metric_main) and tick cadence, but with THIS package's iteration inside:

  reference hot loop (:274-328)                          here
  ------------------------------------------------------ ----------------------------------------------------------------------
  per phase: zero_grad(set_to_none) / requires_grad_     trainer.Trainer: flat parameter / gradient / Adam buffers (flat.py)
  accumulate_gradients over micro-batches                training.loss.StyleGAN2Loss on the sm_100a modules, lanes (lanes.py)
  cat(grads) -> all_reduce -> / gpus -> nan_to_num       NCCL all-reduce of the flat gradient buffer in place, 1 / gpus and
  -> split -> torch.optim.Adam.step                       nan_to_num inside the fused Adam kernel (csrc/optim.cu)
  Greg / Dreg with pl_weight = r1_gamma = 0              nothing to launch (zero_grad + step over parameters without gradients)
  G_ema lerp over all parameters + buffer copy           ld_ema_flat over the trainable range (frozen tensors: lerp(p, p) = p)
  5 eager forward / backward passes per iteration        one CUDA-graph replay per iteration (trainer.GraphedStep), keyed on the
                                                         batch's valid-slot pattern; eager fallback for one-off patterns

Host-side helpers that are outside the hot path — `dnnlib.util.construct_class_by_name`, `legacy.load_network_pkl`,
`metric_main`, the PNG grid export of `util.save_image*` — are used from the reference checkout when it is importable
(under `python -m layoutdetr_b200.dropin.run train.py` it is) and replaced by minimal local equivalents / skipped otherwise, so
the loop also runs stand-alone (bench.py --workload loop, tests).
"""
import copy
import importlib
import json
import os
import pickle
import time

import numpy as np
import torch

from ..torch_utils import misc
from ..torch_utils import training_stats
from . import networks_detr as nd
from .trainer import GraphedStep, Trainer

UP_DETR_CKPT = 'pretrained/up-detr-pre-training-60ep-imagenet.pth'          # reference :138-140


def _construct(*args, class_name=None, **kwargs):
    """dnnlib.util.construct_class_by_name: 'pkg.module.Class' -> Class(*args, **kwargs)."""
    try:
        import dnnlib
        return dnnlib.util.construct_class_by_name(*args, class_name=class_name, **kwargs)
    except ImportError:
        mod, _, cls = class_name.rpartition('.')
        return getattr(importlib.import_module(mod), cls)(*args, **kwargs)


def _optional(modname):
    try:
        return importlib.import_module(modname)
    except Exception:
        return None


def fetch_batch(samples, real_c):
    """A loader item -> the host-side batch dict of one rank (pinned CPU tensors + strings): only what the hot path reads.
    Accepts the reference loader's item (default collate of `LayoutDataset.__getitem__`: keys bboxes / labels / texts / mask /
    background, texts as N tuples of B strings, reference :254-263) and the lean collate's (dataset_layoutganpp.collate_lean).
    `patches` ([B, N, 3, 256, 256] fp32 in the reference's loader, ~2.4 MB per element) is shape-only in G / D
    (training/networks_detr.py:140,286), so a [B, N, 3, 1, 1] placeholder travels instead."""
    if 'bbox_real' in samples:                                    # collate_lean
        batch = dict(samples)
        if 'background' not in batch:
            raise KeyError("lean batches must be normalised first (dataset_layoutganpp.to_device) or carry 'background'")
        batch.pop('background_u8', None)
    else:
        bbox = torch.as_tensor(samples['bboxes']).to(torch.float32)
        B, N = bbox.shape[0], bbox.shape[1]
        texts = [list(t) for t in zip(*samples['texts'])]         # default collate: N tuples of B strings -> B lists of N
        batch = dict(bbox_real=bbox, bbox_class=torch.as_tensor(samples['labels']).to(torch.int64), bbox_text=texts,
                     bbox_patch=torch.zeros((B, N, 3, 1, 1)), padding_mask=~torch.as_tensor(samples['mask']).to(torch.bool),
                     background=torch.as_tensor(samples['background']).to(torch.float32), c=torch.as_tensor(real_c).to(torch.float32))
    if torch.cuda.is_available():
        batch = {k: (v.pin_memory() if (torch.is_tensor(v) and not v.is_cuda and not v.is_pinned()) else v) for k, v in batch.items()}
    return batch


def _split_batch(batch, batch_gpu):
    B = batch['bbox_real'].shape[0]
    if B <= batch_gpu:
        return [batch]
    out = []
    for s in range(0, B, batch_gpu):
        out.append({k: (v[s:s + batch_gpu]) for k, v in batch.items()})
    return out


def load_initial_weights(modules, rank=0):
    """Up-DETR initialisation of G / D / G_ema (reference :138-140, strict=False).  The reference loads the file
    unconditionally; here its absence is an error unless synthetic weights were requested (LAYOUTDETR_SYNTHETIC_WEIGHTS=1)."""
    from .med import synthetic_weights_allowed
    if not os.path.exists(UP_DETR_CKPT):
        if synthetic_weights_allowed():
            if rank == 0:
                print('Up-DETR checkpoint %s not found: keeping the initialisation (synthetic-weight run)' % UP_DETR_CKPT)
            return False
        raise FileNotFoundError('%s not found (the reference initialises the DETR backbone / transformer from it); '
                                'set LAYOUTDETR_SYNTHETIC_WEIGHTS=1 to train from the default initialisation' % UP_DETR_CKPT)
    sd = torch.load(UP_DETR_CKPT, map_location='cpu')['model']
    for m in modules:
        m.load_state_dict(sd, strict=False)
    return True


def training_loop(
    run_dir                 = '.',      # Output directory.
    training_set_kwargs     = {},       # Options for training set.
    validation_set_kwargs   = {},       # Options for validation set.
    data_loader_kwargs      = {},       # Options for torch.utils.data.DataLoader.
    G_kwargs                = {},       # Options for generator network.
    D_kwargs                = {},       # Options for discriminator network.
    G_opt_kwargs            = {},       # Options for generator optimizer.
    D_opt_kwargs            = {},       # Options for discriminator optimizer.
    augment_kwargs          = None,     # Options for augmentation pipeline. None = disable.
    loss_kwargs             = {},       # Options for loss function.
    metrics                 = [],       # Metrics to evaluate during training.
    random_seed             = 0,        # Global random seed.
    num_gpus                = 1,        # Number of GPUs participating in the training.
    rank                    = 0,        # Rank of the current process in [0, num_gpus[.
    batch_size              = 4,        # Total batch size for one training iteration.
    batch_gpu               = 4,        # Number of samples processed at a time by one GPU.
    ema_kimg                = 10,       # Half-life of the exponential moving average (EMA) of generator weights.
    ema_rampup              = 0.05,     # EMA ramp-up coefficient. None = no rampup.
    G_reg_interval          = None,     # How often to perform regularization for G? None = disable lazy regularization.
    D_reg_interval          = 16,       # How often to perform regularization for D? None = disable lazy regularization.
    augment_p               = 0,        # Initial value of augmentation probability.
    ada_target              = None,     # ADA target value. None = fixed p.
    ada_interval            = 4,        # How often to perform ADA adjustment?
    ada_kimg                = 500,      # ADA adjustment speed.
    total_kimg              = 25000,    # Total length of the training, measured in thousands of real images.
    kimg_per_tick           = 4,        # Progress snapshot interval.
    image_snapshot_ticks    = 50,       # How often to save image snapshots? None = disable.
    network_snapshot_ticks  = 50,       # How often to save network snapshots? None = disable.
    resume_pkl              = None,     # Network pickle to resume training from.
    resume_kimg             = 0,        # First kimg to report when resuming training.
    cudnn_benchmark         = True,     # (no cuDNN on this path; accepted for signature compatibility)
    abort_fn                = None,     # Callback function for determining whether to abort training.
    progress_fn             = None,     # Callback function for updating training progress. Called for all ranks.
    use_cuda_graph          = True,     # (extension) replay the captured iteration; False = eager Trainer.iteration
    max_iterations          = None,     # (extension) stop after this many iterations regardless of total_kimg (tests / bench)
    iteration_callback      = None,     # (extension) called as fn(batch_idx) after every iteration's work has been enqueued
):
    start_time = time.time()
    device = torch.device('cuda', rank)
    torch.cuda.set_device(device)
    np.random.seed(random_seed * num_gpus + rank)
    torch.manual_seed(random_seed * num_gpus + rank)
    from .. import rng as RNG
    RNG.manual_seed(random_seed * num_gpus + rank, device)
    if (augment_kwargs is not None) and (augment_p > 0 or ada_target is not None):
        raise NotImplementedError('the ADA augmentation pipe (training/augment.py) is outside the accelerated path; train.py leaves it '
                                  'off for layout training (--aug=noaug)')
    if batch_size % num_gpus != 0:
        raise ValueError('batch_size must be a multiple of num_gpus')

    # Load training set.
    if rank == 0:
        print('Loading training set...')
    training_set = _construct(**training_set_kwargs)
    sampler = misc.InfiniteSampler(dataset=training_set, rank=rank, num_replicas=num_gpus, seed=random_seed)
    loader = torch.utils.data.DataLoader(dataset=training_set, sampler=sampler, batch_size=batch_size // num_gpus, **data_loader_kwargs)
    training_set_iterator = iter(loader)
    validation_set = _construct(**validation_set_kwargs) if validation_set_kwargs else training_set
    if rank == 0:
        print()
        print('Num training images: ', len(training_set))
        print('Num validation images: ', len(validation_set))
        print()

    # Construct networks.
    if rank == 0:
        print('Constructing networks...')
    common_kwargs = dict(num_bbox_labels=training_set.num_bbox_labels, img_channels=training_set.num_channels,
                         img_height=training_set.height, img_width=training_set.width,
                         background_size=training_set.background_size_for_training, c_dim=training_set.label_dim)
    G = _construct(**G_kwargs, **common_kwargs).train().requires_grad_(False).to(device)
    D = _construct(**D_kwargs, **common_kwargs).train().requires_grad_(False).to(device)
    load_initial_weights([G, D], rank)
    G_ema_init = None
    if (resume_pkl is not None) and (rank == 0):
        print(f'Resuming from "{resume_pkl}"')
        legacy, dnnlib = _optional('legacy'), _optional('dnnlib')
        if legacy is not None and dnnlib is not None:
            with dnnlib.util.open_url(resume_pkl) as f:
                resume_data = legacy.load_network_pkl(f)
        else:
            with open(resume_pkl, 'rb') as f:
                resume_data = pickle.load(f)
        for name, module in [('G', G), ('D', D)]:
            misc.copy_params_and_buffers(resume_data[name], module, require_all=False)
        G_ema_init = resume_data.get('G_ema')

    # Distribute across GPUs: every replica starts from rank 0's weights (reference :177-181).
    if num_gpus > 1:
        for module in (G, D):
            for t in misc.params_and_buffers(module):
                torch.distributed.broadcast(t, src=0)

    # Network summary tables (rank 0), on one real batch — also the first, single-stream pass that fills the lazy host caches.
    first_samples, first_c = next(training_set_iterator)
    first = fetch_batch(first_samples, first_c)
    if rank == 0:
        mb = {k: (v[:batch_gpu].to(device) if torch.is_tensor(v) else v[:batch_gpu]) for k, v in first.items()}
        z0 = torch.zeros([mb['bbox_real'].shape[0], mb['bbox_real'].shape[1], G.z_dim], device=device)
        with torch.no_grad():
            outs = misc.print_module_summary(G, [z0, mb['bbox_class'], mb['bbox_real'], mb['bbox_text'], mb['bbox_patch'], mb['padding_mask'],
                                                 mb['background'], mb['c'], True])
            misc.print_module_summary(D, [outs[0], mb['bbox_class'], mb['bbox_text'], mb['bbox_patch'], mb['padding_mask'], mb['background'],
                                          mb['c'], True])

    # Setup training phases: one Trainer holds both optimizers (flat Adam) and G_ema.
    if rank == 0:
        print('Setting up training phases...')
    for kw, who in ((G_opt_kwargs, 'G'), (D_opt_kwargs, 'D')):
        if not str(kw.get('class_name', 'torch.optim.Adam')).endswith('Adam'):
            raise NotImplementedError('%s optimizer %r: the fused flat step implements torch.optim.Adam' % (who, kw.get('class_name')))
    if dict(G_opt_kwargs, class_name=None) != dict(D_opt_kwargs, class_name=None):
        raise NotImplementedError('G and D optimizer options differ; train.py gives both the same lr / betas / eps')
    opt = dict(G_opt_kwargs)
    trainer = Trainer(G, D, device, batch_size=batch_size, num_gpus=num_gpus, lr=opt.get('lr', 1e-5), betas=tuple(opt.get('betas', (0.0, 0.99))),
                      eps=opt.get('eps', 1e-8), G_reg_interval=G_reg_interval, D_reg_interval=D_reg_interval, ema_kimg=ema_kimg,
                      ema_rampup=ema_rampup, loss_kwargs=dict(loss_kwargs, **({} if 'class_name' not in loss_kwargs else {})))
    G_ema = trainer.G_ema
    if G_ema_init is not None:
        misc.copy_params_and_buffers(G_ema_init, G_ema, require_all=False)
    if num_gpus > 1:
        for t in misc.params_and_buffers(G_ema):
            torch.distributed.broadcast(t, src=0)
    per_rank = batch_size // num_gpus
    graphed = GraphedStep(trainer) if (use_cuda_graph and per_rank <= batch_gpu) else None
    reg_on = (trainer.loss.pl_weight != 0) or (trainer.loss.r1_gamma != 0)
    if reg_on:
        raise NotImplementedError('path-length / R1 regularisation (--pl-weight / --gamma != 0) needs second-order gradients through the '
                                  'fused transformer kernels; the reference default is 0 for both')
    n_phases = sum(1 if r is None else 2 for r in (G_reg_interval, D_reg_interval))        # the reference draws z for every phase

    # Export sample images: rendering lives in the reference's util.py (outside the path); used when importable.
    ref_util = _optional('util') if image_snapshot_ticks is not None else None
    grid = None
    if rank == 0 and ref_util is not None and hasattr(ref_util, 'save_image'):
        grid = {k: (v[:batch_gpu].to(device) if torch.is_tensor(v) else v[:batch_gpu]) for k, v in first.items()}
        grid['z'] = torch.randn([grid['bbox_real'].shape[0], grid['bbox_real'].shape[1], G.z_dim], device=device)

    # Initialize logs.
    if rank == 0:
        print('Initializing logs...')
    training_stats.enable(True)
    stats_collector = training_stats.Collector(regex='.*')
    stats_metrics = dict()
    stats_jsonl = open(os.path.join(run_dir, 'stats.jsonl'), 'wt') if rank == 0 else None
    metric_main = _optional('metrics.metric_main') if len(metrics) > 0 else None

    # Train.
    if rank == 0:
        print(f'Training for {total_kimg} kimg...')
        print()
    cur_nimg = resume_kimg * 1000
    trainer.cur_nimg = cur_nimg
    cur_tick = 0
    tick_start_nimg = cur_nimg
    tick_start_time = time.time()
    maintenance_time = tick_start_time - start_time
    batch_idx = 0
    iter_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    if progress_fn is not None:
        progress_fn(0, total_kimg)
    pending = first
    host_prof = [] if os.environ.get('LD_LOOP_PROFILE') else None       # host wall time per phase of an iteration (ms): fetch, pin, z, run
    while True:
        # Fetch training data (host side: pinned tensors + strings; the H2D copies go into the graph's static buffers).
        t_h0 = time.perf_counter()
        if pending is not None:
            batch, pending = pending, None
            t_h1 = time.perf_counter()
        else:
            samples, real_c = next(training_set_iterator)
            t_h1 = time.perf_counter()
            batch = fetch_batch(samples, real_c)
        t_h2 = time.perf_counter()
        N = batch['bbox_real'].shape[1]
        all_gen_z = torch.randn([n_phases * batch_size, N, G.z_dim], dtype=torch.float32, device=device)     # same draw as :267
        phase_z = [zz[:per_rank] for zz in all_gen_z.split(batch_size)]                                       # one slice per phase
        z_g, z_d = phase_z[0], phase_z[1 if G_reg_interval is None else 2]
        for _ in range(n_phases * batch_size):                    # keep numpy's stream aligned with the reference's gen_c draws (:269)
            np.random.randint(len(training_set))

        # Execute training phases: Gmain (+ Greg no-op) + Dmain (+ Dreg no-op) + both optimizer steps + G_ema.
        t_h3 = time.perf_counter()
        iter_ev[0].record(torch.cuda.current_stream(device))
        if graphed is not None:
            graphed.run(batch, z_g, z_d)
        else:
            micro = _split_batch(batch, batch_gpu)
            dev_micro = [{k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in mb.items()} for mb in micro]
            if len(micro) == 1:
                trainer.iteration(dev_micro[0], z_g, z_d)
            else:
                trainer.iteration_multi(dev_micro, list(z_g.split(batch_gpu)), list(z_d.split(batch_gpu)))
        iter_ev[1].record(torch.cuda.current_stream(device))
        if host_prof is not None:
            t_h4 = time.perf_counter()
            host_prof.append((t_h1 - t_h0, t_h2 - t_h1, t_h3 - t_h2, t_h4 - t_h3))

        # Update state.
        cur_nimg += batch_size
        batch_idx += 1
        if iteration_callback is not None:
            iteration_callback(batch_idx)

        # Perform maintenance tasks once per tick.
        done = (cur_nimg >= total_kimg * 1000) or (max_iterations is not None and batch_idx >= max_iterations)
        if (not done) and (cur_tick != 0) and (cur_nimg < tick_start_nimg + kimg_per_tick * 1000):
            continue

        # Print status line, accumulating the same information in training_stats.
        tick_end_time = time.time()
        fields = []
        fields += [f"tick {training_stats.report0('Progress/tick', cur_tick):<5d}"]
        fields += [f"kimg {training_stats.report0('Progress/kimg', cur_nimg / 1e3):<8.1f}"]
        fields += [f"time {_format_time(training_stats.report0('Timing/total_sec', tick_end_time - start_time)):<12s}"]
        fields += [f"sec/tick {training_stats.report0('Timing/sec_per_tick', tick_end_time - tick_start_time):<7.1f}"]
        fields += [f"sec/kimg {training_stats.report0('Timing/sec_per_kimg', (tick_end_time - tick_start_time) / max(cur_nimg - tick_start_nimg, 1) * 1e3):<7.2f}"]
        fields += [f"maintenance {training_stats.report0('Timing/maintenance_sec', maintenance_time):<6.1f}"]
        fields += [f"gpumem {training_stats.report0('Resources/peak_gpu_mem_gb', torch.cuda.max_memory_allocated(device) / 2**30):<6.2f}"]
        fields += [f"reserved {training_stats.report0('Resources/peak_gpu_mem_reserved_gb', torch.cuda.max_memory_reserved(device) / 2**30):<6.2f}"]
        torch.cuda.reset_peak_memory_stats()
        if rank == 0:
            print(' '.join(fields))

        # Check for abort.
        if (not done) and (abort_fn is not None) and abort_fn():
            done = True
            if rank == 0:
                print()
                print('Aborting...')

        # Save image snapshot (G_ema on the fixed grid; rendering by the reference's util.save_image).
        if (rank == 0) and (grid is not None) and (done or cur_tick % image_snapshot_ticks == 0):
            with torch.no_grad():
                fake = G_ema(grid['z'], grid['bbox_class'], grid['bbox_real'], grid['bbox_text'], grid['bbox_patch'], grid['padding_mask'],
                             grid['background'], grid['c'])
            try:
                ref_util.save_image(fake, grid['bbox_class'], ~grid['padding_mask'], training_set.colors,
                                    os.path.join(run_dir, f'train_layouts_fake_{cur_nimg//1000:06d}.png'),
                                    first_samples['W_page'][:batch_gpu], first_samples['H_page'][:batch_gpu])
            except Exception as e:                               # rendering is best effort, never part of the training result
                print('image snapshot skipped:', e)

        # Save network snapshot.
        snapshot_pkl = None
        snapshot_data = None
        if (network_snapshot_ticks is not None) and (done or cur_tick % network_snapshot_ticks == 0):
            snapshot_data = dict(G=G, D=D, G_ema=G_ema, augment_pipe=None, training_set_kwargs=dict(training_set_kwargs))
            for key, value in snapshot_data.items():
                if isinstance(value, torch.nn.Module):
                    value = copy.deepcopy(value).eval().requires_grad_(False)
                    if num_gpus > 1:
                        misc.check_ddp_consistency(value, ignore_regex=r'.*\.[^.]+_(avg|ema)')
                    snapshot_data[key] = value.cpu()
                del value
            snapshot_pkl = os.path.join(run_dir, f'network-snapshot-{cur_nimg//1000:06d}.pkl')
            if rank == 0:
                with open(snapshot_pkl, 'wb') as f:
                    pickle.dump(snapshot_data, f)

        # Evaluate metrics.
        if (snapshot_data is not None) and (len(metrics) > 0) and (metric_main is not None):
            if rank == 0:
                print('Evaluating metrics...')
            for metric in metrics:
                ds_kwargs = training_set_kwargs if '_train' in metric else validation_set_kwargs
                result_dict = metric_main.calc_metric(metric=metric, run_dir=run_dir, G=snapshot_data['G_ema'], dataset_kwargs=ds_kwargs,
                                                      num_gpus=num_gpus, rank=rank, device=device)
                if rank == 0:
                    metric_main.report_metric(result_dict, run_dir=run_dir, snapshot_pkl=snapshot_pkl)
                stats_metrics.update(result_dict.results)
        del snapshot_data

        # Collect statistics.
        iter_ev[1].synchronize()
        training_stats.report0('Timing/iteration_ms', iter_ev[0].elapsed_time(iter_ev[1]))
        for ph, terms in trainer.loss.last.items():
            for k, v in terms.items():
                training_stats.report0('Loss/%s/%s' % ('G' if ph.startswith('G') else 'D', k), v.float().mean())
        stats_collector.update()
        stats_dict = stats_collector.as_dict()
        timestamp = time.time()
        if stats_jsonl is not None:
            stats_jsonl.write(json.dumps(dict(stats_dict, timestamp=timestamp)) + '\n')
            stats_jsonl.flush()
        if progress_fn is not None:
            progress_fn(cur_nimg // 1000, total_kimg)

        # Update state.
        cur_tick += 1
        tick_start_nimg = cur_nimg
        tick_start_time = time.time()
        maintenance_time = tick_start_time - tick_end_time
        if done:
            break

    if stats_jsonl is not None:
        stats_jsonl.close()
    if host_prof and rank == 0:
        tail = host_prof[len(host_prof) // 2:]
        print('host ms per iteration (second half of the run): next(loader) %.2f | fetch_batch (pin) %.2f | z / bookkeeping %.2f | run (tokenise, H2D, replay launch) %.2f'
              % tuple(1e3 * sum(t[i] for t in tail) / len(tail) for i in range(4)))
    if graphed is not None and num_gpus > 1:
        graphed.close()              # graphs holding captured NCCL collectives must be gone before the process group is torn down
    if rank == 0:
        print()
        print('Exiting...')
    return dict(iterations=batch_idx, cur_nimg=cur_nimg, trainer=trainer, graphed=graphed)


def _format_time(seconds):
    s = int(np.rint(seconds))
    if s < 60:
        return '{0}s'.format(s)
    if s < 60 * 60:
        return '{0}m {1:02}s'.format(s // 60, s % 60)
    if s < 24 * 60 * 60:
        return '{0}h {1:02}m {2:02}s'.format(s // (60 * 60), (s // 60) % 60, s % 60)
    return '{0}d {1:02}h {2:02}m'.format(s // (24 * 60 * 60), (s // (60 * 60)) % 24, (s // 60) % 60)

"""Data loader: the LayoutDataset mirror against what the reference's own LayoutDataset returned for the
same zip (tests/golden/tiny_layout.zip -> dataset_ref.pt, tests/golden/gen_golden.py gen_dataset), in full and in lean mode."""
import os

import numpy as np
import pytest
import torch

from helpers import golden, GOLD

ZIP = os.path.join(GOLD, "tiny_layout.zip")


def _ds(**kw):
    from layoutdetr_b200.training.dataset_layoutganpp import LayoutDataset
    return LayoutDataset(path=ZIP, use_labels=False, max_size=None, xflip=False, background_size=32, **kw)


def test_full_mode_matches_reference_loader():
    g = golden("dataset_ref.pt")
    ds = _ds()
    assert len(ds) == g["len"] and ds.patch_shape == g["patch_shape"] and ds.num_bbox_labels == g["num_bbox_labels"]
    assert ds.label_dim == g["label_dim"] and ds.num_channels == 3 and (ds.height, ds.width) == tuple(g["patch_shape"][2:])
    assert ds.background_size_for_training == 32 and len(ds.colors) == ds.num_bbox_labels
    for i, ref in enumerate(g["items"]):
        s, lab = ds[i]
        assert np.array_equal(s["bboxes"], ref["bboxes"].numpy()) and s["bboxes"].dtype == np.float32
        assert np.array_equal(s["labels"], ref["labels"].numpy()) and s["labels"].dtype == np.int64
        assert s["texts"] == ref["texts"] and np.array_equal(s["mask"], ref["mask"].numpy())
        assert (s["name"], s["W_page"], s["H_page"]) == (ref["name"], ref["W_page"], ref["H_page"])
        assert np.array_equal(s["background"], ref["background"].numpy())                       # bit-identical (same PIL resize, same fp32 ops)
        assert list(s["patches"].shape) == [9, 3, 256, 256] and np.array_equal(s["patches"][:, :, ::16, ::16], ref["patches_sub"].numpy())
        assert abs(float(s["patches"].astype(np.float64).sum()) - ref["patches_sum"]) < 1e-6
        assert abs(float(np.abs(s["patches"]).astype(np.float64).sum()) - ref["patches_abs"]) < 1e-6
        assert list(s["patches_orig"].shape) == ref["patches_orig_shape"] and abs(float(s["patches_orig"].astype(np.float64).sum()) - ref["patches_orig_sum"]) < 1e-6
        assert list(s["patch_masks"].shape) == ref["patch_masks_shape"] and abs(float(s["patch_masks"].astype(np.float64).sum()) - ref["patch_masks_sum"]) < 1e-6
        assert abs(float(s["background_orig"].astype(np.float64).sum()) - ref["background_orig_sum"]) < 1e-6
        assert np.array_equal(lab, ref["label"].numpy())


def test_lean_mode_keeps_hot_path_keys_identical_and_skips_the_rest():
    from layoutdetr_b200.training import dataset_layoutganpp as dl
    g = golden("dataset_ref.pt")
    ds = _ds(lean=True)
    opened = []
    orig_open = ds._open_file
    ds._open_file = lambda name: (opened.append(name), orig_open(name))[1]
    items = [ds[i] for i in range(len(ds))]
    assert all(n.endswith("_background_orig.png") for n in opened), "lean mode must not open patch / mask PNGs"
    for (s, lab), ref in zip(items, g["items"]):
        assert np.array_equal(s["bboxes"], ref["bboxes"].numpy()) and np.array_equal(s["labels"], ref["labels"].numpy())
        assert s["texts"] == ref["texts"] and np.array_equal(s["mask"], ref["mask"].numpy())
        assert np.array_equal(s["background"], ref["background"].numpy())
        assert s["background_u8"].dtype == np.uint8 and s["background_u8"].shape == (32, 32, 3)
        assert np.array_equal(dl._normalise(s["background_u8"]).transpose(2, 0, 1), ref["background"].numpy())
        assert s["patches"].shape == (9, 3, 1, 1) and s["patches_orig"].ndim == 4 and s["patch_masks"].ndim == 4
    batch = dl.collate_lean(items)
    assert batch["bbox_real"].shape == (3, 9, 4) and batch["bbox_class"].dtype == torch.int64
    assert batch["padding_mask"].dtype == torch.bool and batch["padding_mask"][0].tolist() == [False] * 3 + [True] * 6
    assert batch["bbox_text"][1][8] == "text 8 of page 1" and batch["bbox_text"][2][1:] == [""] * 8
    assert batch["background_u8"].shape == (3, 32, 32, 3) and batch["bbox_patch"].shape == (3, 9, 3, 1, 1) and batch["c"].shape == (3, 0)


def test_dataset_works_under_a_dataloader_with_workers_and_pickles():
    import pickle
    from layoutdetr_b200.training import dataset_layoutganpp as dl
    ds = pickle.loads(pickle.dumps(_ds(lean=True)))
    loader = torch.utils.data.DataLoader(ds, batch_size=2, num_workers=2, collate_fn=dl.collate_lean, shuffle=False)
    shapes = [b["background_u8"].shape[0] for b in loader]
    assert shapes == [2, 1]


@pytest.mark.gpu
def test_device_normalisation_is_bit_identical_to_the_reference_numpy():
    from layoutdetr_b200 import kernels as K
    from layoutdetr_b200.training import dataset_layoutganpp as dl
    g = torch.Generator().manual_seed(0)
    for (B, H, W) in [(3, 32, 32), (2, 256, 256), (1, 6, 10)]:
        u8 = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
        ref = np.stack([dl._normalise(x).transpose(2, 0, 1) for x in u8.numpy()])
        out = K.normalize_u8_image(u8.cuda(), dl.RGB_MEAN, dl.RGB_STD).cpu().numpy()
        assert out.dtype == np.float32 and np.array_equal(out, ref)
    # every byte value, every channel
    allv = torch.arange(256, dtype=torch.uint8).repeat_interleave(3).reshape(1, 16, 16, 3).contiguous()
    ref = dl._normalise(allv[0].numpy()).transpose(2, 0, 1)[None]
    assert np.array_equal(K.normalize_u8_image(allv.cuda(), dl.RGB_MEAN, dl.RGB_STD).cpu().numpy(), ref)


@pytest.mark.gpu
def test_lean_batch_to_device_matches_reference_backgrounds():
    from layoutdetr_b200.training import dataset_layoutganpp as dl
    g = golden("dataset_ref.pt")
    ds = _ds(lean=True)
    batch = dl.to_device(dl.collate_lean([ds[i] for i in range(len(ds))]), torch.device("cuda"))
    ref = torch.stack([it["background"] for it in g["items"]])
    assert torch.equal(batch["background"].cpu(), ref)
    assert batch["bbox_real"].is_cuda and "background_u8" not in batch


def test_sampler_partition_matches_reference_streams():
    """Rank-strided index streams equal the reference InfiniteSampler's, draw for draw (tests/golden/sampler_ref.pt)."""
    import itertools
    from layoutdetr_b200.training.sampler import InfiniteSampler
    for c in golden("sampler_ref.pt"):
        ds = list(range(c["n"]))
        for r in range(c["world"]):
            s = InfiniteSampler(ds, rank=r, num_replicas=c["world"], shuffle=c["shuffle"], seed=c["seed"], window_size=c["window"])
            assert list(itertools.islice(iter(s), 120)) == c["streams"][r]
        if not c["shuffle"]:          # the ranks' streams interleave into ONE global stream: without shuffling, 0..n-1 cyclically
            merged = [c["streams"][p % c["world"]][p // c["world"]] for p in range(120)]
            assert merged == [p % c["n"] for p in range(120)]


def test_eval_sweep_entry_partitions_the_dataset_once_across_ranks(monkeypatch):
    """`sweep_entry.sweep` (what the overlaid metric modules call): every item goes to exactly one rank, batches carry the keys
    run_sweep reads, results are cached so that the FID metric and the overlap / alignment metric share one pass."""
    import types
    from layoutdetr_b200.metrics import sweep_entry, eval_sweep
    from layoutdetr_b200.training import dataset_layoutganpp as dl
    assert sweep_entry.rank_item_subset(7, 3, 0) == [0, 3, 6] and sweep_entry.rank_item_subset(7, 3, 2) == [2, 5]
    seen, calls = [], []

    def fake_run_sweep(G, net, batches, **kw):
        calls.append(kw)
        for b in batches:
            assert {"bbox_real", "bbox_class", "bbox_text", "bbox_patch", "padding_mask", "background", "c"} <= set(b)
            seen.extend(float(x) for x in b["bbox_real"][:, 0, 0])
        return dict(overlap=1.0, alignment=2.0, layoutwise_iou=3.0, layoutwise_docsim=4.0)

    monkeypatch.setattr(eval_sweep, "run_sweep", fake_run_sweep)
    monkeypatch.setattr(dl, "to_device", lambda b, device: dict({k: v for k, v in b.items() if k != "background_u8"}, background=b["background_u8"]))
    G = torch.nn.Linear(1, 1)
    G.z_dim = 4
    for rank in range(2):
        opts = types.SimpleNamespace(G=G, dataset_kwargs=dict(class_name="training.dataset_layoutganpp.LayoutDataset", path=ZIP, use_labels=False,
                                                               max_size=None, xflip=False, background_size=32),
                                     num_gpus=2, rank=rank, device=torch.device("cpu"), G_kwargs={})
        sweep_entry._cache.clear()
        out = sweep_entry.compute_overlap_alignment_laywise_IoU_layerwise_DocSim(opts, max_real=None, num_gen=50000)
        assert out == ((1.0, 2.0, 3.0, 4.0) if rank == 0 else tuple([float("nan")] * 4)) or rank == 1
        n_calls = len(calls)
        sweep_entry.sweep(opts)                                     # second metric of the same G / dataset: served from the cache
        assert len(calls) == n_calls
        with pytest.raises(FileNotFoundError):
            sweep_entry.compute_layout_fid(opts, None, 50000)       # no LayoutNet checkpoint on this box
    g = golden("dataset_ref.pt")
    assert sorted(seen) == sorted(float(it["bboxes"][0, 0]) for it in g["items"])      # 3 items, each seen exactly once over the 2 ranks

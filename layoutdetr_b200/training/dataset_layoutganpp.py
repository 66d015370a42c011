"""Layout dataset — mirror of the reference's training/dataset_layoutganpp.py (to_dense_batch :28-41, Dataset :45-229,
LayoutDataset :233-351): same zip layout (`non_image.json` + per-sample PNGs), same class / property names, same sample
dict, so `train.py`'s `class_name='training.dataset_layoutganpp.LayoutDataset'` resolves to it unchanged.

What is different: the training hot path reads only boxes, labels, texts, the mask and the
`background_size`² background — `bbox_patch` is consumed for its SHAPE only (training/networks_detr.py:141,287) and
`patches_orig` / `patch_masks` / `background_orig` never reach the GPU — yet the reference decodes, normalises and
collates nine 1024² patches + masks per sample (≈150 MB fp32) for every item.  With `lean=True` those PNGs are not
opened: the heavy keys become one-pixel placeholders with the right rank, and the background additionally travels as
uint8 HWC (`background_u8`, 1 byte per value instead of 4) for the device-side normalisation kernel
(`ld_normalize_u8_image`, bit-identical to the NumPy arithmetic of :333-336).  `lean=False` reproduces every key.
"""
import json
import os
import zipfile

import numpy as np
import PIL.Image
import torch

_LANCZOS = getattr(PIL.Image, "ANTIALIAS", None) or PIL.Image.LANCZOS      # Pillow >= 10 dropped the ANTIALIAS alias
RGB_MEAN = np.array([0.485, 0.456, 0.406]).astype(np.float32)
RGB_STD = np.array([0.229, 0.224, 0.225]).astype(np.float32)
MAX_SLOTS = 9


def to_dense_batch(data, is_str=False):
    """Pad the leading (element) axis to 9 slots; mask marks the real ones (reference :28-41)."""
    if not is_str:
        shape = list(data.shape)
        if shape[0] == MAX_SLOTS:
            out = np.array(data, dtype=data.dtype)
        else:
            out = np.zeros([MAX_SLOTS] + shape[1:], dtype=data.dtype)
            out[:shape[0]] = data
        n = shape[0]
    else:
        out = list(data) + [""] * (MAX_SLOTS - len(data))
        n = len(data)
    mask = np.array([1] * n + [0] * (MAX_SLOTS - n), dtype=bool)
    return out, mask


def _normalise(img_u8):
    """uint8 HWC -> fp32 HWC, ImageNet statistics — the exact NumPy expression of the reference (:282, :333)."""
    return (img_u8.astype(np.float32) / 255.0 - RGB_MEAN.reshape(1, 1, 3)) / RGB_STD.reshape(1, 1, 3)


class Dataset(torch.utils.data.Dataset):
    def __init__(self, name, raw_shape, num_bbox_labels, max_size=None, use_labels=False, background_size=1024, random_seed=0):
        self._name = name
        self._raw_shape = list(raw_shape)
        self._num_bbox_labels = num_bbox_labels
        self._colors = None
        self._use_labels = use_labels
        self.background_size = background_size
        self._raw_labels = None
        self._label_shape = None
        self._raw_idx = np.arange(self._raw_shape[0], dtype=np.int64)
        if (max_size is not None) and (self._raw_idx.size > max_size):
            np.random.RandomState(random_seed).shuffle(self._raw_idx)
            self._raw_idx = np.sort(self._raw_idx[:max_size])

    def _get_raw_labels(self):
        if self._raw_labels is None:
            self._raw_labels = self._load_raw_labels() if self._use_labels else None
            if self._raw_labels is None:
                self._raw_labels = np.zeros([self._raw_shape[0], 0], dtype=np.float32)
            assert isinstance(self._raw_labels, np.ndarray) and self._raw_labels.shape[0] == self._raw_shape[0]
            assert self._raw_labels.dtype in [np.float32, np.int64]
        return self._raw_labels

    def close(self):
        pass

    def _load_raw_data(self, raw_idx):
        raise NotImplementedError

    def _load_raw_labels(self):
        raise NotImplementedError

    def __getstate__(self):
        return dict(self.__dict__, _raw_labels=None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self._raw_idx.size

    def __getitem__(self, idx):
        return self._load_raw_data(self._raw_idx[idx]), self.get_label(idx)

    def get_label(self, idx):
        label = self._get_raw_labels()[self._raw_idx[idx]]
        if label.dtype == np.int64:
            onehot = np.zeros(self.label_shape, dtype=np.float32)
            onehot[label] = 1
            label = onehot
        return label.copy()

    def get_details(self, idx):
        raw_idx = int(self._raw_idx[idx])
        return dict(raw_idx=raw_idx, raw_label=self._get_raw_labels()[raw_idx].copy())

    name = property(lambda self: self._name)
    patch_shape = property(lambda self: list(self._raw_shape[1:]))
    num_assets = property(lambda self: self.patch_shape[0])
    num_channels = property(lambda self: self.patch_shape[1])
    height = property(lambda self: self.patch_shape[2])
    width = property(lambda self: self.patch_shape[3])
    background_size_for_training = property(lambda self: self.background_size)
    num_bbox_labels = property(lambda self: self._num_bbox_labels)

    @property
    def colors(self):
        """One RGB tuple per label (the reference takes seaborn's 'husl' palette, :178-183; an evenly spaced HSV wheel here
        — only the snapshot PNGs use it)."""
        if self._colors is None:
            import colorsys
            n = self._num_bbox_labels
            self._colors = [tuple(int(c * 255) for c in colorsys.hsv_to_rgb(i / max(n, 1), 0.65, 0.9)) for i in range(n)]
        return self._colors

    @property
    def label_shape(self):
        if self._label_shape is None:
            raw = self._get_raw_labels()
            self._label_shape = [int(np.max(raw)) + 1] if raw.dtype == np.int64 else raw.shape[1:]
        return list(self._label_shape)

    @property
    def label_dim(self):
        assert len(self.label_shape) == 1
        return self.label_shape[0]

    has_labels = property(lambda self: any(x != 0 for x in self.label_shape))
    has_onehot_labels = property(lambda self: self._get_raw_labels().dtype == np.int64)


class LayoutDataset(Dataset):
    def __init__(self, path, xflip=False, background_size=1024, lean=False, **super_kwargs):
        self._path = path
        self.background_size = background_size
        self.lean = bool(lean)
        self._zipfile = None
        if os.path.splitext(path)[1].lower() != ".zip":
            raise IOError("Path must point to a zip")
        self._type = "zip"
        self._all_fnames = set(self._get_zipfile().namelist())
        PIL.Image.init()
        assert "non_image.json" in self._all_fnames
        with self._open_file("non_image.json") as f:
            self._samples = json.load(f)["samples"]
        parts = self._path.split("/")
        name = parts[-3] if len(parts) >= 3 else os.path.basename(path)
        # raw_shape = [#samples, 9, 3, H, W] of patches_orig: read from the PNG header instead of decoding sample 0
        with self._open_file(self._samples[0][0] + "_0_patch_orig.png") as f:
            w, h = PIL.Image.open(f).size
        raw_shape = [len(self._samples), MAX_SLOTS, 3, h, w]
        num_bbox_labels = self._samples[0][1]["attr"]["num_bbox_labels"]
        super().__init__(name=name, raw_shape=raw_shape, num_bbox_labels=num_bbox_labels, background_size=background_size, **super_kwargs)

    def _get_zipfile(self):
        if self._zipfile is None:
            self._zipfile = zipfile.ZipFile(self._path)
        return self._zipfile

    def _open_file(self, fname):
        return self._get_zipfile().open(fname, "r")

    def close(self):
        try:
            if self._zipfile is not None:
                self._zipfile.close()
        finally:
            self._zipfile = None

    def __getstate__(self):
        return dict(super().__getstate__(), _zipfile=None)

    # ---- pieces of one sample -------------------------------------------------------------------------------------
    def _background_u8(self, base):
        with self._open_file(base + "_background_orig.png") as f:
            img = PIL.Image.open(f)
            small = np.array(img.resize((self.background_size, self.background_size), _LANCZOS))
            orig = None if self.lean else np.array(img)
        assert small.ndim == 3 and small.shape[2] == 3
        return small, orig

    def _patch_256(self, fname):
        """Aspect-preserving resize into a centred 256 x 256 canvas (reference :268-287)."""
        with self._open_file(fname) as f:
            img = PIL.Image.open(f)
            width, height = img.width, img.height
            if width > height:
                wn, hn = 256, int(float(height) / float(width) * 256.0) // 2 * 2
            else:
                hn, wn = 256, int(float(width) / float(height) * 256.0) // 2 * 2
            tmp = np.array(img.resize((wn, hn), _LANCZOS))
        assert tmp.ndim == 3 and tmp.shape[2] == 3
        patch = np.zeros((256, 256, 3)).astype(np.float32)
        patch[128 - hn // 2:128 + hn // 2, 128 - wn // 2:128 + wn // 2] = _normalise(tmp)
        return patch.transpose(2, 0, 1)

    def _load_raw_data(self, raw_idx):
        base, meta = self._samples[raw_idx][0], self._samples[raw_idx][1]
        bboxes = np.array(meta["bboxes"])
        n = bboxes.shape[0]
        bboxes_batch, mask = to_dense_batch(bboxes)
        labels_batch, _ = to_dense_batch(np.array(meta["labels"]))
        texts_batch, _ = to_dense_batch(meta["texts"], is_str=True)
        bg_u8, bg_orig_u8 = self._background_u8(base)
        out = dict(name=meta["attr"]["name"], W_page=meta["attr"]["width"], H_page=meta["attr"]["height"],
                   bboxes=bboxes_batch.astype(np.float32), labels=labels_batch.astype(np.int64), texts=texts_batch, mask=mask,
                   background=_normalise(bg_u8).transpose(2, 0, 1))
        if self.lean:
            out["background_u8"] = bg_u8                                         # HWC uint8, normalised on the device
            out["patches"] = np.zeros((MAX_SLOTS, 3, 1, 1), dtype=np.float32)    # shape-only consumer: N = patches.shape[1]
            out["patches_orig"] = np.zeros((MAX_SLOTS, 3, 1, 1), dtype=np.float32)
            out["patch_masks"] = np.zeros((MAX_SLOTS, 1, 1, 1), dtype=np.float32)
            out["background_orig"] = np.zeros((3, 1, 1), dtype=np.float32)
            return out
        patches = np.stack([self._patch_256(base + "_%d_patch.png" % i) for i in range(n)], axis=0)
        out["patches"], _ = to_dense_batch(patches)
        orig = []
        for i in range(n):
            with self._open_file(base + "_%d_patch_orig.png" % i) as f:
                p = np.array(PIL.Image.open(f))
            assert p.ndim == 3 and p.shape[2] == 3
            orig.append(_normalise(p).transpose(2, 0, 1))
        out["patches_orig"], _ = to_dense_batch(np.stack(orig, axis=0))
        masks = []
        for i in range(n):
            with self._open_file(base + "_%d_patch_mask.png" % i) as f:
                m = np.array(PIL.Image.open(f))[:, :, np.newaxis]
            masks.append((m.astype(np.float32) / 255.0).transpose(2, 0, 1))
        out["patch_masks"], _ = to_dense_batch(np.stack(masks, axis=0))
        out["background_orig"] = _normalise(bg_orig_u8).transpose(2, 0, 1)
        return out

    def _load_raw_labels(self):
        return None          # as the reference (:343-351): whether or not page labels exist, none are returned (use_labels=False)


# ---------------------------------------------------------------------------------------------------------------------
def collate_lean(items):
    """DataLoader collate for `lean=True` items: uint8 backgrounds stacked into ONE buffer, texts transposed into the
    list-of-lists the networks take (the reference transposes after the default collate, training_loop.py:259)."""
    samples = [it[0] for it in items]
    B = len(samples)
    bg = torch.from_numpy(np.stack([s["background_u8"] for s in samples]))
    out = dict(
        bbox_real=torch.from_numpy(np.stack([s["bboxes"] for s in samples])),
        bbox_class=torch.from_numpy(np.stack([s["labels"] for s in samples])),
        padding_mask=~torch.from_numpy(np.stack([s["mask"] for s in samples])),
        bbox_text=[list(s["texts"]) for s in samples],
        bbox_patch=torch.zeros((B, MAX_SLOTS, 3, 1, 1)),
        background_u8=bg,
        c=torch.from_numpy(np.stack([it[1] for it in items])),
    )
    return out          # pinning is the DataLoader's job (`pin_memory=True`: done in the parent process, never in a forked worker)


def to_device(batch, device):
    """H2D of a lean batch (non-blocking from pinned memory) + device-side normalisation of the uint8 backgrounds."""
    from .. import kernels as K
    out = {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}
    out["background"] = K.normalize_u8_image(out.pop("background_u8"), RGB_MEAN, RGB_STD)
    return out

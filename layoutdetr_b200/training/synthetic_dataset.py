"""A dataset with the interface the training loop reads from `LayoutDataset` (reference training/dataset_layoutganpp.py:
attributes num_bbox_labels / num_channels / height / width / background_size_for_training / label_dim / colors, items =
(dict with bboxes / labels / texts / mask / background / W_page / H_page ..., label)) that serves the seeded synthetic layouts
of SURVEY §8d instead of reading a zip — what bench.py --workload loop and the loop tests train on (no dataset files travel
to the GPU box).  `n_valid` may be an int or a (lo, hi) range drawn per item (ragged layouts as in real data)."""
import numpy as np
import torch

from ..synthetic import make_inputs


class SyntheticLayoutDataset(torch.utils.data.Dataset):
    def __init__(self, num_items=256, n_valid=8, n_slots=9, background_size=256, num_bbox_labels=8, seed=0, **_ignored):
        self.num_items = int(num_items)
        self.n_valid = n_valid
        self.n_slots = n_slots
        self.background_size_for_training = background_size
        self.num_bbox_labels = num_bbox_labels
        self.num_channels = 3
        self.height = self.width = 1024
        self.label_dim = 0
        self.seed = seed
        self.name = "synthetic"
        self.colors = [(31 * i % 255, 67 * i % 255, 101 * i % 255) for i in range(num_bbox_labels)]
        self.patch_shape = [n_slots, 3, 1, 1]
        self.label_shape = [0]

    def __len__(self):
        return self.num_items

    def get_label(self, idx):
        return np.zeros([0], dtype=np.float32)

    def __getitem__(self, idx):
        nv = self.n_valid
        if isinstance(nv, (tuple, list)):
            nv = int(np.random.RandomState(self.seed * 1000003 + idx).randint(nv[0], nv[1] + 1))
        d = make_inputs(1, n_valid=nv, n_slots=self.n_slots, background_size=self.background_size_for_training,
                        num_bbox_labels=self.num_bbox_labels, seed=self.seed * 1000003 + idx + 1)
        sample = dict(name="synthetic_%06d" % idx, W_page=1024, H_page=1024,
                      bboxes=d["bbox_real"][0].numpy().astype(np.float32), labels=d["bbox_class"][0].numpy().astype(np.int64),
                      texts=list(d["bbox_text"][0]), mask=(~d["padding_mask"][0]).numpy(),
                      background=d["background"][0].numpy().astype(np.float32),
                      patches=np.zeros((self.n_slots, 3, 1, 1), dtype=np.float32))
        return sample, self.get_label(idx)

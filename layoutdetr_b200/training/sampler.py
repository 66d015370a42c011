"""Data-parallel partition of the training stream: every rank walks the SAME endless, progressively
re-shuffled order of dataset indices and keeps the positions `p % num_replicas == rank` — the index stream of the
reference's `torch_utils/misc.py:InfiniteSampler` (:114-148), reproduced draw for draw (same RandomState consumption), so
a run on N GPUs sees exactly the samples the reference would give each rank.
"""
import itertools

import numpy as np
import torch


def global_index_stream(n_items, shuffle=True, seed=0, window_size=0.5):
    """The rank-independent stream: position p yields order[p % n]; after each position the visited slot is swapped with a
    slot up to `window` positions behind it (window = round(n * window_size); no swaps below 2)."""
    order = np.arange(n_items)
    rnd, window = None, 0
    if shuffle:
        rnd = np.random.RandomState(seed)
        rnd.shuffle(order)
        window = int(np.rint(order.size * window_size))
    for p in itertools.count():
        i = p % order.size
        yield int(order[i])
        if window >= 2:
            j = (i - rnd.randint(window)) % order.size
            order[i], order[j] = order[j], order[i]


class InfiniteSampler(torch.utils.data.Sampler):
    def __init__(self, dataset, rank=0, num_replicas=1, shuffle=True, seed=0, window_size=0.5):
        assert len(dataset) > 0 and num_replicas > 0 and 0 <= rank < num_replicas and 0 <= window_size <= 1
        self.dataset, self.rank, self.num_replicas = dataset, rank, num_replicas
        self.shuffle, self.seed, self.window_size = shuffle, seed, window_size

    def __iter__(self):
        stream = global_index_stream(len(self.dataset), self.shuffle, self.seed, self.window_size)
        return itertools.islice(stream, self.rank, None, self.num_replicas)

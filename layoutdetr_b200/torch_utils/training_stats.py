"""Minimal stand-in for the reference's torch_utils/training_stats.py (report / Collector API, :54-97, :232-254).
The reference launches three tiny reductions per `report` call (~25 calls per phase); here values are only
recorded when collection is enabled, and reduced lazily."""
import torch

_enabled = False
_store = {}


def enable(flag=True):
    global _enabled
    _enabled = flag


def init_multiprocessing(rank, sync_device):
    return None


def report(name, value):
    if _enabled:
        v = torch.as_tensor(value).detach().float().reshape(-1)
        _store.setdefault(name, []).append(v)
    return value


def report0(name, value):
    return report(name, value)


class Collector:
    def __init__(self, regex=".*", keep_previous=True):
        self._last = {}

    def names(self):
        return list(_store.keys())

    def update(self):
        self._last = {k: torch.cat(v) for k, v in _store.items() if len(v)}
        _store.clear()

    def mean(self, name):
        v = self._last.get(name)
        return float(v.mean()) if v is not None and v.numel() else float("nan")

    def as_dict(self):
        return {k: dict(num=int(v.numel()), mean=float(v.mean()), std=float(v.std()) if v.numel() > 1 else 0.0)
                for k, v in self._last.items()}

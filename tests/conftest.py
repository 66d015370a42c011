import os
import sys

import pytest

# no vocabulary / checkpoint files travel with the repo: every test runs on the synthetic tokenizer and synthetic weights
os.environ.setdefault("LAYOUTDETR_SYNTHETIC_TOKENIZER", "1")
os.environ.setdefault("LAYOUTDETR_SYNTHETIC_WEIGHTS", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

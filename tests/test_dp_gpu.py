"""GPU (>= 2 devices): the overlapped gradient exchange inside the captured iteration vs the plain schedule, and replica identity
after real steps — tools/dp_check.py under torchrun.  Skipped on a single-GPU box (the world-size-2 host logic runs under gloo in
tests/test_host_logic.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_overlapped_exchange_matches_plain_schedule_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads(lines[-1])
    print(out)
    assert out["ok"], out

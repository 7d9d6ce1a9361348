"""Multi-GPU (NCCL) parity: N-GPU sharded transcription == 1-GPU transcription.
Needs >= 2 visible GPUs (skipped on a single-GPU box); the CPU-side logic of the same functions is covered by the
world-size-2 gloo tests in test_parallel_gloo.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_sharded_tokens_equal_single_gpu_tokens():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "nccl_shard_check.py"), "--utterances", "70"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["ok"] and d["world"] == world and all(d["checks"].values())

"""GPU, needs >= 2 devices (skipped on a single-GPU box): BLASes built round-robin across ranks and exchanged over NCCL,
rays sharded, hit records gathered — the result must be bit-identical to the single-GPU one (tools/sharded_scene_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_scene_build_and_trace(ctx):
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29547", os.path.join(ROOT, "tools", "sharded_scene_check.py")],
                       capture_output=True, text=True, timeout=250, cwd=ROOT)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "all ranks ok: True" in r.stdout

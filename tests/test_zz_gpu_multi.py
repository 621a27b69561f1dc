"""GPU (-m gpu), needs two devices (skipped on a one-GPU box): multi-GPU pieces of the hot path."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_bcast_weights_two_ranks():
    """b2sr_bcast_weights: the one collective of the path, at start-up, over NCCL."""
    from upscale_video_b200 import engine as E
    if E.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_bcast.py")], cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ranks agree after the broadcast: True; rank > 0 differed before it: True") == 2, r.stdout[-2000:]

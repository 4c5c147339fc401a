"""The tensor-parallel exchange kernel needs two GPUs on one node: runs tools/check_peer_exchange.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_exchange_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on one node (run under `gpurun --gpus 2`)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_peer_exchange.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for kind in ("push-oneshot", "push-twophase", "push-nomc-oneshot", "push-nomc-twophase"):
        for sync in ("poll", "flags"):      # data-is-the-signal form (default) and the flag handshake
            assert f"{kind}-{sync} exchange ok" in r.stdout, (kind, sync)
    assert "push decoder ok" in r.stdout
    assert "peer exchange ok" in r.stdout
    assert "multicast exchange ok" in r.stdout or "multicast exchange unavailable" in r.stdout

"""The kernels of the N-GPU exchange path, run and checked on ONE GPU: logical ranks = separate buffer sets + streams in one
process (tests/check_exchange_one_gpu.py).  GEMM-epilogue push (reduce-scatter and one-shot) + flag-free finish kernel +
quantising finish kernel, 2 and 4 ranks, bit-exact against oracle.mixq_oracle.tp_exchange and against the library's own
separate activation prologue.  Own process: a protocol bug ends in the library's stall trap, which poisons a CUDA context."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_fused_exchange_logical_ranks_on_one_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "check_exchange_one_gpu.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "ALL OK" in r.stdout
    assert r.stdout.count("one-GPU exchange ok") == 8

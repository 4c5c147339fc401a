#!/usr/bin/env python
"""Stage the reference's OWN Python for the MixLinear path into baseline/_ref/ (git-ignored, but it travels to the GPU box
with the gpurun snapshot), so that tests/test_gpu_reference_shim.py can execute it UNMODIFIED on a B200 on top of this
library's `mixlib` module — the drop-in claim of INTEGRATION.md §1, demonstrated rather than described.

Run in the build container (needs /root/reference; __graft_entry__.build() calls it when the reference is present):
    python tools/stage_reference_py.py

Copied verbatim (never into the tracked tree): mixquant/modules/linear.py, mixquant/Cache.py,
mixquant/modules/fused/norm.py, mixquant/modules/fused/mlp.py.  Nothing else of the reference imports on this image
(transformers 5.5 dropped shard_checkpoint, accelerate is absent), and nothing else is on the path.
"""
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MIXQ_REFERENCE", "/root/reference")
DST = os.path.join(REPO, "baseline", "_ref")
FILES = ["mixquant/modules/linear.py", "mixquant/Cache.py", "mixquant/modules/fused/norm.py", "mixquant/modules/fused/mlp.py"]


def stage(verbose=True) -> bool:
    if not os.path.isdir(REF):
        if verbose:
            print(f"{REF} not present: nothing staged")
        return False
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            print("staged", rel)
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)

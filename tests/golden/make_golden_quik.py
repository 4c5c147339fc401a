#!/usr/bin/env python
"""Generate tests/golden/quik_*.npz by running the REFERENCE'S OWN MixedQLinear (mixquant/modules/qlinear.py) on CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_quik.py

Executed verbatim from /root/reference: MixedQLinear.__init__, from_linear (weight rounding, nibble packing, reduced_w — pure
torch), forward (the shared-input / fp-part / dequantize sequencing).  Substituted: the un-vendored `quik` extension
(qlinear.py:6-7) -> oracle/quik_oracle.py's restatement of the published QUIK arithmetic, `.cuda()` -> no-op.  So the fixtures
pin the reference's real weight arithmetic and call sequence around the oracle's kernel arithmetic (parity of the `quik`
kernels themselves is unpinned: their source is not in the tree).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)


def install():
    from oracle import quik_oracle as Q
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    n = lambda x: x.detach().cpu().numpy()
    quik = types.ModuleType("quik")
    quik.asymmetric = types.SimpleNamespace(
        quantize=lambda x, ii, fi, bits: tuple(t(v) for v in Q.asymmetric_quantize(n(x), n(ii), n(fi), bits)),
        dequantize=lambda acc, meta, ws, rw, fp, bits: t(Q.asymmetric_dequantize(n(acc), n(meta), n(ws), n(rw), n(fp), bits)))
    quik.symmetric = types.SimpleNamespace()
    quik.matmul = types.SimpleNamespace(int4Matmul=lambda q, w: t(Q.int_matmul(n(q), n(w), 4)),
                                        int8Matmul=lambda q, w: t(Q.int_matmul(n(q), n(w), 8)))
    sys.modules["quik"] = quik
    for name, path in (("mixquant", "mixquant"), ("mixquant.modules", "mixquant/modules")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, path)]
        sys.modules[name] = m
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.current_device = lambda: torch.device("cpu")
    torch.cuda.set_device = lambda d: None
    return importlib.import_module("mixquant.modules.qlinear")


def case(ql, bits, seed, M=12, K=512, N=96, n_fp=64, bias=False):
    g = torch.Generator().manual_seed(seed)
    W = (torch.randn(N, K, generator=g) * 0.05).half()
    fp_idx = torch.randperm(K, generator=g)[:n_fp].sort().values
    qmax = 7 if bits == 4 else 127
    mask = torch.ones(K, dtype=torch.bool)
    mask[fp_idx] = False
    ws = (W[:, mask].float().abs().amax(1, keepdim=True) / qmax).half()
    lin = torch.nn.Linear(K, N, bias=False)
    lin.weight.data = W.clone()
    _init = ql.MixedQLinear.__init__

    def cpu_init(self, *a, **k):          # the reference allocates its buffers on 'cuda' by default
        k["dev"] = "cpu"
        _init(self, *a, **k)
    ql.MixedQLinear.__init__ = cpu_init
    m = ql.MixedQLinear.from_linear(lin, W.clone(), ws, None, fp_idx, False, bits)
    out = {"W": W.numpy(), "weights_scales": ws.numpy(), "fp_indices": fp_idx.numpy(), "bits": np.array(bits),
           "int_weight": m.int_weight.numpy(), "reduced_w": m.reduced_w.numpy(), "fp_weight": m.fp_weight.numpy(),
           "int_indices": m.int_indices.numpy()}
    for t in range(2):
        x = torch.randn(M, K, generator=g)
        x[:, fp_idx] *= 15.0
        x = x.half()
        if t == 1:
            x = x.reshape(2, M // 2, K)       # a 3-D input once
        out[f"c{t}_x"] = x.numpy()
        out[f"c{t}_y"] = m(x).numpy()
    # shared input across two Linears (q/k/v style): the second consumer re-uses the first one's quantised activations
    sh = ql.SharedQuantizedInput(2)
    W2 = (torch.randn(N, K, generator=g) * 0.05).half()
    ws2 = (W2[:, mask].float().abs().amax(1, keepdim=True) / qmax).half()
    lin2 = torch.nn.Linear(K, N, bias=False)
    lin2.weight.data = W2.clone()
    a = ql.MixedQLinear.from_linear(lin, W.clone(), ws, sh, fp_idx, False, bits)
    b = ql.MixedQLinear.from_linear(lin2, W2.clone(), ws2, sh, fp_idx, False, bits)
    x = torch.randn(M, K, generator=g).half()
    out["W2"], out["weights_scales2"] = W2.numpy(), ws2.numpy()
    out["sh_x"], out["sh_ya"], out["sh_yb"] = x.numpy(), a(x).numpy(), b(x).numpy()
    assert sh.qint_x is None and sh.cur_group_elem == 0
    ql.MixedQLinear.__init__ = _init
    return out


def main():
    ql = install()
    for name, bits, seed in (("quik_w4", 4, 11), ("quik_w8", 8, 12)):
        d = case(ql, bits, seed)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **d)
        print(f"{name}: {len(d)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

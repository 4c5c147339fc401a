#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE'S OWN Python on CPU.

Run in the build container only (needs /root/reference; the GPU box never runs this):
    python tests/golden/make_golden.py

What is executed verbatim from /root/reference: mixquant/modules/linear.py (MixLinear_GEMM.from_linear,
forward, forward_without_preconditionFusedSilu, FindOutliers, pack_to_i4), mixquant/Cache.py (MixLibCache),
mixquant/modules/fused/norm.py (FasterTransformerRMSNorm) and mixquant/modules/fused/mlp.py (MixLlamaMLP).
What is substituted: the un-vendored CUDA extension `mixlib` -> oracle/mixlib_cpu.py (the per-kernel
arithmetic of the oracle), `EETQ` -> a stub, `.cuda()` / `.to('cuda')` -> no-ops, and
torch.cuda.get_device_capability -> (10, 0) (B200: the reference's `arch != 9` branch).

So the fixtures pin (a) the reference's real weight-quantisation arithmetic (from_linear is pure torch) and
(b) the reference's real control flow — discovery state machine, hstack order, in-place zeroing, cache
reuse between up_proj and gate_proj — around the oracle's kernel arithmetic.  tests/test_oracle_golden.py
replays them against oracle/mixq_oracle.py (an independent restatement of that control flow), the GPU
tests replay them against the CUDA path.
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


def install_reference():
    from oracle import mixlib_cpu
    sys.modules["mixlib"] = mixlib_cpu
    eetq = types.ModuleType("EETQ")
    for n in ("quant_weights", "preprocess_weights", "w8_a16_gemm"):
        setattr(eetq, n, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("EETQ stub")))
    sys.modules["EETQ"] = eetq
    # packages as bare namespaces so that mixquant/__init__.py (HF model zoo imports) is not executed
    for name, path in (("mixquant", "mixquant"), ("mixquant.modules", "mixquant/modules"),
                       ("mixquant.modules.fused", "mixquant/modules/fused")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, path)]
        sys.modules[name] = m
    # device no-ops
    torch.Tensor.cuda = lambda self, *a, **k: self
    _to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k.pop("device")
        return _to(self, *a, **k) if (a or k) else self
    torch.Tensor.to = to
    torch.cuda.get_device_capability = lambda *a, **k: (10, 0)
    linear = importlib.import_module("mixquant.modules.linear")
    cache = importlib.import_module("mixquant.Cache")
    norm = importlib.import_module("mixquant.modules.fused.norm")
    mlp = importlib.import_module("mixquant.modules.fused.mlp")
    return linear, cache, norm, mlp


def np16(t):
    return t.detach().cpu().contiguous().numpy().copy()


def make_x(g, M, K, cols, scale=20.0):
    x = torch.randn(M, K, generator=g)
    if len(cols):
        x[:, cols] *= scale
    return x.half()


def snap(prefix, out, lin, cache, y, x_after, M):
    out[f"{prefix}_y"] = np16(y)
    out[f"{prefix}_x_after"] = np16(x_after)
    out[f"{prefix}_ind"] = np16(lin.ind).astype(np.int32)
    out[f"{prefix}_x_scale"] = np16(cache.x_scale[:M])
    out[f"{prefix}_q_x"] = np16(cache.q_xcache)
    if lin.ind.shape[0]:
        out[f"{prefix}_weight_cache"] = np16(lin.weight_cache)
        out[f"{prefix}_act_outliers"] = np16(cache.activation_outliers)
    out[f"{prefix}_add_outliers"] = np.array(int(lin.add_outliers))


def case_unfused(linear_mod, cache_mod, bit, bias, seed, M=16, K=256, N=128, fp=32):
    g = torch.Generator().manual_seed(seed)
    W = (torch.randn(N, K, generator=g) * 0.05).half()
    b = (torch.randn(N, generator=g) * 0.1).half() if bias else None
    lin = torch.nn.Linear(K, N, bias=bias)
    lin.weight.data = W.clone()
    if bias:
        lin.bias.data = b.clone()
    cache = cache_mod.MixLibCache(inputdim=32, sigma=6, bit=bit)
    out = {"W": np16(W), "M": np.array(M), "K": np.array(K), "N": np.array(N), "bit": np.array(bit), "fp": np.array(fp)}
    if bias:
        out["bias"] = np16(b)
    perm = torch.randperm(K, generator=g)
    s0, s1, s2 = perm[:3].sort().values, perm[3:5].sort().values, perm[5:7].sort().values
    layer_scales = None
    if bit == 4:
        layer_scales = torch.rand(K, generator=g)
        layer_scales[s0] += 10.0   # the calibrated outlier channels are among the static top-`fp`
        out["layer_scales"] = np16(layer_scales)
    q = linear_mod.MixLinear_GEMM.from_linear(lin, bit, cache=cache, layer_scales=layer_scales, dev="cpu",
                                              fp_features_num=fp)
    out["q_weight"] = np16(q.q_weight)
    out["scale_col"] = np16(q.scale_col)
    if bit == 4:
        out["ind0"] = np16(q.ind).astype(np.int32)
        out["weight_cache0"] = np16(q.weight_cache)
    # call 0: outliers s0; call 1: s0 + s1 (s1 new); call 2/3: s0+s1+s2 (s2 arrives after discovery stopped)
    sets = [s0, torch.cat([s0, s1]), torch.cat([s0, s1, s2]), torch.cat([s0, s1, s2])]
    for t, cols in enumerate(sets):
        x = make_x(g, M, K, cols)
        if t == 3:
            x[5] = 0  # an all-zero row: scale 0, guarded division
        out[f"c{t}_x"] = np16(x)
        xin = x.clone().reshape(M // 2, 2, K) if t == 2 else x.clone()   # a 3-D input once: cache.shape handling
        y = q(xin, None, True)
        snap(f"c{t}", out, q, cache, y, xin.reshape(M, K), M)
    out["ncalls"] = np.array(len(sets))
    return out


def case_fused_mlp(linear_mod, cache_mod, norm_mod, mlp_mod, bit, seed, M=8, K=256, I=384, fp=32):
    g = torch.Generator().manual_seed(seed)
    mk = lambda n, k: (torch.randn(n, k, generator=g) * 0.05).half()
    Wu, Wg, Wd = mk(I, K), mk(I, K), mk(K, I)
    nw = (1 + 0.1 * torch.randn(K, generator=g)).half()
    cache = cache_mod.MixLibCache(inputdim=32, sigma=6, bit=bit)
    perm = torch.randperm(K, generator=g)
    s0, s1 = perm[:3].sort().values, perm[3:5].sort().values
    nw_b = nw.clone()
    out = {"Wu": np16(Wu), "Wg": np16(Wg), "Wd": np16(Wd), "norm_w": np16(nw_b), "eps": np.array(1e-5),
           "M": np.array(M), "K": np.array(K), "I": np.array(I), "bit": np.array(bit), "fp": np.array(fp)}
    layer_scales = None
    if bit == 4:
        layer_scales = torch.rand(K, generator=g)
        layer_scales[s0] += 10.0
        out["layer_scales"] = np16(layer_scales)

    def ql(W, b, scales=None):
        lin = torch.nn.Linear(W.shape[1], W.shape[0], bias=False)
        lin.weight.data = W.clone()
        return linear_mod.MixLinear_GEMM.from_linear(lin, b, cache=cache, layer_scales=scales, dev="cpu", fp_features_num=fp)
    up, gate = ql(Wu, bit, layer_scales), ql(Wg, bit, layer_scales)
    down = ql(Wd, 8)   # down_proj stays 8-bit in 4-bit models (utils/module.py:2, base.py:308-312)
    norm = norm_mod.FasterTransformerRMSNorm(nw_b, 1e-5, cache)
    norm.next_layer = up
    mlp = mlp_mod.MixLlamaMLP(gate, down, up, cache)
    sets = [s0, torch.cat([s0, s1]), torch.cat([s0, s1])]
    for t, cols in enumerate(sets):
        x = torch.randn(M, K, generator=g)
        x[:, cols] *= 30.0
        x = x.half()
        out[f"c{t}_x"] = np16(x)
        h = norm(x.clone())
        out[f"c{t}_normed"] = np16(h)
        y = mlp(h)
        out[f"c{t}_y"] = np16(y)
        out[f"c{t}_up_ind"] = np16(up.ind).astype(np.int32)
        out[f"c{t}_gate_ind"] = np16(gate.ind).astype(np.int32)
        out[f"c{t}_down_ind"] = np16(down.ind).astype(np.int32)
    out["ncalls"] = np.array(len(sets))
    return out


def main():
    linear_mod, cache_mod, norm_mod, mlp_mod = install_reference()
    os.makedirs(OUT, exist_ok=True)
    cases = {
        "w8_unfused": case_unfused(linear_mod, cache_mod, 8, False, 1),
        "w8_unfused_bias": case_unfused(linear_mod, cache_mod, 8, True, 2),
        "w4_unfused": case_unfused(linear_mod, cache_mod, 4, False, 3),
        "w8_fused_mlp": case_fused_mlp(linear_mod, cache_mod, norm_mod, mlp_mod, 8, 4),
        "w4_fused_mlp": case_fused_mlp(linear_mod, cache_mod, norm_mod, mlp_mod, 4, 5),
    }
    for name, d in cases.items():
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **d)
        print(f"{name}: {len(d)} arrays, {os.path.getsize(path)/1024:.0f} KiB")
    # the reference's checkpoint layout: state_dict() keys / dtypes / shapes of its own MixLinear_GEMM (linear.py:39-65),
    # which is what save_quantized writes (base.py:78-119); mixq_b200/checkpoint.py must produce exactly these entries
    import json
    manifest = {}
    for tag, bit, bias in (("w8", 8, False), ("w8_bias", 8, True), ("w4", 4, False)):
        K, N = 256, 96
        cache = cache_mod.MixLibCache(inputdim=16, sigma=6, bit=bit)
        m = linear_mod.MixLinear_GEMM(K, N, bias, "cpu", bit, False, cache)
        manifest[tag] = {"in_features": K, "out_features": N,
                         "state": {k: [str(v.dtype).replace("torch.", ""), list(v.shape)] for k, v in m.state_dict().items()}}
    with open(os.path.join(OUT, "state_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("state_manifest.json:", {k: sorted(v["state"]) for k, v in manifest.items()})


if __name__ == "__main__":
    main()

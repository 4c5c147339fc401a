"""The reference's OWN Python on top of this library, on the GPU: the drop-in claim, executed.

`baseline/_ref/` holds verbatim copies of /root/reference/mixquant/{modules/linear.py, Cache.py, modules/fused/norm.py,
modules/fused/mlp.py} (staged by tools/stage_reference_py.py in the build container; git-ignored, ships with the gpurun
snapshot).  Here `sys.modules["mixlib"] = mixq_b200.mixlib` — nothing else is substituted except the
un-vendored EETQ import (a stub: weight_only is never taken for Llama, utils/module.py:6) — and the five golden fixtures
(recorded from the same reference files running over the CPU oracle, tests/golden/make_golden.py) are replayed through the
REFERENCE's MixLinear_GEMM / FasterTransformerRMSNorm / MixLlamaMLP (linear.py:165-289, :291-376, norm.py:14-39,
mlp.py:57-70), every mixlib.* call landing in libmixq_sm100.so.

Bars: outlier index sets, in-place zeroing, q_x, x_scale, weight_cache, gathered outliers bit-exact; y <= 1e-2 relative
(the reference computes the outlier GEMM with torch.mm here, cuBLAS accumulation order).
"""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
_NAMES = ("mixquant", "mixquant.modules", "mixquant.modules.fused", "mixquant.modules.linear", "mixquant.Cache",
          "mixquant.modules.fused.norm", "mixquant.modules.fused.mlp", "mixlib", "EETQ")


def host(t):
    return t.detach().cpu().numpy()


def bits_equal(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: {a.shape} vs {b.shape}"
    if a.dtype == np.float16:
        ua, ub = a.view(np.uint16), np.asarray(b, np.float16).view(np.uint16)
        bad = (ua != ub) & ~(((ua | ub) & 0x7FFF) == 0)
    else:
        bad = a != b
    assert not bad.any(), f"{what}: {int(bad.sum())} of {a.size} differ; first at {np.argwhere(bad)[:4].tolist()}"


def rel_close(a, b, what, tol=1e-2):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    rel = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
    assert rel <= tol, f"{what}: rel Frobenius error {rel:.3e} > {tol}"


@pytest.fixture(scope="module")
def ref_mods():
    if not os.path.exists(os.path.join(REF, "mixquant", "modules", "linear.py")):
        pytest.skip("baseline/_ref not staged (python tools/stage_reference_py.py in the build container)")
    from mixq_b200 import mixlib as shim
    saved = {n: sys.modules.get(n) for n in _NAMES}
    sys.modules["mixlib"] = shim
    eetq = types.ModuleType("EETQ")
    for n in ("quant_weights", "preprocess_weights", "w8_a16_gemm"):
        setattr(eetq, n, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("EETQ is outside the MixLinear path")))
    sys.modules["EETQ"] = eetq
    # bare namespace packages: mixquant/__init__.py (the HF model zoo) is not on this path and does not import on this image
    for name, path in (("mixquant", "mixquant"), ("mixquant.modules", "mixquant/modules"),
                       ("mixquant.modules.fused", "mixquant/modules/fused")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, path)]
        sys.modules[name] = m
    mods = tuple(importlib.import_module(n) for n in ("mixquant.modules.linear", "mixquant.Cache",
                                                      "mixquant.modules.fused.norm", "mixquant.modules.fused.mlp"))
    assert all(os.path.realpath(m.__file__).startswith(os.path.realpath(REF)) for m in mods)
    yield mods
    for n, v in saved.items():
        if v is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = v


def _linear(W, b=None):
    lin = torch.nn.Linear(W.shape[1], W.shape[0], bias=b is not None)
    lin.weight.data = torch.from_numpy(W.copy())
    if b is not None:
        lin.bias.data = torch.from_numpy(b.copy())
    return lin


@pytest.mark.parametrize("case", ["w8_unfused", "w8_unfused_bias", "w4_unfused"])
def test_reference_mixlinear_runs_on_the_shim(ref_mods, golden, case):
    linear_mod, cache_mod, _, _ = ref_mods
    from mixq_b200 import _lib
    d = golden(case)
    bit, M, K = int(d["bit"]), int(d["M"]), int(d["K"])
    cache = cache_mod.MixLibCache(inputdim=32, sigma=6, bit=bit)
    ls = torch.from_numpy(d["layer_scales"]) if bit == 4 else None
    n0 = _lib.launch_count()
    q = linear_mod.MixLinear_GEMM.from_linear(_linear(d["W"], d["bias"] if "bias" in d.files else None), bit, cache=cache,
                                              layer_scales=ls, dev="cuda", fp_features_num=int(d["fp"]))
    assert type(q).__module__ == "mixquant.modules.linear" and q.arch == 10      # the reference class, the arch != 9 branch
    bits_equal(host(q.q_weight), d["q_weight"], "q_weight")
    bits_equal(host(q.scale_col), d["scale_col"], "scale_col")
    for t in range(int(d["ncalls"])):
        x = torch.from_numpy(d[f"c{t}_x"].copy()).cuda()
        xin = x.reshape(M // 2, 2, K) if t == 2 else x
        y = q(xin, None, True)
        torch.cuda.synchronize()
        assert tuple(y.shape) == tuple(d[f"c{t}_y"].shape)
        bits_equal(host(q.ind), d[f"c{t}_ind"], f"call {t} outlier index set")
        bits_equal(host(x), d[f"c{t}_x_after"], f"call {t} x zeroed in place")
        bits_equal(host(cache.x_scale[:M]), d[f"c{t}_x_scale"], f"call {t} x_scale")
        bits_equal(host(cache.q_xcache), d[f"c{t}_q_x"], f"call {t} q_x")
        if q.ind.shape[0]:
            bits_equal(host(q.weight_cache), d[f"c{t}_weight_cache"], f"call {t} weight_cache")
            bits_equal(host(cache.activation_outliers), d[f"c{t}_act_outliers"], f"call {t} activation_outliers")
        assert int(q.add_outliers) == int(d[f"c{t}_add_outliers"])
        rel_close(host(y), d[f"c{t}_y"], f"call {t} y")
    assert _lib.launch_count() - n0 >= 2 * int(d["ncalls"]), "the reference's mixlib calls must land in libmixq_sm100.so"


@pytest.mark.parametrize("case", ["w8_fused_mlp", "w4_fused_mlp"])
def test_reference_fused_norm_mlp_runs_on_the_shim(ref_mods, golden, case):
    linear_mod, cache_mod, norm_mod, mlp_mod = ref_mods
    d = golden(case)
    bit, M, K = int(d["bit"]), int(d["M"]), int(d["K"])
    cache = cache_mod.MixLibCache(inputdim=32, sigma=6, bit=bit)
    ls = torch.from_numpy(d["layer_scales"]) if bit == 4 else None
    fp = int(d["fp"])
    mk = lambda W, b, s=None: linear_mod.MixLinear_GEMM.from_linear(_linear(W), b, cache=cache, layer_scales=s, dev="cuda",
                                                                     fp_features_num=fp)
    up, gate, down = mk(d["Wu"], bit, ls), mk(d["Wg"], bit, ls), mk(d["Wd"], 8)
    norm = norm_mod.FasterTransformerRMSNorm(torch.from_numpy(d["norm_w"]), float(d["eps"]), cache)
    norm.next_layer = up
    mlp = mlp_mod.MixLlamaMLP(gate, down, up, cache)
    for t in range(int(d["ncalls"])):
        x = torch.from_numpy(d[f"c{t}_x"].copy()).cuda()
        h = norm(x)
        ref_n = d[f"c{t}_normed"]
        assert (np.abs(host(h).astype(np.float32) - ref_n.astype(np.float32)) <= np.abs(np.spacing(ref_n)).astype(np.float32)).all()
        y = mlp(h)
        torch.cuda.synchronize()
        bits_equal(host(up.ind), d[f"c{t}_up_ind"], f"call {t} up_proj outlier set")
        bits_equal(host(gate.ind), d[f"c{t}_gate_ind"], f"call {t} gate_proj outlier set")
        bits_equal(host(down.ind), d[f"c{t}_down_ind"], f"call {t} down_proj outlier set")
        rel_close(host(y), d[f"c{t}_y"], f"call {t} y")

"""CPU: the oracle (oracle/mixq_oracle.py) against the golden vectors recorded from the reference's own
Python (tests/golden/make_golden.py).  Everything here is bit-exact: integer / fp16-bit comparisons."""
import numpy as np
import pytest

from oracle import mixq_oracle as O


def eq(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype == np.float16:
        a, b = a.view(np.uint16), np.asarray(b, np.float16).view(np.uint16)
        # +0 / -0 compare equal
        bad = (a != b) & ~(((a | b) & 0x7FFF) == 0)
    else:
        bad = a != b
    assert not bad.any(), f"{what}: {int(bad.sum())} of {a.size} elements differ"


def near(a, b, what, frac=0.01, rtol=2e-3):
    """y goes through the fp16 outlier GEMM, whose accumulation order is implementation-defined (torch CPU mm in
    the fixture, float64 in the oracle, fp32 tensor cores on the GPU): results may differ by an fp16 rounding of
    the outlier term on a few elements.  Everything else is compared bit-exactly with eq()."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    tol = rtol * max(1.0, float(np.abs(b).max()))
    diff = np.abs(a - b)
    assert diff.max() <= tol, f"{what}: max |diff| {diff.max()} > {tol}"
    nbad = int((diff > 0).sum())
    assert nbad <= frac * a.size, f"{what}: {nbad} of {a.size} elements differ"


@pytest.mark.parametrize("case", ["w8_unfused", "w8_unfused_bias", "w4_unfused"])
def test_weight_quant_matches_reference(golden, case):
    """from_linear (linear.py:111-143) is pure torch in the reference: this pins the oracle's weight arithmetic."""
    d = golden(case)
    bit = int(d["bit"])
    if bit == 8:
        q, s = O.quant_weight_w8(d["W"])
        eq(q, d["q_weight"], "q_weight")
        eq(s, d["scale_col"], "scale_col")
    else:
        q, s, wc, ind = O.quant_weight_w4(d["W"], d["layer_scales"], int(d["fp"]))
        eq(q, d["q_weight"], "q_weight (packed)")
        eq(s, d["scale_col"], "scale_col")
        eq(ind, d["ind0"], "static ind")
        eq(wc, d["weight_cache0"], "weight_cache")


@pytest.mark.parametrize("case", ["w8_unfused", "w8_unfused_bias", "w4_unfused"])
def test_forward_state_machine_matches_reference(golden, case):
    """linear.py:165-289 run by the reference itself vs MixLinearOracle.forward: outlier sets, in-place zeroing,
    hstack order, stop after cache.stop calls, bias add, 3-D input reshape."""
    d = golden(case)
    bit, M, K = int(d["bit"]), int(d["M"]), int(d["K"])
    cache = O.MixLibCacheOracle(32, 6, bit)
    bias = d["bias"] if "bias" in d.files else None
    lin = O.MixLinearOracle.from_linear(d["W"], bit, bias=bias, cache=cache,
                                        layer_scales=d["layer_scales"] if bit == 4 else None,
                                        fp_features_num=int(d["fp"]))
    for t in range(int(d["ncalls"])):
        x = d[f"c{t}_x"].copy()
        xin = x.reshape(M // 2, 2, K) if t == 2 else x
        y = lin.forward(xin, None, True)
        eq(lin.ind, d[f"c{t}_ind"], f"call {t} ind")
        eq(xin.reshape(M, K), d[f"c{t}_x_after"], f"call {t} x zeroed in place")
        eq(cache.x_scale[:M], d[f"c{t}_x_scale"], f"call {t} x_scale")
        eq(cache.q_xcache, d[f"c{t}_q_x"], f"call {t} q_x")
        if len(lin.ind):
            eq(lin.weight_cache, d[f"c{t}_weight_cache"], f"call {t} weight_cache")
            eq(cache.activation_outliers, d[f"c{t}_act_outliers"], f"call {t} activation_outliers")
        assert int(lin.add_outliers) == int(d[f"c{t}_add_outliers"])
        near(y, d[f"c{t}_y"], f"call {t} y")


@pytest.mark.parametrize("case", ["w8_fused_mlp", "w4_fused_mlp"])
def test_fused_norm_mlp_matches_reference(golden, case):
    """norm.py:14-39 + mlp.py:57-70 + linear.py:291-376 (gate_proj re-using up_proj's quantised input)."""
    d = golden(case)
    bit, M, K = int(d["bit"]), int(d["M"]), int(d["K"])
    cache = O.MixLibCacheOracle(32, 6, bit)
    ls = d["layer_scales"] if bit == 4 else None
    fp = int(d["fp"])
    up = O.MixLinearOracle.from_linear(d["Wu"], bit, cache=cache, layer_scales=ls, fp_features_num=fp)
    gate = O.MixLinearOracle.from_linear(d["Wg"], bit, cache=cache, layer_scales=ls, fp_features_num=fp)
    down = O.MixLinearOracle.from_linear(d["Wd"], 8, cache=cache)
    for t in range(int(d["ncalls"])):
        x = d[f"c{t}_x"].copy()
        h, ao, q_x, xs = O.rmsnorm_extract_outliers(x, d["norm_w"], float(d["eps"]), up.ind, bit)
        cache.activation_outliers, cache.q_xcache = ao, q_x
        cache.x_scale[:M] = xs
        eq(h, d[f"c{t}_normed"], f"call {t} normed")
        u = up.forward(h, cache)
        g = gate.forward_without_precondition_fused_silu(h, cache)
        g = (g.astype(np.float32) * u.astype(np.float32)).astype(np.float16)   # gate_output *= up_output (fp16)
        y = down.forward(g, None, True)
        eq(up.ind, d[f"c{t}_up_ind"], f"call {t} up ind")
        eq(gate.ind, d[f"c{t}_gate_ind"], f"call {t} gate ind")
        eq(down.ind, d[f"c{t}_down_ind"], f"call {t} down ind")
        near(y, d[f"c{t}_y"], f"call {t} y", frac=0.05, rtol=5e-3)


def test_sample_py_algorithm_agrees_with_forward():
    """models/sample.py:5-12 (the algorithm in one screen) == steady-state forward."""
    rng = np.random.default_rng(0)
    M, K, N = 8, 128, 64
    W = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
    x = rng.standard_normal((M, K)).astype(np.float16)
    cols = np.array([3, 77], np.int32)
    x[:, cols] *= np.float16(25)
    cache = O.MixLibCacheOracle(32, 6, 8)
    lin = O.MixLinearOracle.from_linear(W, 8, cache=cache)
    y = lin.forward(x.copy(), None, True)
    y2 = O.mixgemm_sample(x, lin.q_weight, lin.scale_col, lin.ind)
    eq(lin.ind, cols, "ind")
    eq(y, y2, "y")


def test_pack_unpack_roundtrip_and_edges():
    q = np.arange(-8, 8, dtype=np.int8).reshape(2, 8)
    p = O.pack_to_i4(q)
    assert p.dtype == np.uint8 and p.shape == (2, 4)
    eq(O.unpack_i4(p), q, "roundtrip")
    assert p[0, 0] == ((16 - 8) | ((16 - 7) << 4))   # low nibble = even column
    # empty / zero rows / int4 activations
    q_x, xs = O.find_row_scale(np.zeros((3, 16), np.float16), 8)
    assert not q_x.any() and not xs.any()
    q4, xs4 = O.find_row_scale(np.array([[7, -3.5, 0.1, 0]], np.float16), 4)
    assert q4.tolist() == [[7, -4, 0, 0]] or q4.tolist() == [[7, -3, 0, 0]]
    assert np.abs(q4).max() <= 7
    assert O.find_outliers(np.zeros((2, 8), np.float16), 6).shape == (0,)


def test_error_budget_vs_fp32_linear():
    """the quantised oracle vs the un-quantised fp32 Linear — inherent W8A8 error ~1e-2."""
    rng = np.random.default_rng(0)
    M, K, N = 32, 1024, 256
    W = (rng.standard_normal((N, K)) * 0.02).astype(np.float16)
    x = rng.standard_normal((M, K)).astype(np.float16)
    cols = rng.permutation(K)[:10]
    x[:, cols] *= np.float16(20)
    cache = O.MixLibCacheOracle(32, 6, 8)
    lin = O.MixLinearOracle.from_linear(W, 8, cache=cache)
    ref = O.linear_fp32(x, W)
    y = lin.forward(x.copy(), None, True).astype(np.float32)
    rel = np.linalg.norm(y - ref) / np.linalg.norm(ref)
    assert sorted(lin.ind.tolist()) == sorted(cols.tolist())
    assert rel < 2e-2, rel

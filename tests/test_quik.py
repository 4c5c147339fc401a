"""QUIK MixedQLinear (mixquant/modules/qlinear.py:41-211): the oracle against fixtures recorded from the reference's own class
(tests/golden/make_golden_quik.py), and the CUDA path (mixq_b200.qlinear.MixedQLinear through the C ABI) against both."""
import numpy as np
import pytest
import torch

from oracle import quik_oracle as Q


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("case", ["quik_w4", "quik_w8"])
def test_quik_oracle_replays_reference_golden(golden, case):
    """from_linear (the reference's real torch arithmetic: rounding, nibble packing, reduced_w) bit-exact; forward equal."""
    d = golden(case)
    bits = int(d["bits"])
    st = Q.from_linear(d["W"], d["weights_scales"], d["fp_indices"], bits)
    assert np.array_equal(st["int_weight"], d["int_weight"])
    assert np.array_equal(st["int_indices"], d["int_indices"])
    assert np.array_equal(st["reduced_w"].view(np.uint16), d["reduced_w"].view(np.uint16))
    assert np.array_equal(st["fp_weight"].view(np.uint16), d["fp_weight"].view(np.uint16))
    for t in range(2):
        y, aux = Q.mixed_qlinear_forward(d[f"c{t}_x"], st, bits)
        assert y.shape == d[f"c{t}_y"].shape
        assert np.array_equal(y.view(np.uint16), d[f"c{t}_y"].view(np.uint16))
        half = 2 ** (bits - 1)
        assert aux["q"].min() >= -half and aux["q"].max() <= half - 1
    # against the un-quantised Linear: W4A4 a few percent, W8A8 well below one percent
    x = d["c0_x"]
    ref = x.astype(np.float32) @ d["W"].astype(np.float32).T
    y, _ = Q.mixed_qlinear_forward(x, st, bits)
    assert _rel(y, ref) < (0.06 if bits == 4 else 0.01)


def test_quik_oracle_properties():
    """x = scale (q + 2^(b-1)) + zero reconstructs the int columns to half a step; a constant row has scale 0 and is exact."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 64)).astype(np.float16)
    x[3] = np.float16(0.75)
    ii, fi = np.arange(8, 64), np.arange(8)
    for bits in (4, 8):
        q, meta, fp_x = Q.asymmetric_quantize(x, ii, fi, bits)
        s, z = meta[0].astype(np.float32), meta[1].astype(np.float32)
        rec = s[:, None] * (q.astype(np.float32) + 2 ** (bits - 1)) + z[:, None]
        err = np.abs(rec - x[:, ii].astype(np.float32))
        assert (err <= 0.5 * s[:, None] + 2e-3 * np.abs(x[:, ii].astype(np.float32)) + 1e-3).all()
        assert s[3] == 0 and (rec[3] == 0.75).all()
        assert np.array_equal(fp_x, x[:, fi])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["quik_w4", "quik_w8"])
def test_mixed_qlinear_gpu_replays_reference_golden(golden, case):
    from mixq_b200.qlinear import MixedQLinear, SharedQuantizedInput
    d = golden(case)
    bits = int(d["bits"])

    def lin(W):
        l = torch.nn.Linear(W.shape[1], W.shape[0], bias=False)
        l.weight.data = torch.from_numpy(W.copy())
        return l
    fp_idx = torch.from_numpy(d["fp_indices"])
    m = MixedQLinear.from_linear(lin(d["W"]), torch.from_numpy(d["W"].copy()), torch.from_numpy(d["weights_scales"].copy()), None,
                                 fp_idx, False, bits)
    assert sorted(m.state_dict()) == sorted(["weights_scales", "int_weight", "int_indices", "fp_indices", "fp_weight", "reduced_w"])
    assert np.array_equal(m.int_weight.cpu().numpy(), d["int_weight"])
    assert np.array_equal(m.reduced_w.cpu().numpy().view(np.uint16), d["reduced_w"].view(np.uint16))
    st = Q.from_linear(d["W"], d["weights_scales"], d["fp_indices"], bits)
    for t in range(2):
        x = torch.from_numpy(d[f"c{t}_x"].copy()).cuda()
        y = m(x)
        assert tuple(y.shape) == d[f"c{t}_y"].shape
        assert _rel(y.cpu().numpy(), d[f"c{t}_y"]) <= 1e-2          # north-star tolerance; the fp part is a cuBLAS GEMM
    # the kernels alone, bit-exact against the oracle: quantise, addend, int GEMM + dequantise
    import ctypes as C
    from mixq_b200 import _lib
    lib = _lib.load()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    x = d["c0_x"]
    M, K = x.shape
    n_int, n_fp, N = len(st["int_indices"]), len(st["fp_indices"]), d["W"].shape[0]
    q_ref, meta_ref, fpx_ref = Q.asymmetric_quantize(x, st["int_indices"], st["fp_indices"], bits)
    xd = torch.from_numpy(x.copy()).cuda()
    q = torch.zeros(M, n_int, dtype=torch.int8, device="cuda")
    meta = torch.zeros(2, M, dtype=torch.float16, device="cuda")
    fpx = torch.zeros(M, n_fp, dtype=torch.float16, device="cuda")
    _lib.check(lib.mixq_quik_quantize(xd.data_ptr(), m.int_indices.data_ptr(), n_int, m.fp_indices.data_ptr(), n_fp, bits, q.data_ptr(),
                                      meta.data_ptr(), fpx.data_ptr(), M, K, stream), "quantize")
    assert np.array_equal(q.cpu().numpy(), q_ref)
    assert np.array_equal(meta.cpu().numpy().view(np.uint16), meta_ref.view(np.uint16))
    assert np.array_equal(fpx.cpu().numpy().view(np.uint16), fpx_ref.view(np.uint16))
    fp_res = Q.fp_linear(fpx_ref, st["fp_weight"])
    add_ref = Q.asymmetric_addend(meta_ref, st["reduced_w"], fp_res, bits)
    fpr = torch.from_numpy(fp_res).cuda()
    add = torch.zeros(M, N, dtype=torch.float16, device="cuda")
    _lib.check(lib.mixq_quik_addend(meta.data_ptr(), m.reduced_w.data_ptr(), fpr.data_ptr(), N, add.data_ptr(), M, N, bits, stream), "addend")
    assert np.array_equal(add.cpu().numpy().view(np.uint16), add_ref.view(np.uint16))
    y = torch.zeros(M, N, dtype=torch.float16, device="cuda")
    fn = lib.mixq_int4_fused_dequantize if bits == 4 else lib.mixq_int8_fused_dequantize
    _lib.check(fn(q.data_ptr(), m.int_weight.data_ptr(), meta.data_ptr(), m.weights_scales.data_ptr(), add.data_ptr(), N, y.data_ptr(),
                  M, N, n_int, 0, stream), "int matmul + dequantize")
    y_ref = Q.asymmetric_dequantize(Q.int_matmul(q_ref, st["int_weight"], bits), meta_ref, st["weights_scales"], st["reduced_w"], fp_res, bits)
    assert np.array_equal(y.cpu().numpy().view(np.uint16), y_ref.view(np.uint16)), "y bit-exact against the oracle"
    # shared quantised input across two Linears (qlinear.py:22-38)
    sh = SharedQuantizedInput(2)
    a = MixedQLinear.from_linear(lin(d["W"]), torch.from_numpy(d["W"].copy()), torch.from_numpy(d["weights_scales"].copy()), sh, fp_idx, False, bits)
    b = MixedQLinear.from_linear(lin(d["W2"]), torch.from_numpy(d["W2"].copy()), torch.from_numpy(d["weights_scales2"].copy()), sh, fp_idx, False, bits)
    xs = torch.from_numpy(d["sh_x"].copy()).cuda()
    ya = a(xs)
    assert sh.qint_x is not None and sh.cur_group_elem == 1
    yb = b(xs)
    assert sh.qint_x is None and sh.cur_group_elem == 0
    assert _rel(ya.cpu().numpy(), d["sh_ya"]) <= 1e-2 and _rel(yb.cpu().numpy(), d["sh_yb"]) <= 1e-2


@pytest.mark.gpu
def test_mixed_qlinear_llama_shape():
    """C3-sized QUIK Linear (K = 4096 with 256 fp16 columns, N = 11008, M = 512) against the oracle."""
    from mixq_b200.qlinear import MixedQLinear
    rng = np.random.default_rng(3)
    M, K, N, n_fp = 512, 4096, 11008, 256
    W = (rng.standard_normal((N, K)) * 0.02).astype(np.float16)
    fp = np.sort(rng.permutation(K)[:n_fp])
    mask = np.ones(K, bool)
    mask[fp] = False
    ws = (np.abs(W[:, mask]).max(1, keepdims=True) / 7).astype(np.float16)
    l = torch.nn.Linear(K, N, bias=False)
    l.weight.data = torch.from_numpy(W.copy())
    m = MixedQLinear.from_linear(l, torch.from_numpy(W.copy()), torch.from_numpy(ws.copy()), None, torch.from_numpy(fp), False, 4)
    x = rng.standard_normal((M, K)).astype(np.float16)
    x[:, fp] = (x[:, fp].astype(np.float32) * 20).astype(np.float16)
    st = Q.from_linear(W, ws, fp, 4)
    y_ref, _ = Q.mixed_qlinear_forward(x, st, 4)
    y = m(torch.from_numpy(x.copy()).cuda())
    assert _rel(y.cpu().numpy(), y_ref) <= 1e-2

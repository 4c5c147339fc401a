"""CPU restatement of the QUIK path of the reference: `MixedQLinear` (mixquant/modules/qlinear.py:41-211) and the `quik.*`
kernels it calls.  TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg may import it).

`quik` is a third-party CUDA extension (IST-DASLab/QUIK) that the reference imports from a hard-coded home directory
(qlinear.py:6-7: sys.path.append('/home/chenyidong/QUIK/')); it is NOT in /root/reference, has no pinned version, and the
reference holds no tests or golden vectors for it: **parity unpinned**.  What is pinned by the reference's own call sites:
  * qlinear.py:117-120  quik.asymmetric.quantize(x, int_indices, fp_indices, bits) -> (qint_x, meta, fp_x)
  * qlinear.py:104-106  symmetric: qscale_x = rowabsmax(int_x) / (1 << (bits-1) - 1)  [sic: Python precedence gives
                        1 << ((bits-1) - 1) = 2^(bits-2), i.e. 4 for bits = 4]; quik.symmetric.quantize(int_x, qscale_x)
  * qlinear.py:142-144  quik.matmul.int4Matmul / int8Matmul(qint_x, int_weight) -> int32 [M, N]
  * qlinear.py:147-150  quik.symmetric.dequantize(int_result, qscale_x, weights_scales, fp_result) /
                        quik.asymmetric.dequantize(int_result, meta, weights_scales, reduced_w, fp_result, bits)
  * qlinear.py:186-196  weights: round(W[:, int_idx] / weights_scales) packed two's-complement nibbles (low = even column),
                        reduced_w = sum_k W[:, int_idx] (fp32 sum of the UN-quantised weights, rounded to fp16)
Published algorithm restated (QUIK paper, arXiv 2310.09259 section 3.2, eq. for asymmetric activations): per token row
    zero = min_k x, scale = (max_k x - min_k x) / (2^bits - 1), q = rn((x - zero) / scale) - 2^(bits-1)   in [-2^(b-1), 2^(b-1)-1]
    x ~ scale * (q + 2^(bits-1)) + zero
    y[m,n] = scale[m] * ws[n] * sum_k q[m,k] qw[n,k]  +  (zero[m] + 2^(bits-1) * scale[m]) * reduced_w[n]  +  fp_result[m,n]
Where the source is silent this oracle fixes: scale / zero stored as fp16 (meta is an fp16 tensor at the call site), IEEE fp32
arithmetic, round-half-even, correction + fp part rounded to fp16 once (the addend), result rounded to fp16 once.
"""
import numpy as np

F16, F32 = np.float16, np.float32


def pack_to_i4(x_i8):
    """qlinear.py:16-19 (identical to linear.py:14-18)."""
    u = np.where(x_i8 < 0, 16 + x_i8.astype(np.int16), x_i8.astype(np.int16)).astype(np.uint8)
    return (u[:, 0::2] | (u[:, 1::2] << 4)).astype(np.uint8)


def unpack_i4(p):
    lo = (p & 0xF).astype(np.int8)
    hi = (p >> 4).astype(np.int8)
    lo = np.where(lo > 7, lo - 16, lo)
    hi = np.where(hi > 7, hi - 16, hi)
    out = np.empty((p.shape[0], p.shape[1] * 2), np.int8)
    out[:, 0::2], out[:, 1::2] = lo, hi
    return out


def from_linear(weight, weights_scales, fp_indices, bits=4, symm=False):
    """MixedQLinear.from_linear (qlinear.py:153-211).  weight fp16 [N,K], weights_scales fp16 [N,1].
    Returns dict(int_weight, weights_scales, int_indices, fp_indices, fp_weight, reduced_w)."""
    W = np.asarray(weight, F16)
    N, K = W.shape
    fp_indices = np.asarray(fp_indices, np.int64)
    mask = np.ones(K, bool)
    mask[fp_indices] = False
    int_indices = np.nonzero(mask)[0].astype(np.int64)
    ws = np.asarray(weights_scales, F16).reshape(N, 1)
    q = np.rint((W[:, int_indices].astype(F32) / ws.astype(F32)).astype(F16).astype(F32))     # fp16 division on the GPU, then round
    out = {"weights_scales": ws, "int_indices": int_indices, "fp_indices": fp_indices,
           "fp_weight": W[:, fp_indices].copy(),
           "int_weight": pack_to_i4(q.astype(np.int8)) if bits == 4 else q.astype(np.int8)}
    if not symm:
        out["reduced_w"] = W[:, int_indices].astype(F32).sum(axis=1, dtype=F32).astype(F16).reshape(1, N)
    return out


def asymmetric_quantize(x, int_indices, fp_indices, bits=4):
    """quik.asymmetric.quantize (qlinear.py:117-120): -> (q int8 [M, n_int] in [-2^(b-1), 2^(b-1)-1], meta fp16 [2, M] =
    (scale, zero), fp_x fp16 [M, n_fp])."""
    x = np.asarray(x, F16)
    xi = x[:, int_indices].astype(F32)
    mn, mx = xi.min(axis=1), xi.max(axis=1)
    levels = F32(2 ** bits - 1)
    scale = ((mx - mn) / levels).astype(F16)
    zero = mn.astype(F16)
    s = scale.astype(F32)
    r = np.where(s > 0, F32(1) / np.where(s > 0, s, F32(1)), F32(0)).astype(F32)
    q = np.rint(((xi - zero.astype(F32)[:, None]) * r[:, None]).astype(F32))
    half = 2 ** (bits - 1)
    q = np.clip(q - half, -half, half - 1).astype(np.int8)
    return q, np.stack([scale, zero]), x[:, fp_indices].copy()


def int_matmul(q_x, int_weight, bits=4):
    """quik.matmul.int4Matmul / int8Matmul (qlinear.py:142-144): exact int32 sums (float64 BLAS is exact here)."""
    w = unpack_i4(int_weight) if bits == 4 else int_weight
    return (q_x.astype(np.float64) @ w.astype(np.float64).T).astype(np.int64).astype(np.int32)


def asymmetric_addend(meta, reduced_w, fp_result, bits=4):
    """(zero + 2^(b-1) scale) * reduced_w + fp_result, fp32, one rounding to fp16."""
    scale, zero = meta[0].astype(F32), meta[1].astype(F32)
    shift = (zero + F32(2 ** (bits - 1)) * scale).astype(F32)
    v = shift[:, None] * np.asarray(reduced_w, F16).reshape(1, -1).astype(F32)
    if fp_result is not None:
        v = v + np.asarray(fp_result, F16).astype(F32)
    return v.astype(F16)


def asymmetric_dequantize(int_result, meta, weights_scales, reduced_w, fp_result, bits=4):
    """quik.asymmetric.dequantize (qlinear.py:149-150)."""
    add = asymmetric_addend(meta, reduced_w, fp_result, bits).astype(F32)
    v = (int_result.astype(F32) * meta[0].astype(F32)[:, None]) * np.asarray(weights_scales, F16).reshape(1, -1).astype(F32)
    return (v + add).astype(F16)


def fp_linear(fp_x, fp_weight, bias=None):
    """torch.nn.functional.linear(fp_x, fp_weight, bias) on fp16 tensors (qlinear.py:129): fp32 accumulate, fp16 result."""
    y = fp_x.astype(F32) @ fp_weight.astype(F32).T
    if bias is not None:
        y = y + bias.astype(F32)[None]
    return y.astype(F16)


def mixed_qlinear_forward(x, st, bits=4, bias=None):
    """MixedQLinear.forward, asymmetric branch (qlinear.py:82-152)."""
    shape = x.shape
    x2 = np.asarray(x, F16).reshape(-1, shape[-1])
    q, meta, fp_x = asymmetric_quantize(x2, st["int_indices"], st["fp_indices"], bits)
    fp_result = fp_linear(fp_x, st["fp_weight"], bias) if len(st["fp_indices"]) else None
    acc = int_matmul(q, st["int_weight"], bits)
    y = asymmetric_dequantize(acc, meta, st["weights_scales"], st["reduced_w"], fp_result, bits)
    return y.reshape(shape[:-1] + (y.shape[-1],)), dict(q=q, meta=meta, fp_x=fp_x, acc=acc)

"""CPU oracle for the MixLinear hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this module; the product path (mixq_b200/) never does and has no CPU fallback.

It restates, in numpy with explicit IEEE fp16/fp32 steps, the arithmetic of the reference's hot path:
  /root/reference/mixquant/modules/linear.py   (MixLinear_GEMM: from_linear, FindOutliers, forward,
                                                forward_without_preconditionFusedSilu, pack_to_i4)
  /root/reference/mixquant/Cache.py            (MixLibCache: sigma = 6, stop = 2)
  /root/reference/mixquant/modules/fused/norm.py (RMSNorm fused with extract + quantise)
  /root/reference/mixquant/models/sample.py:5-12 (the algorithm in one screen)

PARITY PINNING.  The control flow (discovery state machine, hstack order, which buffers are reused)
is pinned: tests/golden/make_golden.py imports the reference's own linear.py in the build container and
records its outputs; tests/test_oracle_golden.py replays them against this file.  The per-kernel
arithmetic of `mixlib.*` is NOT pinned by the reference: those kernels live in the un-vendored, un-pinned
repository github.com/Qcompiler/QComplier (quantkernel -> `mixlib`; README.md:39-49 says "git clone",
no version), the reference tree holds no tests, golden vectors or fixtures for them (SURVEY.md §4,
§8c), and `import mixlib` fails here.  For that layer: "parity unpinned"; the choices below are the
ones the call sites force (linear.py:201 fixes x_scale*(2^(bit-1)-1) == row absmax) plus, where the tree
is silent: round-half-even, symmetric clamp, IEEE fp32 division, one rounding to fp16 at the end.
"""
from __future__ import annotations

import numpy as np

F16, F32, F64 = np.float16, np.float32, np.float64


# --------------------------------------------------------------------------- weights (offline)
def quant_weight_w8(weight: np.ndarray):
    """linear.py:111-119.  weight [N,K] fp16 (or fp32).  Returns (q_weight int8 [N,K], scale_col fp16 [1,N]).

    scale = (rowabsmax / 127).to(fp16); tmp = W; tmp /= scale.T (in W's dtype); tmp.round().to(int8).
    torch computes fp16 elementwise ops in fp32 and rounds the result to fp16; numpy float16 does the same.
    """
    w = np.asarray(weight)
    amax = np.abs(w).max(axis=1)
    scale = (amax.astype(F32) / F32(127)).astype(w.dtype).astype(F16)  # division in W's dtype, then .to(fp16)
    tmp = (w.astype(F32) / scale.astype(F32)[:, None]).astype(w.dtype)
    q = np.rint(tmp.astype(F32)).astype(np.int8)
    return q, scale.reshape(1, -1)


def pack_to_i4(x_i8: np.ndarray) -> np.ndarray:
    """linear.py:12-18: two's-complement nibbles, low nibble = even column, high nibble = odd column."""
    u = np.where(x_i8 < 0, 16 + x_i8.astype(np.int16), x_i8.astype(np.int16)).astype(np.uint8)
    return (u[:, 0::2] | (u[:, 1::2] << 4)).astype(np.uint8)


def unpack_i4(packed: np.ndarray) -> np.ndarray:
    lo = (packed & 0xF).astype(np.int8)
    hi = (packed >> 4).astype(np.int8)
    lo = np.where(lo >= 8, lo - 16, lo).astype(np.int8)
    hi = np.where(hi >= 8, hi - 16, hi).astype(np.int8)
    out = np.empty((packed.shape[0], packed.shape[1] * 2), np.int8)
    out[:, 0::2] = lo
    out[:, 1::2] = hi
    return out


def quant_weight_w4(weight: np.ndarray, layer_scales: np.ndarray, fp_features: int = 128):
    """linear.py:121-143.  Static outliers = the fp_features columns with the largest layer_scales
    (torch.sort ascending, last fp_features entries, in that order); they stay fp16 in weight_cache and
    are zeroed before 4-bit quantisation with scale = rowabsmax/10, clamp [-8,7]."""
    w = np.asarray(weight).copy()
    ind = np.argsort(np.asarray(layer_scales), kind="stable")[-fp_features:].astype(np.int32)
    weight_cache = w[:, ind].copy()
    w[:, ind] = 0
    amax = np.abs(w).max(axis=1)
    scale = (amax.astype(F32) / F32(10)).astype(w.dtype).astype(F16)
    tmp = (w.astype(F32) / scale.astype(F32)[:, None]).astype(w.dtype)
    q = np.clip(np.rint(tmp.astype(F32)), -8, 7).astype(np.int8)
    return pack_to_i4(q), scale.reshape(1, -1), weight_cache.astype(F16), ind


# --------------------------------------------------------------------------- mixlib.* restated
def find_row_scale(x: np.ndarray, bit: int = 8):
    """mixlib.FindRowScale (linear.py:190-193): x fp16 [M,K] -> (q_x int8 [M,K], x_scale fp16 [M,1]).
    x_scale = fp16(absmax / qmax) (fp32 divide); q = clamp(rint(x / x_scale)) (fp32 divide); a zero row
    gives scale 0 and q 0."""
    qmax = F32(2 ** (bit - 1) - 1)
    xf = np.asarray(x, F16).astype(F32)
    amax = np.abs(xf).max(axis=1) if xf.shape[1] else np.zeros(xf.shape[0], F32)
    xs = (amax / qmax).astype(F16)
    xs32 = xs.astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.where(xs32[:, None] > 0, np.rint(xf / xs32[:, None]), F32(0))
    q = np.clip(q, -qmax, qmax).astype(np.int8)
    return q, xs.reshape(-1, 1)


def extract_outliers_and_set_to_zeros(ind: np.ndarray, x: np.ndarray) -> np.ndarray:
    """mixlib.ExtractOutliersAndSetToZeros (linear.py:189, :205): returns x[:, ind] and zeroes those
    columns of x IN PLACE."""
    ind = np.asarray(ind, np.int64)
    out = x[:, ind].copy()
    x[:, ind] = 0
    return out


def find_outliers(x: np.ndarray, sigma) -> np.ndarray:
    """MixLinear_GEMM.FindOutliers (linear.py:157-161): sorted unique column ids with any |x| > sigma."""
    cols = np.nonzero((np.abs(np.asarray(x, F16)) > F16(sigma)).any(axis=0))[0]
    return cols.astype(np.int32)


_GEMM_CHUNK = 1024     # 1024 * 127^2 < 2^24: every partial sum of a chunk is an integer that fp32 holds exactly


def weights_f32_chunks(q_w: np.ndarray):
    """The int8 weight matrix as fp32 K-chunks (a cache for gemm_i8: weights are constants, converting them per call would
    dominate the CPU baseline)."""
    return [np.ascontiguousarray(q_w[:, c:c + _GEMM_CHUNK].astype(F32).T) for c in range(0, q_w.shape[1], _GEMM_CHUNK)]


def gemm_i8(q_x: np.ndarray, q_w: np.ndarray, w_chunks=None) -> np.ndarray:
    """mixlib.gemm (linear.py:235): exact int32 accumulation, on the host's fp32 BLAS: K is cut into chunks of 1024 so that
    every partial sum (|sum| <= 1024 * 127^2 < 2^24) is exactly representable whatever order the BLAS adds in; the chunk
    results are added as int32.  Bit-identical to an integer GEMM (and to the float64 product it replaced, which was 2x slower
    for nothing)."""
    if w_chunks is None:
        w_chunks = weights_f32_chunks(q_w)
    acc = None
    for i, wc in enumerate(w_chunks):
        part = (q_x[:, i * _GEMM_CHUNK:(i + 1) * _GEMM_CHUNK].astype(F32) @ wc).astype(np.int32)
        acc = part if acc is None else acc + part
    return acc


def silu(v: np.ndarray) -> np.ndarray:
    v = v.astype(F32)
    return (v / (F32(1) + np.exp(-v))).astype(F32)


def dequantize(acc_i32, x_scale, scale_col, outl=None, bias=None, act: int = 0, outl_f32=None) -> np.ndarray:
    """Epilogue of mixlib.int8FusedDequantize / dequantizeInt8 (linear.py:238, :251):
    y = act(fp16( (f32(acc) * x_scale[m]) * scale_col[n] + outl[m,n] + bias[n] )), all in IEEE fp32,
    one rounding to fp16.  `outl` is the fp16 [M,N] addend the reference passes; `outl_f32` is the
    un-rounded fp32 outlier product the fused kernel keeps in TMEM."""
    M, N = acc_i32.shape
    xs = np.asarray(x_scale, F16).reshape(-1)[:M].astype(F32)[:, None]
    ws = np.asarray(scale_col, F16).reshape(-1).astype(F32)[None, :]
    v = (acc_i32.astype(F32) * xs) * ws
    if outl_f32 is not None:
        v = v + outl_f32.astype(F32)
    if outl is not None:
        v = v + np.asarray(outl, F16)[:M, :N].astype(F32)
    if bias is not None:
        v = v + np.asarray(bias, F16).astype(F32)[None, :]
    if act == 1:
        v = silu(v)
    return v.astype(F16)


def int8_fused_dequantize(q_x, q_w, x_scale, scale_col, outl=None, act: int = 0):
    """mixlib.int8FusedDequantize[Silu] (linear.py:251-256, :337-342)."""
    return dequantize(gemm_i8(q_x, q_w), x_scale, scale_col, outl=outl, act=act)


def int4_fused_dequantize(q_x, q_w_packed, x_scale, scale_col, outl=None, act: int = 0):
    """mixlib.int4FusedDequantize[Silu] (linear.py:259-265): packed-nibble weights, activations in [-7,7]."""
    return dequantize(gemm_i8(q_x, unpack_i4(q_w_packed)), x_scale, scale_col, outl=outl, act=act)


def unpack_int4_to_fp16(q_w_packed: np.ndarray, ind: np.ndarray) -> np.ndarray:
    """mixlib.unpack_int4_to_fp16 (linear.py:20-22): sign-extended nibbles of the ind columns, fp16, un-scaled."""
    return unpack_i4(q_w_packed)[:, np.asarray(ind, np.int64)].astype(F16)


def weight_cache_columns(q_weight, scale_col, ind, bit: int = 8) -> np.ndarray:
    """linear.py:207 (bit 8) / :209-210 (bit 4): q_weight[:, ind].half() * scale_col.T — one fp16 multiply."""
    ind = np.asarray(ind, np.int64)
    q = q_weight[:, ind] if bit == 8 else unpack_i4(q_weight)[:, ind]
    return (q.astype(F16).astype(F32) * np.asarray(scale_col, F16).reshape(-1, 1).astype(F32)).astype(F16)


def rmsnorm(x: np.ndarray, w: np.ndarray, eps: float) -> np.ndarray:
    """mixlib.layernorm_forward_cuda (norm.py:21) — RMSNorm: fp16((x * rsqrt(mean(x^2)+eps)) * w).
    Floating-point: the variance sum is order dependent on the GPU; compare with a tolerance."""
    xf = np.asarray(x, F16).astype(F64)
    var = (xf * xf).mean(axis=1, keepdims=True)
    rstd = (1.0 / np.sqrt(var + eps)).astype(F32)
    return ((xf.astype(F32) * rstd) * np.asarray(w, F16).astype(F32)[None, :]).astype(F16)


def rmsnorm_extract_outliers(x, w, eps, ind, bit: int = 8):
    """mixlib.layernorm_forward_cuda_extract_outliers[_int4] (norm.py:25-33):
    returns (out, activation_outliers, q_x, x_scale); `out` has the ind columns zeroed (oracle choice,
    the only one consistent with linear.py:203-205 re-scanning `out` for NEW outliers)."""
    out = rmsnorm(x, w, eps)
    ao = extract_outliers_and_set_to_zeros(ind, out) if len(ind) else np.zeros((out.shape[0], 0), F16)
    q_x, xs = find_row_scale(out, bit)
    return out, ao, q_x, xs


def outlier_gemm_f32(act_outliers: np.ndarray, weight_cache: np.ndarray) -> np.ndarray:
    """fp16 x fp16 products accumulated in fp32-or-better: what torch.mm(fp16) / tcgen05 kind::f16 compute
    before the final rounding (linear.py:248)."""
    return (np.asarray(act_outliers, F16).astype(F64) @ np.asarray(weight_cache, F16).astype(F64).T).astype(F32)


# --------------------------------------------------------------------------- state carriers
class MixLibCacheOracle:
    """Cache.py:5-25: shared scratch; sigma = 6, stop = 2 by default."""

    def __init__(self, inputdim: int = 1024, sigma: float = 6, bit: int = 8):
        self.x_scale = np.zeros((inputdim, 1), F16)
        self.sigma = F16(sigma)
        self.ind = None
        self.new_ind = None
        self.shape = None
        self.activation_outliers = None
        self.q_xcache = None
        self.bit = bit
        self.max_outliers = 256
        self.stop = 2


class MixLinearOracle:
    """MixLinear_GEMM (linear.py:26-376), arch != 9 branch (B200 reports major 10)."""

    def __init__(self, q_weight, scale_col, bit=8, bias=None, cache=None, ind=None, weight_cache=None):
        self.q_weight = q_weight
        self.scale_col = np.asarray(scale_col, F16).reshape(1, -1)
        self.bit = bit
        self.bias = bias
        self.cache = cache
        self.out_features = q_weight.shape[0]
        self.in_features = q_weight.shape[1] * (2 if bit == 4 else 1)
        self.ind = np.zeros((0,), np.int32) if ind is None else np.asarray(ind, np.int32)
        self.weight_cache = weight_cache
        self.cnt = 0
        self.add_outliers = True
        self.forward_without_precondition_len = -1 if bit == 8 else len(self.ind)
        self.sigma = F16(cache.sigma) if cache is not None else F16(6)

    @classmethod
    def from_linear(cls, weight, bit=8, bias=None, cache=None, layer_scales=None, fp_features_num=128):
        if bit == 8:
            q, s = quant_weight_w8(weight)
            return cls(q, s, 8, bias, cache)
        q, s, wc, ind = quant_weight_w4(weight, layer_scales, fp_features_num)
        return cls(q, s, 4, bias, cache, ind=ind, weight_cache=wc)

    def _gemm(self, cache, M, act):
        if len(self.ind):
            outl32 = outlier_gemm_f32(cache.activation_outliers, self.weight_cache)
            outl = outl32.astype(F16)  # torch.mm returns fp16 (linear.py:248)
        else:
            outl = None
        if getattr(self, "_w_chunks", None) is None or self._w_chunks_src is not self.q_weight:
            qw = self.q_weight if self.bit == 8 else unpack_i4(self.q_weight)
            self._w_chunks, self._w_chunks_src = weights_f32_chunks(qw), self.q_weight       # converted once per module
        y = dequantize(gemm_i8(cache.q_xcache, None, self._w_chunks), cache.x_scale[:M], self.scale_col, outl=outl, act=act)
        if self.bias is not None:  # linear.py:284-285: fp16 in-place add after the kernel
            y = (y.astype(F32) + np.asarray(self.bias, F16).astype(F32)[None, :]).astype(F16)
        return y

    def forward(self, x, cache=None, unfused=False):
        """linear.py:165-289.  x is modified in place exactly where the reference modifies it."""
        if cache is None:
            cache = self.cache
        cache.shape = x.shape[:-1] + (self.out_features,)
        inputs = x.reshape(-1, x.shape[-1])
        M = inputs.shape[0]
        if unfused:
            if len(self.ind):
                cache.activation_outliers = extract_outliers_and_set_to_zeros(self.ind, inputs)
            cache.q_xcache, xs = find_row_scale(inputs, self.bit)
            cache.x_scale[:M] = xs
        cache.ind = self.ind
        if self.add_outliers:
            qmax = 2 ** (self.bit - 1) - 1
            thr = F16(F32(self.sigma) / F32(qmax))  # fp16 tensor / python int -> fp16 (linear.py:201)
            if cache.x_scale[:M].max() > thr:
                ind = find_outliers(inputs, self.sigma)
                cache.new_ind = ind
                ao = extract_outliers_and_set_to_zeros(ind, inputs)
                wc = weight_cache_columns(self.q_weight, self.scale_col, ind, self.bit)
                if len(self.ind) == 0:
                    cache.activation_outliers = ao
                    self.weight_cache = wc
                else:
                    cache.activation_outliers = np.hstack((cache.activation_outliers, ao))
                    self.weight_cache = np.hstack((self.weight_cache, wc))
                self.ind = np.hstack((self.ind, ind)).astype(np.int32)
                cache.ind = self.ind
                cache.q_xcache, xs = find_row_scale(inputs, self.bit)
                cache.x_scale[:M] = xs
            self.cnt += 1
            if self.cnt >= self.cache.stop or len(self.ind) > 128:
                self.add_outliers = False
        return self._gemm(cache, M, 0).reshape(cache.shape)

    def forward_without_precondition_fused_silu(self, x, cache):
        """linear.py:291-376: gate_proj — reuses cache.q_xcache / activation_outliers / ind of up_proj."""
        inputs = x.reshape(-1, x.shape[-1])
        M = inputs.shape[0]
        if self.forward_without_precondition_len != len(cache.ind):
            if len(cache.ind):
                ind = cache.new_ind
                wc = weight_cache_columns(self.q_weight, self.scale_col, ind, self.bit)
                self.weight_cache = wc if len(self.ind) == 0 else np.hstack((self.weight_cache, wc))
                self.ind = cache.ind
                self.forward_without_precondition_len = len(self.ind)
        if self.bit == 4 and len(self.ind) == 0:
            raise RuntimeError("int4 mod should have outliers !")
        return self._gemm(cache, M, 1).reshape(cache.shape)


def mixgemm_sample(a, q_w, scale_col, ind):
    """models/sample.py:5-12 — the algorithm in one screen (no relu): used as a cross-check of forward()."""
    a = np.asarray(a, F16).copy()
    afp = extract_outliers_and_set_to_zeros(ind, a)
    bfp = weight_cache_columns(q_w, scale_col, ind, 8)
    q_x, xs = find_row_scale(a, 8)
    outl = outlier_gemm_f32(afp, bfp).astype(F16) if len(ind) else None
    return dequantize(gemm_i8(q_x, q_w), xs, scale_col, outl=outl)


def tp_exchange(partials, residual=None) -> np.ndarray:
    """The tensor-parallel exchange of the row-parallel Linears (no counterpart in the reference, which is single-GPU:
    models/base.py:196-225 only places layers): h = fp16( fp16(sum over ranks of partial) + residual ), the sum taken in fp32
    in RANK ORDER (so every rank gets the same bits) and rounded to fp16 once — what an fp16 all-reduce returns — before the
    decoder's residual add, a separate fp16 op.  mixq_b200/csrc/mixq_kernels.cu: allreduce_residual_kernel."""
    acc = np.zeros(partials[0].shape, F32)
    for part in partials:
        acc = (acc + part.astype(F32)).astype(F32)
    y = acc.astype(F16)
    if residual is not None:
        y = (y.astype(F32) + residual.astype(F32)).astype(F16)
    return y


def linear_fp32(x, weight, bias=None):
    """The un-quantised fp32 Linear MixLinear replaces: ground truth for the error budget."""
    y = np.asarray(x).astype(F32) @ np.asarray(weight).astype(F32).T
    if bias is not None:
        y = y + np.asarray(bias).astype(F32)[None, :]
    return y

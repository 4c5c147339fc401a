"""`mixlib` — the reference's native operator module, re-implemented on libmixq_sm100.so.

The reference (`/root/reference/mixquant/modules/linear.py`, `fused/norm.py`) does `import mixlib` and calls
twelve functions that live in an un-vendored CUDA extension (github.com/Qcompiler/QComplier, quantkernel).
This module presents the same twelve names with the same argument order, allocation behaviour (outputs are
allocated here with torch and returned) and error behaviour (Python exception on failure), each forwarding
to one `extern "C"` symbol of include/mixq.h through ctypes.  To run the reference's own Python on top of
it: `sys.modules["mixlib"] = mixq_b200.mixlib`.

All tensors must be CUDA tensors on the current device; work is enqueued on torch's current stream.
There is no CPU path: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

__all__ = [
    "FindRowScale", "ExtractOutliersAndSetToZeros", "int8FusedDequantize", "int8FusedDequantizeSilu",
    "int4FusedDequantize", "int4FusedDequantizeSilu", "gemm", "dequantizeInt8", "dequantizeInt8Silu",
    "unpack_int4_to_fp16", "layernorm_forward_cuda", "layernorm_forward_cuda_extract_outliers",
    "layernorm_forward_cuda_extract_outliers_int4",
]


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MixqError("mixlib: expected a CUDA tensor (there is no CPU path)")


def _contig(t, what):
    if not t.is_contiguous():
        raise _lib.MixqError(f"mixlib: {what} must be contiguous")
    return t


def FindRowScale(x, x_scale, M, K, bit=8):
    """linear.py:190-193, :221 — writes x_scale[:M], returns q_x int8 [M,K]."""
    _need_cuda(x, x_scale)
    _contig(x, "x")
    q_x = torch.empty((M, K), dtype=torch.int8, device=x.device)
    _lib.check(_lib.load().mixq_find_row_scale(_p(x), _p(x_scale), _p(q_x), M, K, bit, _stream()), "FindRowScale")
    return q_x


def ExtractOutliersAndSetToZeros(ind, x):
    """linear.py:189, :205 — returns x[:, ind] (fp16 [M,n]) and zeroes those columns of x in place."""
    _need_cuda(ind, x)
    _contig(x, "x")
    x2 = x.view(-1, x.shape[-1])
    M, K = x2.shape
    n = int(ind.shape[0])
    out = torch.empty((M, n), dtype=torch.float16, device=x.device)
    if n:
        ind = ind.to(torch.int32).contiguous()
        _lib.check(_lib.load().mixq_extract_outliers_and_set_to_zeros(_p(ind), n, _p(x2), _p(out), n, M, K, _stream()),
                   "ExtractOutliersAndSetToZeros")
    return out


def _outl_args(outl, M, N):
    """The reference passes either torch.mm's [M,N] result or the big cache.zeros buffer (a zero addend)."""
    if outl is None:
        return _p(None), 0
    _need_cuda(outl)
    if outl.dim() != 2 or outl.shape[0] < M or outl.shape[1] < N or outl.stride(1) != 1:
        raise _lib.MixqError("mixlib: outliers tensor must be [>=M, >=N] with unit column stride")
    return _p(outl), int(outl.stride(0))


def _fused(name, sym, q_x, q_w, x_scale, scale_col, outl, M, N, K, act):
    _need_cuda(q_x, q_w, x_scale, scale_col)
    y = torch.empty((M, N), dtype=torch.float16, device=q_x.device)
    po, ld = _outl_args(outl, M, N)
    fn = getattr(_lib.load(), sym)
    _lib.check(fn(_p(_contig(q_x, "q_x")), _p(_contig(q_w, "q_w")), _p(x_scale), _p(scale_col), po, ld, _p(y), M, N, K,
                  act, _stream()), name)
    return y


def int8FusedDequantize(q_x, q_w, x_scale, scale_col, outl, M, N, K):
    """linear.py:251-256, :268-273."""
    return _fused("int8FusedDequantize", "mixq_int8_fused_dequantize", q_x, q_w, x_scale, scale_col, outl, M, N, K, 0)


def int8FusedDequantizeSilu(q_x, q_w, x_scale, scale_col, outl, M, N, K):
    """linear.py:337-351."""
    return _fused("int8FusedDequantizeSilu", "mixq_int8_fused_dequantize", q_x, q_w, x_scale, scale_col, outl, M, N, K, 1)


def int4FusedDequantize(q_x, q_w, x_scale, scale_col, outl, M, N, Khalf):
    """linear.py:259-265, :278-283 — the last argument is K/2, as the reference passes it."""
    return _fused("int4FusedDequantize", "mixq_int4_fused_dequantize", q_x, q_w, x_scale, scale_col, outl, M, N,
                  2 * Khalf, 0)


def int4FusedDequantizeSilu(q_x, q_w, x_scale, scale_col, outl, M, N, Khalf):
    """linear.py:360-366."""
    return _fused("int4FusedDequantizeSilu", "mixq_int4_fused_dequantize", q_x, q_w, x_scale, scale_col, outl, M, N,
                  2 * Khalf, 1)


def gemm(q_x, q_w, M, N, K):
    """linear.py:235, :321 — int8 x int8 -> int32 [M,N]."""
    _need_cuda(q_x, q_w)
    y = torch.empty((M, N), dtype=torch.int32, device=q_x.device)
    _lib.check(_lib.load().mixq_gemm_i8(_p(_contig(q_x, "q_x")), _p(_contig(q_w, "q_w")), _p(y), M, N, K, _stream()), "gemm")
    return y


def _dequant(name, y_i32, x_scale, scale_col, outl, bit, M, N, act):
    _need_cuda(y_i32, x_scale, scale_col)
    if bit != 8:
        raise _lib.MixqError("mixlib.dequantizeInt8: bit must be 8")
    y = torch.empty((M, N), dtype=torch.float16, device=y_i32.device)
    po, ld = _outl_args(outl, M, N)
    _lib.check(_lib.load().mixq_dequantize_int8(_p(_contig(y_i32, "y")), _p(x_scale), _p(scale_col), po, ld, _p(y), M, N,
                                                act, _stream()), name)
    return y


def dequantizeInt8(y, x_scale, scale_col, outl, bit, M, N):
    """linear.py:238, :241."""
    return _dequant("dequantizeInt8", y, x_scale, scale_col, outl, bit, M, N, 0)


def dequantizeInt8Silu(y, x_scale, scale_col, outl, bit, M, N):
    """linear.py:324, :327."""
    return _dequant("dequantizeInt8Silu", y, x_scale, scale_col, outl, bit, M, N, 1)


def unpack_int4_to_fp16(q_w, ind):
    """linear.py:20-22 — q_w uint8 [N,K/2], returns the sign-extended nibbles of columns `ind`, fp16 [N,n]."""
    _need_cuda(q_w, ind)
    N, Kh = q_w.shape
    n = int(ind.shape[0])
    out = torch.empty((N, n), dtype=torch.float16, device=q_w.device)
    if n:
        ind = ind.to(torch.int32).contiguous()
        _lib.check(_lib.load().mixq_unpack_int4_to_fp16(_p(_contig(q_w, "q_w")), _p(ind), n, _p(out), n, N, 2 * Kh, _stream()),
                   "unpack_int4_to_fp16")
    return out


def layernorm_forward_cuda(x, w, out, eps):
    """fused/norm.py:21 — RMSNorm into `out`."""
    _need_cuda(x, w, out)
    K = x.shape[-1]
    M = x.numel() // K
    _lib.check(_lib.load().mixq_rmsnorm(_p(_contig(x, "x")), _p(w), _p(_contig(out, "out")), float(eps), M, K, _stream()),
               "layernorm_forward_cuda")


def _norm_extract(name, x, w, out, eps, ind, x_scale, bit):
    _need_cuda(x, w, out, ind, x_scale)
    K = x.shape[-1]
    M = x.numel() // K
    n = int(ind.shape[0])
    ao = torch.empty((M, n), dtype=torch.float16, device=x.device)
    q_x = torch.empty((M, K), dtype=torch.int8, device=x.device)
    ind = ind.to(torch.int32).contiguous()
    _lib.check(_lib.load().mixq_rmsnorm_extract_outliers(_p(_contig(x, "x")), _p(w), _p(_contig(out, "out")), float(eps),
                                                         _p(ind), n, _p(x_scale), _p(ao), n, _p(q_x), M, K, bit,
                                                         _stream()), name)
    return ao, q_x


def layernorm_forward_cuda_extract_outliers(x, w, out, eps, ind, x_scale):
    """fused/norm.py:25-28 — returns (activation_outliers, q_x); writes `out` and x_scale."""
    return _norm_extract("layernorm_forward_cuda_extract_outliers", x, w, out, eps, ind, x_scale, 8)


def layernorm_forward_cuda_extract_outliers_int4(x, w, out, eps, ind, x_scale):
    """fused/norm.py:30-33."""
    return _norm_extract("layernorm_forward_cuda_extract_outliers_int4", x, w, out, eps, ind, x_scale, 4)

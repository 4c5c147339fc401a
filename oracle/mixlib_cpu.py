"""`mixlib`-shaped CPU module built on the oracle — TEST INFRASTRUCTURE ONLY.

tests/golden/make_golden.py installs this as `sys.modules["mixlib"]` so that the reference's own
/root/reference/mixquant/modules/linear.py (and fused/norm.py) run UNMODIFIED on CPU tensors: the
control flow is then the reference's, the per-kernel arithmetic is the oracle's (the real mixlib is an
un-vendored CUDA extension, see oracle/mixq_oracle.py).  Signatures follow the reference call sites
(linear.py:22,189,190,205,221,235-283,321-366; norm.py:21,25,30).
"""
from __future__ import annotations

import numpy as np
import torch

from . import mixq_oracle as O


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().numpy()


def _t(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a))


def FindRowScale(x, x_scale, M, K, bit=8):
    q, xs = O.find_row_scale(_np(x)[:M, :K], bit)
    x_scale[:M] = _t(xs)
    return _t(q)


def ExtractOutliersAndSetToZeros(ind, x):
    idx = ind.long()
    out = x[:, idx].clone()
    x[:, idx] = 0  # in place on the caller's tensor, like the CUDA kernel
    return out


def _outl(outl, M, N):
    return None if outl is None else _np(outl)[:M, :N]


def int8FusedDequantize(q_x, q_w, x_scale, scale_col, outl, M, N, K):
    return _t(O.int8_fused_dequantize(_np(q_x), _np(q_w), _np(x_scale)[:M], _np(scale_col), _outl(outl, M, N), 0))


def int8FusedDequantizeSilu(q_x, q_w, x_scale, scale_col, outl, M, N, K):
    return _t(O.int8_fused_dequantize(_np(q_x), _np(q_w), _np(x_scale)[:M], _np(scale_col), _outl(outl, M, N), 1))


def int4FusedDequantize(q_x, q_w, x_scale, scale_col, outl, M, N, Khalf):
    return _t(O.int4_fused_dequantize(_np(q_x), _np(q_w), _np(x_scale)[:M], _np(scale_col), _outl(outl, M, N), 0))


def int4FusedDequantizeSilu(q_x, q_w, x_scale, scale_col, outl, M, N, Khalf):
    return _t(O.int4_fused_dequantize(_np(q_x), _np(q_w), _np(x_scale)[:M], _np(scale_col), _outl(outl, M, N), 1))


def gemm(q_x, q_w, M, N, K):
    return _t(O.gemm_i8(_np(q_x), _np(q_w)))


def dequantizeInt8(y, x_scale, scale_col, outl, bit, M, N):
    return _t(O.dequantize(_np(y), _np(x_scale)[:M], _np(scale_col), outl=_outl(outl, M, N), act=0))


def dequantizeInt8Silu(y, x_scale, scale_col, outl, bit, M, N):
    return _t(O.dequantize(_np(y), _np(x_scale)[:M], _np(scale_col), outl=_outl(outl, M, N), act=1))


def unpack_int4_to_fp16(q_w, ind):
    return _t(O.unpack_int4_to_fp16(_np(q_w), _np(ind)))


def layernorm_forward_cuda(x, w, out, eps):
    shp = x.shape
    out.copy_(_t(O.rmsnorm(_np(x).reshape(-1, shp[-1]), _np(w), eps)).reshape(shp))


def _norm_extract(x, w, out, eps, ind, x_scale, bit):
    shp = x.shape
    o, ao, q, xs = O.rmsnorm_extract_outliers(_np(x).reshape(-1, shp[-1]), _np(w), eps, _np(ind), bit)
    out.copy_(_t(o).reshape(shp))
    x_scale[: q.shape[0]] = _t(xs)
    return _t(ao), _t(q)


def layernorm_forward_cuda_extract_outliers(x, w, out, eps, ind, x_scale):
    return _norm_extract(x, w, out, eps, ind, x_scale, 8)


def layernorm_forward_cuda_extract_outliers_int4(x, w, out, eps, ind, x_scale):
    return _norm_extract(x, w, out, eps, ind, x_scale, 4)

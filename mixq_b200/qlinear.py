"""MixedQLinear — the QUIK-style Linear of the reference (W4A4 / W8A8 with asymmetric per-token activations, a zero-point
correction term and a static set of full-precision columns) on libmixq_sm100.

Mirror of /root/reference/mixquant/modules/qlinear.py:22-211 (`SharedQuantizedInput`, `MixedQLinear` with the same constructor,
buffers — weights_scales [N,1], int_weight (uint8 [N, n_int/2] packed nibbles | int8 [N, n_int]), int_indices / fp_indices
(torch.long), fp_weight [N, n_fp], reduced_w [1, N] — `from_linear` and `forward`).  The reference calls the un-vendored `quik`
CUDA extension (qlinear.py:6-7); here the same steps are this library's kernels:

    quik.asymmetric.quantize(x, int_indices, fp_indices, bits)      -> mixq_quik_quantize (one launch: min/max, meta, q, fp_x)
    F.linear(fp_x, fp_weight, bias)                                  -> torch (cuBLAS), as in the reference (qlinear.py:129)
    quik.matmul.int4Matmul / int8Matmul + quik.asymmetric.dequantize -> mixq_quik_addend + mixq_int4/int8_fused_dequantize:
        the packed-nibble tcgen05 GEMM with the dequant epilogue  y = fp16(acc * scale[m] * ws[n] + addend[m,n]),
        addend = (zero[m] + 2^(bits-1) scale[m]) * reduced_w[n] + fp_result[m,n]

`qint_x` is kept one value per byte (int8 in [-2^(b-1), 2^(b-1)-1]) instead of QUIK's packed activations: Blackwell has no int4
MMA, the GEMM consumes int8 activations.  The symmetric branch (qlinear.py:96-106) is implemented with the MixQ row quantiser
only for bits == 4 where its scale convention (rowabsmax / 4, see oracle/quik_oracle.py) is meaningful — it raises otherwise.
No CPU path.  Parity: oracle/quik_oracle.py (unpinned: `quik` is not in the reference tree).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def two_compl(x: torch.Tensor, bits: int) -> torch.Tensor:
    return torch.where(x < 0, 2 ** bits + x, x)


def pack_to_i4(X: torch.Tensor):
    X_i8 = two_compl(X.to(dtype=torch.int8), 4).to(torch.uint8)
    return X_i8[:, 0::2] | (X_i8[:, 1::2] << 4)


class SharedQuantizedInput:
    """qlinear.py:22-38: q/k/v (or up/gate) share one quantised activation; the last consumer of the group clears it."""

    def __init__(self, group_size):
        self.qint_x = None
        self.fp_x = None
        self.qscale_x = None
        self.meta = None
        self.group_size = group_size
        self.cur_group_elem = 0

    def finish(self):
        self.cur_group_elem += 1
        if self.cur_group_elem == self.group_size:
            self.qint_x = None
            self.qscale_x = None
            self.meta = None
            self.fp_x = None
            self.cur_group_elem = 0


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class MixedQLinear(torch.nn.Module):
    def __init__(self, in_features, out_features, shared_input=None, fp_features_num=0, symm=False, bits=4,
                 dtype=torch.float16, dev="cuda"):
        super().__init__()
        if bits not in (4, 8):
            raise ValueError("bits must be 4 or 8")
        if symm:
            raise NotImplementedError("symmetric QUIK activations (qlinear.py:96-106) are not on the benchmarked path "
                                      "(quantize_QUIK builds asymmetric modules); use MixLinear_GEMM for symmetric W4A4")
        self.fp_features_num = fp_features_num
        self.int_features_num = in_features - fp_features_num
        self.in_features = in_features
        self.out_features = out_features
        self.symmetric = symm
        self.bits = bits
        self.shared_input = shared_input
        self.dtype = dtype
        self.register_buffer("weights_scales", torch.zeros((out_features, 1), dtype=dtype, device=dev))
        if bits == 4:
            self.register_buffer("int_weight", torch.zeros((out_features, self.int_features_num // 2), dtype=torch.uint8, device=dev))
        else:
            self.register_buffer("int_weight", torch.zeros((out_features, self.int_features_num), dtype=torch.int8, device=dev))
        self.bias = None
        self.register_buffer("int_indices", torch.zeros((self.int_features_num,), dtype=torch.long, device=dev))
        self.register_buffer("fp_indices", torch.zeros((self.fp_features_num,), dtype=torch.long, device=dev))
        if self.fp_features_num > 0:
            self.register_buffer("fp_weight", torch.zeros((out_features, self.fp_features_num), dtype=dtype, device=dev))
        self.register_buffer("reduced_w", torch.zeros((1, out_features), dtype=dtype, device=dev))

    @torch.no_grad()
    def forward(self, x):
        if self.int_features_num <= 0:
            return torch.nn.functional.linear(x, self.fp_weight, self.bias)
        if not x.is_cuda:
            raise _lib.MixqError("MixedQLinear.forward needs CUDA tensors: there is no CPU path")
        lib = _lib.load()
        shared = self.shared_input if self.shared_input is not None else SharedQuantizedInput(1)
        out_shape = x.shape[:-1] + (self.out_features,)
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M, K = x2.shape
        N, n_int, n_fp = self.out_features, self.int_features_num, self.fp_features_num
        if shared.qint_x is None:
            # qlinear.py:117-120: (qint_x, meta, fp_x) = quik.asymmetric.quantize(x, int_indices, fp_indices, bits)
            q = torch.empty((M, n_int), dtype=torch.int8, device=x.device)
            meta = torch.empty((2, M), dtype=torch.float16, device=x.device)
            fp_x = torch.empty((M, n_fp), dtype=torch.float16, device=x.device) if n_fp else None
            _lib.check(lib.mixq_quik_quantize(x2.data_ptr(), self.int_indices.data_ptr(), n_int,
                                              self.fp_indices.data_ptr() if n_fp else 0, n_fp, self.bits, q.data_ptr(),
                                              meta.data_ptr(), fp_x.data_ptr() if n_fp else 0, M, K, _st()), "quik.asymmetric.quantize")
            shared.qint_x, shared.meta, shared.fp_x = q, meta, fp_x
        # qlinear.py:126-140: the full-precision part
        if n_fp > 0:
            fp_result = torch.nn.functional.linear(shared.fp_x, self.fp_weight, self.bias)
        elif self.bias is not None:
            fp_result = self.bias.repeat(M, 1)
        else:
            fp_result = None
        # qlinear.py:142-150: int matmul + asymmetric dequantize (zero-point correction through reduced_w)
        addend = torch.empty((M, N), dtype=torch.float16, device=x.device)
        _lib.check(lib.mixq_quik_addend(shared.meta.data_ptr(), self.reduced_w.data_ptr(), 0 if fp_result is None else fp_result.data_ptr(),
                                        N, addend.data_ptr(), M, N, self.bits, _st()), "quik addend")
        y = torch.empty((M, N), dtype=torch.float16, device=x.device)
        fn = lib.mixq_int4_fused_dequantize if self.bits == 4 else lib.mixq_int8_fused_dequantize
        _lib.check(fn(shared.qint_x.data_ptr(), self.int_weight.data_ptr(), shared.meta.data_ptr(), self.weights_scales.data_ptr(),
                      addend.data_ptr(), N, y.data_ptr(), M, N, n_int, 0, _st()), "quik int matmul + dequantize")
        shared.finish()
        return y.reshape(out_shape)

    @classmethod
    def from_linear(cls, module, weight_matrix, weights_scales=None, shared_input=None, fp_indices=None, symm=False, bits=4,
                    init_only=False, fp_features_num=256):
        """qlinear.py:153-211.  `weight_matrix` [N,K] fp16 (the GPTQ-updated weights in quantize_QUIK), `weights_scales` [N,1]."""
        if init_only:
            return cls(module.in_features, module.out_features, shared_input=None, fp_features_num=fp_features_num, symm=symm,
                       bits=bits, dtype=torch.float16)
        assert weights_scales is not None
        assert weights_scales.shape == (module.out_features, 1), "weights_scales should have shape (out_features, 1)"
        assert weight_matrix.shape == (module.out_features, module.in_features)
        assert (symm and bits == 4) or not symm, "Symmetric quantization with 8 bits is not supported"
        int_indices = torch.arange(module.in_features)
        if fp_indices is None or len(fp_indices) == 0:
            fp_indices = torch.tensor([], dtype=int_indices.dtype)
        else:
            fp_indices = fp_indices.to("cpu", torch.long)
            int_indices = int_indices[~torch.isin(int_indices, fp_indices)]
        assert torch.numel(int_indices) + torch.numel(fp_indices) == module.in_features, "There are some duplication in the fp_indices!"
        m = cls(module.in_features, module.out_features, shared_input, fp_features_num=torch.numel(fp_indices), symm=symm,
                bits=bits, dtype=weight_matrix.dtype)
        weight_matrix = weight_matrix.cuda()
        m.weights_scales.copy_(weights_scales.to(weight_matrix.dtype))
        q = (weight_matrix[:, int_indices] / weights_scales.to(weight_matrix.device)).round()
        if bits == 4:
            m.int_weight.copy_(pack_to_i4(q.to(torch.int8)).cpu())
        else:
            m.int_weight.copy_(q.to(torch.int8).cpu())
        if not symm:
            reduced_w = torch.sum(weight_matrix[:, int_indices].float(), dim=1, keepdim=True).to(weight_matrix.dtype)
            m.reduced_w.copy_(reduced_w.t().cpu())
        if module.bias is not None:
            m.bias = module.bias.detach().to(weight_matrix.device, weight_matrix.dtype)
        m.int_indices.copy_(int_indices)
        m.fp_indices.copy_(fp_indices)
        if m.fp_features_num > 0:
            m.fp_weight.copy_(weight_matrix[:, fp_indices].to(weight_matrix.dtype))
        return m

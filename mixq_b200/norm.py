"""FasterTransformerRMSNorm — RMSNorm fused with the NEXT MixLinear's activation prologue.

Mirror of /root/reference/mixquant/modules/fused/norm.py:6-39 (same constructor, `next_layer` wiring done
by LlamaMixQForCausalLM.fuse_layers, llama.py:20-22).  With a `next_layer`, one launch normalises, gathers
the next layer's outlier columns, computes the per-row scale and writes the int8 activations into the
shared MixLibCache (`activation_outliers`, `q_xcache`, `x_scale`); the returned fp16 `output` has the
outlier columns zeroed, exactly what the unfused path leaves behind (linear.py:189).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib


class FasterTransformerRMSNorm(nn.Module):
    def __init__(self, weight, eps=1e-6, cache=None):
        super().__init__()
        self.weight = weight.to(torch.float16)
        self.variance_epsilon = eps
        self.cache = cache
        self.next_layer = None

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda:
            raise _lib.MixqError("FasterTransformerRMSNorm needs CUDA tensors: there is no CPU path")
        if self.weight.device != x.device:
            self.weight = self.weight.to(x.device)
        x = x.contiguous()
        output = torch.empty_like(x)
        K = x.shape[-1]
        M = x.numel() // K
        lib = _lib.load()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if self.next_layer is None:
            _lib.check(lib.mixq_rmsnorm(x.data_ptr(), self.weight.data_ptr(), output.data_ptr(),
                                        float(self.variance_epsilon), M, K, stream), "layernorm_forward_cuda")
            return output
        nl = self.next_layer
        if nl.bit not in (4, 8):
            raise NotImplementedError
        cache = self.cache
        n = nl._n_ind
        ao = cache.ao_buffer(n)
        q_x = cache.q_x_buffer(M, K)
        _lib.check(lib.mixq_rmsnorm_extract_outliers(x.data_ptr(), self.weight.data_ptr(), output.data_ptr(),
                                                     float(self.variance_epsilon), nl._ind_buf.data_ptr(), n,
                                                     cache.x_scale.data_ptr(), ao.data_ptr(), ao.shape[1],
                                                     q_x.data_ptr(), M, K, nl.bit, stream),
                   "layernorm_forward_cuda_extract_outliers")
        cache.activation_outliers = ao[:M, :n]
        cache.q_xcache = q_x
        return output

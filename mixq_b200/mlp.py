"""MixLlamaMLP — call sequencing of the three MLP MixLinears around one shared quantised activation.

Mirror of /root/reference/mixquant/modules/fused/mlp.py:36-70: up_proj consumes what the preceding
FasterTransformerRMSNorm left in the cache, gate_proj re-uses the same q_x / outliers with a SiLU epilogue
(linear.py:291-376), `gate *= up`, down_proj runs unfused on the product.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib


class MixLlamaMLP(nn.Module):
    def __init__(self, gate_proj, down_proj, up_proj, MixGemmCache=None):
        super().__init__()
        self.down_proj_ = down_proj
        self.gate_proj_ = gate_proj
        self.up_proj_ = up_proj
        self.out_features = down_proj.out_features
        self.MLPCache = MixGemmCache

    @torch.no_grad()
    def forward(self, x, residual=None):
        up_output = self.up_proj_(x, self.MLPCache)
        gate_output = self.gate_proj_.forward_without_preconditionFusedSilu(x, self.MLPCache)
        # gate_output *= up_output   (mlp.py:64)
        _lib.check(_lib.load().mixq_mul_inplace(gate_output.data_ptr(), up_output.data_ptr(), gate_output.numel(),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mul_inplace")
        return self.down_proj_(gate_output, None, True, residual=residual)

"""QuantAttentionFused — the fused attention block of a MixQ Llama layer: W_pack -> RoPE -> KV cache -> attention -> o_proj.

Mirror of /root/reference/mixquant/modules/fused/attn.py:76-278 (same class name, constructor and `forward` signature, same
call sequencing of the two MixLinears: `W_pack(hidden_states)` in the reference's FUSED call mode — it consumes the
q_xcache / x_scale / activation_outliers the preceding FasterTransformerRMSNorm left in the MixLibCache, attn.py:219 — and
`o_proj(attn_output, None, True)` in UNFUSED mode, attn.py:263).

What differs behind the same names:
  * decode (q_len == 1) runs this library's kernel: RoPE (HF rotate_half) + KV-cache append + single-query attention in one
    launch (`mixq_rope_attention_decode`); once o_proj's outlier discovery has finished, the kernel also quantises its own
    output rows for o_proj (`mixq_rope_attention_decode_quant`), which then runs without an activation prologue;
  * the KV cache is owned by the module ([batch, kv_heads, max_seq_len, head_dim] per layer, allocated on first use) instead
    of an HF `Cache` object whose API changes between transformers releases (`get_usable_length` — attn.py:238 — is gone in
    transformers 5); `past_key_value` is accepted and returned untouched;
  * prefill / q_len > 1 takes the library path the reference takes (flash-attn there, torch SDPA here: attention math is
    outside the quantised hot path).
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


def rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


class QuantAttentionFused(nn.Module):
    def __init__(self, hidden_size, n_heads, n_kv_heads, qkv_layer, o_proj, dev, max_seq_len, use_alibi=False,
                 attention_shapes=None, MixGemmCache=None, layer_idx=None, rope_theta: float = 10000.0):
        super().__init__()
        if use_alibi:
            raise NotImplementedError("alibi attention (MPT / Falcon) is outside the Llama MixLinear path")
        self.layer_idx = layer_idx
        self.hidden_size = hidden_size
        self.num_heads = n_heads
        self.num_kv_heads = n_kv_heads
        self.head_dim = hidden_size // n_heads
        if self.head_dim * n_heads != hidden_size:
            raise ValueError(f"hidden_size must be divisible by num_heads (got `hidden_size`: {hidden_size} and `num_heads`: {n_heads}).")
        if self.head_dim not in (64, 128):
            raise ValueError("head_dim must be 64 or 128")
        self.max_position_embeddings = max_seq_len
        self.cache_batch_size = int(os.getenv("BATCH_SIZE", "1"))     # attn.py:95 (AutoForCausalLM.from_quantized sets it)
        self.W_pack = qkv_layer
        self.o_proj = o_proj
        self.MixGemmCache = MixGemmCache
        self.rope_theta = float(rope_theta)
        self.dev = dev
        self.k_cache = self.v_cache = None     # [batch, kv_heads, max_seq_len, head_dim], allocated by the first cached call
        self.start_pos = 0                     # tokens already in the cache

    # ------------------------------------------------------------------ KV cache
    def reset_cache(self):
        self.start_pos = 0

    def _ensure_cache(self, bsz, need, device):
        """Capacity grows geometrically up to max_seq_len (an HF DynamicCache grows too): a benchflops-style call sequence
        (fresh cache every forward) never holds more than a few positions."""
        if self.k_cache is not None and self.k_cache.shape[0] == bsz and self.k_cache.shape[2] >= need:
            return
        cap = 8
        while cap < need:
            cap *= 2
        cap = min(cap, self.max_position_embeddings)
        shape = (bsz, self.num_kv_heads, cap, self.head_dim)
        k = torch.zeros(shape, dtype=torch.float16, device=device)
        v = torch.zeros(shape, dtype=torch.float16, device=device)
        if self.k_cache is not None and self.k_cache.shape[0] == bsz and self.start_pos:
            k[:, :, :self.start_pos] = self.k_cache[:, :, :self.start_pos]
            v[:, :, :self.start_pos] = self.v_cache[:, :, :self.start_pos]
        else:
            self.start_pos = 0
        self.k_cache, self.v_cache = k, v

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, attention_mask=None, position_ids=None, past_key_value=None,
                output_attentions: bool = False, use_cache: bool = False, padding_mask=None, *args, **kwargs):
        bsz, q_len, _ = hidden_states.size()
        if not hidden_states.is_cuda:
            raise _lib.MixqError("QuantAttentionFused needs CUDA tensors: there is no CPU path")
        proj = self.W_pack(hidden_states)                 # attn.py:219, fused call mode: the preceding norm quantised hidden_states
        H, Hkv, D = self.num_heads, self.num_kv_heads, self.head_dim
        past = self.start_pos if use_cache else 0
        if use_cache:
            if self.start_pos + q_len > self.max_position_embeddings:
                raise _lib.MixqError("KV cache full: raise max_seq_len")
            self._ensure_cache(bsz, self.start_pos + q_len, hidden_states.device)
            past = self.start_pos
        lib = _lib.load()
        o = self.o_proj
        if q_len == 1:
            qkv = proj.reshape(bsz, (H + 2 * Hkv) * D)
            kc = self.k_cache.data_ptr() if use_cache else 0
            vc = self.v_cache.data_ptr() if use_cache else 0
            cap = self.k_cache.shape[2] if use_cache else 0
            if not o.add_outliers and o.cache is not None:
                # steady state: the attention kernel owns whole token rows -> it quantises them for o_proj itself
                oc = o.cache
                n = o._n_ind
                ao, q_x = oc.ao_buffer(n), oc.q_x_buffer(bsz, H * D)
                _lib.check(lib.mixq_rope_attention_decode_quant(qkv.data_ptr(), kc, vc, cap, past, 0, bsz, H, Hkv, D, self.rope_theta,
                                                                o._ind_buf.data_ptr(), n, ao.data_ptr(), ao.shape[1], q_x.data_ptr(),
                                                                oc.x_scale.data_ptr(), o.bit, self._stream()),
                           "rope_attention_decode_quant")
                oc.q_xcache, oc.activation_outliers = q_x, ao[:bsz, :n]
                attn_output = o.forward_quantized(bsz).reshape(bsz, 1, -1)
            else:
                out = torch.empty((bsz, H * D), dtype=torch.float16, device=hidden_states.device)
                _lib.check(lib.mixq_rope_attention_decode(qkv.data_ptr(), kc, vc, cap, past, out.data_ptr(), bsz, H, Hkv, D,
                                                          self.rope_theta, self._stream()), "rope_attention_decode")
                attn_output = o(out.reshape(bsz, 1, H * D), None, True)
        else:
            # prefill: library attention, as the reference (flash-attn, attn.py:256-258); fp16 tensors, causal
            xqkv = proj.view(bsz, q_len, H + 2 * Hkv, D)
            xq, xk, xv = xqkv[:, :, :H], xqkv[:, :, H:H + Hkv], xqkv[:, :, H + Hkv:]
            pos = torch.arange(past, past + q_len, device=proj.device, dtype=torch.float32)
            inv = 1.0 / (self.rope_theta ** (torch.arange(0, D, 2, device=proj.device, dtype=torch.float32) / D))
            ang = torch.outer(pos, inv)
            cos, sin = torch.cat((ang, ang), -1).cos()[None, :, None, :], torch.cat((ang, ang), -1).sin()[None, :, None, :]
            xq = (xq.float() * cos + rotate_half(xq.float()) * sin).half()
            xk = (xk.float() * cos + rotate_half(xk.float()) * sin).half()
            q, k, v = xq.transpose(1, 2), xk.transpose(1, 2), xv.transpose(1, 2)
            if use_cache:
                self.k_cache[:, :, past:past + q_len] = k
                self.v_cache[:, :, past:past + q_len] = v
                k, v = self.k_cache[:, :, :past + q_len], self.v_cache[:, :, :past + q_len]
            if Hkv != H:
                k, v = k.repeat_interleave(H // Hkv, 1), v.repeat_interleave(H // Hkv, 1)
            mask = None
            if past:
                i = torch.arange(q_len, device=proj.device)[:, None] + past
                j = torch.arange(past + q_len, device=proj.device)[None, :]
                mask = j <= i
            att = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, is_causal=(mask is None))
            attn_output = att.transpose(1, 2).reshape(bsz, q_len, self.hidden_size).contiguous()
            attn_output = o(attn_output, None, True)
        if use_cache:
            self.start_pos = past + q_len
        return attn_output, None, past_key_value

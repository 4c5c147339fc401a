"""CPU restatement of one Llama decode step built from MixLinearOracle — TEST INFRASTRUCTURE ONLY.

Follows the reference's fused call sequence: fused/norm.py:14-39 (RMSNorm + extract + quantise for
next_layer), fused/attn.py:206-278 (W_pack -> RoPE -> attention -> o_proj unfused), fused/mlp.py:57-70
(up_proj, gate_proj with SiLU on the shared q_x, gate *= up, down_proj unfused), HF LlamaDecoderLayer's
residual adds, final norm, fp16 lm_head.  With an empty KV cache and q_len = 1 (benchflops.py:124 never
passes past_key_values) the softmax is over one key, so attention returns v.
"""
from __future__ import annotations

import numpy as np

from . import mixq_oracle as O

F16, F32 = np.float16, np.float32


def h16_add(a, b):
    return (a.astype(F32) + b.astype(F32)).astype(F16)


def rope_rotate(x, pos, theta, D):
    """HF apply_rotary_pos_emb (rotate_half convention), fp32 math, one rounding to fp16.  x: [M, heads, D]."""
    i = np.arange(D // 2, dtype=F32)
    inv = theta ** (-2.0 * i / D)
    ang = np.float32(pos) * inv
    cos = np.concatenate([np.cos(ang), np.cos(ang)]).astype(F32)
    sin = np.concatenate([np.sin(ang), np.sin(ang)]).astype(F32)
    xf = x.astype(F32)
    rot = np.concatenate([-xf[..., D // 2:], xf[..., :D // 2]], -1)
    return (xf * cos + rot * sin).astype(F16)


def attention_decode(qkv, H, Hkv, D, theta, past_k=None, past_v=None):
    """qkv [M, (H+2Hkv)*D] fp16 -> [M, H*D]; past_k/past_v [M, Hkv, L, D] (already rotated) or None."""
    M = qkv.shape[0]
    q = qkv[:, :H * D].reshape(M, H, D)
    k = qkv[:, H * D:(H + Hkv) * D].reshape(M, Hkv, D)
    v = qkv[:, (H + Hkv) * D:].reshape(M, Hkv, D)
    L = 0 if past_k is None else past_k.shape[2]
    q = rope_rotate(q, L, theta, D)
    k = rope_rotate(k, L, theta, D)
    keys = k[:, :, None, :] if L == 0 else np.concatenate([past_k, k[:, :, None, :]], 2)
    vals = v[:, :, None, :] if L == 0 else np.concatenate([past_v, v[:, :, None, :]], 2)
    rep = H // Hkv
    keys = np.repeat(keys, rep, 1).astype(F32)
    vals = np.repeat(vals, rep, 1).astype(F32)
    s = np.einsum("mhd,mhld->mhl", q.astype(F32), keys) / np.sqrt(F32(D))
    s = s - s.max(-1, keepdims=True)
    p = np.exp(s)
    p = p / p.sum(-1, keepdims=True)
    o = np.einsum("mhl,mhld->mhd", p, vals)
    return o.astype(F16).reshape(M, H * D)


class LlamaLayerOracle:
    def __init__(self, ln1, ln2, w_pack, wo, wg, wu, wd, cache, bit=8, scales1=None, scales2=None, fp=128):
        mk = lambda W, b, s=None: O.MixLinearOracle.from_linear(W, b, cache=cache, layer_scales=s, fp_features_num=fp)
        self.ln1, self.ln2 = ln1, ln2
        self.W_pack, self.o_proj = mk(w_pack, bit, scales1), mk(wo, 8)
        self.gate, self.up, self.down = mk(wg, bit, scales2), mk(wu, bit, scales2), mk(wd, 8)


def decode_step(h, layers, cache, cfg, allreduce=None):
    """h fp16 [M, hidden] (embedded tokens) -> fp16 [M, hidden] after all layers.  cfg: dict(heads, kv_heads,
    head_dim, theta, eps) with LOCAL head counts; allreduce(x) sums row-parallel partial outputs over ranks."""
    M = h.shape[0]
    ar = allreduce or (lambda t: t)
    for L in layers:
        out, ao, q_x, xs = O.rmsnorm_extract_outliers(h, L.ln1, cfg["eps"], L.W_pack.ind, L.W_pack.bit)
        cache.activation_outliers, cache.q_xcache = ao, q_x
        cache.x_scale[:M] = xs
        qkv = L.W_pack.forward(out, cache)
        attn = attention_decode(qkv, cfg["heads"], cfg["kv_heads"], cfg["head_dim"], cfg["theta"])
        h = h16_add(ar(L.o_proj.forward(attn, None, True)), h)
        out, ao, q_x, xs = O.rmsnorm_extract_outliers(h, L.ln2, cfg["eps"], L.up.ind, L.up.bit)
        cache.activation_outliers, cache.q_xcache = ao, q_x
        cache.x_scale[:M] = xs
        up = L.up.forward(out, cache)
        gate = L.gate.forward_without_precondition_fused_silu(out, cache)
        gate = (gate.astype(F32) * up.astype(F32)).astype(F16)
        h = h16_add(ar(L.down.forward(gate, None, True)), h)
    return h

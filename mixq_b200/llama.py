"""Llama decode-step harness around MixLinear — the measurement definition of the headline metric.

What it reproduces: one timed iteration of /root/reference/benchflops.py:112-128, i.e. `model(inputs[B,1],
use_cache=True)` on a `from_quantized(..., fuse_layers=True)` Llama (call sequence: fused/attn.py:206-278,
fused/mlp.py:57-70, fused/norm.py:14-39, models/llama.py:9-22), with random-init weights of the named
architecture and synthetic tokens (no checkpoint / dataset is reachable).  benchflops never carries
past_key_values between iterations (benchflops.py:124), so every step is an independent [B,1] forward with
an empty KV cache; `past_len > 0` with a real cache is supported for completeness.

Per decoder layer, steady state (after the two outlier-discovery calls), 7 launches (5 with the SwiGLU pair fused):
    W_pack    = RMSNorm + extract + quantise + int8 GEMM + fp16 outlier GEMM + dequant     (1 launch)
    attention = RoPE + single-query attention                                              (1 launch)
    o_proj    = extract + quantise + GEMMs + dequant + residual add                        (1 launch)
    up_proj   = RMSNorm + ... (1), gate_proj = GEMMs on the shared q_x + SiLU (1), gate *= up (1)
    down_proj = extract + quantise + GEMMs + dequant + residual add                        (1 launch)
The whole step is captured into one CUDA graph.

Column-/row-parallel sharding over `world_size` ranks: W_pack / gate / up shard N (by
heads / intermediate channels), o_proj / down_proj shard K and all-reduce their fp16 output once.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import _lib
from .cache import MixLibCache
from .linear import MixLinear_GEMM
from .tp import all_reduce_sum, make_exchange, pack_qkv_shard, shard_cols, shard_rows


@dataclass
class LlamaConfig:
    name: str
    hidden: int
    intermediate: int
    layers: int
    heads: int
    kv_heads: int
    vocab: int
    rope_theta: float = 10000.0
    eps: float = 1e-5

    @property
    def head_dim(self):
        return self.hidden // self.heads


CONFIGS = {
    "llama-2-7b": LlamaConfig("llama-2-7b", 4096, 11008, 32, 32, 32, 32000),
    "llama-3-8b": LlamaConfig("llama-3-8b", 4096, 14336, 32, 32, 8, 128256, rope_theta=500000.0),
    "llama-2-70b": LlamaConfig("llama-2-70b", 8192, 28672, 80, 64, 8, 32000),
    # CPU-oracle-sized configuration for tests / smoke
    "tiny": LlamaConfig("tiny", 256, 512, 2, 4, 4, 512),
    "tiny8": LlamaConfig("tiny8", 1024, 2048, 2, 8, 8, 512),   # shards over 8 ranks
}


class _W:
    """Minimal stand-in for nn.Linear handed to MixLinear_GEMM.from_linear."""

    def __init__(self, weight):
        self.weight = weight
        self.bias = None
        self.out_features, self.in_features = weight.shape


def _forced(n: int, frac: float, gen: torch.Generator) -> torch.Tensor:
    k = max(1, int(round(frac * n))) if frac > 0 else 0
    return torch.randperm(n, generator=gen)[:k].sort().values


class LlamaDecoder:
    """Random-init Llama with every decoder-layer Linear replaced by MixLinear_GEMM (base.py:273-347)."""

    def __init__(self, cfg: LlamaConfig, batch: int, bit: int = 8, device="cuda", seed: int = 0,
                 outlier_frac: float = 0.01, rank: int = 0, world_size: int = 1, group=None, layers: int | None = None):
        self.cfg, self.batch, self.bit, self.device = cfg, batch, bit, device
        self.rank, self.world, self.group = rank, world_size, group
        self.n_layers = cfg.layers if layers is None else layers
        H, I, D = cfg.hidden, cfg.intermediate, cfg.head_dim
        if cfg.heads % world_size or cfg.kv_heads % world_size or I % world_size:
            raise ValueError("heads, kv_heads and intermediate must divide by world_size")
        self.h_loc, self.kv_loc, self.i_loc = cfg.heads // world_size, cfg.kv_heads // world_size, I // world_size
        self.cache = MixLibCache(inputdim=batch, sigma=6, bit=bit, device=device)
        gen = torch.Generator(device="cpu").manual_seed(seed)          # same on every rank
        dgen = torch.Generator(device=device).manual_seed(seed)        # same on every rank: full weights, then shard
        f16 = torch.float16

        def rand_w(n, k, row_boost=None, forced_in=0.0):
            # Synthetic statistics modelled on real LLM activations: inlier activations O(0.35) with ~1 % "massive"
            # channels.  Linears fed by a normed x whose forced channels carry 20x the inlier magnitude get their
            # weights scaled down accordingly (forced_in); boosted rows (x40) make ~1 % of the OUTPUT channels massive,
            # which is what the next unfused Linear (o_proj / down_proj) then discovers as its outlier columns.
            std = 0.35 / (k * (1.0 + 399.0 * forced_in)) ** 0.5
            w = torch.randn((n, k), generator=dgen, device=device, dtype=torch.float32) * std
            if row_boost is not None and row_boost.numel():
                w[row_boost.to(device)] *= 40.0
            return w.to(f16)

        self.embed = torch.randn((cfg.vocab, H), generator=dgen, device=device, dtype=torch.float32).to(f16)
        self.layers = []
        eight_only = ("o_proj", "down_proj")  # utils/module.py:2: these stay 8-bit in 4-bit models (base.py:308-312)
        for li in range(self.n_layers):
            ln1 = torch.ones(H, dtype=f16)
            ln1[_forced(H, outlier_frac, gen)] = 20.0
            ln2 = torch.ones(H, dtype=f16)
            ln2[_forced(H, outlier_frac, gen)] = 20.0
            v_boost = _forced(cfg.kv_heads * D, outlier_frac, gen)
            up_boost = _forced(I, outlier_frac, gen)
            wq = rand_w(cfg.heads * D, H, None, outlier_frac)
            wk = rand_w(cfg.kv_heads * D, H, None, outlier_frac)
            wv = rand_w(cfg.kv_heads * D, H, v_boost, outlier_frac)
            wo = rand_w(H, cfg.heads * D)
            wg = rand_w(I, H, None, outlier_frac)
            wu = rand_w(I, H, up_boost, outlier_frac)
            wd = rand_w(H, I)
            r, w_ = rank, world_size
            w_pack = pack_qkv_shard(wq, wk, wv, r, w_)
            wo_s = shard_cols(wo, r, w_).contiguous()
            wd_s = shard_cols(wd, r, w_).contiguous()
            sl = lambda t, n: shard_rows(t, r, w_)
            mk = lambda w, nm, b, scales=None: MixLinear_GEMM.from_linear(
                _W(w), b, cache=self.cache, dev=device, name=f"L{li}.{nm}", layer_scales=scales)
            if bit == 4:
                # static outliers need per-input-channel activation scales (mixquant.py:201-208): synthetic ones that
                # single out the forced channels, padded by the largest-index channels
                s_h1 = torch.arange(H, dtype=torch.float32) * 1e-6 + (ln1.float() > 1) * 10
                s_h2 = torch.arange(H, dtype=torch.float32) * 1e-6 + (ln2.float() > 1) * 10
            layer = {
                "ln1": ln1.to(device), "ln2": ln2.to(device),
                "W_pack": mk(w_pack, "W_pack", bit, s_h1 if bit == 4 else None),
                "o_proj": mk(wo_s, "o_proj", 8),
                "gate_proj": mk(sl(wg, I).contiguous(), "gate_proj", bit, s_h2 if bit == 4 else None),
                "up_proj": mk(sl(wu, I).contiguous(), "up_proj", bit, s_h2 if bit == 4 else None),
                "down_proj": mk(wd_s, "down_proj", 8),
            }
            del wq, wk, wv, wo, wg, wu, wd, w_pack, wo_s, wd_s
            self.layers.append(layer)
        _ = eight_only
        self.norm_f = torch.ones(H, dtype=f16, device=device)
        self.lm_head = rand_w(cfg.vocab, H)   # fp16, never quantised (base.py:285-288 only walks decoder layers)
        self.discovered = False
        # one launch for up_proj + gate_proj + SiLU + gate*up (needs the 2-CTA kernel: M > 128)
        self.fuse_swiglu = batch > 128
        # attention + o_proj's activation prologue in one launch (MIXQ_FUSE_ATTN_QUANT=0: separate launches)
        self.fuse_attn_quant = os.environ.get("MIXQ_FUSE_ATTN_QUANT", "1") != "0"
        self.fuse_xchg_quant = os.environ.get("MIXQ_TP_FUSE_QUANT", "1") != "0"
        self.vocab_parallel = False
        self.graph = None
        self._static_tokens = None
        self._static_logits = None
        self.kv = None
        self.lib = _lib.load()
        # row-parallel exchange: ONE kernel of this library (all-reduce + residual) — through the NVSwitch (NVLS multicast) when
        # MIXQ_TP_EXCHANGE = auto (= push: reduce-scatter fused into the GEMM epilogue + finish kernel) | push-nomc | multicast | peer | nccl
        self.xchg = None
        kind = os.environ.get("MIXQ_TP_EXCHANGE", "auto")
        if world_size > 1 and kind != "nccl":
            self.xchg = make_exchange(batch, H, rank, world_size, group=group, device=device, kind=kind)

    def close(self):
        """Release the peer-exchange buffers (CUDA-IPC mappings + cudaMalloc'd partial / result buffers) and the graph."""
        self.graph = None
        if getattr(self, "xchg", None) is not None:
            self.xchg.close()
            self.xchg = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ pieces
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _attention(self, qkv, past_len=0, layer_idx=0):
        cfg = self.cfg
        M = qkv.shape[0]
        out = torch.empty((M, self.h_loc * cfg.head_dim), dtype=torch.float16, device=qkv.device)
        kc = vc = None
        cap = 0
        if self.kv is not None:
            kc, vc = self.kv[layer_idx]
            cap = kc.shape[2]
        _lib.check(self.lib.mixq_rope_attention_decode(qkv.data_ptr(), 0 if kc is None else kc.data_ptr(),
                                                       0 if vc is None else vc.data_ptr(), cap, past_len, out.data_ptr(),
                                                       M, self.h_loc, self.kv_loc, cfg.head_dim, cfg.rope_theta,
                                                       self._stream()), "rope_attention_decode")
        return out

    def _attention_sdpa(self, qkv, past_len, layer_idx):
        """Attention over a real KV cache through the LIBRARY path, as the reference does (flash-attn at fused/attn.py:256-258;
        torch SDPA here): RoPE at position past_len (HF rotate_half), cache append, single-query attention over past_len + 1
        keys.  Used for the KV-cache variant of the metric; attention math is outside the quantised hot path."""
        cfg = self.cfg
        M, H, Hkv, D = qkv.shape[0], self.h_loc, self.kv_loc, cfg.head_dim
        kc, vc = self.kv[layer_idx]
        x = qkv.view(M, H + 2 * Hkv, D)
        q, k, v = x[:, :H], x[:, H:H + Hkv], x[:, H + Hkv:]
        if past_len > 0:
            inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, D, 2, device=qkv.device, dtype=torch.float32) / D))
            ang = past_len * inv
            cos, sin = torch.cat((ang, ang)).cos(), torch.cat((ang, ang)).sin()
            rot = lambda t: torch.cat((-t[..., D // 2:], t[..., :D // 2]), -1)
            q = (q.float() * cos + rot(q.float()) * sin).half()
            k = (k.float() * cos + rot(k.float()) * sin).half()
        kc[:, :, past_len] = k
        vc[:, :, past_len] = v
        out = torch.nn.functional.scaled_dot_product_attention(q.unsqueeze(2), kc[:, :, :past_len + 1], vc[:, :, :past_len + 1],
                                                               enable_gqa=(H != Hkv))
        return out.reshape(M, H * D)

    def _attention_quant(self, qkv, lin, past_len=0, layer_idx=0):
        """Attention + the activation prologue of `lin` (o_proj) in ONE launch: the kernel owns whole token rows, so it
        gathers lin's outlier columns, takes the row abs-max and quantises its own output (linear.py:187-193 on the
        attention output); lin then runs in the reference's fused call mode (forward_quantized).  Steady state only."""
        cfg, cache = self.cfg, self.cache
        M, K = qkv.shape[0], self.h_loc * cfg.head_dim
        kc = vc = None
        cap = 0
        if self.kv is not None:
            kc, vc = self.kv[layer_idx]
            cap = kc.shape[2]
        n = lin._n_ind
        ao, q_x = cache.ao_buffer(n), cache.q_x_buffer(M, K)
        _lib.check(self.lib.mixq_rope_attention_decode_quant(
            qkv.data_ptr(), 0 if kc is None else kc.data_ptr(), 0 if vc is None else vc.data_ptr(), cap, past_len, 0,
            M, self.h_loc, self.kv_loc, cfg.head_dim, cfg.rope_theta, lin._ind_buf.data_ptr(), n, ao.data_ptr(), ao.shape[1],
            q_x.data_ptr(), cache.x_scale.data_ptr(), lin.bit, self._stream()), "rope_attention_decode_quant")
        cache.q_xcache = q_x
        cache.activation_outliers = ao[:M, :n]

    def _allreduce(self, t):
        return all_reduce_sum(t, self.group) if self.world > 1 else t

    def alloc_kv(self, capacity: int):
        cfg = self.cfg
        self.kv = [(torch.zeros((self.batch, self.kv_loc, capacity, cfg.head_dim), dtype=torch.float16, device=self.device),
                    torch.zeros((self.batch, self.kv_loc, capacity, cfg.head_dim), dtype=torch.float16, device=self.device))
                   for _ in range(self.n_layers)]

    # ------------------------------------------------------------------ one decode step
    @torch.no_grad()
    def step(self, tokens: torch.Tensor, past_len: int = 0) -> torch.Tensor:
        """tokens int64 [B,1] (or [B]) on the device -> logits fp16 [B, vocab]."""
        cfg = self.cfg
        h = torch.nn.functional.embedding(tokens.reshape(-1), self.embed)   # [B, H] fp16
        steady = self.discovered
        tp = self.world > 1
        M = h.shape[0]
        fused_x = tp and self.xchg is not None and getattr(self.xchg, "fused", False)
        # the exchange's finish kernel can run the next Linear's activation prologue on the rows it has just assembled
        xq = steady and fused_x and self.fuse_xchg_quant and getattr(self.xchg, "can_quantize", False)
        prequant = False          # q_x / x_scale / outliers of the coming norm-fused Linear are already in the cache

        def row_parallel(lin, x=None, quant=None):
            """o_proj / down_proj + the exchange (+ residual): x = fp16 activations, or None when a producer kernel has
            already quantised them into the cache."""
            nonlocal h
            if tp and self.xchg is not None:
                kw = {"push": self.xchg.push_targets()} if fused_x else {"out": self.xchg.next_partial()}
                if x is None:
                    lin.forward_quantized(M, **kw)
                else:
                    lin(x, None, True, **kw)
                h = self.xchg.reduce(h, quant=quant) if quant is not None else self.xchg.reduce(h)
            elif tp:
                h = h + self._allreduce(lin.forward_quantized(M) if x is None else lin(x, None, True))
            else:
                h = lin.forward_quantized(M, residual=h) if x is None else lin(x, None, True, residual=h)

        for li, L in enumerate(self.layers):
            if prequant:
                qkv = L["W_pack"].forward_quantized(M)
            elif steady:
                qkv = L["W_pack"].forward_norm_fused(h, L["ln1"], cfg.eps)
            else:
                qkv = self._norm_then_linear(h, L["ln1"], L["W_pack"])
            mlp_in = L["gate_proj"] if (steady and self.fuse_swiglu) else L["up_proj"]
            q_mlp = (L["ln2"], cfg.eps, mlp_in, self.cache) if xq else None
            if self.kv is not None and getattr(self, "kv_library_attention", False):
                row_parallel(L["o_proj"], self._attention_sdpa(qkv, past_len, li), q_mlp)
            elif steady and self.fuse_attn_quant:
                # attention quantises its own output rows for o_proj: o_proj has no activation prologue left
                self._attention_quant(qkv, L["o_proj"], past_len, li)
                row_parallel(L["o_proj"], None, q_mlp)
            else:
                row_parallel(L["o_proj"], self._attention(qkv, past_len, li), q_mlp)
            if xq and self.fuse_swiglu:
                gate = L["gate_proj"].forward_swiglu_quantized(L["up_proj"], M)
            elif steady and self.fuse_swiglu:
                gate = L["gate_proj"].forward_swiglu_fused(L["up_proj"], h, L["ln2"], cfg.eps)
            else:
                if xq:
                    up = L["up_proj"].forward_quantized(M)
                elif steady:
                    up = L["up_proj"].forward_norm_fused(h, L["ln2"], cfg.eps)
                else:
                    up = self._norm_then_linear(h, L["ln2"], L["up_proj"])
                gate = L["gate_proj"].forward_without_preconditionFusedSilu(h, self.cache)
                _lib.check(self.lib.mixq_mul_inplace(gate.data_ptr(), up.data_ptr(), gate.numel(), self._stream()), "mul")
            nxt = self.layers[li + 1] if li + 1 < len(self.layers) else None
            prequant = xq and nxt is not None
            row_parallel(L["down_proj"], gate, (nxt["ln1"], cfg.eps, nxt["W_pack"], self.cache) if prequant else None)
            if li == 0 and getattr(self, "probe_layer0", False):
                self.hidden_after_layer0 = h.clone()     # parity probe (bench.py tp_parity): the residual stream after layer 0
        hn = torch.empty_like(h)
        _lib.check(self.lib.mixq_rmsnorm(h.data_ptr(), self.norm_f.data_ptr(), hn.data_ptr(), cfg.eps, h.shape[0],
                                         cfg.hidden, self._stream()), "final norm")
        return torch.matmul(hn, self.lm_head.t())

    # ------------------------------------------------------------------ vocab-parallel lm_head (tensor parallel only)
    def shard_lm_head(self):
        """Keep only this rank's vocab / world rows of the fp16 lm_head: step() then returns the logits shard [B, vocab / world]
        (columns rank * v .. (rank + 1) * v of the full logits) — the Megatron vocab-parallel head; `gather_logits` rebuilds the
        full tensor where somebody needs it, `argmax` picks the next token without moving the logits."""
        if self.world == 1 or self.vocab_parallel:
            return
        if self.cfg.vocab % self.world:
            raise ValueError("vocab must be divisible by the tensor-parallel world size")
        v = self.cfg.vocab // self.world
        self.lm_head = self.lm_head[self.rank * v:(self.rank + 1) * v].contiguous()
        self.vocab_parallel = True
        self.graph = None          # a captured step still holds the full head

    def gather_logits(self, local: torch.Tensor) -> torch.Tensor:
        if not self.vocab_parallel:
            return local
        import torch.distributed as dist
        parts = [torch.empty_like(local) for _ in range(self.world)]
        dist.all_gather(parts, local.contiguous(), group=self.group)
        return torch.cat(parts, dim=-1)

    def argmax(self, logits: torch.Tensor) -> torch.Tensor:
        """Next-token ids [B] from step()'s output: plain argmax, or — vocab-parallel — the best of the ranks' local maxima
        (ties go to the lower index, like torch.argmax over the full row)."""
        if not self.vocab_parallel:
            return torch.argmax(logits, dim=-1)
        from .tp import vocab_parallel_argmax
        return vocab_parallel_argmax(logits, self.rank, self.world, self.group)

    def _norm_then_linear(self, h, ln_w, lin):
        """Discovery-phase path: the reference's two-step sequence (norm.py:24-28 then linear.py:165, fused mode)."""
        from .norm import FasterTransformerRMSNorm
        norm = FasterTransformerRMSNorm(ln_w, self.cfg.eps, self.cache)
        norm.next_layer = lin
        out = norm(h)
        return lin(out, self.cache)

    @torch.no_grad()
    def discover(self, tokens: torch.Tensor, calls: int | None = None):
        """The reference's first `cache.stop` forwards: online outlier discovery with host syncs."""
        for _ in range(self.cache.stop if calls is None else calls):
            self._rank_barrier()      # discovery calls synchronise with the host: ranks drift apart between them
            self.step(tokens)
        # gate_proj never runs discovery itself: it follows up_proj's outlier set (linear.py:299-315)
        self.discovered = all(not L[k].add_outliers for L in self.layers for k in ("W_pack", "o_proj", "up_proj", "down_proj"))
        return self.discovered

    def _rank_barrier(self):
        """Process-group barrier before an exchange that follows a rank-asymmetric phase (host syncs, graph capture,
        rank-0-only work): the peer-exchange kernel spins on its peers' flags and reports a stall after a timeout."""
        if self.world > 1 and torch.distributed.is_initialized():
            torch.cuda.synchronize()
            torch.distributed.barrier(group=self.group)

    def capture(self, tokens: torch.Tensor, past_len: int = 0):
        """Capture the steady-state step into a CUDA graph (tokens are copied into a static buffer per replay)."""
        assert self.discovered, "run discover() first"
        self._rank_barrier()
        self._static_tokens = tokens.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self.step(self._static_tokens, past_len)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._rank_barrier()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._static_logits = self.step(self._static_tokens, past_len)
        self._rank_barrier()
        return self.graph

    def replay(self, tokens: torch.Tensor | None = None) -> torch.Tensor:
        if tokens is not None:
            self._static_tokens.copy_(tokens, non_blocking=True)
        self.graph.replay()
        return self._static_logits

    # ------------------------------------------------------------------ checkpoints (the reference's on-disk layout)
    def _split_qkv(self, w_pack: MixLinear_GEMM):
        """W_pack -> (q_proj, k_proj, v_proj): the inverse of models/llama.py:98-166 (per-row scales make it lossless)."""
        cfg, D = self.cfg, self.cfg.head_dim
        sizes = [cfg.heads * D, cfg.kv_heads * D, cfg.kv_heads * D]
        out, r0 = [], 0
        fpn = getattr(w_pack, "fp_features_num", 128)
        for n in sizes:
            m = MixLinear_GEMM(w_pack.in_features, n, False, w_pack.q_weight.device, w_pack.bit, cache=self.cache,
                               fp_features_num=fpn)
            m.q_weight.copy_(w_pack.q_weight[r0:r0 + n])
            m.scale_col.copy_(w_pack.scale_col[:, r0:r0 + n])
            if w_pack.bit == 4:
                m._wc_buf[:, :fpn] = w_pack._wc_buf[r0:r0 + n, :fpn]
                m._ind_buf[:fpn] = w_pack._ind_buf[:fpn]
                m._n_ind = fpn
            out.append(m)
            r0 += n
        return out

    def save_quantized(self, save_dir: str, safetensors: bool = False, shard_size="10GB"):
        """Write this model in the layout of the reference's save_quantized (models/base.py:78-119; HF Llama module names),
        so that `from_quantized` here — or the reference's own loader — reads it back.  Single rank only."""
        from . import checkpoint as ck
        if self.world != 1:
            raise ValueError("save_quantized writes the un-sharded model: use world_size 1")
        mods, extra = {}, {"model.embed_tokens.weight": self.embed, "model.norm.weight": self.norm_f, "lm_head.weight": self.lm_head}
        for i, L in enumerate(self.layers):
            p = f"model.layers.{i}."
            q, k, v = self._split_qkv(L["W_pack"])
            mods[p + "self_attn.q_proj"], mods[p + "self_attn.k_proj"], mods[p + "self_attn.v_proj"] = q, k, v
            mods[p + "self_attn.o_proj"] = L["o_proj"]
            mods[p + "mlp.gate_proj"], mods[p + "mlp.up_proj"], mods[p + "mlp.down_proj"] = L["gate_proj"], L["up_proj"], L["down_proj"]
            extra[p + "input_layernorm.weight"] = L["ln1"]
            extra[p + "post_attention_layernorm.weight"] = L["ln2"]
        files = list(ck.save_quantized(save_dir, mods, {"w_bit": self.bit, "version": "MIX", "q_group_size": 128}, extra=extra,
                                       safetensors=safetensors, shard_size=shard_size))
        # the HF config.json the reference's loader (base.py:_load_config -> AutoConfig.from_pretrained) and
        # mixq_b200.auto.AutoForCausalLM.from_quantized read the architecture from
        import json
        cfg = self.cfg
        hf = {"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": cfg.hidden,
              "intermediate_size": cfg.intermediate, "num_hidden_layers": self.n_layers, "num_attention_heads": cfg.heads,
              "num_key_value_heads": cfg.kv_heads, "vocab_size": cfg.vocab, "rms_norm_eps": cfg.eps, "rope_theta": cfg.rope_theta,
              "hidden_act": "silu", "max_position_embeddings": 4096, "torch_dtype": "float16", "tie_word_embeddings": False}
        path = os.path.join(save_dir, "config.json")
        with open(path, "w") as f:
            json.dump(hf, f, indent=2)
        return files + [path]

    @classmethod
    def from_quantized(cls, save_dir: str, cfg: LlamaConfig, batch: int, device="cuda", safetensors: bool = False):
        """models/base.py:162-229 for the decode harness: load a MixQ checkpoint directory (q/k/v fused into W_pack like
        models/llama.py:98-166, o_proj / down_proj kept 8-bit in 4-bit models) instead of random-initialising."""
        from . import checkpoint as ck
        qc = ck.load_quant_config(save_dir)
        if int(qc.get("w_bit", 0)) not in (4, 8):
            raise ValueError(f"{save_dir}: not a quantised checkpoint (no quant_config.json with w_bit 4 or 8)")
        self = cls.__new__(cls)
        self.cfg, self.batch, self.bit, self.device = cfg, batch, int(qc["w_bit"]), device
        self.rank, self.world, self.group = 0, 1, None
        self.h_loc, self.kv_loc, self.i_loc = cfg.heads, cfg.kv_heads, cfg.intermediate
        self.cache = MixLibCache(inputdim=batch, sigma=6, bit=self.bit, device=device)
        mods, rest, _ = ck.load_quantized(save_dir, cache=self.cache, dev=device, safetensors=safetensors, fuse_layers=True)
        n_layers = 1 + max(int(k.split(".")[2]) for k in mods)
        self.n_layers = n_layers
        self.layers = []
        for i in range(n_layers):
            p = f"model.layers.{i}."
            self.layers.append({
                "ln1": rest[p + "input_layernorm.weight"].to(device), "ln2": rest[p + "post_attention_layernorm.weight"].to(device),
                "W_pack": mods[p + "self_attn.W_pack"], "o_proj": mods[p + "self_attn.o_proj"],
                "gate_proj": mods[p + "mlp.gate_proj"], "up_proj": mods[p + "mlp.up_proj"], "down_proj": mods[p + "mlp.down_proj"]})
        self.embed = rest["model.embed_tokens.weight"].to(device)
        self.norm_f = rest["model.norm.weight"].to(device)
        self.lm_head = rest["lm_head.weight"].to(device)
        self.discovered = False
        self.fuse_swiglu = batch > 128
        self.fuse_attn_quant = os.environ.get("MIXQ_FUSE_ATTN_QUANT", "1") != "0"
        self.fuse_xchg_quant = os.environ.get("MIXQ_TP_FUSE_QUANT", "1") != "0"
        self.vocab_parallel = False
        self.graph = self._static_tokens = self._static_logits = self.kv = None
        self.lib = _lib.load()
        self.xchg = None
        return self

    # ------------------------------------------------------------------ accounting
    def linear_shapes(self):
        """(name, N, K, bit, n_outliers) of the five MixLinears of layer 0 as sharded on this rank."""
        return [(k, m.out_features, m.in_features, m.bit, m._n_ind) for k, m in self.layers[0].items()
                if isinstance(m, MixLinear_GEMM)]

    def algorithmic_work(self):
        """(flops, bytes) of the quantised Linears of one decode step on this rank."""
        M = self.batch
        fl = by = 0
        for L in self.layers:
            for m in L.values():
                if not isinstance(m, MixLinear_GEMM):
                    continue
                N, K, n = m.out_features, m.in_features, m._n_ind
                fl += 2 * M * N * K
                by += N * K * m.bit // 8 + 2 * M * K + 2 * M * N + 2 * N + 2 * n * N
        return fl, by

    def launches_per_step(self):
        return (5 if self.fuse_swiglu else 7) * self.n_layers + 1   # + final RMSNorm; embedding / lm_head are library calls

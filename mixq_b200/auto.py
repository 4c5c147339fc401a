"""AutoForCausalLM / LlamaMixQForCausalLM — the reference's model-level surface over the MixLinear hot path.

`basic_quant_mix.py -> AutoForCausalLM.from_quantized(...) -> benchflops.py` is the user journey BASELINE.json names; this
module keeps those names and call signatures (/root/reference/mixquant/models/auto.py:26-53, models/base.py:162-229,
models/llama.py:9-22) on top of this package's checkpoint reader and fused modules:

    model = AutoForCausalLM.from_quantized(quant_path, quant_file, fuse_layers=True, mix=True, cache=MixLibCache(bit=8),
                                           batch_size=512)
    out = model(input_ids, use_cache=True)        # benchflops.py:100, :124 — out.logits / out[0]

What is NOT here (out of scope): the HF model zoo plumbing (accelerate device maps, remote code), the other
model families of CAUSAL_LM_MODEL_MAP, and offline quantisation from an fp16 HF checkpoint beyond `quantize_state_dict`
(MixQuantizer's per-layer loop, mixquant/quantize/mixquant.py:164-267, is `MixLinear_GEMM.from_linear` per Linear).
A checkpoint directory is: config.json (HF Llama keys), quant_config.json, and the sharded state dict (checkpoint.py).
"""
from __future__ import annotations

import json
import os
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from . import checkpoint as ck
from .attn import QuantAttentionFused
from .cache import MixLibCache
from .linear import MixLinear_GEMM
from .mlp import MixLlamaMLP
from .norm import FasterTransformerRMSNorm

SUPPORTED_MODEL_TYPES = ("llama", "aquila")     # auto.py:6-15 maps both to LlamaMixQForCausalLM; the rest is out of scope


def check_and_get_model_type(model_dir, trust_remote_code=True):
    """auto.py:17-24 without the transformers dependency: read config.json's model_type."""
    path = os.path.join(model_dir, "config.json")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path}: a MixQ checkpoint directory holds the HF config.json of the model")
    cfg = json.load(open(path))
    mt = cfg.get("model_type")
    if mt not in SUPPORTED_MODEL_TYPES:
        raise TypeError(f"{mt} isn't supported yet.")
    return mt


class CausalLMOutput(tuple):
    """`out.logits`, `out[0]`, `out.past_key_values` — what benchflops / generate read from the HF output object."""

    def __new__(cls, logits, past_key_values=None):
        self = super().__new__(cls, (logits, past_key_values))
        self.logits, self.past_key_values = logits, past_key_values
        return self


class _LlamaLayer(nn.Module):
    def __init__(self, input_layernorm, self_attn, post_attention_layernorm, mlp):
        super().__init__()
        self.input_layernorm, self.self_attn = input_layernorm, self_attn
        self.post_attention_layernorm, self.mlp = post_attention_layernorm, mlp

    @torch.no_grad()
    def forward(self, h, use_cache):
        # HF LlamaDecoderLayer with the fused modules of models/llama.py:9-22: norm -> attn -> +residual -> norm -> mlp -> +residual
        a, _, _ = self.self_attn(self.input_layernorm(h), None, None, None, False, use_cache)
        h = h + a
        return h + self.mlp(self.post_attention_layernorm(h))


class LlamaMixQForCausalLM(nn.Module):
    """models/llama.py:4-22 + base.py:24-39: the wrapped, layer-fused model.  `.model` is the callable benchflops uses."""

    layer_type = "LlamaDecoderLayer"
    max_new_tokens_key = "max_position_embeddings"

    def __init__(self, config: dict, layers, embed, norm_f, lm_head, cache: MixLibCache, quant_config: dict,
                 model_type="llama", is_quantized=True):
        super().__init__()
        self.config, self.quant_config, self.model_type, self.is_quantized = config, quant_config, model_type, is_quantized
        self.layers = nn.ModuleList(layers)
        self.embed_tokens = nn.Parameter(embed, requires_grad=False)
        self.norm = norm_f
        self.lm_head = nn.Parameter(lm_head, requires_grad=False)
        self.cache = cache
        self.model = self            # base.py wraps the HF model in `.model`; callers use either

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def forward(self, input_ids, use_cache: bool = False, past_key_values=None, **kwargs):
        """input_ids [B, q_len] -> CausalLMOutput(logits [B, q_len, vocab]).  As in HF, a call WITHOUT past_key_values starts
        from an empty KV cache (benchflops.py:124 never passes one: every timed iteration is an independent forward); pass the
        returned `past_key_values` back to continue a sequence."""
        if not input_ids.is_cuda:
            raise _lib.MixqError("the MixQ model needs CUDA tensors: there is no CPU path")
        if past_key_values is None:
            for L in self.layers:
                L.self_attn.reset_cache()
        h = torch.nn.functional.embedding(input_ids, self.embed_tokens)
        for L in self.layers:
            h = L(h, use_cache)
        logits = torch.matmul(self.norm(h), self.lm_head.t())
        return CausalLMOutput(logits, ("mixq-kv", self.layers[0].self_attn.start_pos) if use_cache else None)

    def generate(self, input_ids, max_new_tokens=16):
        """Greedy decode (generate.py's use): prefill, then one token per step on the module-owned KV caches."""
        out = self(input_ids, use_cache=True)
        toks = [input_ids]
        for _ in range(max_new_tokens):
            nxt = out.logits[:, -1].argmax(-1, keepdim=True)
            toks.append(nxt)
            out = self(nxt, use_cache=True, past_key_values=out.past_key_values)
        return torch.cat(toks, 1)


def build_llama(config: dict, mods: dict, rest: dict, cache: MixLibCache, quant_config: dict, dev="cuda", max_seq_len=None):
    """LlamaFuser (models/llama.py:49-178) for modules loaded by checkpoint.load_quantized(fuse_layers=True)."""
    H, nh = config["hidden_size"], config["num_attention_heads"]
    nkv = config.get("num_key_value_heads", nh)
    eps, theta = config.get("rms_norm_eps", 1e-6), config.get("rope_theta", 10000.0)
    msl = max_seq_len or config.get("max_new_tokens") or config.get("max_position_embeddings", 2048)
    layers = []
    for i in range(config["num_hidden_layers"]):
        p = f"model.layers.{i}."
        W_pack, o_proj = mods[p + "self_attn.W_pack"], mods[p + "self_attn.o_proj"]
        gate, up, down = mods[p + "mlp.gate_proj"], mods[p + "mlp.up_proj"], mods[p + "mlp.down_proj"]
        ln1 = FasterTransformerRMSNorm(rest[p + "input_layernorm.weight"].to(dev), eps, cache)
        ln1.next_layer = W_pack                                         # llama.py:20-22
        ln2 = FasterTransformerRMSNorm(rest[p + "post_attention_layernorm.weight"].to(dev), eps, cache)
        ln2.next_layer = up
        attn = QuantAttentionFused(H, nh, nkv, W_pack, o_proj, dev, msl, MixGemmCache=cache, layer_idx=i, rope_theta=theta)
        layers.append(_LlamaLayer(ln1, attn, ln2, MixLlamaMLP(gate, down, up, cache)))
    norm_f = FasterTransformerRMSNorm(rest["model.norm.weight"].to(dev), eps, cache)
    embed = rest["model.embed_tokens.weight"].to(dev, torch.float16)
    lm_head = rest.get("lm_head.weight", rest["model.embed_tokens.weight"]).to(dev, torch.float16)
    return LlamaMixQForCausalLM(config, layers, embed, norm_f, lm_head, cache, quant_config)


def quantize_state_dict(sd: dict, config: dict, w_bit: int = 8, act_scales: Optional[dict] = None, dev="cuda", cache=None):
    """quantize_mix (base.py:41-57 -> MixQuantizer.quantize, mixquant.py:164-267) on an fp16 HF Llama state dict: every decoder
    Linear -> MixLinear_GEMM.from_linear; down_proj / o_proj stay 8-bit in 4-bit models (utils/module.py:2, base.py:308-312);
    bit 4 needs per-input-channel activation scales (mixquant.py:201-208: act_scales[<module name>]).  Returns
    ({prefix: MixLinear_GEMM}, remaining tensors)."""
    mods, rest = {}, {}
    names = ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj",
             "mlp.down_proj")

    class _W:
        def __init__(self, w):
            self.weight, self.bias = w, None
            self.out_features, self.in_features = w.shape
    quantised = set()
    for i in range(config["num_hidden_layers"]):
        for nm in names:
            p = f"model.layers.{i}.{nm}"
            w = sd[p + ".weight"].to(dev, torch.float16)
            bit = 8 if (w_bit == 4 and any(e in nm for e in ("down_proj", "o_proj", "fc_out"))) else w_bit
            scales = None
            if bit == 4:
                if act_scales is None or p not in act_scales:
                    raise ValueError(f"4-bit quantisation needs act_scales['{p}'] (mixquant.py:201-208)")
                scales = act_scales[p]
            mods[p] = MixLinear_GEMM.from_linear(_W(w), bit, cache=cache, layer_scales=scales, dev=dev, name=p)
            quantised.add(p + ".weight")
    rest = {k: v for k, v in sd.items() if k not in quantised}
    return mods, rest


class AutoForCausalLM:
    def __init__(self):
        raise EnvironmentError("You must instantiate AutoForCausalLM with\n"
                               "AutoForCausalLM.from_quantized or AutoForCausalLM.from_pretrained")

    @classmethod
    def from_pretrained(cls, model_path, trust_remote_code=True, safetensors=False, device_map=None, mix=False,
                        **model_init_kwargs):
        """auto.py:31-39: an fp16 HF Llama directory -> an object with `.quantize_mix(...)` and `.save_quantized(dir)`
        (examples/basic_quant_mix.py).  Only the tensors are read (config.json + the sharded state dict)."""
        check_and_get_model_type(model_path, trust_remote_code)
        return _Pretrained(model_path, safetensors)

    @classmethod
    def from_quantized(cls, quant_path, quant_filename="", max_new_tokens=None, trust_remote_code=True, fuse_layers=True,
                       batch_size=1, safetensors=False, max_memory=None, offload_folder=None, mix=False, cache=None):
        """auto.py:42-53 -> base.py:162-229.  `mix=True` is the only implemented branch there too (`else: raise
        NotImplementedError`, base.py:193-194); QUIK checkpoints (version "QUIK") load MixedQLinear modules (qlinear.py)."""
        model_type = check_and_get_model_type(quant_path, trust_remote_code)
        os.environ["BATCH_SIZE"] = str(batch_size)                      # auto.py:48
        if not mix:
            raise NotImplementedError
        config = json.load(open(os.path.join(quant_path, "config.json")))
        config["max_new_tokens"] = 2048 if max_new_tokens is None else max_new_tokens     # base.py:253-260
        qc = ck.load_quant_config(quant_path)
        if int(qc.get("w_bit", 0)) not in (4, 8):
            raise ValueError(f"{quant_path}: no quant_config.json with w_bit 4 or 8 — not a quantised checkpoint (the reference "
                             "quantises online in that case, base.py:252-258: use from_pretrained(...).quantize(...))")
        if qc.get("version") == "QUIK":
            raise NotImplementedError("QUIK checkpoints: build MixedQLinear modules with mixq_b200.qlinear (fuse_layers is False there: base.py:186-188)")
        if cache is None:
            cache = MixLibCache(inputdim=max(batch_size, 1024), bit=int(qc["w_bit"]))
        if not fuse_layers:
            raise NotImplementedError("un-fused MixQ models run the HF modules (outside this package); use fuse_layers=True")
        weights_dir = quant_filename if quant_filename and os.path.isdir(quant_filename) else quant_path   # base.py:245-248
        mods, rest, qc = ck.load_quantized(weights_dir, cache=cache, dev="cuda", safetensors=safetensors, fuse_layers=True)
        model = build_llama(config, mods, rest, cache, qc)
        model.model_type = model_type
        return model


class _Pretrained:
    """What `AutoForCausalLM.from_pretrained` returns: quantize_mix + save_quantized (base.py:41-119)."""

    def __init__(self, path, safetensors):
        self.path, self.safetensors = path, safetensors
        self.config = json.load(open(os.path.join(path, "config.json")))
        self.mods = self.rest = self.quant_config = None

    @torch.no_grad()
    def quantize_mix(self, tokenizer=None, quant_config=None, calib_data=None, act_scales=None, dev="cuda", **kw):
        quant_config = dict(quant_config or {})
        w_bit = int(quant_config.get("w_bit", 8))
        sd = ck.load_state_dict(self.path, self.safetensors)
        self.mods, self.rest = quantize_state_dict(sd, self.config, w_bit, act_scales, dev=dev)
        self.quant_config = {"w_bit": w_bit, "version": quant_config.get("version", "MIX"), "q_group_size": int(quant_config.get("q_group_size", 128))}

    def save_quantized(self, save_dir, safetensors=False, shard_size="10GB"):
        if self.mods is None:
            raise RuntimeError("call quantize_mix first")
        files = list(ck.save_quantized(save_dir, self.mods, self.quant_config, extra=self.rest, safetensors=safetensors,
                                       shard_size=shard_size))
        with open(os.path.join(save_dir, "config.json"), "w") as f:
            json.dump(self.config, f, indent=2)
        return files

"""On-disk format of a MixQ-quantised model: read and write what the reference's `save_quantized` / `from_quantized` use.

Reference (paths under /root/reference):
  * mixquant/models/base.py:78-119 `save_quantized`: HF-style sharded state dict — `pytorch_model.bin` (or
    `pytorch_model-0000i-of-0000n.bin` + `pytorch_model.bin.index.json`) via `torch.save`, or `model.safetensors`
    (+ `model.safetensors.index.json`) — next to `quant_config.json` = {"w_bit", "version" ("MIX"/"QUIK"), "q_group_size"}.
  * mixquant/modules/linear.py:39-65: per MixLinear the state-dict keys are
        <module>.q_weight   int8 [N, K]            (bit 4: uint8 [N, K/2], low nibble = even column)
        <module>.scale_col  fp16 [1, N]
        <module>.bias       fp16 [N]               (only when the Linear has one)
        <module>.weight_cache fp16 [N, 128], <module>.ind int32 [128]      (bit 4 only: static outliers)
    For bit 8, `ind` / `weight_cache` are plain attributes (linear.py:42-44): the online-discovered outlier set is NOT
    checkpointed by the reference; `save_outlier_state=True` stores it under two extra keys so that a process can skip
    the two host-synchronising discovery calls (files written without it load in the reference unchanged).
  * mixquant/models/llama.py:98-166 `_fuse_qkv`: at load time q/k/v are concatenated along N into one `W_pack`.

Pure host-side plumbing (torch CPU tensors, json); the kernels never see it.
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, Iterable, Mapping, Optional, Tuple

import torch

from .linear import MixLinear_GEMM

QUANT_CONFIG_NAME = "quant_config.json"
_EXTRA_IND = "outlier_ind"          # bit-8 extension keys (see module docstring)
_EXTRA_WC = "outlier_weight_cache"


# ----------------------------------------------------------------------------------------------- one Linear
def linear_state(m: MixLinear_GEMM, prefix: str = "", save_outlier_state: bool = False) -> Dict[str, torch.Tensor]:
    """The reference's state-dict entries of one MixLinear (CPU tensors, contiguous)."""
    p = prefix + "." if prefix and not prefix.endswith(".") else prefix
    sd = {p + "q_weight": m.q_weight.detach().cpu().contiguous(),
          p + "scale_col": m.scale_col.detach().cpu().contiguous()}
    if m.bias is not None:
        sd[p + "bias"] = m.bias.detach().cpu().contiguous()
    if m.bit == 4:
        n = m.fp_features_num
        sd[p + "weight_cache"] = m._wc_buf[:, :n].detach().cpu().contiguous()
        sd[p + "ind"] = m._ind_buf[:n].detach().cpu().contiguous()
    elif save_outlier_state and m._n_ind > 0:
        sd[p + _EXTRA_IND] = m.ind.detach().cpu().contiguous()
        sd[p + _EXTRA_WC] = m.weight_cache.detach().cpu().contiguous()
    return sd


def linear_from_state(sd: Mapping[str, torch.Tensor], prefix: str, bit: int, cache=None, dev="cuda",
                      name: Optional[str] = None) -> MixLinear_GEMM:
    """Rebuild one MixLinear from its checkpoint entries (shapes and dtypes are checked against the reference layout)."""
    p = prefix + "." if prefix and not prefix.endswith(".") else prefix
    qw, sc = sd[p + "q_weight"], sd[p + "scale_col"]
    N = qw.shape[0]
    K = qw.shape[1] * (2 if bit == 4 else 1)
    want = torch.uint8 if bit == 4 else torch.int8
    if qw.dtype != want:
        raise ValueError(f"{p}q_weight is {qw.dtype}, a w_bit={bit} checkpoint stores {want}")
    if tuple(sc.shape) != (1, N) or sc.dtype != torch.float16:
        raise ValueError(f"{p}scale_col must be fp16 [1, {N}], got {sc.dtype} {tuple(sc.shape)}")
    bias = sd.get(p + "bias")
    fpn = sd[p + "ind"].shape[0] if bit == 4 else 128
    m = MixLinear_GEMM(K, N, bias is not None, dev, bit, cache=cache, name=name or prefix, fp_features_num=fpn)
    m.q_weight.copy_(qw)
    m.scale_col.copy_(sc)
    if bias is not None:
        m.bias.copy_(bias)
    if bit == 4:
        wc, ind = sd[p + "weight_cache"], sd[p + "ind"]
        if tuple(wc.shape) != (N, fpn) or ind.dtype != torch.int32:
            raise ValueError(f"{p}weight_cache / ind do not match the bit-4 layout")
        m._wc_buf[:, :fpn] = wc.to(m._wc_buf.device)
        m._ind_buf[:fpn] = ind.to(m._ind_buf.device)
        m._n_ind = fpn
    elif p + _EXTRA_IND in sd:
        m.weight_cache = sd[p + _EXTRA_WC]
        m.ind = sd[p + _EXTRA_IND]
        m.add_outliers = False            # the stored set replaces the discovery calls (linear.py:200-226)
        m.forward_without_precondition_len = m._n_ind
    return m


def fuse_qkv(q: MixLinear_GEMM, k: MixLinear_GEMM, v: MixLinear_GEMM, cache=None, dev=None) -> MixLinear_GEMM:
    """models/llama.py:98-166: W_pack = q|k|v concatenated along N (q_weight rows, scale_col columns, weight_cache rows);
    bit 4 takes q_proj's static outlier columns (all three were quantised with the same layer_scales)."""
    if not (q.in_features == k.in_features == v.in_features and q.bit == k.bit == v.bit):
        raise ValueError("q/k/v must share in_features and bit width")
    if q.bias is not None or k.bias is not None or v.bias is not None:
        raise NotImplementedError("fused qkv with bias (the reference raises here too: llama.py:147)")
    dev = dev if dev is not None else q.q_weight.device
    fpn = getattr(q, "fp_features_num", 128)
    w = MixLinear_GEMM(q.in_features, q.out_features + k.out_features + v.out_features, False, dev, q.bit,
                       cache=cache if cache is not None else q.cache, name="W_pack", fp_features_num=fpn)
    w.q_weight.copy_(torch.cat([q.q_weight, k.q_weight, v.q_weight], dim=0))
    w.scale_col.copy_(torch.cat([q.scale_col, k.scale_col, v.scale_col], dim=1))
    if q.bit == 4:
        if not (torch.equal(q.ind, k.ind) and torch.equal(q.ind, v.ind)):
            raise ValueError("bit-4 q/k/v must share their static outlier columns")
        w._wc_buf[:, :fpn] = torch.cat([q._wc_buf[:, :fpn], k._wc_buf[:, :fpn], v._wc_buf[:, :fpn]], dim=0)
        w._ind_buf[:fpn] = q._ind_buf[:fpn]
        w._n_ind = fpn
    return w


# ----------------------------------------------------------------------------------------------- whole checkpoints
def _parse_size(s) -> int:
    if isinstance(s, int):
        return s
    m = re.fullmatch(r"\s*(\d+(?:\.\d+)?)\s*([KMG]i?B)\s*", str(s))
    if not m:
        raise ValueError(f"bad shard size {s!r}")
    unit = {"KB": 10**3, "MB": 10**6, "GB": 10**9, "KiB": 2**10, "MiB": 2**20, "GiB": 2**30}[m.group(2)]
    return int(float(m.group(1)) * unit)


def shard_state_dict(sd: Mapping[str, torch.Tensor], max_shard_size="10GB", weights_name="pytorch_model.bin"
                     ) -> Tuple[Dict[str, Dict[str, torch.Tensor]], Optional[dict]]:
    """HF `shard_checkpoint` layout (base.py:99-103): greedy split in key order; one shard keeps `weights_name`,
    several are named `<stem>-0000i-of-0000n<ext>` and come with an index {"metadata": {"total_size"}, "weight_map"}."""
    limit = _parse_size(max_shard_size)
    shards, cur, cur_size, total = [], {}, 0, 0
    for k, t in sd.items():
        sz = t.numel() * t.element_size()
        if cur and cur_size + sz > limit:
            shards.append(cur)
            cur, cur_size = {}, 0
        cur[k] = t
        cur_size += sz
        total += sz
    shards.append(cur)
    if len(shards) == 1:
        return {weights_name: shards[0]}, None
    stem, ext = os.path.splitext(weights_name)
    out, weight_map = {}, {}
    for i, sh in enumerate(shards):
        fn = f"{stem}-{i + 1:05d}-of-{len(shards):05d}{ext}"
        out[fn] = sh
        for k in sh:
            weight_map[k] = fn
    return out, {"metadata": {"total_size": total}, "weight_map": weight_map}


def save_quantized(save_dir: str, modules: Mapping[str, MixLinear_GEMM], quant_config: Mapping, extra: Mapping[str, torch.Tensor] = (),
                   safetensors: bool = False, shard_size="10GB", save_outlier_state: bool = False) -> Iterable[str]:
    """Write the quantised Linears (`modules`: state-dict prefix -> module, e.g. "model.layers.0.mlp.up_proj") plus any
    un-quantised tensors (`extra`: embeddings, norms, lm_head) in the reference's layout.  Returns the files written."""
    os.makedirs(save_dir, exist_ok=True)
    sd: Dict[str, torch.Tensor] = {}
    for k, t in dict(extra).items():
        sd[k] = t.detach().cpu().contiguous()
    for prefix, m in modules.items():
        sd.update(linear_state(m, prefix, save_outlier_state))
    name = "model.safetensors" if safetensors else "pytorch_model.bin"
    shards, index = shard_state_dict(sd, shard_size, name)
    written = []
    for fn, shard in shards.items():
        path = os.path.join(save_dir, fn)
        if safetensors:
            from safetensors.torch import save_file
            save_file({k: v.clone().contiguous() for k, v in shard.items()}, path, metadata={"format": "pt"})
        else:
            torch.save(shard, path)
        written.append(path)
    if index is not None:
        path = os.path.join(save_dir, name + ".index.json")
        with open(path, "w") as f:
            f.write(json.dumps(index, indent=4))
        written.append(path)
    qc = {"w_bit": int(quant_config.get("w_bit", 8)), "version": quant_config.get("version", "MIX"),
          "q_group_size": int(quant_config.get("q_group_size", 128))}
    if qc["version"] != "MIX":
        raise NotImplementedError("only the MIX layout is on the MixLinear path (QUIK's MixedQLinear is out of scope)")
    path = os.path.join(save_dir, QUANT_CONFIG_NAME)
    with open(path, "w") as f:
        f.write(json.dumps(qc, indent=4))
    written.append(path)
    return written


def load_state_dict(save_dir: str, safetensors: bool = False) -> Dict[str, torch.Tensor]:
    """All tensors of a (possibly sharded) checkpoint directory, on the CPU."""
    name = "model.safetensors" if safetensors else "pytorch_model.bin"
    index_path = os.path.join(save_dir, name + ".index.json")
    files = sorted(set(json.load(open(index_path))["weight_map"].values())) if os.path.exists(index_path) else [name]
    sd: Dict[str, torch.Tensor] = {}
    for fn in files:
        path = os.path.join(save_dir, fn)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        if safetensors:
            from safetensors.torch import load_file
            sd.update(load_file(path))
        else:
            sd.update(torch.load(path, map_location="cpu", weights_only=True))
    return sd


def load_quant_config(save_dir: str, version: str = "Mix") -> dict:
    """base.py:249-258: quant_config.json, or {"w_bit": 0, "version": version} when the file is missing — the reference's
    marker for "not a quantised checkpoint, quantise online"."""
    path = os.path.join(save_dir, QUANT_CONFIG_NAME)
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {"w_bit": 0, "version": version}


def load_quantized(save_dir: str, cache=None, dev="cuda", safetensors: bool = False, fuse_layers: bool = False,
                   eight_bit_names: Iterable[str] = ("down_proj", "o_proj", "fc_out")
                   ) -> Tuple[Dict[str, MixLinear_GEMM], Dict[str, torch.Tensor], dict]:
    """from_quantized (base.py:162-229) for the quantised Linears: returns ({prefix: MixLinear}, remaining tensors, quant_config).
    `eight_bit_names`: modules that stay 8-bit in a 4-bit model — matched as SUBSTRINGS of the module name, as the reference
    does (utils/module.py:2 `eightbit_only_name`, base.py:308-312).  With `fuse_layers`
    every `<p>.q_proj / k_proj / v_proj` triple is replaced by `<p>.W_pack` (llama.py:98-166)."""
    qc = load_quant_config(save_dir)
    if int(qc.get("w_bit", 0)) not in (4, 8):
        raise ValueError(f"{save_dir}: no {QUANT_CONFIG_NAME} with w_bit 4 or 8 (w_bit 0 = not a quantised checkpoint; quantise it "
                         "with AutoForCausalLM.from_pretrained(...).quantize(...) first)")
    sd = load_state_dict(save_dir, safetensors)
    prefixes = sorted({k[: -len(".q_weight")] for k in sd if k.endswith(".q_weight")})
    mods: Dict[str, MixLinear_GEMM] = {}
    used = set()
    for p in prefixes:
        bit = qc["w_bit"]
        if bit == 4 and any(key in p for key in eight_bit_names):
            bit = 8
        mods[p] = linear_from_state(sd, p, bit, cache=cache, dev=dev)
        used.update(k for k in sd if k.startswith(p + "."))
    if fuse_layers:
        for p in sorted({k[: -len(".q_proj")] for k in mods if k.endswith(".q_proj")}):
            q, k_, v = mods.pop(p + ".q_proj"), mods.pop(p + ".k_proj"), mods.pop(p + ".v_proj")
            mods[p + ".W_pack"] = fuse_qkv(q, k_, v, cache=cache, dev=dev)
    rest = {k: v for k, v in sd.items() if k not in used}
    return mods, rest, qc

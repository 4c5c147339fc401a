"""mixq_b200/checkpoint.py against the reference's on-disk layout (base.py:78-119, linear.py:39-65, llama.py:98-166).
CPU only: the format is host-side plumbing."""
import json
import os

import pytest
import torch

from mixq_b200 import checkpoint as ck
from mixq_b200.cache import MixLibCache
from mixq_b200.linear import MixLinear_GEMM

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class _Lin:
    def __init__(self, n, k, bias, seed):
        g = torch.Generator().manual_seed(seed)
        self.weight = torch.nn.Parameter((torch.randn(n, k, generator=g) * 0.02).half(), requires_grad=False)
        self.bias = torch.nn.Parameter(torch.randn(n, generator=g).half(), requires_grad=False) if bias else None
        self.in_features, self.out_features = k, n


def _make(bit, n=96, k=256, bias=False, seed=0, cache=None):
    cache = cache or MixLibCache(inputdim=16, sigma=6, bit=bit, device="cpu")
    scales = torch.rand(k, generator=torch.Generator().manual_seed(99)) if bit == 4 else None
    return MixLinear_GEMM.from_linear(_Lin(n, k, bias, seed), bit, cache=cache, dev="cpu", layer_scales=scales)


@pytest.mark.parametrize("tag,bit,bias", [("w8", 8, False), ("w8_bias", 8, True), ("w4", 4, False)])
def test_state_entries_match_reference_manifest(tag, bit, bias):
    """Keys, dtypes and shapes equal those of the reference's own MixLinear_GEMM.state_dict() (fixture written by
    tests/golden/make_golden.py from /root/reference/mixquant/modules/linear.py)."""
    man = json.load(open(os.path.join(GOLDEN, "state_manifest.json")))[tag]
    m = _make(bit, man["out_features"], man["in_features"], bias)
    got = {k: [str(v.dtype).replace("torch.", ""), list(v.shape)] for k, v in ck.linear_state(m).items()}
    assert got == man["state"]


@pytest.mark.parametrize("safetensors", [False, True])
@pytest.mark.parametrize("bit", [8, 4])
def test_save_load_round_trip_sharded(tmp_path, bit, safetensors):
    cache = MixLibCache(inputdim=16, sigma=6, bit=bit, device="cpu")
    names = ["model.layers.0.self_attn.q_proj", "model.layers.0.self_attn.k_proj", "model.layers.0.self_attn.v_proj",
             "model.layers.0.self_attn.o_proj", "model.layers.0.mlp.up_proj"]
    mods = {}
    for i, nme in enumerate(names):
        b = 8 if nme.endswith("o_proj") else bit     # o_proj / down_proj stay 8-bit in 4-bit models (utils/module.py:2)
        mods[nme] = _make(b, 96, 256, bias=False, seed=i, cache=cache)
    extra = {"model.embed_tokens.weight": torch.randn(32, 256).half(), "model.norm.weight": torch.ones(256).half()}
    files = ck.save_quantized(str(tmp_path), mods, {"w_bit": bit, "version": "MIX", "q_group_size": 128}, extra=extra,
                              safetensors=safetensors, shard_size="40KB")
    name = "model.safetensors" if safetensors else "pytorch_model.bin"
    index = json.load(open(tmp_path / (name + ".index.json")))
    assert set(index) == {"metadata", "weight_map"} and len(set(index["weight_map"].values())) > 1
    assert all(os.path.exists(tmp_path / f) for f in index["weight_map"].values())
    assert json.load(open(tmp_path / "quant_config.json")) == {"w_bit": bit, "version": "MIX", "q_group_size": 128}
    assert str(tmp_path / "quant_config.json") in files

    loaded, rest, qc = ck.load_quantized(str(tmp_path), cache=cache, dev="cpu", safetensors=safetensors)
    assert qc["w_bit"] == bit and set(loaded) == set(names) and set(rest) == set(extra)
    for nme, m in mods.items():
        l = loaded[nme]
        assert l.bit == m.bit and l.in_features == m.in_features and l.out_features == m.out_features
        assert torch.equal(l.q_weight, m.q_weight) and torch.equal(l.scale_col, m.scale_col)
        if m.bit == 4:
            assert torch.equal(l.ind, m.ind) and torch.equal(l.weight_cache, m.weight_cache) and l._n_ind == 128
        else:
            assert l._n_ind == 0 and l.add_outliers           # the reference does not checkpoint the discovered set
    for k, v in extra.items():
        assert torch.equal(rest[k], v)


def test_single_file_when_small(tmp_path):
    m = _make(8, bias=True)
    ck.save_quantized(str(tmp_path), {"lin": m}, {"w_bit": 8})
    assert os.path.exists(tmp_path / "pytorch_model.bin") and not os.path.exists(tmp_path / "pytorch_model.bin.index.json")
    loaded, rest, qc = ck.load_quantized(str(tmp_path), dev="cpu")
    assert torch.equal(loaded["lin"].bias, m.bias) and not rest and qc["version"] == "MIX"


@pytest.mark.parametrize("bit", [8, 4])
def test_fuse_qkv_layout(tmp_path, bit):
    cache = MixLibCache(inputdim=16, sigma=6, bit=bit, device="cpu")
    q, k, v = (_make(bit, n, 256, seed=s, cache=cache) for n, s in ((96, 1), (32, 2), (32, 3)))
    mods = {"model.layers.3.self_attn.q_proj": q, "model.layers.3.self_attn.k_proj": k, "model.layers.3.self_attn.v_proj": v}
    ck.save_quantized(str(tmp_path), mods, {"w_bit": bit})
    loaded, _, _ = ck.load_quantized(str(tmp_path), cache=cache, dev="cpu", fuse_layers=True)
    assert set(loaded) == {"model.layers.3.self_attn.W_pack"}
    w = loaded["model.layers.3.self_attn.W_pack"]
    assert w.out_features == 160 and w.in_features == 256
    assert torch.equal(w.q_weight, torch.cat([q.q_weight, k.q_weight, v.q_weight], 0))          # llama.py:131
    assert torch.equal(w.scale_col, torch.cat([q.scale_col, k.scale_col, v.scale_col], 1))      # llama.py:132
    if bit == 4:
        assert torch.equal(w.ind, q.ind)                                                          # llama.py:145
        assert torch.equal(w.weight_cache, torch.cat([q.weight_cache, k.weight_cache, v.weight_cache], 0))


def test_outlier_state_extension_round_trip(tmp_path):
    m = _make(8)
    m.ind = torch.tensor([3, 17, 200], dtype=torch.int32)
    m.weight_cache = (m.q_weight[:, [3, 17, 200]].half() * m.scale_col.T)
    ck.save_quantized(str(tmp_path), {"l": m}, {"w_bit": 8}, save_outlier_state=True)
    l = ck.load_quantized(str(tmp_path), dev="cpu")[0]["l"]
    assert torch.equal(l.ind, m.ind) and torch.equal(l.weight_cache, m.weight_cache) and not l.add_outliers
    # without the flag the file holds only what the reference writes
    ck.save_quantized(str(tmp_path), {"l": m}, {"w_bit": 8})
    assert set(ck.load_state_dict(str(tmp_path))) == {"l.q_weight", "l.scale_col"}


def test_layout_errors(tmp_path):
    m = _make(8)
    sd = ck.linear_state(m, "l")
    with pytest.raises(ValueError):
        ck.linear_from_state(sd, "l", 4, dev="cpu")                   # int8 tensor in a w_bit=4 checkpoint
    sd["l.scale_col"] = sd["l.scale_col"].reshape(-1)
    with pytest.raises(ValueError):
        ck.linear_from_state(sd, "l", 8, dev="cpu")                   # scale_col must be [1, N]
    with pytest.raises(NotImplementedError):
        ck.save_quantized(str(tmp_path), {"l": m}, {"w_bit": 4, "version": "QUIK"})
    with pytest.raises(FileNotFoundError):
        ck.load_state_dict(str(tmp_path / "nowhere"))


@pytest.mark.parametrize("bit", [8, 4])
def test_decoder_save_and_from_quantized_round_trip(tmp_path, bit):
    """The decode harness writes the reference's layout (HF Llama module names, q/k/v separate) and reads it back with
    q/k/v fused into W_pack: every tensor identical.  (Construction is device-agnostic; the step itself needs the GPU.)"""
    from mixq_b200.llama import CONFIGS, LlamaDecoder
    cfg = CONFIGS["tiny"]
    m = LlamaDecoder(cfg, batch=8, bit=bit, device="cpu", seed=3)
    files = m.save_quantized(str(tmp_path), safetensors=(bit == 4), shard_size="200KB")
    assert any(f.endswith("quant_config.json") for f in files)
    sd = ck.load_state_dict(str(tmp_path), safetensors=(bit == 4))
    assert "model.layers.0.self_attn.q_proj.q_weight" in sd and "model.layers.1.mlp.down_proj.scale_col" in sd
    assert sd["model.layers.0.self_attn.k_proj.q_weight"].shape[0] == cfg.kv_heads * cfg.head_dim
    r = LlamaDecoder.from_quantized(str(tmp_path), cfg, batch=8, device="cpu", safetensors=(bit == 4))
    assert r.bit == bit and r.n_layers == m.n_layers == cfg.layers
    assert torch.equal(r.embed, m.embed) and torch.equal(r.lm_head, m.lm_head) and torch.equal(r.norm_f, m.norm_f)
    for a, b in zip(m.layers, r.layers):
        assert torch.equal(a["ln1"], b["ln1"]) and torch.equal(a["ln2"], b["ln2"])
        for k in ("W_pack", "o_proj", "gate_proj", "up_proj", "down_proj"):
            assert a[k].bit == b[k].bit and a[k].out_features == b[k].out_features, k
            assert torch.equal(a[k].q_weight, b[k].q_weight) and torch.equal(a[k].scale_col, b[k].scale_col), k
            if a[k].bit == 4:
                assert torch.equal(a[k].ind, b[k].ind) and torch.equal(a[k].weight_cache, b[k].weight_cache), k

"""GPU: the reference-facing surface (MixLinear_GEMM / MixLibCache / FasterTransformerRMSNorm / MixLlamaMLP, and the
`mixlib` module) replayed against the golden fixtures recorded from the reference's own Python, plus the Llama
decode step against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import llama_oracle as LO
from oracle import mixq_oracle as O

pytestmark = pytest.mark.gpu


def bits_equal(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: {a.shape} vs {b.shape}"
    if a.dtype == np.float16:
        ua, ub = a.view(np.uint16), np.asarray(b, np.float16).view(np.uint16)
        bad = (ua != ub) & ~(((ua | ub) & 0x7FFF) == 0)
    else:
        bad = a != b
    assert not bad.any(), f"{what}: {int(bad.sum())} of {a.size} differ"


def rel_close(a, b, what, tol=1e-2):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape
    rel = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
    assert rel <= tol, f"{what}: rel {rel:.3e}"


def host(t):
    return t.detach().cpu().numpy()


class _Lin:
    def __init__(self, W, b=None):
        self.weight = torch.nn.Parameter(torch.from_numpy(W.copy()).cuda(), requires_grad=False)
        self.bias = None if b is None else torch.nn.Parameter(torch.from_numpy(b.copy()).cuda(), requires_grad=False)
        self.out_features, self.in_features = W.shape


@pytest.mark.parametrize("case", ["w8_unfused", "w8_unfused_bias", "w4_unfused"])
def test_mixlinear_replays_reference_golden(golden, case):
    """linear.py:165-289 as run by the reference vs the product on the GPU: discovery state machine, outlier index
    sets (bit-exact), in-place zeroing, q_x / x_scale / weight_cache bits, y within 1e-2."""
    from mixq_b200.cache import MixLibCache
    from mixq_b200.linear import MixLinear_GEMM
    d = golden(case)
    bit, M, K = int(d["bit"]), int(d["M"]), int(d["K"])
    cache = MixLibCache(inputdim=32, sigma=6, bit=bit)
    ls = torch.from_numpy(d["layer_scales"]).cuda() if bit == 4 else None
    q = MixLinear_GEMM.from_linear(_Lin(d["W"], d["bias"] if "bias" in d.files else None), bit, cache=cache,
                                   layer_scales=ls, fp_features_num=int(d["fp"]))
    bits_equal(host(q.q_weight), d["q_weight"], "q_weight")
    bits_equal(host(q.scale_col), d["scale_col"], "scale_col")
    for t in range(int(d["ncalls"])):
        x = torch.from_numpy(d[f"c{t}_x"].copy()).cuda()
        xin = x.reshape(M // 2, 2, K) if t == 2 else x
        y = q(xin, None, True)
        assert tuple(y.shape) == tuple(d[f"c{t}_y"].shape)
        bits_equal(host(q.ind), d[f"c{t}_ind"], f"call {t} outlier index set")
        bits_equal(host(x), d[f"c{t}_x_after"], f"call {t} x zeroed in place")
        bits_equal(host(cache.x_scale[:M]), d[f"c{t}_x_scale"], f"call {t} x_scale")
        bits_equal(host(cache.q_xcache), d[f"c{t}_q_x"], f"call {t} q_x")
        if q.ind.shape[0]:
            bits_equal(host(q.weight_cache), d[f"c{t}_weight_cache"], f"call {t} weight_cache")
            bits_equal(host(cache.activation_outliers), d[f"c{t}_act_outliers"], f"call {t} activation_outliers")
        assert int(q.add_outliers) == int(d[f"c{t}_add_outliers"])
        rel_close(host(y), d[f"c{t}_y"], f"call {t} y")


@pytest.mark.parametrize("case", ["w8_fused_mlp", "w4_fused_mlp"])
def test_fused_norm_mlp_replays_reference_golden(golden, case):
    """norm.py:14-39 -> mlp.py:57-70 -> linear.py:291-376 through the product's modules."""
    from mixq_b200.cache import MixLibCache
    from mixq_b200.linear import MixLinear_GEMM
    from mixq_b200.mlp import MixLlamaMLP
    from mixq_b200.norm import FasterTransformerRMSNorm
    d = golden(case)
    bit, M, K = int(d["bit"]), int(d["M"]), int(d["K"])
    cache = MixLibCache(inputdim=32, sigma=6, bit=bit)
    ls = torch.from_numpy(d["layer_scales"]).cuda() if bit == 4 else None
    fp = int(d["fp"])
    mk = lambda W, b, s=None: MixLinear_GEMM.from_linear(_Lin(W), b, cache=cache, layer_scales=s, fp_features_num=fp)
    up, gate, down = mk(d["Wu"], bit, ls), mk(d["Wg"], bit, ls), mk(d["Wd"], 8)
    norm = FasterTransformerRMSNorm(torch.from_numpy(d["norm_w"]).cuda(), float(d["eps"]), cache)
    norm.next_layer = up
    mlp = MixLlamaMLP(gate, down, up, cache)
    for t in range(int(d["ncalls"])):
        x = torch.from_numpy(d[f"c{t}_x"].copy()).cuda()
        h = norm(x)
        ref_n = d[f"c{t}_normed"]
        assert (np.abs(host(h).astype(np.float32) - ref_n.astype(np.float32)) <= np.abs(np.spacing(ref_n)).astype(np.float32)).all()
        y = mlp(h)
        bits_equal(host(up.ind), d[f"c{t}_up_ind"], f"call {t} up_proj outlier set")
        bits_equal(host(gate.ind), d[f"c{t}_gate_ind"], f"call {t} gate_proj outlier set")
        bits_equal(host(down.ind), d[f"c{t}_down_ind"], f"call {t} down_proj outlier set")
        rel_close(host(y), d[f"c{t}_y"], f"call {t} y")


def test_mixlib_module_signatures():
    """The twelve `mixlib` names with the reference's argument order (linear.py / norm.py call sites)."""
    from mixq_b200 import mixlib
    rng = np.random.default_rng(0)
    M, K, N = 24, 256, 136
    x = rng.standard_normal((M, K)).astype(np.float16)
    x[:, [5, 9]] *= np.float16(30)
    W = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
    qw, ws = O.quant_weight_w8(W)
    xd = torch.from_numpy(x.copy()).cuda()
    ind = torch.tensor([5, 9], dtype=torch.int32, device="cuda")
    x_scale = torch.zeros(64, 1, dtype=torch.float16, device="cuda")
    ao = mixlib.ExtractOutliersAndSetToZeros(ind, xd)
    q_x = mixlib.FindRowScale(xd, x_scale, M, K, 8)
    xr = x.copy()
    ao_ref = O.extract_outliers_and_set_to_zeros(np.array([5, 9]), xr)
    q_ref, xs_ref = O.find_row_scale(xr, 8)
    bits_equal(host(ao), ao_ref, "ExtractOutliersAndSetToZeros")
    bits_equal(host(q_x), q_ref, "FindRowScale q_x")
    bits_equal(host(x_scale[:M]), xs_ref, "FindRowScale x_scale")
    qwd, wsd = torch.from_numpy(qw).cuda(), torch.from_numpy(ws).cuda()
    wc = (qwd[:, ind.long()].half() * wsd.T)
    outl = torch.mm(ao, wc.T)
    zeros = torch.zeros(64, 4 * N, dtype=torch.float16, device="cuda")   # like cache.zeros: bigger than [M,N]
    y = mixlib.int8FusedDequantize(q_x, qwd, x_scale, wsd, outl, M, N, K)
    y0 = mixlib.int8FusedDequantize(q_x, qwd, x_scale, wsd, zeros, M, N, K)
    ys = mixlib.int8FusedDequantizeSilu(q_x, qwd, x_scale, wsd, zeros, M, N, K)
    bits_equal(host(y), O.int8_fused_dequantize(q_ref, qw, xs_ref, ws, host(outl)), "int8FusedDequantize")
    bits_equal(host(y0), O.int8_fused_dequantize(q_ref, qw, xs_ref, ws), "int8FusedDequantize(zeros)")
    rel_close(host(ys), O.int8_fused_dequantize(q_ref, qw, xs_ref, ws, act=1), "int8FusedDequantizeSilu", 2e-3)
    acc = mixlib.gemm(q_x, qwd, M, N, K)
    bits_equal(host(acc), O.gemm_i8(q_ref, qw), "gemm")
    bits_equal(host(mixlib.dequantizeInt8(acc, x_scale, wsd, outl, 8, M, N)), host(y), "dequantizeInt8 == fused")
    rel_close(host(mixlib.dequantizeInt8Silu(acc, x_scale, wsd, zeros, 8, M, N)), host(ys), "dequantizeInt8Silu", 1e-3)
    # int4
    q4 = rng.integers(-8, 8, (N, K), dtype=np.int8)
    qwp = O.pack_to_i4(q4)
    q_x4 = mixlib.FindRowScale(xd, x_scale, M, K, 4)
    q4_ref, xs4_ref = O.find_row_scale(xr, 4)
    bits_equal(host(q_x4), q4_ref, "FindRowScale bit 4")
    y4 = mixlib.int4FusedDequantize(q_x4, torch.from_numpy(qwp).cuda(), x_scale, wsd, zeros, M, N, K // 2)
    bits_equal(host(y4), O.int4_fused_dequantize(q4_ref, qwp, xs4_ref, ws), "int4FusedDequantize")
    bits_equal(host(mixlib.unpack_int4_to_fp16(torch.from_numpy(qwp).cuda(), ind)), O.unpack_int4_to_fp16(qwp, [5, 9]), "unpack")
    # norm family
    w = torch.ones(K, dtype=torch.float16, device="cuda")
    xin = torch.from_numpy(x.copy()).cuda()
    out = torch.empty_like(xin)
    mixlib.layernorm_forward_cuda(xin, w, out, 1e-5)
    out2 = torch.empty_like(xin)
    ao2, qx2 = mixlib.layernorm_forward_cuda_extract_outliers(xin, w, out2, 1e-5, ind, x_scale)
    normed = host(out).copy()
    ao2_ref = O.extract_outliers_and_set_to_zeros(np.array([5, 9]), normed)
    bits_equal(host(out2), normed, "normed, outlier columns zeroed")
    bits_equal(host(ao2), ao2_ref, "norm activation_outliers")
    bits_equal(host(qx2), O.find_row_scale(normed, 8)[0], "norm q_x")
    ao3, qx3 = mixlib.layernorm_forward_cuda_extract_outliers_int4(xin, w, out2, 1e-5, ind, x_scale)
    bits_equal(host(qx3), O.find_row_scale(normed, 4)[0], "norm q_x int4")


@pytest.mark.parametrize("bit", [8, 4])
def test_llama_decode_step_vs_oracle(bit):
    """Two discovery steps + steady state + CUDA-graph replay of the tiny Llama against oracle/llama_oracle.py."""
    from mixq_b200 import _lib
    from mixq_b200.llama import CONFIGS, LlamaDecoder
    cfg = CONFIGS["tiny"]
    B = 24
    m = LlamaDecoder(cfg, batch=B, bit=bit, seed=3, outlier_frac=0.02)
    tok = torch.randint(0, cfg.vocab, (B, 1), generator=torch.Generator().manual_seed(0)).cuda()
    # oracle twin built from the same quantised weights is impossible (weights are quantised from fp16 we no longer
    # hold) — rebuild the oracle layers from the product's integer weights instead
    cache = O.MixLibCacheOracle(B, 6, bit)
    layers = []
    for L in m.layers:
        lay = LO.LlamaLayerOracle.__new__(LO.LlamaLayerOracle)
        lay.ln1, lay.ln2 = host(L["ln1"]), host(L["ln2"])

        def mk(q):
            o = O.MixLinearOracle(host(q.q_weight), host(q.scale_col), q.bit, None, cache,
                                  ind=host(q.ind).copy() if q.bit == 4 else None,
                                  weight_cache=host(q.weight_cache).copy() if q.bit == 4 else None)
            return o
        lay.W_pack, lay.o_proj = mk(L["W_pack"]), mk(L["o_proj"])
        lay.gate, lay.up, lay.down = mk(L["gate_proj"]), mk(L["up_proj"]), mk(L["down_proj"])
        layers.append(lay)
    ocfg = dict(heads=cfg.heads, kv_heads=cfg.kv_heads, head_dim=cfg.head_dim, theta=cfg.rope_theta, eps=cfg.eps)
    emb = host(m.embed)[host(tok).reshape(-1)]
    lm = host(m.lm_head).astype(np.float32)

    def oracle_logits():
        h = LO.decode_step(emb.copy(), layers, cache, ocfg)
        hn = O.rmsnorm(h, host(m.norm_f), cfg.eps)
        return hn.astype(np.float32) @ lm.T

    for call in range(3):
        if call < 2:
            logits = m.step(tok)
            if call == 1:
                m.discovered = all(not L[k].add_outliers for L in m.layers for k in ("W_pack", "o_proj", "up_proj", "down_proj"))
                assert m.discovered
        else:
            n0 = _lib.launch_count()
            logits = m.step(tok)
            assert _lib.launch_count() - n0 == 7 * cfg.layers + 1, "steady state: 7 launches per layer + final norm"
        ref = oracle_logits()
        for L, lay in zip(m.layers, layers):
            for k, o in (("W_pack", lay.W_pack), ("o_proj", lay.o_proj), ("up_proj", lay.up), ("gate_proj", lay.gate), ("down_proj", lay.down)):
                bits_equal(host(L[k].ind), o.ind, f"call {call} {k} outlier index set")
        rel_close(host(logits), ref, f"call {call} logits", 2e-2)
    assert any(L["W_pack"].ind.shape[0] > 0 for L in m.layers), "forced outlier channels must have been discovered"
    m.capture(tok)
    out = m.replay(tok).clone()
    torch.cuda.synchronize()
    assert torch.equal(out, logits), "graph replay == eager steady-state step"


def test_llama_swiglu_fused_step_matches_unfused():
    """M > 128: the decode step with the SwiGLU pair in one launch (5 launches per layer) vs the reference's
    up_proj / gate_proj / gate*=up sequence (7 launches per layer) on the same model: same logits."""
    from mixq_b200 import _lib
    from mixq_b200.llama import CONFIGS, LlamaDecoder
    cfg = CONFIGS["tiny"]
    B = 160
    m = LlamaDecoder(cfg, batch=B, bit=8, seed=5, outlier_frac=0.02)
    assert m.fuse_swiglu
    tok = torch.randint(0, cfg.vocab, (B, 1), generator=torch.Generator().manual_seed(1)).cuda()
    assert m.discover(tok)
    n0 = _lib.launch_count()
    fused = m.step(tok).clone()
    assert _lib.launch_count() - n0 == 5 * cfg.layers + 1
    m.fuse_swiglu = False
    n0 = _lib.launch_count()
    plain = m.step(tok).clone()
    assert _lib.launch_count() - n0 == 7 * cfg.layers + 1
    rel_close(host(fused), host(plain), "logits fused vs unfused", 2e-3)
    # attention quantising its own output for o_proj (one launch) vs attention, then o_proj with its own prologue: same bits
    m.fuse_swiglu = True
    m.fuse_attn_quant = False
    sep = m.step(tok).clone()
    m.fuse_attn_quant = True
    aq = m.step(tok).clone()
    assert torch.equal(sep, aq), "attention-fused o_proj prologue must not change a bit"


def test_attention_decode_with_kv_cache():
    """RoPE + single-query attention kernel against the numpy restatement, with a non-empty KV cache and GQA."""
    from mixq_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    M, H, Hkv, D, L = 5, 8, 2, 128, 7
    rng = np.random.default_rng(0)
    qkv = rng.standard_normal((M, (H + 2 * Hkv) * D)).astype(np.float16)
    pk = rng.standard_normal((M, Hkv, L, D)).astype(np.float16)
    pv = rng.standard_normal((M, Hkv, L, D)).astype(np.float16)
    ref = LO.attention_decode(qkv, H, Hkv, D, 10000.0, pk, pv)
    cap = 16
    kc = torch.zeros(M, Hkv, cap, D, dtype=torch.float16, device="cuda")
    vc = torch.zeros_like(kc)
    kc[:, :, :L] = torch.from_numpy(pk).cuda()
    vc[:, :, :L] = torch.from_numpy(pv).cuda()
    out = torch.zeros(M, H * D, dtype=torch.float16, device="cuda")
    qkv_d = torch.from_numpy(qkv).cuda()
    _lib.check(lib.mixq_rope_attention_decode(qkv_d.data_ptr(), kc.data_ptr(), vc.data_ptr(), cap, L,
                                              out.data_ptr(), M, H, Hkv, D, 10000.0,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "attn")
    rel_close(host(out), ref, "attention", 5e-3)
    # the new key (rotated) and value were appended at position L
    k_new = LO.rope_rotate(qkv[:, H * D:(H + Hkv) * D].reshape(M, Hkv, D), L, 10000.0, D)
    rel_close(host(kc[:, :, L]), k_new, "appended key", 2e-3)
    bits_equal(host(vc[:, :, L]), qkv[:, (H + Hkv) * D:].reshape(M, Hkv, D), "appended value")
    # empty cache, q_len = 1: softmax over one key -> attention returns v (benchflops.py:124 regime)
    out0 = torch.zeros(M, H * D, dtype=torch.float16, device="cuda")
    _lib.check(lib.mixq_rope_attention_decode(qkv_d.data_ptr(), 0, 0, 0, 0, out0.data_ptr(), M, H, Hkv, D,
                                              10000.0, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "attn0")
    v = qkv[:, (H + Hkv) * D:].reshape(M, Hkv, 1, D).repeat(H // Hkv, 2).reshape(M, H * D)
    bits_equal(host(out0), v, "attention over a single key == v")


@pytest.mark.parametrize("M,H,Hkv,D,L,n,bit", [(5, 8, 2, 128, 7, 9, 8), (33, 32, 32, 128, 0, 41, 8), (4, 4, 1, 64, 3, 0, 8), (6, 8, 8, 128, 0, 5, 4)])
def test_attention_quant_fused_matches_separate_prologue(M, H, Hkv, D, L, n, bit):
    """mixq_rope_attention_decode_quant == mixq_rope_attention_decode followed by the oracle's ExtractOutliersAndSetToZeros +
    FindRowScale (linear.py:187-193) on that attention output: q_x / x_scale / gathered outliers / zeroed fp16 copy bit-exact."""
    from mixq_b200 import _lib
    import ctypes as C
    from oracle import mixq_oracle as O
    lib = _lib.load()
    rng = np.random.default_rng(M + H + n)
    qkv = rng.standard_normal((M, (H + 2 * Hkv) * D)).astype(np.float16)
    cols = np.sort(rng.permutation(H * D)[:n]).astype(np.int32)
    vcols = (H + Hkv) * D   # make some value channels "massive" so that the gathered outlier columns are real outliers
    qkv[:, vcols:] = (qkv[:, vcols:].astype(np.float32) * np.where(rng.random(Hkv * D) < 0.02, 30.0, 1.0)).astype(np.float16)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cap = 16
    mk = lambda: (torch.from_numpy(rng.standard_normal((M, Hkv, cap, D)).astype(np.float16)).cuda() if L else None)
    kc, vc = mk(), mk()
    kc2, vc2 = (kc.clone(), vc.clone()) if L else (None, None)
    p = lambda t: 0 if t is None else t.data_ptr()
    qkv_d = torch.from_numpy(qkv).cuda()
    out = torch.zeros(M, H * D, dtype=torch.float16, device="cuda")
    _lib.check(lib.mixq_rope_attention_decode(qkv_d.data_ptr(), p(kc), p(vc), cap if L else 0, L, out.data_ptr(), M, H, Hkv, D,
                                              10000.0, st), "attn")
    x = host(out).copy()
    ao_ref = O.extract_outliers_and_set_to_zeros(cols, x) if n else None
    q_ref, xs_ref = O.find_row_scale(x, bit)
    capo = max(64, (n + 63) // 64 * 64)
    ao = torch.zeros(M, capo, dtype=torch.float16, device="cuda")
    q_x = torch.zeros(M, H * D, dtype=torch.int8, device="cuda")
    xs = torch.zeros(M, dtype=torch.float16, device="cuda")
    ind = torch.from_numpy(cols).cuda()
    for keep_fp16 in (True, False):
        out2 = torch.full((M, H * D), 7.0, dtype=torch.float16, device="cuda")
        q_x.zero_(); xs.zero_(); ao.zero_()
        _lib.check(lib.mixq_rope_attention_decode_quant(qkv_d.data_ptr(), p(kc2), p(vc2), cap if L else 0, L,
                                                        out2.data_ptr() if keep_fp16 else 0, M, H, Hkv, D, 10000.0,
                                                        ind.data_ptr() if n else 0, n, ao.data_ptr(), capo, q_x.data_ptr(),
                                                        xs.data_ptr(), bit, st), "attn_quant")
        torch.cuda.synchronize()
        bits_equal(host(q_x), q_ref, "q_x")
        bits_equal(host(xs), xs_ref.reshape(-1), "x_scale")
        if n:
            bits_equal(host(ao)[:, :n], ao_ref, "activation_outliers")
        if keep_fp16:
            bits_equal(host(out2), x, "fp16 copy with the outlier columns zeroed")
    if L:
        bits_equal(host(kc2), host(kc), "k cache append")
        bits_equal(host(vc2), host(vc), "v cache append")


@pytest.mark.parametrize("bit", [8, 4])
def test_auto_from_quantized_model_surface(bit, tmp_path):
    """basic_quant_mix.py -> AutoForCausalLM.from_quantized -> benchflops.py with the reference's names (auto.py:42-53,
    base.py:162-229, llama.py:9-22, attn.py:206-278): a checkpoint directory written in the reference's layout is loaded into
    QuantAttentionFused / MixLlamaMLP / FasterTransformerRMSNorm modules; `model(input_ids, use_cache=True).logits` must agree
    with the decode harness (itself pinned to the oracle above) through the discovery calls and in steady state."""
    import json
    from mixq_b200 import AutoForCausalLM, MixLibCache
    from mixq_b200.attn import QuantAttentionFused
    from mixq_b200.llama import CONFIGS, LlamaDecoder
    cfg = CONFIGS["tiny"]
    B = 24
    m = LlamaDecoder(cfg, batch=B, bit=bit, seed=3, outlier_frac=0.02)
    d = str(tmp_path / "ckpt")
    m.save_quantized(d)
    assert json.load(open(f"{d}/quant_config.json"))["w_bit"] == bit and json.load(open(f"{d}/config.json"))["model_type"] == "llama"
    model = AutoForCausalLM.from_quantized(d, "", fuse_layers=True, mix=True, cache=MixLibCache(inputdim=B, bit=bit), batch_size=B)
    assert isinstance(model.layers[0].self_attn, QuantAttentionFused) and model.layers[0].self_attn.cache_batch_size == B
    tok = torch.randint(0, cfg.vocab, (B, 1), generator=torch.Generator().manual_seed(0)).cuda()
    for call in range(4):
        la = m.step(tok)
        if call == 1:
            m.discovered = True
        out = model(tok, use_cache=True)              # benchflops.py:124: no past_key_values -> an empty KV cache every call
        assert tuple(out.logits.shape) == (B, 1, cfg.vocab) and out[0] is out.logits
        for L, Lm in zip(m.layers, model.layers):
            bits_equal(host(L["W_pack"].ind), host(Lm.self_attn.W_pack.ind), f"call {call} W_pack outlier set")
            bits_equal(host(L["down_proj"].ind), host(Lm.mlp.down_proj_.ind), f"call {call} down_proj outlier set")
        rel_close(host(out.logits[:, 0]), host(la), f"call {call} logits", 2e-3)
    # KV cache: three tokens one by one on the module-owned cache (own decode kernel) == the three tokens as one prefill (SDPA)
    seq = torch.randint(0, cfg.vocab, (B, 3), generator=torch.Generator().manual_seed(1)).cuda()
    big = AutoForCausalLM.from_quantized(d, "", fuse_layers=True, mix=True, cache=MixLibCache(inputdim=3 * B, bit=bit), batch_size=B)
    for La, Lb in zip(model.layers, big.layers):      # the same discovered outlier state on both
        for a, b in ((La.self_attn.W_pack, Lb.self_attn.W_pack), (La.self_attn.o_proj, Lb.self_attn.o_proj),
                     (La.mlp.up_proj_, Lb.mlp.up_proj_), (La.mlp.gate_proj_, Lb.mlp.gate_proj_), (La.mlp.down_proj_, Lb.mlp.down_proj_)):
            assert not a.add_outliers or a is La.mlp.gate_proj_
            if a._n_ind:
                b.weight_cache = a.weight_cache
                b.ind = a.ind
            b.add_outliers = a.add_outliers
            b.forward_without_precondition_len = a.forward_without_precondition_len
    out = model(seq[:, :1], use_cache=True)
    for t in (1, 2):
        out = model(seq[:, t:t + 1], use_cache=True, past_key_values=out.past_key_values)
    assert model.layers[0].self_attn.start_pos == 3
    pre = big(seq, use_cache=True)
    rel_close(host(out.logits[:, -1]), host(pre.logits[:, -1]), "decode with KV cache vs prefill", 2e-2)
    g = model.generate(seq[:, :1], max_new_tokens=3)
    assert tuple(g.shape) == (B, 4)

"""CPU: host-side logic of the product (no kernels are launched) and the C-ABI surface."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from mixq_b200 import _lib
from mixq_b200.linear import MixLinear_GEMM, pack_to_i4

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Lin:
    def __init__(self, W, b=None):
        self.weight = torch.nn.Parameter(torch.from_numpy(W.copy()), requires_grad=False)
        self.bias = None if b is None else torch.nn.Parameter(torch.from_numpy(b.copy()), requires_grad=False)
        self.out_features, self.in_features = W.shape


class _Cache:
    sigma = torch.tensor([[6.0]], dtype=torch.float16)
    stop = 2


@pytest.mark.parametrize("case", ["w8_unfused", "w8_unfused_bias", "w4_unfused"])
def test_from_linear_matches_reference(golden, case):
    """The product's from_linear (offline torch arithmetic, any device) vs the reference's, bit for bit."""
    d = golden(case)
    bit = int(d["bit"])
    ls = torch.from_numpy(d["layer_scales"]) if bit == 4 else None
    q = MixLinear_GEMM.from_linear(_Lin(d["W"], d["bias"] if "bias" in d.files else None), bit, cache=_Cache(),
                                   layer_scales=ls, dev="cpu", fp_features_num=int(d["fp"]))
    assert np.array_equal(q.q_weight.numpy(), d["q_weight"])
    assert np.array_equal(q.scale_col.numpy().view(np.uint16), d["scale_col"].view(np.uint16))
    if bit == 4:
        assert np.array_equal(q.ind.numpy(), d["ind0"])
        assert np.array_equal(q.weight_cache.numpy().view(np.uint16), d["weight_cache0"].view(np.uint16))
        assert q.forward_without_precondition_len == int(d["fp"])
    else:
        assert q.ind.shape[0] == 0 and q.weight_cache is None
    if "bias" in d.files:
        assert np.array_equal(q.bias.numpy().view(np.uint16), d["bias"].view(np.uint16))
    # the caller's weights are left intact
    assert np.array_equal(d["W"].view(np.uint16), np.asarray(_Lin(d["W"]).weight.numpy()).view(np.uint16))


def test_pack_to_i4_layout():
    q = torch.arange(-8, 8, dtype=torch.int8).reshape(2, 8)
    p = pack_to_i4(q)
    assert p.dtype == torch.uint8 and tuple(p.shape) == (2, 4)
    assert int(p[0, 0]) == ((16 - 8) | ((16 - 7) << 4))


def test_ind_weight_cache_views_and_setters():
    q = MixLinear_GEMM(128, 32, False, "cpu", 8, cache=_Cache())
    assert q.ind.dtype == torch.int32 and q.ind.shape[0] == 0 and q.weight_cache is None
    q.weight_cache = torch.ones(32, 70, dtype=torch.float16)
    q.ind = torch.arange(70)
    assert q.ind.shape[0] == 70 and tuple(q.weight_cache.shape) == (32, 70)
    assert q._wc_buf.shape[1] % 64 == 0 and q._wc_buf.shape[1] >= 70     # TMA pitch: multiples of 64 fp16
    assert q._wc_buf.stride(0) * 2 % 16 == 0
    with pytest.raises(ValueError):
        q.ind = torch.arange(129)
    with pytest.raises(NotImplementedError):
        MixLinear_GEMM(64, 32, False, "cpu", 8, weight_only=True)
    with pytest.raises(ValueError):
        MixLinear_GEMM(64, 32, False, "cpu", 3)


def test_no_cpu_path():
    """The product path must fail loudly instead of computing on the CPU."""
    from mixq_b200.cache import MixLibCache
    cache = MixLibCache(inputdim=8, device="cpu")
    q = MixLinear_GEMM(64, 32, False, "cpu", 8, cache=cache)
    q.add_outliers = False
    with pytest.raises(_lib.MixqError):
        q(torch.zeros(4, 64, dtype=torch.float16), None, True)
    from mixq_b200 import mixlib
    with pytest.raises(_lib.MixqError):
        mixlib.FindRowScale(torch.zeros(4, 64, dtype=torch.float16), torch.zeros(4, 1, dtype=torch.float16), 4, 64, 8)
    from mixq_b200.norm import FasterTransformerRMSNorm
    with pytest.raises(_lib.MixqError):
        FasterTransformerRMSNorm(torch.ones(64))(torch.zeros(4, 64, dtype=torch.float16))


def test_library_missing_is_loud(monkeypatch, tmp_path):
    monkeypatch.setenv("MIXQ_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.MixqError, match="no CPU or PyTorch fallback"):
        _lib.load()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mixq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mixq_[a-z0-9_]+)\s*\(", src)))


def test_c_abi_exports_every_declared_symbol():
    """include/mixq.h <-> libmixq_sm100.so <-> the ctypes table: no drift, no compute call (no GPU here)."""
    lib = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mixq.h but not exported"
    assert set(_lib.SIGNATURES) == set(declared), set(_lib.SIGNATURES) ^ set(declared)
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared) <= exported
    assert lib.mixq_version() == 100
    assert lib.mixq_launch_count() == 0
    # argument validation happens before any CUDA call
    assert lib.mixq_set_tile_n(100) == -1
    assert b"tile_n" in lib.mixq_last_error()
    assert lib.mixq_set_tile_n(0) == 0


def test_linear_args_struct_layout_matches_header():
    """sizeof/field order of mixq_linear_args as the C compiler sees it vs the ctypes mirror."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "mixq.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(mixq_linear_args), offsetof(mixq_linear_args, q_weight),
         offsetof(mixq_linear_args, ind), offsetof(mixq_linear_args, q_x), offsetof(mixq_linear_args, residual),
         offsetof(mixq_linear_args, y), offsetof(mixq_linear_args, tile_n), offsetof(mixq_linear_args, y_peer),
         offsetof(mixq_linear_args, peer_cols));
  return 0;
}'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    A = _lib.LinearArgs
    want = [ctypes.sizeof(A), A.q_weight.offset, A.ind.offset, A.q_x.offset, A.residual.offset, A.y.offset, A.tile_n.offset,
            A.y_peer.offset, A.peer_cols.offset]
    assert got == want


def test_allreduce_args_struct_layout_and_validation():
    """mixq_allreduce_args (the tensor-parallel exchange) as the C compiler sees it vs the ctypes mirror; bad arguments are
    rejected before any CUDA call."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "mixq.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(mixq_allreduce_args), offsetof(mixq_allreduce_args, partial1),
         offsetof(mixq_allreduce_args, flags), offsetof(mixq_allreduce_args, result1), offsetof(mixq_allreduce_args, epoch),
         offsetof(mixq_allreduce_args, out), offsetof(mixq_allreduce_args, n), offsetof(mixq_allreduce_args, buf));
  return 0;
}'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    A = _lib.AllReduceArgs
    want = [ctypes.sizeof(A), A.partial1.offset, A.flags.offset, A.result1.offset, A.epoch.offset, A.out.offset, A.n.offset,
            A.buf.offset]
    assert got == want
    lib = _lib.load()
    a = A()
    a.world, a.rank, a.n = 1, 0, 64            # a single rank has nothing to exchange
    assert lib.mixq_allreduce_residual(ctypes.byref(a), None) != 0
    assert b"all-reduce" in lib.mixq_last_error()
    a.world, a.n = 2, 12                        # n % 8 != 0
    assert lib.mixq_allreduce_residual(ctypes.byref(a), None) != 0


def test_exchange_args_struct_layouts_and_validation():
    """mixq_exchange_finish_args / mixq_mc_allreduce_args (the fused and the NVLS exchanges) vs their ctypes mirrors."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "mixq.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(mixq_exchange_finish_args), offsetof(mixq_exchange_finish_args, result),
         offsetof(mixq_exchange_finish_args, mc_result), offsetof(mixq_exchange_finish_args, flags),
         offsetof(mixq_exchange_finish_args, epoch), offsetof(mixq_exchange_finish_args, residual), offsetof(mixq_exchange_finish_args, rank));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(mixq_mc_allreduce_args), offsetof(mixq_mc_allreduce_args, partial_off),
         offsetof(mixq_mc_allreduce_args, flags_off), offsetof(mixq_mc_allreduce_args, epoch), offsetof(mixq_mc_allreduce_args, n),
         offsetof(mixq_mc_allreduce_args, buf));
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(mixq_exchange_poll_args), offsetof(mixq_exchange_poll_args, result),
         offsetof(mixq_exchange_poll_args, mc_result), offsetof(mixq_exchange_poll_args, reset),
         offsetof(mixq_exchange_poll_args, residual), offsetof(mixq_exchange_poll_args, M), offsetof(mixq_exchange_poll_args, one_shot));
  return 0;
}'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    A, B = _lib.ExchangeFinishArgs, _lib.McAllReduceArgs
    want = [ctypes.sizeof(A), A.result.offset, A.mc_result.offset, A.flags.offset, A.epoch.offset, A.residual.offset, A.rank.offset,
            ctypes.sizeof(B), B.partial_off.offset, B.flags_off.offset, B.epoch.offset, B.n.offset, B.buf.offset]
    P = _lib.ExchangePollArgs
    want += [ctypes.sizeof(P), P.result.offset, P.mc_result.offset, P.reset.offset, P.residual.offset, P.M.offset, P.one_shot.offset]
    assert got == want
    lib = _lib.load()
    pa = P()
    pa.world, pa.N, pa.M = 3, 4096, 8
    assert lib.mixq_exchange_finish_poll(ctypes.byref(pa), None) != 0 and b"exchange_finish_poll" in lib.mixq_last_error()
    assert lib.mixq_exchange_finish_poll_quant(ctypes.byref(pa), None, 0.0, None, 0, None, 0, None, None, 8, None) != 0
    a = A()
    a.world, a.N, a.M = 3, 4096, 8            # N % (8 * world) != 0
    assert lib.mixq_exchange_finish(ctypes.byref(a), None) != 0 and b"exchange_finish" in lib.mixq_last_error()
    b = B()
    b.world = 1
    assert lib.mixq_allreduce_multicast(ctypes.byref(b), None) != 0


def _plan(M, N, K, bit=8, n=0, pair=0, tile=0, sms=148):
    lib = _lib.load()
    pl = _lib.LinearPlan()
    rc = lib.mixq_plan_linear(M, N, K, bit, n, pair, tile, sms, ctypes.byref(pl))
    return rc, pl


def test_launch_plan_headline_shapes():
    """The plans the Llama-2-7B batch-512 step runs with: tile widths, k-atoms, TMEM plan."""
    rc, p = _plan(512, 12288, 4096, n=41)                       # W_pack: one 256 x 352 tile per CTA pair
    assert rc == 0 and (p.two_cta, p.tile_w, p.k_atoms, p.tiles, p.tiles_per_unit) == (1, 352, 1, 70, 1)
    assert (p.acc_slots, p.passes, p.pass_cols, p.pass_buffers, p.tmem_cols) == (1, 6, 64, 2, 480)
    rc, p = _plan(512, 4096, 4096, n=41)                        # o_proj: narrow tile, two k-atoms per TMA op
    assert rc == 0 and (p.two_cta, p.tile_w, p.k_atoms, p.nstages, p.passes, p.tmem_cols) == (1, 128, 2, 4, 1, 256)
    rc, p = _plan(512, 4096, 11008, n=110)                      # down_proj: two resident outlier k-blocks still fit
    assert rc == 0 and (p.tile_w, p.k_atoms, p.nstages) == (128, 2, 4)
    rc, p = _plan(512, 11008, 4096, n=41, pair=1)               # SwiGLU pair: 2 x 11008 columns, two tiles per CTA pair
    assert rc == 0 and (p.tile_w, p.k_atoms, p.tiles, p.tiles_per_unit, p.acc_slots) == (320, 1, 138, 2, 1)
    assert (p.passes, p.pass_cols, p.pass_buffers) == (4, 96, 2)
    rc, p = _plan(32, 4096, 4096, n=41)                         # C1: M <= 128 runs the 1-CTA kernel
    assert rc == 0 and p.two_cta == 0 and p.tile_w in (128, 256) and p.units == 148
    rc, p = _plan(512, 12288, 4096, bit=4, n=128)               # W4: the same 2-CTA kernel, 3 main stages + the packed-row ring
    assert rc == 0 and p.two_cta == 1 and p.nstages == 3 and p.acc_slots == 1
    rc, _ = _plan(64, 11008, 4096, pair=1)                      # the pair launch needs the 2-CTA kernel
    assert rc != 0 and b"pair" in _lib.load().mixq_last_error()


def test_split_k_plan():
    """The planner's split-K choices for the small-M shapes (host-only; csrc/mixq_api.cu: plan_gemm's cost model fitted to
    profiles/r02_trace_splitk_v3.log): long-K few-tile launches split, the rest — and everything without a workspace — do not."""
    lib = _lib.load()
    ws = 16 * 1024 * 1024
    S = lambda M, N, K, w=ws: lib.mixq_plan_split_k(M, N, K, 8, 41, 148, w)
    assert S(128, 1280, 8192) == 4 and S(128, 3584, 8192) == 4            # C5 W_pack / up / gate per rank: 10 and 28 tiles
    assert S(128, 8192, 3584) == 1 and S(128, 8192, 1024) == 1            # 64 tiles, short K: the fixed cost would not pay
    assert S(32, 4096, 4096) == 1 and S(32, 4096, 11008) == 4             # C1: 4096^2 measured slower split, down_proj faster
    assert S(512, 4096, 4096) == 1                                         # M > 128: the 2-CTA kernel never splits
    assert S(128, 1280, 8192, 0) == 1                                      # no workspace
    assert S(128, 1280, 8192, 4096 + 10 * 65536) == 2                      # a small workspace caps the factor
    assert lib.mixq_plan_split_k(128, 1280, 8190, 8, 0, 148, ws) < 0


def test_launch_plan_invariants_over_all_configs():
    """Every Linear shape of BASELINE.json's configs (7B / 8B / 70B, TP 1..8, batch 32..512, 0..200 outlier columns) gets a
    plan that fits the hardware: TMEM <= 512 columns, pipeline <= 192 KB, resident outlier stages, passes cover the tile."""
    models = {"7b": (4096, 11008, 32, 32), "8b": (4096, 14336, 32, 8), "70b": (8192, 28672, 64, 8)}
    seen = 0
    for H, I, heads, kv in models.values():
        D = H // heads
        for tp in (1, 2, 4, 8):
            shapes = [((heads + 2 * kv) * D // tp, H, 0), (H, heads * D // tp, 0), (I // tp, H, 0), (I // tp, H, 1), (H, I // tp, 0)]
            for N, K, pair in shapes:
                if K % 16 or N % 16:
                    continue
                for M in (32, 128, 129, 256, 512):
                    for n in (0, 41, 110, 128, 200):
                        if n > K:
                            continue
                        rc, p = _plan(M, N, K, n=n, pair=pair)
                        if pair and M <= 128:
                            assert rc != 0
                            continue
                        if pair and rc != 0:            # too many outlier columns for the resident stages: documented limit
                            assert n > 128
                            continue
                        assert rc == 0, (M, N, K, n, pair, _lib.load().mixq_last_error())
                        seen += 1
                        assert 1 <= p.tmem_cols <= 512, (M, N, K, n, pair, p.tmem_cols)
                        assert p.tiles >= 1 and p.tiles_per_unit == -(-p.tiles // p.units)
                        assert p.passes * p.pass_cols >= p.tile_w and p.pass_buffers in (1, 2)
                        if p.two_cta:
                            assert M > 128 and p.tile_w % 32 == 0 and 128 <= p.tile_w <= (448 if n else 512)
                            assert p.stage_bytes == p.k_atoms * (16384 + p.tile_w // 2 * 128)
                            assert 2 <= p.nstages <= 8 and p.nstages * p.stage_bytes <= 192 * 1024
                            assert -(-n // 64) <= p.nstages - 1              # outlier k-blocks stay resident during the passes
                            assert p.pass_cols % 32 == 0
                            if p.k_atoms == 2:
                                assert K % 128 == 0 and p.tile_w <= 256 and p.nstages >= 3
                            if p.acc_slots == 2:
                                assert p.tiles_per_unit > 1 and 2 * p.tile_w + (64 if n else 0) <= 512
                        else:
                            assert p.tile_w in (128, 256) and p.k_atoms == 1
    assert seen > 500


def test_llama_accounting_formulas():
    """per-step algorithmic work of Llama-2-7B at M=512 = 6.63 TFLOP / 8.77 GB."""
    M, H, I, L = 512, 4096, 11008, 32
    shapes = [(3 * H, H), (H, H), (I, H), (I, H), (H, I)]
    fl = sum(2 * M * n * k for n, k in shapes) * L
    by = sum(n * k + 2 * M * k + 2 * M * n + 2 * n for n, k in shapes) * L
    assert abs(fl / 1e12 - 6.63) < 0.01
    assert abs(by / 1e9 - 8.77) < 0.3


@pytest.mark.parametrize("tag,bit,bias", [("w8", 8, False), ("w8_bias", 8, True), ("w4", 4, False)])
def test_module_state_dict_matches_reference_manifest(tag, bit, bias):
    """module.state_dict() has exactly the reference class's keys / dtypes / shapes (tests/golden/state_manifest.json,
    recorded from /root/reference/mixquant/modules/linear.py:27-87) and load_state_dict round-trips — for bit 4 that
    includes the registered `weight_cache` / `ind` buffers (what models/base.py's load_checkpoint_in_model sets by name)."""
    import json
    import os
    man = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "state_manifest.json")))[tag]
    K, N = man["in_features"], man["out_features"]
    q = MixLinear_GEMM(K, N, bias, "cpu", bit, cache=_Cache())
    sd = q.state_dict()
    assert sorted(sd) == sorted(man["state"])
    for k, (dt, shape) in man["state"].items():
        assert str(sd[k].dtype) == "torch." + dt and list(sd[k].shape) == shape, k
    g = torch.Generator().manual_seed(3)
    src = {k: (torch.randint(0, 100, v.shape, generator=g).to(v.dtype)) for k, v in sd.items()}
    q2 = MixLinear_GEMM(K, N, bias, "cpu", bit, cache=_Cache())
    res = q2.load_state_dict(src, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in q2.state_dict().items():
        assert torch.equal(v, src[k]), k
    if bit == 4:
        assert torch.equal(q2.ind, src["ind"]) and torch.equal(q2.weight_cache, src["weight_cache"])
        bad = dict(src)
        del bad["ind"]
        with pytest.raises(RuntimeError):
            MixLinear_GEMM(K, N, bias, "cpu", bit, cache=_Cache()).load_state_dict(bad, strict=True)

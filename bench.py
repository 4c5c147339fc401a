#!/usr/bin/env python
"""bench.py — the headline metric of BASELINE.json on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mixq|reference] [--model llama-2-7b] [--batch 512] [--bit 8]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

metric: Llama-2-7B W8A8O16 decode tokens/s at batch 512 (/root/reference/benchflops.py:96-128, :222 — every step is one
independent [512,1] forward; "decode length 1024" there is the number of timed iterations, not a KV length), random-init
weights with ~1 % forced outlier channels, synthetic tokens.  A "step" = one pass of the MixLinear hot path over one batch:
the whole decode step (32 layers x 5 MixLinears + norms + attention glue + fp16 lm_head), replayed from one CUDA graph.

N > 1 shards every Linear column-/row-wise over N ranks; the one exchange per row-parallel Linear is this library's
all-reduce + residual kernel over NVLink peer memory (MIXQ_TP_EXCHANGE=nccl: NCCL all-reduce + add).  Total work fixed:
"scaling": "strong".

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how each field is obtained.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mixq", choices=["mixq", "reference"])
    ap.add_argument("--model", default="llama-2-7b")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--bit", type=int, default=8, choices=[4, 8])
    ap.add_argument("--layers", type=int, default=None, help="debug: fewer layers (the number is reported, never default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-lm-head", action="store_true",
                    help="N > 1: keep the fp16 lm_head replicated (default: vocab-parallel, every rank computes its slice of the logits)")
    ap.add_argument("--kv-len", type=int, default=-1,
                    help="KV-cache variant of the metric: every sequence already holds this many tokens (-1: the largest power of "
                         "two minus one, up to 1023, whose cache fits beside the model at N=1; 0: off)")
    return ap.parse_args()


def peaks():
    """Roofline denominators.  HBM / bf16: MEASURED_PEAKS.json (driver-written) else the profiling guide's fallback.  INT8 tensor
    pipe: profiles/r02_int8_peak.json, measured on this pool's B200 by tools/int8_peak.py (torch._int_mm 8192^3 and this
    library's own prologue-less kernel; burst = best single launch, sustained = 4 s back to back under the power cap)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        pk = dict(hbm_gbs=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                  source="measured")
    else:
        pk = dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")
    q = os.path.join(ROOT, "profiles", "r02_int8_peak.json")
    if os.path.exists(q):
        d = json.load(open(q))
        pk.update(int8_burst=d["int8_tops_burst"], int8_sustained=d["int8_tops_sustained"],
                  int8_source="measured int8 (profiles/r02_int8_peak.json: torch._int_mm 8192^3 / own kernel without prologue)")
    else:
        pk.update(int8_burst=2.0 * pk["bf16"], int8_sustained=2.0 * pk["bf16_sustained"], int8_source="2 x bf16 (no int8 measurement found)")
    return pk


# ----------------------------------------------------------------------------------------------- CPU arms
def cpu_layer_sample(cfg_name, batch, bit, budget_s=12.0, build_only=False):
    """One decoder layer's five MixLinears + both fused norms, steady state, through the CPU oracle (numpy/BLAS, all host
    threads).  Returns (seconds per layer, cores, description).  This is bench.py's `cpu_baseline` / `--impl reference`
    leg — the one place outside tests/ that executes oracle/ (as the thing timed beside the product, never as the product)."""
    import numpy as np
    from mixq_b200.llama import CONFIGS
    from oracle import mixq_oracle as O
    cfg = CONFIGS[cfg_name]
    H, I, D = cfg.hidden, cfg.intermediate, cfg.head_dim
    rng = np.random.default_rng(0)
    cache = O.MixLibCacheOracle(batch, 6, bit)
    n_out = max(1, round(0.01 * H))

    def lin(n, k, b):
        q = rng.integers(-127, 128, (n, k), dtype=np.int8) if b == 8 else rng.integers(0, 256, (n, k // 2), dtype=np.uint8)
        s = (rng.random((1, n)) * 2e-4 + 1e-4).astype(np.float16)   # keeps the synthetic activations inside fp16
        m = O.MixLinearOracle(q, s, b, None, cache)
        m.add_outliers = False
        m.ind = np.sort(rng.permutation(k)[:max(1, round(0.01 * k))]).astype(np.int32)
        m.weight_cache = O.weight_cache_columns(q, s, m.ind, b)
        m.forward_without_precondition_len = len(m.ind)
        return m
    qkv_n = (cfg.heads + 2 * cfg.kv_heads) * D
    W_pack, o_proj = lin(qkv_n, H, bit), lin(H, cfg.heads * D, 8)
    up, gate, down = lin(I, H, bit), lin(I, H, bit), lin(H, I, 8)
    gate.ind, gate.weight_cache = up.ind, O.weight_cache_columns(gate.q_weight, gate.scale_col, up.ind, bit)
    ln = np.ones(H, np.float16)
    h = rng.standard_normal((batch, H)).astype(np.float16)
    _ = n_out

    def one_layer(h):
        out, ao, q_x, xs = O.rmsnorm_extract_outliers(h, ln, cfg.eps, W_pack.ind, bit)
        cache.activation_outliers, cache.q_xcache = ao, q_x
        cache.x_scale[:batch] = xs
        qkv = W_pack.forward(out, cache)
        attn = qkv[:, -cfg.heads * D:] if cfg.kv_heads == cfg.heads else np.repeat(
            qkv[:, -cfg.kv_heads * D:].reshape(batch, cfg.kv_heads, D), cfg.heads // cfg.kv_heads, 1).reshape(batch, -1)
        h = (o_proj.forward(np.ascontiguousarray(attn).copy(), None, True).astype(np.float32) + h).astype(np.float16)
        out, ao, q_x, xs = O.rmsnorm_extract_outliers(h, ln, cfg.eps, up.ind, bit)
        cache.activation_outliers, cache.q_xcache = ao, q_x
        cache.x_scale[:batch] = xs
        u = up.forward(out, cache)
        g = gate.forward_without_precondition_fused_silu(out, cache)
        g = (g.astype(np.float32) * u.astype(np.float32)).astype(np.float16)
        return (down.forward(g, None, True).astype(np.float32) + h).astype(np.float16)

    desc = f"1 of {cfg.layers} decoder layers (5 MixLinears + 2 fused norms, M={batch}), numpy on the host BLAS (exact int GEMM in fp32 K-chunks)"
    if build_only:
        return (lambda: one_layer(h)), os.cpu_count(), desc
    one_layer(h)   # warm BLAS
    t0 = time.perf_counter()
    reps = 0
    while True:
        one_layer(h)
        reps += 1
        el = time.perf_counter() - t0
        if (reps >= 3 and el > budget_s / 2) or el > budget_s or reps >= 20:
            break
    return el / reps, os.cpu_count(), desc + f", {reps} reps"


def torch_cpu_linear_sample(cfg_name, batch):
    """torch-CPU fp32 F.linear (the un-quantised Linear MixLinear replaces) on the five Linear shapes of one layer."""
    import torch
    from mixq_b200.llama import CONFIGS
    cfg = CONFIGS[cfg_name]
    H, I, D = cfg.hidden, cfg.intermediate, cfg.head_dim
    torch.set_num_threads(os.cpu_count())
    shapes = [((cfg.heads + 2 * cfg.kv_heads) * D, H), (H, H), (I, H), (I, H), (H, I)]
    tot = 0.0
    for n, k in shapes:
        x, w = torch.randn(batch, k), torch.randn(n, k)
        torch.nn.functional.linear(x, w)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            torch.nn.functional.linear(x, w)
            best = min(best, time.perf_counter() - t0)
        tot += best
    return tot


def run_reference(args):
    """`--impl reference`: the path on the host cores (the reference has no CPU implementation of its CUDA kernels and its
    mixlib/EETQ sources are not in the tree, so this is the oracle port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mixq_b200.llama import CONFIGS
    cfg = CONFIGS[args.model]
    times = []
    # each step = a bounded sample of the workload: ONE decoder layer on all host threads; the step time is that layer
    # time x the number of layers (every layer does identical work)
    run_once, cores, desc = cpu_layer_sample(args.model, args.batch, args.bit, build_only=True)
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        run_once()
        if i >= args.warmup:
            times.append((time.perf_counter() - t0) * cfg.layers)
    step_s = sum(times) / len(times)
    val = args.batch / step_s
    line = {
        "impl": "reference", "metric": "llama_decode_tokens_per_s", "value": val, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        # the GPU arm's workload, word for word; only `parallelism` says where it ran
        "config": {"workload": f"{args.model} W{args.bit}A{args.bit}O16 decode step, batch {args.batch}, q_len 1, empty KV cache "
                               "(benchflops.py:96-128), ~1% forced outlier channels",
                   "layers": cfg.layers, "global_batch": args.batch, "parallelism": "host cpu (rank 0)", "bit": args.bit},
        "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": desc + f", x{cfg.layers} layers extrapolated"},
        "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------- TP parity
def tp_parity_check(args, cfg, model, tok0, rank, world):
    """Before anything is timed at N > 1: the same seeded model at world_size 1 on rank 0 (steady state, same tokens).
    Gate (north star, <= 1e-2 relative on the fp16 result of the sharded Linears): the residual stream after decoder layer 0 —
    one row-parallel o_proj and one down_proj, each quantising its K-shard with its own per-row scale and summed across ranks by
    the exchange — against the single-GPU stream, and the column-parallel Linears (replicated input) must have found exactly
    the single-GPU outlier index sets.  Reported beside it: the relative difference of the final logits and the argmax
    agreement (per-shard row scales change the quantisation-noise realisation of 2 Linears per layer, so the difference
    accumulates over the depth: it is statistical, not bitwise, SURVEY.md section 7 "Hard parts").  Models too large to rebuild
    whole beside the shard (> 40 layers) are compared on their first 4 layers (same seeds => the same first layers)."""
    import torch
    import torch.distributed as dist
    from mixq_b200.llama import LlamaDecoder
    B = args.batch
    n_par = model.n_layers if model.n_layers <= 40 else 4
    tp_model = model
    if n_par != model.n_layers:     # every rank takes part: the truncated TP model has exchanges
        tp_model = LlamaDecoder(cfg, batch=B, bit=args.bit, seed=0, outlier_frac=0.01, rank=rank, world_size=world, layers=n_par)
        assert tp_model.discover(tok0)
    tp_model._rank_barrier()
    tp_model.probe_layer0 = True
    lt = tp_model.step(tok0)
    tp_model.probe_layer0 = False
    h0_tp = tp_model.hidden_after_layer0
    torch.cuda.synchronize()
    out = None
    if rank == 0:
        ref = LlamaDecoder(cfg, batch=B, bit=args.bit, seed=0, outlier_frac=0.01, rank=0, world_size=1, layers=n_par)
        assert ref.discover(tok0)
        ref.probe_layer0 = True
        lr = ref.step(tok0).float()
        h0 = ref.hidden_after_layer0.float()
        rel0 = float((h0_tp.float() - h0).norm() / h0.norm())
        rel = float((lt.float() - lr).norm() / lr.norm())
        col = ("W_pack", "up_proj", "gate_proj")
        col_same = all(torch.equal(a[k].ind, b[k].ind) for a, b in zip(tp_model.layers, ref.layers) for k in col)
        # row-parallel: this rank's shard of the single-GPU index set (informative: the shard sees its own x_scale)
        hits = tot = 0
        for a, b in zip(tp_model.layers, ref.layers):
            for k in ("o_proj", "down_proj"):
                ks = a[k].in_features
                mine = set((b[k].ind[(b[k].ind >= rank * ks) & (b[k].ind < (rank + 1) * ks)] - rank * ks).tolist())
                got = set(a[k].ind.tolist())
                hits += len(mine & got)
                tot += len(mine | got)
        out = {"rel": rel0, "tol": 1e-2, "ok": bool(rel0 <= 1e-2 and col_same), "what": "residual stream after decoder layer 0, TP vs 1 GPU",
               "rel_final_logits": rel, "column_parallel_outlier_sets_identical": bool(col_same),
               "row_parallel_outlier_set_overlap": (hits / tot if tot else 1.0), "layers_compared": n_par,
               "argmax_agreement": float((lt.argmax(-1) == lr.argmax(-1)).float().mean())}
        del ref, lr
        torch.cuda.empty_cache()
    if tp_model is not model:
        tp_model.close()
        del tp_model
    torch.cuda.synchronize()
    dist.barrier()
    return out


def time_exchanges(model, n):
    """Per-exchange device time the exchange ADDS to the step: one CUDA graph of n exchanges on the real buffers (every rank
    replays it at the same time).  For the fused exchange the reduce-scatter half rides inside the row-parallel GEMM's
    epilogue, so the graph holds layer 0's o_proj pushing its tiles + the finish kernel, and the same o_proj alone is timed
    and subtracted."""
    import torch
    h = torch.zeros((model.batch, model.cfg.hidden), dtype=torch.float16, device="cuda")
    fused = getattr(model.xchg, "fused", False)
    if fused:
        lin = model.layers[0]["o_proj"]
        x = torch.randn((model.batch, lin.in_features), device="cuda").half()

        def one(push=True):
            lin(x, None, True, push=model.xchg.push_targets() if push else None)
            if push:
                model.xchg.reduce(h)
    else:
        def one(push=True):
            model.xchg.next_partial()
            model.xchg.reduce(h)

    def timed(push):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        model._rank_barrier()
        with torch.cuda.stream(side):
            for _ in range(2):
                one(push)
        torch.cuda.current_stream().wait_stream(side)
        model._rank_barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                one(push)
        model._rank_barrier()
        g.replay()
        model._rank_barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * n)
        del g
        model._rank_barrier()
        return us
    us = timed(True)
    if fused:
        us -= timed(False)
    return us


# ----------------------------------------------------------------------------------------------- GPU arm
def run_mixq(args):
    import torch
    import torch.distributed as dist
    from mixq_b200 import _lib
    from mixq_b200.linear import MixLinear_GEMM
    from mixq_b200.llama import CONFIGS, LlamaDecoder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.load()   # fails loudly when the extension is missing: there is no fallback
    cfg = CONFIGS[args.model]
    B = args.batch

    model = LlamaDecoder(cfg, batch=B, bit=args.bit, seed=0, outlier_frac=0.01, rank=rank, world_size=world, layers=args.layers)
    gen = torch.Generator().manual_seed(0)
    n_tok_sets = 8
    host_tokens = [torch.randint(0, cfg.vocab, (B, 1), generator=gen).pin_memory() for _ in range(n_tok_sets)]
    host_next = torch.empty(B, dtype=torch.int64).pin_memory()
    tok0 = host_tokens[0].cuda()
    # the reference's first cache.stop (=2) forwards: online outlier discovery, host-synchronising (linear.py:200-226)
    ok = model.discover(tok0)
    assert ok, "outlier discovery did not converge after cache.stop calls"
    n0 = _lib.launch_count()
    model._rank_barrier()
    logits_steady = model.step(tok0)
    launches_per_step = _lib.launch_count() - n0
    tp_parity = tp_parity_check(args, cfg, model, tok0, rank, world) if world > 1 else None
    vocab_parallel = None
    if world > 1 and not args.full_lm_head and cfg.vocab % world == 0:
        # Megatron-style vocab-parallel head: every rank keeps vocab / N rows of the fp16 lm_head and computes its slice of the
        # logits; the next token is the best of the ranks' local maxima.  Checked here against the full head.
        model.shard_lm_head()
        model._rank_barrier()
        shard = model.step(tok0)
        v = cfg.vocab // world
        want = logits_steady[:, rank * v:(rank + 1) * v].float()
        rel_v = float((shard.float() - want).norm() / want.norm())
        same_tok = bool(torch.equal(model.argmax(shard), torch.argmax(model.gather_logits(shard), dim=-1)))
        vocab_parallel = {"rows_per_rank": v, "rel_vs_full_head_slice": rel_v, "distributed_argmax_equals_argmax_of_gathered": same_tok}
        assert rel_v <= 1e-3 and same_tok, vocab_parallel
    model.capture(tok0)
    dev_tokens = [t.cuda() for t in host_tokens]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, K graph replays, device-timed
    for i in range(args.warmup):
        model.replay(dev_tokens[i % n_tok_sets])
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        model.replay(dev_tokens[i % n_tok_sets])
        evs[i + 1].record()
    barrier()
    ms = evs[0].elapsed_time(evs[-1])
    per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps))
    ms_median = per_step[len(per_step) // 2]
    # ---- e2e: host tokens -> H2D -> step -> argmax -> D2H, every step, through the public call
    for i in range(2):
        model.replay(dev_tokens[0])
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        logits = model.replay(host_tokens[i % n_tok_sets])          # pinned host -> static device buffer (H2D)
        host_next.copy_(model.argmax(logits), non_blocking=True)   # result D2H (vocab-parallel head: + one tiny all-gather)
        torch.cuda.current_stream().synchronize()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_median], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_median = float(t[0]), float(t[1]), float(t[2])

    # ---- roofline of the dominant kernel (mixq_linear2_kernel: every MixLinear launch of the step): per-launch CUDA-event
    # time of each of the step's Linear launches, taken over all layers' distinct weights
    pk = peaks()
    pair = model.fuse_swiglu
    kinds = ["W_pack", "o_proj"] + (["swiglu_pair"] if pair else ["up_proj", "gate_proj"]) + ["down_proj"]
    h = torch.randn(B, cfg.hidden, device="cuda").half()
    per_kind = {}
    tot_t = tot_fl = tot_by = 0.0
    for kind in kinds:
        mods = [L["gate_proj" if kind == "swiglu_pair" else kind] for L in model.layers]
        ups = [L["up_proj"] for L in model.layers]
        K_in = mods[0].in_features
        xin = torch.randn(B, K_in, device="cuda").half()
        xw = xin.clone()

        def launch(i):
            m = mods[i]
            if kind in ("W_pack", "up_proj"):
                return m.forward_norm_fused(h, model.layers[0]["ln1"], cfg.eps)
            if kind == "swiglu_pair":
                return m.forward_swiglu_fused(ups[i], h, model.layers[0]["ln2"], cfg.eps)
            if kind == "gate_proj":
                return m.forward_without_preconditionFusedSilu(h, model.cache)
            if kind == "o_proj" and o_quantized:
                # as the step runs it: the attention kernel has quantised o_proj's input rows, no activation prologue left
                return m.forward_quantized(B, residual=h) if world == 1 else m.forward_quantized(B)
            return m(xw, None, True, residual=h) if world == 1 else m(xw, None, True)
        o_quantized = kind == "o_proj" and getattr(model, "fuse_attn_quant", False)
        if o_quantized:           # every layer's o_proj shares the module cache: one ordinary call leaves q_x / x_scale / outliers there
            mods[0](xw, None, True)
            n_max = max(m._n_ind for m in mods)          # (timing only: every layer reads the same gathered-outlier buffer)
            model.cache.activation_outliers = model.cache.ao_buffer(n_max)[:B, :n_max] if n_max else None
        if kind == "gate_proj":   # needs up_proj's q_x in the cache
            model.layers[0]["up_proj"].forward_norm_fused(h, model.layers[0]["ln2"], cfg.eps)
        # one CUDA graph holding this Linear of every layer (distinct weights: L2-cold, as in the step), so that the
        # events bracket device time only and not the Python launch path
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(min(3, len(mods))):
                launch(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gk = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gk):
            for i in range(len(mods)):
                launch(i)
        gk.replay()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        g0.record()
        for _ in range(reps):
            gk.replay()
        g1.record()
        torch.cuda.synchronize()
        del gk
        t_launch = g0.elapsed_time(g1) * 1e-3 / (reps * len(mods))
        m0 = mods[0]
        N_, n_ = m0.out_features, m0._n_ind
        nlin = 2 if kind == "swiglu_pair" else 1      # the pair launch holds two Linears (and writes one y)
        fl = 2.0 * B * N_ * K_in * nlin
        by = nlin * (N_ * K_in * m0.bit / 8 + 2 * N_ + 2 * n_ * N_) + 2 * B * K_in + 2 * B * N_
        per_kind[kind] = {"N": N_, "K": K_in, "n_outliers": n_, "linears": nlin, "us": t_launch * 1e6, "tflops": fl / t_launch / 1e12,
                          "gbs": by / t_launch / 1e9}
        tot_t += t_launch
        tot_fl += fl
        tot_by += by
    # burst figure for a region that ran at full clocks without the power cap, else the sustained one
    clk = sampler.result()
    at_full_clock = (clk["sm_mhz"] is not None and clk["sm_max_mhz"] and clk["sm_mhz"] >= 0.97 * clk["sm_max_mhz"]
                     and "sw_power_cap" not in clk["reasons"])
    int8_peak = pk["int8_burst"] if at_full_clock else pk["int8_sustained"]
    ach_tf, ach_gb = tot_fl / tot_t / 1e12, tot_by / tot_t / 1e9
    tensor_bound = (tot_fl / (int8_peak * 1e12)) >= (tot_by / (pk["hbm_gbs"] * 1e9))
    # DRAM bytes per launch from the committed `ncu --set full` capture of the same four launches (profiles/)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and world == 1 and args.model == "llama-2-7b" and args.bit == 8 and B == 512:
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["dram_bytes_per_launch_avg"], tj["source"]
    roofline = {
        "kernel": f"mixq_linear2_kernel ({len(kinds)} launches per layer: " + ", ".join(kinds) + ")",
        "bound": "tensor" if tensor_bound else "hbm",
        "achieved": ach_tf if tensor_bound else ach_gb, "peak": int8_peak if tensor_bound else pk["hbm_gbs"],
        "unit": "TFLOP/s" if tensor_bound else "GB/s",
        "frac": (ach_tf / int8_peak) if tensor_bound else (ach_gb / pk["hbm_gbs"]),
        "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch_avg": tot_by / len(kinds), "algorithmic_flops_per_launch_avg": tot_fl / len(kinds),
        "avg_launch_us": tot_t / len(kinds) * 1e6,
        "peak_source": f"{pk['int8_source']}, {'burst' if at_full_clock else 'sustained'} figure (timed region at "
                       f"{clk['sm_mhz']} MHz, reasons {clk['reasons']}); hbm_gbs {pk['hbm_gbs']} ({pk['source']})",
        "int8_peaks": {"burst": pk["int8_burst"], "sustained": pk["int8_sustained"]},
        "other_bound": {"tflops": ach_tf, "frac_int8": ach_tf / int8_peak, "gbs": ach_gb, "frac_hbm": ach_gb / pk["hbm_gbs"]},
        "per_linear": per_kind, "linear_share_of_step": tot_t * len(model.layers) / (ms * 1e-3 / args.steps),
    }

    # ---- the KV-cache variant ("also report a variant with a real growing KV cache"): the same step with
    # every sequence at position kv_len, attention over the cache through the library path the reference uses (flash-attn there,
    # torch SDPA here).  benchflops itself never carries a cache (benchflops.py:124), so this is reported beside the headline.
    kv_variant = None
    if world == 1 and args.kv_len != 0:
        per_tok = 2 * B * cfg.kv_heads * cfg.head_dim * 2 * len(model.layers)          # bytes per cached position
        free = torch.cuda.mem_get_info()[0] - (8 << 30)
        L = args.kv_len
        if L < 0:
            L = 1023
            while L > 0 and (L + 1) * per_tok > free:
                L = (L + 1) // 2 - 1
        if L > 0 and (L + 1) * per_tok <= free:
            model.graph = None
            model.alloc_kv(L + 1)
            model.kv_library_attention = True
            for kc, vc in model.kv:
                kc.normal_(0, 1)
                vc.normal_(0, 1)
            model.capture(tok0, past_len=L)
            for i in range(2):
                model.replay(dev_tokens[i])
            torch.cuda.synchronize()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ksteps = max(3, min(args.steps, 10))
            k0.record()
            for i in range(ksteps):
                model.replay(dev_tokens[i % n_tok_sets])
            k1.record()
            torch.cuda.synchronize()
            kms = k0.elapsed_time(k1) / ksteps
            kv_variant = {"past_len": L, "tokens_per_s": B / (kms * 1e-3), "ms_per_step": kms, "steps": ksteps,
                          "kv_cache_gb": (L + 1) * per_tok / 1e9,
                          "note": "same step with every sequence at position past_len; attention over the KV cache = torch SDPA "
                                  "(library, as the reference's flash-attn); HBM floor for reading the cache once: "
                                  f"{(L + 1) * per_tok / (pk['hbm_gbs'] * 1e9) * 1e3:.2f} ms"}
            model.graph = None
            model.kv = None
            model.kv_library_attention = False
            torch.cuda.empty_cache()
    exchange_us = None
    if world > 1 and model.xchg is not None:
        exchange_us = time_exchanges(model, 2 * len(model.layers))
        t = torch.tensor([exchange_us], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exchange_us = float(t[0])
    model_xchg = model.xchg is not None
    model_xchg_two_shot = bool(model.xchg.two_shot) if model_xchg else False
    if rank == 0:
        step_s = ms * 1e-3 / args.steps
        fl_step, by_step = model.algorithmic_work()
        line = {
            "metric": "llama_decode_tokens_per_s", "value": B / step_s, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_s * 1e3, "ms_per_step_median": ms_median,
            "value_median": B / (ms_median * 1e-3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int8" if args.bit == 8 else "int4->int8", "data": "synthetic",
            "config": {"workload": f"{args.model} W{args.bit}A{args.bit}O16 decode step, batch {B}, q_len 1, empty KV cache "
                                   "(benchflops.py:96-128), ~1% forced outlier channels",
                       "layers": len(model.layers), "global_batch": B, "parallelism": f"tp{world}", "bit": args.bit,
                       "l2": "weights (>= 6 GB per step) exceed the 126 MB L2: inputs larger than L2, no flush",
                       "outliers_layer0": {k: m._n_ind for k, m in model.layers[0].items() if isinstance(m, MixLinear_GEMM)},
                       "cuda_graph": True, "programmatic_dependent_launch": True,
                       "lm_head": ("fp16, replicated" if vocab_parallel is None else
                                   f"fp16, vocab-parallel: {vocab_parallel['rows_per_rank']} rows per rank, logits stay sharded; next token = best of "
                                   "the ranks' local maxima (e2e includes that exchange)"),
                       "exchange": (None if world == 1 else ((type(model.xchg).__name__ + " kernel, " + ("two-shot" if model_xchg_two_shot else "one-shot"))
                                                              if model_xchg else "nccl all-reduce + add"))},
            "e2e": {"value": B / (ms_e2e * 1e-3 / args.steps), "unit": "tokens/s", "h2d_bytes_per_step": B * 8,
                    "d2h_bytes_per_step": B * 8, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "tp_parity": tp_parity,
            "vocab_parallel_check": vocab_parallel,
            "kv_variant": kv_variant,
            "step_breakdown_us": {"linear_us": tot_t * len(model.layers) * 1e6,
                                  "exchange_us": None if exchange_us is None else exchange_us * 2 * len(model.layers),
                                  "per_exchange_us": exchange_us, "step_us": step_s * 1e6},
            "clocks": sampler.result(),
            "roofline": roofline,
            "step_work": {"linear_tflop": fl_step / 1e12, "linear_gb": by_step / 1e9,
                          "tflops_whole_step": fl_step / step_s / 1e12},
        }
        if world == 1 and not args.no_cpu_baseline:
            t_layer, cores, desc = cpu_layer_sample(args.model, B, args.bit)
            t_lin = torch_cpu_linear_sample(args.model, B)
            line["cpu_baseline"] = {"value": B / (t_layer * cfg.layers), "unit": "tokens/s", "cores": cores, "kind": "port",
                                    "sample": desc + f", x{cfg.layers} layers extrapolated",
                                    "torch_fp32_linear_tokens_per_s": B / (t_lin * cfg.layers)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL communicators that CUDA graphs still reference can hang in destroy_process_group (seen at N=2: the line was
        # printed, then the ranks never exited).  No collective follows the max-over-ranks reduction above, so every rank
        # drops its graphs and leaves on its own, without the NCCL teardown.
        model.graph = None
        del model
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_mixq(args)


if __name__ == "__main__":
    main()

"""Two or more ranks on one node: the exchange kernels (PeerExchange: all-reduce + residual over NVLink peer memory;
MulticastExchange: the same through the NVSwitch, NVLS multimem) against the same arithmetic in torch (fp32 accumulation in
rank order, one rounding to fp16, residual added as a separate fp16 op), eagerly and replayed from a CUDA graph; then the
tensor-parallel decoder step against the NCCL exchange AND against the same seeded model at world_size 1 (the single-GPU
product, itself pinned to the oracle by tests/test_gpu_module.py).  CHECK_KIND = peer | multicast (default: both, multicast
only where the platform maps it).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_peer_exchange.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixq_b200.llama import CONFIGS, LlamaDecoder  # noqa: E402
from mixq_b200.tp import MulticastExchange, PeerExchange, PushExchange  # noqa: E402


def check_kind(kind, rank, world):
    M, H = 512, 4096
    two_shot = {"0": False, "1": True}.get(os.environ.get("MIXQ_TP_TWO_SHOT"), None)
    if kind == "peer":
        ex = PeerExchange(M, H, rank, world, two_shot=two_shot)
    else:
        try:
            ex = MulticastExchange(M, H, rank, world)
        except Exception as e:
            if rank == 0:
                print(f"multicast exchange unavailable here: {type(e).__name__}: {e}")
            return None
    exact = kind == "peer"     # the switch's summation order is its own: <= 1 fp16 ulp of the sum against the rank-order sum
    kernel_copy = os.environ.get("CHECK_MEMCPY") != "1"       # a kernel (not a memcpy node) fills the partial buffer

    def reference(parts, res):                # == oracle.mixq_oracle.tp_exchange, restated in torch on the device
        acc = torch.zeros(M, H, dtype=torch.float32, device="cuda")
        for p_ in parts:                      # rank order, fp32 — what the kernel does (exact for 2 ranks, rounds beyond)
            acc += p_.float()
        ref = (acc.half().float() + res.float()).half()
        ref.sum_abs = acc.abs()          # magnitude of the sum before the residual add (for the ulp tolerance)
        return ref

    def compare(out, ref, what):
        if exact:
            bad = int((out != ref).sum())
        else:
            ulp = torch.maximum(torch.maximum(ref.float().abs(), ref.sum_abs), torch.full((), 2.0 ** -14, device="cuda")) * 2.0 ** -10
            bad = int(((out.float() - ref.float()).abs() > 2 * ulp).sum())
        assert bad == 0, f"rank {rank} {kind} {what}: {bad} of {out.numel()} elements differ, max |diff| {float((out.float() - ref.float()).abs().max()):.3e}"
        # every rank holds the same bits
        mine = out.contiguous().view(torch.int16).to(torch.int32)
        lo, hi = mine.clone(), mine.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), f"rank {rank} {kind} {what}: ranks disagree"

    def fill(dst, src_t):
        if kernel_copy:
            torch.add(src_t, 0, out=dst)
        else:
            dst.copy_(src_t)
    g = torch.Generator(device="cuda").manual_seed(1234)          # same stream of numbers on every rank
    for it in range(6):
        parts = [torch.randn(M, H, generator=g, device="cuda").half() for _ in range(world)]
        res = torch.randn(M, H, generator=g, device="cuda").half()
        fill(ex.next_partial(), parts[rank])
        out = ex.reduce(res).clone()
        compare(out, reference(parts, res), f"exchange {it}")
    # replayed from a graph (two exchanges per replay: the buffers alternate)
    src = [torch.zeros(M, H, dtype=torch.float16, device="cuda") for _ in range(2)]
    res = torch.randn(M, H, generator=g, device="cuda").half()
    outs = [None, None]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for j in range(2):
            fill(ex.next_partial(), src[j])
            ex.reduce(res)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for j in range(2):
            fill(ex.next_partial(), src[j])
            outs[j] = ex.reduce(res).clone()
    for rep in range(5):
        allp = [[torch.randn(M, H, generator=g, device="cuda").half() for _ in range(world)] for _ in range(2)]
        for j in range(2):
            src[j].copy_(allp[j][rank])
        gr.replay()
        torch.cuda.synchronize()
        for j in range(2):
            compare(outs[j], reference(allp[j], res), f"graph replay {rep} exchange {j}")
    del gr
    rel, rel1 = check_decoder(kind, rank, world)
    models = {}
    for m in models.values():
        m.close()
    ex.close()
    dist.barrier()
    if rank == 0:
        how = "bit-exact" if exact else "<= 1 fp16 ulp"
        print(f"{kind} exchange ok on {world} ranks ({'two' if ex.two_shot else 'one'}-shot, {'kernel' if kernel_copy else 'memcpy'} fill): "
              f"{how} vs the torch restatement and identical on all ranks (eager + graph replay); decoder step rel diff vs NCCL {rel:.2e}, "
              f"vs the single-GPU model {rel1:.2e}, column-parallel outlier sets identical")
    return True


def check_decoder(kind, rank, world):
    """The tensor-parallel decoder step with this exchange vs NCCL all-reduce + add (both sum fp16 partials; orders differ ->
    tolerance), and vs the same seeded model on ONE GPU (north star: <= 1e-2 relative; column-parallel outlier sets identical)."""
    cfg = CONFIGS["tiny"] if world <= 2 else CONFIGS["tiny8"]
    tok = torch.randint(0, cfg.vocab, (256, 1), generator=torch.Generator().manual_seed(0)).cuda()
    logits, models = {}, {}
    for mode in (kind, "nccl"):
        os.environ["MIXQ_TP_EXCHANGE"] = mode
        m = LlamaDecoder(cfg, batch=256, bit=8, rank=rank, world_size=world)
        assert m.discover(tok)
        m._rank_barrier()
        logits[mode] = m.step(tok).float()
        models[mode] = m
    rel = float((logits[kind] - logits["nccl"]).norm() / logits["nccl"].norm())
    assert rel < 5e-3, rel
    one = LlamaDecoder(cfg, batch=256, bit=8, rank=0, world_size=1)      # every rank: no collective inside
    assert one.discover(tok)
    l1 = one.step(tok).float()
    rel1 = float((logits[kind] - l1).norm() / l1.norm())
    assert rel1 <= 1e-2, f"TP logits vs single GPU: {rel1}"
    for a, b in zip(models[kind].layers, one.layers):
        for k in ("W_pack", "up_proj", "gate_proj"):
            assert torch.equal(a[k].ind, b[k].ind), f"column-parallel {k}: outlier index set differs from the single-GPU one"
    # steady-state graph replay == eager step
    m = models[kind]
    eager = m.step(tok).clone()
    if getattr(m.xchg, "can_quantize", False) and m.fuse_xchg_quant:
        # the finish kernel running the next Linear's activation prologue == the Linear running it itself, bit for bit
        m.fuse_xchg_quant = False
        m._rank_barrier()
        plain = m.step(tok).clone()
        m.fuse_xchg_quant = True
        m._rank_barrier()
        assert torch.equal(plain, eager), f"rank {rank}: quantising finish kernel changes the logits (max diff {float((plain.float() - eager.float()).abs().max()):.3e})"
    m.capture(tok)
    rep = m.replay(tok).clone()
    torch.cuda.synchronize()
    assert torch.equal(rep, eager), "graph replay == eager steady-state step"
    for m in models.values():
        m.close()
    dist.barrier()
    if rank == 0:
        print(f"{kind} decoder ok on {world} ranks: rel diff vs NCCL {rel:.2e}, vs the single-GPU model {rel1:.2e}, column-parallel "
              f"outlier sets identical, graph replay == eager")
    return rel, rel1


def check_push(rank, world, multicast, one_shot=None, sync="poll"):
    """The fused exchange: a real row-parallel MixLinear pushes its partial from the GEMM epilogue (both GEMM kernels: M = 512
    and M = 128), the finish kernel reduces + broadcasts; reference = the same Linear's partials (all-gathered over NCCL) summed
    in fp32 in rank order, rounded to fp16, + residual — bit-exact, identical on all ranks, eager and graph-replayed."""
    from mixq_b200.cache import MixLibCache
    from mixq_b200.linear import MixLinear_GEMM
    kind = ("push" if multicast else "push-nomc") + ("" if one_shot is None else ("-oneshot" if one_shot else "-twophase")) + "-" + sync
    for M, N, Ktot in ((512, 4096, 4096), (128, 2048, 1024)):
        Kr = Ktot // world
        g = torch.Generator(device="cuda").manual_seed(77 + rank)
        ex = PushExchange(M, N, rank, world, multicast=multicast, one_shot=one_shot, sync=sync)
        cache = MixLibCache(inputdim=M, sigma=6, bit=8)

        class W:
            weight = (torch.randn(N, Kr, generator=g, device="cuda") * 0.02).half()
            bias = None
            out_features, in_features = N, Kr
        lin = MixLinear_GEMM.from_linear(W, 8, cache=cache)

        def make_x():
            x = torch.randn(M, Kr, generator=g, device="cuda")
            x[:, 3::97] *= 20
            return x.half()
        for _ in range(2):                       # discovery calls
            lin(make_x(), None, True)
        assert not lin.add_outliers and lin._n_ind > 0

        def reference(x, res):
            y = lin(x.clone(), None, True)      # this rank's partial (same kernel, same bits as the pushed tiles)
            parts = [torch.empty_like(y) for _ in range(world)]
            dist.all_gather(parts, y.contiguous())
            acc = torch.zeros(M, N, dtype=torch.float32, device="cuda")
            for p_ in parts:
                acc += p_.float()
            return (acc.half().float() + res.float()).half()

        def same_on_all_ranks(out, what):
            mine = out.contiguous().view(torch.int16).to(torch.int32)
            lo, hi = mine.clone(), mine.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), f"rank {rank} {kind} {what}: ranks disagree"
        gr = torch.Generator(device="cuda").manual_seed(5)      # residual: same on every rank
        for it in range(4):
            x, res = make_x(), torch.randn(M, N, generator=gr, device="cuda").half()
            ref = reference(x, res)
            torch.cuda.synchronize()
            dist.barrier()
            lin(x.clone(), None, True, push=ex.push_targets())
            out = ex.reduce(res).clone()
            bad = int((out != ref).sum())
            assert bad == 0, f"rank {rank} {kind} M={M} exchange {it}: {bad} of {out.numel()} differ, max {float((out.float() - ref.float()).abs().max()):.3e}"
            same_on_all_ranks(out, f"exchange {it}")
        # graph replay: two exchanges per graph
        xs = [make_x(), make_x()]
        res = torch.randn(M, N, generator=gr, device="cuda").half()
        refs = [reference(x, res) for x in xs]
        xw = [x.clone() for x in xs]
        outs = [None, None]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        torch.cuda.synchronize()
        dist.barrier()
        with torch.cuda.stream(s):
            for j in range(2):
                lin(xw[j], None, True, push=ex.push_targets())
                ex.reduce(res)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        dist.barrier()
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            for j in range(2):
                lin(xw[j], None, True, push=ex.push_targets())
                outs[j] = ex.reduce(res).clone()
        torch.cuda.synchronize()
        dist.barrier()
        for rep in range(3):
            for j in range(2):
                xw[j].copy_(xs[j])              # the Linear zeroes its outlier columns in place
            gph.replay()
            torch.cuda.synchronize()
            for j in range(2):
                assert torch.equal(outs[j], refs[j]), f"rank {rank} {kind} M={M} graph replay {rep} exchange {j}"
        # ranks free-running: 50 replays back to back with no host synchronisation, rank r delayed by r * 200 us at the start
        torch.cuda._sleep(int(rank * 200e-6 * 1.9e9))
        for rep in range(50):
            for j in range(2):
                xw[j].copy_(xs[j])
            gph.replay()
        torch.cuda.synchronize()
        for j in range(2):
            assert torch.equal(outs[j], refs[j]), f"rank {rank} {kind} M={M} free-running replays exchange {j}"
        del gph
        ex.close()
        dist.barrier()
    if rank == 0:
        print(f"{kind} exchange ok on {world} ranks: GEMM-epilogue push + finish kernel bit-exact vs the rank-order fp32 sum of the "
              f"Linear's partials, identical on all ranks (eager + graph replay, M = 512 and M = 128)")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kinds = [os.environ["CHECK_KIND"]] if os.environ.get("CHECK_KIND") else ["push", "push-nomc", "peer", "multicast"]
    for kind in kinds:
        if kind.startswith("push"):
            for sync in ("poll", "flags"):
                for one_shot in (True, False):
                    check_push(rank, world, multicast=(kind == "push"), one_shot=one_shot, sync=sync)
            check_decoder(kind, rank, world)
        else:
            check_kind(kind, rank, world)
        torch.cuda.synchronize()
        dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()

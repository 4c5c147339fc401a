"""Latency anatomy of the tensor-parallel exchange on N ranks of one node (run under torchrun).

  * NVLink flag ping-pong: one-way latency of a release-store / multimem.red seen by a spinning ld.acquire.sys on the peer;
  * per-exchange device time of each exchange kernel alone (CUDA graph of 64 exchanges, CUDA events);
  * phase stamps inside the finish kernel of the fused exchange (entry, signal 0 sent, wait 0 done, data done, fence done,
    last block, signal 1 sent, wait 1 done), microseconds relative to entry, median over the graph's launches.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/bench_exchange.py
"""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixq_b200 import _lib  # noqa: E402
from mixq_b200.tp import MulticastExchange, PeerExchange, PushExchange, _symmetric_alloc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {"world": world}
    M, H = 512, 4096

    # ---- flag ping-pong between ranks 0 and 1
    loc, ptrs, mc, keep, close = _symmetric_alloc(4096, rank, world)
    ns = torch.zeros(1, dtype=torch.int64, device="cuda")
    iters = 2000
    for name, use_mc in (("st.release.sys", False), ("multimem.red", True)):
        if use_mc and (not mc or world != 2):
            continue
        off = 256 if use_mc else 0
        torch.cuda.synchronize()
        dist.barrier()
        if rank < 2:
            _lib.check(lib.mixq_debug_pingpong(loc + off, ptrs[1 - rank] + off, (mc + off) if use_mc else 0, iters, rank, ns.data_ptr(), st()), "pingpong")
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            out[f"pingpong_{name}_one_way_us"] = float(ns.item()) / iters / 2 / 1e3
    close()

    # ---- exchange kernels alone
    def time_graph(fn, n=64, reps=3):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        torch.cuda.synchronize()
        dist.barrier()
        with torch.cuda.stream(s):
            fn()
            fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                fn()
        torch.cuda.synchronize()
        dist.barrier()
        g.replay()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * n)
        t = torch.tensor([us], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        return float(t[0]), g

    h = torch.zeros(M, H, dtype=torch.float16, device="cuda")
    kinds = [("peer one-shot", lambda: PeerExchange(M, H, rank, world, two_shot=False)),
             ("peer two-shot", lambda: PeerExchange(M, H, rank, world, two_shot=True)),
             ("multicast all-reduce", lambda: MulticastExchange(M, H, rank, world)),
             ("push finish (multicast)", lambda: PushExchange(M, H, rank, world, multicast=True, one_shot=False, sync="flags")),
             ("push finish (peer stores)", lambda: PushExchange(M, H, rank, world, multicast=False, one_shot=False, sync="flags")),
             ("push finish (one-shot)", lambda: PushExchange(M, H, rank, world, one_shot=True, sync="flags"))]
    for name, mk in kinds:
        try:
            ex = mk()
        except Exception as e:
            out[name] = f"unavailable: {type(e).__name__}"
            continue
        us, g = time_graph(lambda: ex.reduce(h))
        out[name + " us"] = round(us, 2)
        if name.startswith("push finish (multicast)"):
            tb = torch.zeros(4096, dtype=torch.int64, device="cuda")
            lib.mixq_set_trace_buffer(tb.data_ptr())
            stamps = []
            for _ in range(40):
                ex.reduce(h)
                torch.cuda.synchronize()
                t = tb[:8].cpu().tolist()
                stamps.append([(v - t[0]) / 1e3 for v in t])
            lib.mixq_set_trace_buffer(None)
            med = [sorted(s[i] for s in stamps)[len(stamps) // 2] for i in range(8)]
            out["push finish phases us (entry, sig0 sent, wait0 done, data done, fence done, last block, sig1 sent, wait1 done)"] = [round(v, 2) for v in med]
        del g
        ex.close()
        dist.barrier()
    # the whole fused exchange as the decoder runs it: row-parallel o_proj (K = 4096 / world) pushing from its epilogue + the
    # finish kernel, against the same Linear alone -> what one exchange adds to the step
    from mixq_b200.cache import MixLibCache
    from mixq_b200.linear import MixLinear_GEMM
    Kr = 4096 // world
    gg = torch.Generator(device="cuda").manual_seed(3 + rank)

    class W:
        weight = (torch.randn(H, Kr, generator=gg, device="cuda") * 0.02).half()
        bias = None
        out_features, in_features = H, Kr
    cache = MixLibCache(inputdim=M, sigma=6, bit=8)
    lin = MixLinear_GEMM.from_linear(W, 8, cache=cache)
    x = torch.randn(M, Kr, generator=gg, device="cuda")
    x[:, 3::97] *= 20
    x = x.half()
    for _ in range(2):
        lin(x.clone(), None, True)
    us_lin, g = time_graph(lambda: lin(x, None, True))
    out["row-parallel o_proj alone us"] = round(us_lin, 2)
    del g
    for name, kw in (("poll two-phase multicast", dict(sync="poll", one_shot=False)),
                     ("poll two-phase peer stores", dict(sync="poll", one_shot=False, multicast=False)),
                     ("poll one-shot", dict(sync="poll", one_shot=True)),
                     ("flags two-phase multicast", dict(sync="flags", one_shot=False)),
                     ("flags one-shot", dict(sync="flags", one_shot=True))):
        ex = PushExchange(M, H, rank, world, **kw)

        def both():
            lin(x, None, True, push=ex.push_targets())
            ex.reduce(h)
        us, g = time_graph(both)
        out[f"o_proj + fused exchange ({name}) us"] = round(us, 2)
        out[f"fused exchange ({name}) adds us"] = round(us - us_lin, 2)
        del g
        ex.close()
        dist.barrier()
    # NCCL for scale
    t = torch.zeros(M, H, dtype=torch.float16, device="cuda")
    us, g = time_graph(lambda: dist.all_reduce(t))
    out["nccl all_reduce 4 MB us"] = round(us, 2)
    if rank == 0:
        print(json.dumps(out), flush=True)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()

"""The fused tensor-parallel exchange on ONE GPU: `world` logical ranks live in this process as separate buffer sets and
streams on the same device (a peer pointer is just another device pointer), so the very kernels of the N-GPU path run and are
checked where only one GPU is available (the driver's test box):

  * a real row-parallel MixLinear per rank pushes column slice j of its partial from the GEMM epilogue into rank j's receive
    slot (mixq_linear_args.y_peer / peer_cols), or the whole partial into every rank's slot (peer_bcast, one-shot);
  * mixq_exchange_finish_poll per rank (no flags: the data is its own signal) — result vs oracle.mixq_oracle.tp_exchange
    (rank-order fp32 sum, one rounding, residual as a separate fp16 add) of the Linears' own partials, BIT-EXACT, on all ranks;
  * mixq_exchange_finish_poll_quant — the same result, plus q_x / x_scale / act_outliers bit-identical to
    mixq_rmsnorm_extract_outliers run on that result (the next Linear's activation prologue, fused/norm.py:24-33);
  * three exchanges in a row through the two alternating buffer sets (consumers re-arm what they read).

The ranks' finish kernels wait for each other's data, so they must be co-resident: rows <= 64 keeps every grid small.
Run in its own process by tests/test_gpu_exchange_one_gpu.py (a protocol bug would end in the library's stall trap); it lives
under tests/ because it calls the oracle, which is test infrastructure.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixq_b200 import _lib  # noqa: E402
from mixq_b200.cache import MixLibCache  # noqa: E402
from mixq_b200.linear import MixLinear_GEMM  # noqa: E402
from oracle import mixq_oracle as O  # noqa: E402  (the checker)


def run(world, one_shot, quant, M=64, N=1024, Ktot=2048):
    lib = _lib.load()
    lib.mixq_set_peer_timeout_ms(20000)
    dev = "cuda"
    Kr, Ns = Ktot // world, N // world
    g = torch.Generator(device=dev).manual_seed(100 * world + 10 * one_shot + quant)
    streams = [torch.cuda.Stream() for _ in range(world)]
    # per rank: 2 buffer sets of receive slots + result buffers, armed with the sentinel (fp16 0xFFFF)
    slot = M * (N if one_shot else Ns)
    recv = [[torch.full((world * slot,), -1, dtype=torch.int16, device=dev) for _ in range(2)] for _ in range(world)]
    result = [[torch.full((M * N,), -1, dtype=torch.int16, device=dev) for _ in range(2)] for _ in range(world)]
    lins, caches = [], []
    for r in range(world):
        cache = MixLibCache(inputdim=M, sigma=6, bit=8)

        class W:
            weight = (torch.randn(N, Kr, generator=g, device=dev) * 0.02).half()
            bias = None
            out_features, in_features = N, Kr
        lin = MixLinear_GEMM.from_linear(W, 8, cache=cache)
        lins.append(lin)
        caches.append(cache)

    def make_x():
        x = torch.randn(M, Kr, generator=g, device=dev)
        x[:, 3::97] *= 20
        return x.half()
    for r in range(world):                      # outlier discovery
        for _ in range(2):
            lins[r](make_x(), None, True)
        assert not lins[r].add_outliers
    # the consumer of the quantising finish: a norm weight with a few loud channels and a fixed outlier set
    norm_w = torch.ones(N, dtype=torch.float16, device=dev)
    ind = torch.tensor(sorted(np.random.default_rng(1).choice(N, 9, replace=False)), dtype=torch.int32, device=dev)
    norm_w[ind.long()] = 12.0
    residual = torch.randn(M, N, generator=g, device=dev).half()
    for ex in range(3):
        b = ex & 1
        xs = [make_x() for _ in range(world)]
        # reference partials: the same Linear, same bits as the pushed tiles
        parts = [lins[r](xs[r].clone(), None, True).cpu().numpy() for r in range(world)]
        want = O.tp_exchange(parts, residual.cpu().numpy())
        torch.cuda.synchronize()
        outs = []
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                if one_shot:
                    ptrs = [recv[j][b].data_ptr() + r * slot * 2 for j in range(world)]
                    push = (ptrs, 0, world)
                else:
                    ptrs = [recv[j][b].data_ptr() + r * slot * 2 for j in range(world)]
                    push = (ptrs, Ns, 0)
                lins[r](xs[r].clone(), None, True, push=push)
        # On one device the ranks' GEMMs (one CTA per SM, grid barrier) must not share the SMs with another rank's spinning finish
        # kernel: let the pushes land first.  (On N GPUs every device runs one rank and the finish kernel follows its GEMM directly.)
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                a = _lib.ExchangePollArgs()
                a.recv = recv[r][b].data_ptr()
                for p_ in range(world):
                    a.result[p_] = result[p_][b].data_ptr()
                a.mc_result = 0
                a.reset = result[r][b ^ 1].data_ptr()
                a.residual = residual.data_ptr()
                a.M, a.N, a.world, a.rank, a.one_shot = M, N, world, r, 1 if one_shot else 0
                st = C.c_void_p(streams[r].cuda_stream)
                if quant:
                    q_x = torch.zeros(M, N, dtype=torch.int8, device=dev)
                    xsc = torch.zeros(M, dtype=torch.float16, device=dev)
                    ao = torch.zeros(M, 64, dtype=torch.float16, device=dev)
                    _lib.check(lib.mixq_exchange_finish_poll_quant(C.byref(a), norm_w.data_ptr(), 1e-5, ind.data_ptr(), len(ind),
                                                                   ao.data_ptr(), 64, q_x.data_ptr(), xsc.data_ptr(), 8, st), "finish_poll_quant")
                    outs.append((q_x, xsc, ao))
                else:
                    _lib.check(lib.mixq_exchange_finish_poll(C.byref(a), st), "finish_poll")
        torch.cuda.synchronize()
        for r in range(world):
            got = result[r][b].view(torch.float16).view(M, N).cpu().numpy()
            bad = int((got.view(np.uint16) != want.view(np.uint16)).sum())
            assert bad == 0, f"world {world} one_shot {one_shot} exchange {ex} rank {r}: {bad} of {got.size} elements differ from the oracle"
            # consumed buffers are re-armed: the slots of this exchange, and the other result buffer
            assert int((recv[r][b] != -1).sum()) == 0, "receive slots not re-armed"
            assert int((result[r][b ^ 1] != -1).sum()) == 0, "previous result buffer not re-armed"
            if quant:
                h = result[r][b].view(torch.float16).view(M, N)
                q2 = torch.zeros(M, N, dtype=torch.int8, device=dev)
                xs2 = torch.zeros(M, dtype=torch.float16, device=dev)
                ao2 = torch.zeros(M, 64, dtype=torch.float16, device=dev)
                nout = torch.zeros(M, N, dtype=torch.float16, device=dev)
                _lib.check(lib.mixq_rmsnorm_extract_outliers(h.data_ptr(), norm_w.data_ptr(), nout.data_ptr(), 1e-5, ind.data_ptr(),
                                                             len(ind), xs2.data_ptr(), ao2.data_ptr(), 64, q2.data_ptr(), M, N, 8,
                                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)), "rmsnorm_extract")
                torch.cuda.synchronize()
                q_x, xsc, ao = outs[r]
                assert torch.equal(q_x, q2) and torch.equal(xsc, xs2) and torch.equal(ao[:, :len(ind)], ao2[:, :len(ind)]), \
                    f"rank {r}: quantising finish differs from the separate prologue"
    print(f"one-GPU exchange ok: world {world}, {'one-shot' if one_shot else 'two-phase'}, {'quantising' if quant else 'plain'} finish, "
          f"3 exchanges bit-exact vs the oracle on every rank")


if __name__ == "__main__":
    for world in (2, 4):
        for one_shot in (False, True):
            for quant in (False, True):
                run(world, one_shot, quant)
    print("ALL OK")

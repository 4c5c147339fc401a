"""Kernel-level MixLinear benchmark (the reference's examples/benchbitsand.py:515-566 convention: TFLOPS = 2*M*N*K / t).

    python tools/bench_linear.py [--M 512] [--shapes 7b|8b|70b|NxK,...] [--modes norm,plain,skip,pairnorm,pairskip] [--tile 0] [--bit 8]
                                 [--copies 12] [--nout 41] [--reps 5] [--eager]

Each (shape, mode) is timed as ONE CUDA graph holding `copies` launches on distinct weight copies (so the weights come
from HBM, not L2, exactly as inside a decode step) between CUDA events.  `--eager` launches without a graph (for ncu).
"""
import argparse
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
from mixq_b200 import _lib  # noqa: E402

SHAPES = {
    "7b": [(12288, 4096), (4096, 4096), (11008, 4096), (4096, 11008)],
    "8b": [(6144, 4096), (4096, 4096), (14336, 4096), (4096, 14336)],
    "70b": [(10240, 8192), (8192, 8192), (28672, 8192), (8192, 28672)],
    "70b-tp8": [(1280, 8192), (8192, 1024), (3584, 8192), (8192, 3584)],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=512)
    ap.add_argument("--shapes", default="7b")
    ap.add_argument("--modes", default="norm,plain,skip")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--bit", type=int, default=8)
    ap.add_argument("--copies", type=int, default=12)
    ap.add_argument("--nout", type=int, default=41)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--no-splitk", action="store_true", help="M <= 128: do not hand the library a split-K workspace")
    args = ap.parse_args()
    lib = _lib.load()
    shapes = SHAPES[args.shapes] if args.shapes in SHAPES else [tuple(int(v) for v in s.split("x")) for s in args.shapes.split(",")]
    M, dev = args.M, "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for N, K in shapes:
        n = min(args.nout, K)
        cap = max(64, (n + 63) // 64 * 64)
        cols = torch.randperm(K, generator=g, device=dev)[:n].sort().values.int()
        x0 = torch.randn(M, K, generator=g, device=dev)
        x0[:, cols.long()] *= 20
        x0 = x0.half()
        nw = torch.ones(K, dtype=torch.float16, device=dev)
        copies = []
        for _ in range(args.copies):
            if args.bit == 8:
                qw = torch.randint(-127, 128, (N, K), generator=g, device=dev, dtype=torch.int8)
            else:
                qw = torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8)
            ws = (torch.rand(N, generator=g, device=dev) * 1e-3 + 1e-4).half()
            wc = (torch.randn(N, cap, generator=g, device=dev) * 0.02).half()
            copies.append((qw, ws, wc))
        pair_modes = [m for m in args.modes.split(",") if m.startswith("pair")]
        ups = []
        if pair_modes:   # SwiGLU pair: a second (up_proj) weight set per copy
            for _ in range(args.copies):
                ups.append((torch.randint(-127, 128, (N, K), generator=g, device=dev, dtype=torch.int8),
                            (torch.rand(N, generator=g, device=dev) * 1e-3 + 1e-4).half(),
                            (torch.randn(N, cap, generator=g, device=dev) * 0.02).half()))
        q_x = torch.zeros(M, K, dtype=torch.int8, device=dev)
        xs = torch.zeros(M, dtype=torch.float16, device=dev)
        ao = torch.zeros(M, cap, dtype=torch.float16, device=dev)
        y = torch.zeros(M, N, dtype=torch.float16, device=dev)
        sync = torch.zeros(1, dtype=torch.int32, device=dev)
        skws = torch.zeros(16 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
        x = x0.clone()
        for mode in args.modes.split(","):
            arglist = []
            pair = mode.startswith("pair")     # pairnorm / pairplain / pairskip
            sub = mode[4:] if pair else mode
            for ci, (qw, ws, wc) in enumerate(copies):
                a = _lib.LinearArgs()
                if pair:
                    a.q_weight_up, a.scale_col_up, a.weight_cache_up = ups[ci][0].data_ptr(), ups[ci][1].data_ptr(), ups[ci][2].data_ptr()
                a.x = x.data_ptr(); a.M, a.N, a.K = M, N, K
                a.norm_weight = nw.data_ptr() if sub == "norm" else 0
                a.norm_out = 0; a.eps = 1e-5
                a.q_weight = qw.data_ptr(); a.scale_col = ws.data_ptr(); a.bit = args.bit
                a.ind = cols.data_ptr(); a.n_ind = n
                a.weight_cache = wc.data_ptr(); a.ld_wc = cap
                a.q_x = q_x.data_ptr(); a.x_scale = xs.data_ptr(); a.act_outliers = ao.data_ptr(); a.ld_ao = cap
                a.sigma = 6.0; a.y = y.data_ptr(); a.grid_sync = sync.data_ptr(); a.tile_n = args.tile
                a.skip_prologue = 1 if sub == "skip" else 0
                if M <= 128 and not args.no_splitk:
                    a.splitk_ws, a.splitk_ws_bytes = skws.data_ptr(), skws.numel() * 4
                arglist.append(a)

            def run_all():
                for a in arglist:
                    _lib.check(lib.mixq_linear_fused(C.byref(a), st()), "linear_fused")
            x.copy_(x0)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run_all()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if args.eager:
                run_all()
                torch.cuda.synchronize()
                continue
            # host cost of one eager C-ABI call (argument checks + launch plan + tensor maps + cudaLaunchKernelEx), launch queue
            # not full: wall clock over the enqueue loop only
            import time
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(8):
                run_all()
            host_us = (time.perf_counter() - t0) * 1e6 / (8 * len(arglist))
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                run_all()
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (args.reps * len(arglist))
            nl = 2 if pair else 1
            fl = 2.0 * M * N * K * nl
            by = nl * (N * K * args.bit / 8 + 2 * N + 2 * n * N) + 2 * M * K + 2 * M * N
            print(json.dumps({"M": M, "N": N, "K": K, "bit": args.bit, "mode": mode, "n_out": n, "tile": args.tile,
                              "us": round(us, 2), "host_us_per_eager_call": round(host_us, 2), "tflops": round(fl / us / 1e6, 1), "gbs": round(by / us / 1e3, 1)}), flush=True)
            del gr
        del copies


if __name__ == "__main__":
    main()

"""GPU bring-up checks for the tcgen05 GEMM (run under gpurun; one case per process so a device
trap in one case cannot poison the next).  Not a test-suite file: tests/ holds the real parity tests.

usage: python tools/bringup_gemm.py CASE M N K TILE [n_out]
"""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from mixq_b200 import _lib  # noqa: E402


def ptr(t):
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def report(name, got, ref, exact):
    got = got.double()
    ref = ref.double()
    diff = (got - ref).abs()
    bad = diff > (0 if exact else 1e-2 * ref.abs().clamp_min(1e-3))
    nbad = int(bad.sum())
    rel = float((got - ref).norm() / ref.norm().clamp_min(1e-30))
    print(f"[{name}] mismatches={nbad}/{got.numel()} max_abs={float(diff.max()):.4g} rel_fro={rel:.3e}")
    if nbad:
        idx = bad.nonzero()[:8].tolist()
        print("   first bad (row,col,got,ref):", [(r, c, float(got[r, c]), float(ref[r, c])) for r, c in idx])
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print(f"   bad rows: {len(rows)} (first {rows[:16].tolist()})  bad cols: {len(cols)} (first {cols[:16].tolist()})")
    return nbad == 0


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    case, M, N, K, tile = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    n_out = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    lib = _lib.load()
    _lib.check(lib.mixq_set_tile_n(tile), "set_tile_n")
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(0)
    qx = torch.randint(-127, 128, (M, K), generator=g, dtype=torch.int8).to(dev)
    qw = torch.randint(-127, 128, (N, K), generator=g, dtype=torch.int8).to(dev)
    ref_i = qx.double() @ qw.double().T
    ok = True
    t0 = time.time()
    if case == "gemm":
        y = torch.full((M, N), -7, dtype=torch.int32, device=dev)
        fn = lambda: _lib.check(lib.mixq_gemm_i8(ptr(qx), ptr(qw), ptr(y), M, N, K, stream()), "gemm_i8")
        fn()
        torch.cuda.synchronize()
        ok = report(f"gemm M{M} N{N} K{K} t{tile}", y, ref_i, True)
    elif case == "dequant":
        xs = (torch.rand(M, 1, generator=g) * 0.05 + 0.01).half().to(dev)
        ws = (torch.rand(1, N, generator=g) * 0.01 + 0.001).half().to(dev)
        outl = torch.randn(M, N, generator=g).half().to(dev)
        y = torch.zeros((M, N), dtype=torch.float16, device=dev)
        fn = lambda: _lib.check(
            lib.mixq_int8_fused_dequantize(ptr(qx), ptr(qw), ptr(xs), ptr(ws), ptr(outl), N, ptr(y), M, N, K, 0, stream()),
            "fused_dequant")
        fn()
        torch.cuda.synchronize()
        ref = ((ref_i.float() * xs.float()) * ws.float() + outl.float()).half()
        ok = report(f"dequant M{M} N{N} K{K} t{tile}", y, ref, True)
    elif case in ("fused", "fused4"):
        bit = 4 if case == "fused4" else 8
        x = torch.randn(M, K, generator=g).half()
        cols = torch.randperm(K, generator=g)[:n_out].sort().values
        x[:, cols] *= 20
        x = x.to(dev)
        w = (torch.randn(N, K, generator=g) * 0.02).half().to(dev)
        ind = cols.int().to(dev)
        if bit == 8:
            ws = (w.float().abs().amax(1) / 127).half()
            qw = (w.float() / ws.float()[:, None]).round().clamp(-127, 127).to(torch.int8)
            wc = torch.zeros(N, 256, dtype=torch.float16, device=dev)
            wc[:, :n_out] = qw[:, cols.to(dev)].half() * ws[:, None]
            qwp = qw
            wdq = qw.float() * ws.float()[:, None]
        else:
            wz = w.clone()
            wc = torch.zeros(N, 256, dtype=torch.float16, device=dev)
            wc[:, :n_out] = w[:, cols.to(dev)]
            wz[:, cols.to(dev)] = 0
            ws = (wz.float().abs().amax(1) / 10).half()
            q4 = (wz.float() / ws.float()[:, None]).round().clamp(-8, 7).to(torch.int8)
            u = torch.where(q4 < 0, q4 + 16, q4).to(torch.uint8)
            qwp = (u[:, 0::2] | (u[:, 1::2] << 4)).contiguous()
            wdq = q4.float() * ws.float()[:, None]
        x_ref = x.clone()
        ao_ref = x_ref[:, cols.to(dev)].clone()
        x_ref[:, cols.to(dev)] = 0
        qmax = 127 if bit == 8 else 7
        xs_ref = (x_ref.float().abs().amax(1) / qmax).half()
        qx_ref = (x_ref.float() / xs_ref.float()[:, None]).round().clamp(-qmax, qmax)
        y_ref = ((qx_ref.double() @ (wdq / ws.float()[:, None]).double().T).float() * xs_ref.float()[:, None]) * ws.float()[None, :] \
            + (ao_ref.double() @ wc[:, :n_out].double().T).float()
        q_x = torch.zeros(M, K, dtype=torch.int8, device=dev)
        x_scale = torch.zeros(M, dtype=torch.float16, device=dev)
        ao = torch.zeros(M, 256, dtype=torch.float16, device=dev)
        y = torch.zeros(M, N, dtype=torch.float16, device=dev)
        sync = torch.zeros(1, dtype=torch.int32, device=dev)
        a = _lib.LinearArgs()
        xw = x.clone()
        a.x = xw.data_ptr(); a.M, a.N, a.K = M, N, K
        a.q_weight = qwp.data_ptr(); a.scale_col = ws.data_ptr(); a.bit = bit
        a.ind = ind.data_ptr(); a.n_ind = n_out
        a.weight_cache = wc.data_ptr(); a.ld_wc = 256
        a.q_x = q_x.data_ptr(); a.x_scale = x_scale.data_ptr(); a.act_outliers = ao.data_ptr(); a.ld_ao = 256
        a.sigma = 6.0; a.y = y.data_ptr(); a.grid_sync = sync.data_ptr(); a.tile_n = tile

        def fn():
            xw.copy_(x)
            _lib.check(lib.mixq_linear_fused(C.byref(a), stream()), "linear_fused")
        fn()
        torch.cuda.synchronize()
        ok = report(f"{case} qx M{M} N{N} K{K} t{tile} n_out{n_out}", q_x, qx_ref, True)
        ok &= report(f"{case} xs", x_scale[:, None], xs_ref[:, None], True)
        ok &= report(f"{case} ao", ao[:, :max(n_out, 1)], ao_ref if n_out else ao[:, :1], True)
        ok &= report(f"{case} y", y, y_ref, False)
    else:
        raise SystemExit("unknown case")
    us = timeit(fn)
    tflops = 2.0 * M * N * K / us / 1e6
    print(f"   time {us:.2f} us  {tflops:.1f} TOPS  ({'OK' if ok else 'FAIL'}) wall {time.time()-t0:.1f}s", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

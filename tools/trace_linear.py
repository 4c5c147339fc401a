"""Per-phase timeline of one mixq_linear_fused launch from in-kernel %globaltimer stamps."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, ".")
from mixq_b200 import _lib
lib = _lib.load()
M = int(os.environ.get('M', '512'))
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
trace = torch.zeros(16384, dtype=torch.int64, device=dev)
NOUT = int(os.environ.get('NOUT', '41'))
PAIR = os.environ.get('PAIR') == '1'   # SwiGLU pair launch (gate + up)
NORM = os.environ.get('NORM') == '1'
BIT = int(os.environ.get('BIT', '8'))
TILE = int(os.environ.get('TILE', '0'))
MODES = os.environ.get('MODES', 'plain,skip').split(',')
SHAPES = [tuple(int(v) for v in t.split('x')) for t in os.environ.get('SHAPES', '12288x4096,4096x4096,4096x11008').split(',')]
for (N, K) in SHAPES:
    n, cap = NOUT, max(64, (NOUT + 63) // 64 * 64)
    cols = torch.randperm(K, generator=g, device=dev)[:max(n, 1)].sort().values.int()
    x0 = torch.randn(M, K, generator=g, device=dev); x0[:, cols.long()] *= 20; x0 = x0.half()
    ws_l = [((torch.randint(-127, 128, (N, K), generator=g, device=dev, dtype=torch.int8) if BIT == 8 else
              torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8)),
             (torch.rand(N, generator=g, device=dev) * 1e-3 + 1e-4).half(),
             (torch.randn(N, cap, generator=g, device=dev) * 0.02).half()) for _ in range(4)]
    q_x = torch.zeros(M, K, dtype=torch.int8, device=dev); xs = torch.zeros(M, dtype=torch.float16, device=dev)
    ao = torch.zeros(M, cap, dtype=torch.float16, device=dev); y = torch.zeros(M, N, dtype=torch.float16, device=dev)
    skws = torch.zeros(16 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
    sync = torch.zeros(1, dtype=torch.int32, device=dev); x = x0.clone(); nw = torch.ones(K, dtype=torch.float16, device=dev)
    for mode in MODES:
        for it, (qw, ws, wc) in enumerate(ws_l):
            a = _lib.LinearArgs()
            a.x = x.data_ptr(); a.M, a.N, a.K = M, N, K
            a.q_weight = qw.data_ptr(); a.scale_col = ws.data_ptr(); a.bit = BIT
            a.ind = cols.data_ptr(); a.n_ind = n; a.weight_cache = wc.data_ptr(); a.ld_wc = cap
            a.q_x = q_x.data_ptr(); a.x_scale = xs.data_ptr(); a.act_outliers = ao.data_ptr(); a.ld_ao = cap
            if PAIR:
                qw2, ws2, wc2 = ws_l[(it + 1) % 4]
                a.q_weight_up = qw2.data_ptr(); a.scale_col_up = ws2.data_ptr(); a.weight_cache_up = wc2.data_ptr()
            if NORM:
                a.norm_weight = nw.data_ptr(); a.eps = 1e-5
            if os.environ.get('SPLITK') == '1':
                a.splitk_ws, a.splitk_ws_bytes = skws.data_ptr(), skws.numel() * 4
            a.sigma = 6.0; a.y = y.data_ptr(); a.grid_sync = sync.data_ptr(); a.skip_prologue = 1 if mode == "skip" else 0; a.tile_n = TILE
            lib.mixq_set_trace_buffer(trace.data_ptr() if it == 3 else 0)
            trace.zero_()
            torch.cuda.synchronize()
            _lib.check(lib.mixq_linear_fused(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "fused")
            torch.cuda.synchronize()
        full = trace.cpu().double()
        t = full[:148 * 8].view(148, 8)
        t0 = t[:, 0][t[:, 0] > 0].min()
        def col(i):
            v = t[:, i][t[:, i] > 0]
            return (f"{(v.min()-t0)/1e3:6.2f}..{(v.max()-t0)/1e3:6.2f}" if len(v) else "      -       ")
        print(f"N={N} K={K} tile={TILE} {mode:5s} us since first CTA start: start {col(0)} | mask built {col(6)} | row0 absmax {col(7)} | prologue done {col(1)} | barrier passed {col(2)} | "
              f"first MMA {col(3)} | int MMAs issued {col(6)} | last MMA {col(4)} | epilogue done {col(5)}")
        if os.environ.get('SPLITK') == '1':
            print(f"    split-K finishers: partials arrived {col(6)} | folded {col(7)} | done (per CTA, us): " + " ".join(f"{(v - t0) / 1e3:.1f}" for v in t[:, 5][t[:, 5] > 0].tolist()[:48]))
        base0 = t[0, 0]
        pp = full[1536:1568].view(16, 2)
        if pp[0, 0] > 0:
            print("    CTA 0 warp 4 passes (start..end us): " + " ".join(f"{(a - base0) / 1e3:.2f}..{(b - base0) / 1e3:.2f}" for a, b in pp.tolist() if a > 0))
            mm = full[1600:1616]
            print("    CTA 0 outlier-pass MMA issue times: " + " ".join(f"{(a - base0) / 1e3:.2f}" for a in mm.tolist() if a > 0))
        gb = full[1700:1700 + 148 * 4].view(148, 4)
        if gb[:, 0].sum() > 0:
            f = lambda i: f"{(gb[:, i].min() - t0) / 1e3:.2f}..{(gb[:, i].max() - t0) / 1e3:.2f}"
            print(f"    grid barrier (all CTAs, us): enter {f(0)} | fence done {f(1)} | atomic done {f(2)} | flip seen {f(3)}")
        ncall = int(full[1799])
        if 0 < ncall <= 24:      # built with EXTRA=-DMIXQ_EPI_TRACE: stamps inside the epilogue runs of CTA 0 / warp 4
            et = full[1800:1800 + 8 * ncall].view(ncall, 8)
            pw = full[2000:2032].view(16, 2)
            c0 = pw[0, 0] if pw[0, 0] > 0 else et[et > 0].min()
            print("    epilogue runs of CTA 0 warp 4, SM clocks since the first pass wait (ld0 start, ld0 done, math0 done, ld1 start, ld1 done, math1 done, out start, out done):")
            for row in et.tolist():
                print("      " + " ".join(f"{int(v - c0):6d}" if v > 0 else "     -" for v in row))
            print("    pass waits (before, after): " + " | ".join(f"{int(a - c0)} {int(b - c0)}" for a, b in pw.tolist() if a > 0))
        if os.environ.get("CADENCE"):
            base = t[0, 0]
            up = full[1200:1200 + 160].view(40, 4)
            if up[:, 0].sum() > 0:
                print("    unpack warp 4 (enter, stage free, packed landed, done) us: " + " | ".join(
                    " ".join(f"{(v - base) / 1e3:.2f}" if v > 0 else "-" for v in row) for row in up.tolist() if row[0] > 0))
            for name, off, cnt in (("mma", 2048, 96), ("wgt-tma", 2048 + 256, 48), ("act-tma", 2048 + 512, 48)):
                v = full[off:off + cnt]
                v = v[v > 0]
                print("   ", name, " ".join(f"{(x - base) / 1e3:.2f}" for x in v.tolist()))
lib.mixq_set_trace_buffer(0)

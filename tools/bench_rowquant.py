"""Standalone activation-prologue kernel timing (graph of 20 launches, CUDA events)."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from mixq_b200 import _lib
lib = _lib.load()
dev = "cuda"
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
for M, K in [(512, 4096), (512, 11008), (128, 8192), (2048, 4096)]:
    x = torch.randn(M, K, device=dev).half()
    w = torch.ones(K, device=dev).half()
    out = torch.empty_like(x)
    xs = torch.zeros(M, 1, dtype=torch.float16, device=dev)
    q = torch.zeros(M, K, dtype=torch.int8, device=dev)
    ind = torch.randperm(K, device=dev)[:41].sort().values.int()
    ao = torch.zeros(M, 64, dtype=torch.float16, device=dev)
    fns = {
        "find_row_scale": lambda: lib.mixq_find_row_scale(x.data_ptr(), xs.data_ptr(), q.data_ptr(), M, K, 8, st()),
        "rmsnorm": lambda: lib.mixq_rmsnorm(x.data_ptr(), w.data_ptr(), out.data_ptr(), 1e-5, M, K, st()),
        "rmsnorm_extract": lambda: lib.mixq_rmsnorm_extract_outliers(x.data_ptr(), w.data_ptr(), out.data_ptr(), 1e-5, ind.data_ptr(), 41,
                                                                      xs.data_ptr(), ao.data_ptr(), 64, q.data_ptr(), M, K, 8, st()),
    }
    for name, fn in fns.items():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                _lib.check(fn(), name)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 100
        by = M * K * 3 + (M * K * 2 if "rmsnorm" in name else 0)
        print(f"M={M} K={K} {name:16s} {us:7.2f} us  {by/us/1e3:7.1f} GB/s")

"""Measure the INT8 tensor-pipe peak of this B200: the roofline denominator of every
MixLinear number.

    python tools/int8_peak.py [--out profiles/r02_int8_peak.json]

Three measurements, CUDA-event timed, clocks sampled through NVML during each:
  * library proxy: torch._int_mm (cuBLASLt IMMA) 8192^3, best of 10 launches (burst) and back to back for 4 s (sustained);
  * library bf16 for the same shape (cross-check against MEASURED_PEAKS.json);
  * this library's own kernel with the activation prologue skipped and no outlier columns (mixq_linear_fused,
    skip_prologue = 1, n_ind = 0), M = 4096, N = K = 8192: burst and sustained.
ops = 2*M*N*K.  Writes one JSON object; bench.py reads `int8_tops_burst` / `int8_tops_sustained` from the committed copy.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Clocks(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.s, self.p, self.reasons, self.stop_flag = [], [], set(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                self.s.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.p.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for name, bit in (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal", 0x20), ("hw_thermal", 0x40)):
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        s = sorted(self.s)
        return {"sm_mhz_median": s[len(s) // 2] if s else None, "sm_mhz_min": s[0] if s else None,
                "power_w_max": max(self.p) if self.p else None, "reasons": sorted(self.reasons), "samples": len(s)}


def timed(fn, ops, sustained_s=4.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    ck = Clocks()
    ck.start()
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    ck.stop_flag = True
    ck.join(timeout=1)
    burst = {"tops": ops / best / 1e12, "us": best * 1e6, "clocks": ck.result()}
    n = max(10, int(sustained_s / best))
    ck = Clocks()
    ck.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ck.stop_flag = True
    ck.join(timeout=1)
    t = e0.elapsed_time(e1) * 1e-3 / n
    return burst, {"tops": ops / t / 1e12, "us": t * 1e6, "launches": n, "clocks": ck.result()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/r02_int8_peak.json")
    args = ap.parse_args()
    from mixq_b200 import _lib
    lib = _lib.load()
    dev = "cuda"
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    n = 8192
    a = torch.randint(-127, 128, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-127, 128, (n, n), dtype=torch.int8, device=dev)
    bt = b.t()   # _int_mm wants the second operand column-major
    res["int_mm_8192"] = dict(zip(("burst", "sustained"), timed(lambda: torch._int_mm(a, bt), 2.0 * n ** 3)))
    x16, w16 = torch.randn(n, n, device=dev, dtype=torch.bfloat16), torch.randn(n, n, device=dev, dtype=torch.bfloat16)
    res["bf16_matmul_8192"] = dict(zip(("burst", "sustained"), timed(lambda: torch.matmul(x16, w16), 2.0 * n ** 3)))
    del x16, w16

    # own kernel, prologue skipped, no outliers: the pure tcgen05 kind::i8 mainloop + dequant epilogue
    M, N, K = 4096, 8192, 8192
    q_x = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    qw = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
    xs = (torch.rand(M, device=dev) * 1e-2 + 1e-3).half()
    ws = (torch.rand(N, device=dev) * 1e-3 + 1e-4).half()
    y = torch.zeros(M, N, dtype=torch.float16, device=dev)
    sync = torch.zeros(1, dtype=torch.int32, device=dev)
    ar = _lib.LinearArgs()
    ar.M, ar.N, ar.K = M, N, K
    ar.q_weight, ar.scale_col, ar.bit = qw.data_ptr(), ws.data_ptr(), 8
    ar.q_x, ar.x_scale, ar.y = q_x.data_ptr(), xs.data_ptr(), y.data_ptr()
    ar.skip_prologue, ar.grid_sync, ar.sigma = 1, sync.data_ptr(), 6.0
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def own():
        _lib.check(lib.mixq_linear_fused(C.byref(ar), st), "linear_fused")
    res["mixq_linear_skip_4096x8192x8192"] = dict(zip(("burst", "sustained"), timed(own, 2.0 * M * N * K)))
    # the exact int32 sums agree with the library GEMM (the peak run computes something real)
    acc = torch._int_mm(q_x[:32].contiguous(), qw.t()).float()
    ref = ((acc * xs[:32, None].float()) * ws[None, :].float()).half().float()
    res["own_kernel_rel_err_vs_int_mm"] = float((y[:32].float() - ref).norm() / ref.norm())
    res["int8_tops_burst"] = max(res["int_mm_8192"]["burst"]["tops"], res["mixq_linear_skip_4096x8192x8192"]["burst"]["tops"])
    res["int8_tops_sustained"] = max(res["int_mm_8192"]["sustained"]["tops"], res["mixq_linear_skip_4096x8192x8192"]["sustained"]["tops"])
    res["how"] = ("ops = 2*M*N*K; burst = best of 10 single launches, sustained = back to back for ~4 s; CUDA events; "
                  "the higher of the cuBLASLt proxy and this library's own prologue-less kernel is the peak used by bench.py")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()

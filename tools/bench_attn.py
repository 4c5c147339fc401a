"""Time the decode attention kernels alone (CUDA graph of `copies` launches on distinct qkv buffers, CUDA events):
mixq_rope_attention_decode vs mixq_rope_attention_decode_quant (attention + o_proj's activation prologue).
    python tools/bench_attn.py [--M 512] [--H 32] [--Hkv 32] [--D 128] [--nout 41] [--eager]
"""
import argparse
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
from mixq_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=512)
    ap.add_argument("--H", type=int, default=32)
    ap.add_argument("--Hkv", type=int, default=32)
    ap.add_argument("--D", type=int, default=128)
    ap.add_argument("--nout", type=int, default=41)
    ap.add_argument("--copies", type=int, default=16)
    ap.add_argument("--eager", action="store_true")
    a = ap.parse_args()
    lib = _lib.load()
    M, H, Hkv, D = a.M, a.H, a.Hkv, a.D
    K = H * D
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    qkvs = [torch.randn(M, (H + 2 * Hkv) * D, device="cuda").half() for _ in range(a.copies)]
    out = torch.zeros(M, K, dtype=torch.float16, device="cuda")
    ind = torch.randperm(K, device="cuda")[:a.nout].sort().values.int()
    cap = max(64, (a.nout + 63) // 64 * 64)
    ao = torch.zeros(M, cap, dtype=torch.float16, device="cuda")
    q_x = torch.zeros(M, K, dtype=torch.int8, device="cuda")
    xs = torch.zeros(M, dtype=torch.float16, device="cuda")

    def plain():
        for q in qkvs:
            _lib.check(lib.mixq_rope_attention_decode(q.data_ptr(), 0, 0, 0, 0, out.data_ptr(), M, H, Hkv, D, 10000.0, st()), "attn")

    def quant():
        for q in qkvs:
            _lib.check(lib.mixq_rope_attention_decode_quant(q.data_ptr(), 0, 0, 0, 0, 0, M, H, Hkv, D, 10000.0, ind.data_ptr(), a.nout,
                                                            ao.data_ptr(), cap, q_x.data_ptr(), xs.data_ptr(), 8, st()), "attn_quant")

    def rowquant():
        for q in qkvs:
            _lib.check(lib.mixq_find_row_scale(out.data_ptr(), xs.data_ptr(), q_x.data_ptr(), M, K, 8, st()), "rowquant")

    for name, fn in (("attention", plain), ("attention+quant", quant), ("FindRowScale alone", rowquant)):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if a.eager:
            fn()
            torch.cuda.synchronize()
            continue
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (5 * a.copies)
        byts = M * (H + 2 * Hkv) * D * 2 + M * K * (2 if name == "attention" else 1)
        print(json.dumps({"kernel": name, "M": M, "H": H, "Hkv": Hkv, "D": D, "us": round(us, 2), "gbs": round(byts / us / 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()

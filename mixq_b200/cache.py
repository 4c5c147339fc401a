"""MixLibCache — the per-model scratch/state carrier every MixLinear of a model shares.

Mirror of /root/reference/mixquant/Cache.py:5-25 (same constructor, same public fields, same defaults
sigma = 6, stop = 2, max_outliers = 256).  Differences, all behind the same names:
  * `zeros` ([inputdim, 36864] fp16, 75 MB in the reference, only ever passed as the "no outliers" addend
    of int8FusedDequantize) is None: the C ABI takes a NULL addend instead;
  * the buffers the reference allocates on every call (q_xcache, activation_outliers) are views into
    caller-owned capacity buffers kept here, so the steady state allocates nothing and can be captured in a
    CUDA graph;
  * a few device words the fused kernel needs (grid barrier word, outlier-scan flags).
"""
from __future__ import annotations

import torch

_PAD = 64  # outlier-column capacity granularity (one fp16 SWIZZLE_128B k-block)


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


class MixLibCache:
    def __init__(self, inputdim=1024, sigma=6, bit=8, eval_ppl=False, locality=False, device="cuda"):
        self.device = device
        self.x_scale = torch.zeros((inputdim, 1), dtype=torch.float16, device=device)
        self.sigma = torch.zeros((1, 1), dtype=torch.float16, device=device)
        self.sigma[0] = sigma
        self.zeros = None

        self.ind = None
        self.new_ind = None
        self.shape = None
        self.activation_outliers = None
        self.q_xcache = None
        self.is_prefill = False
        self.bit = bit

        self.max_outliers = 256
        self.stop = 2

        self.eval_ppl = eval_ppl
        self.locality = locality

        # capacity buffers (grown on demand, never in the steady state)
        self.inputdim = inputdim
        self._q_x = None          # int8 [inputdim, kcap]
        self._ao = None           # fp16 [inputdim, ocap]
        self._col_over = None     # uint8 [kcap]
        self.grid_sync = torch.zeros(1, dtype=torch.int32, device=device)
        self.over_flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.n_new = torch.zeros(1, dtype=torch.int32, device=device)
        self.sigma_f = float(torch.tensor(sigma, dtype=torch.float16))

    # ------------------------------------------------------------------ scratch management
    def q_x_buffer(self, M: int, K: int) -> torch.Tensor:
        """int8 [M,K] contiguous scratch for the quantised activations."""
        if M > self.inputdim:
            raise ValueError(f"batch of {M} rows exceeds MixLibCache(inputdim={self.inputdim})")
        if self._q_x is None or self._q_x.numel() < self.inputdim * K:
            self._q_x = torch.empty(self.inputdim * K, dtype=torch.int8, device=self.device)
        return self._q_x[: M * K].view(M, K)

    def ao_buffer(self, n_out: int) -> torch.Tensor:
        """fp16 [inputdim, ocap] scratch for the gathered outlier activations, ocap % 64 == 0, ocap >= n_out."""
        need = max(_PAD, _round_up(n_out, _PAD))
        if self._ao is None or self._ao.shape[1] < need:
            new = torch.zeros((self.inputdim, need), dtype=torch.float16, device=self.device)
            if self._ao is not None:
                new[:, : self._ao.shape[1]] = self._ao
            self._ao = new
        return self._ao

    def splitk_buffer(self) -> torch.Tensor:
        """Zero-filled workspace for the split-K launches of small-M shapes (include/mixq.h: mixq_linear_args.splitk_ws):
        4 KB of per-tile counters + up to three int32 partial-sum slices.  16 MB covers M <= 128 with N <= 10 922 at the
        full split; larger N runs with fewer splits (the library checks the size)."""
        if getattr(self, "_splitk", None) is None:
            self._splitk = torch.zeros(16 * 1024 * 1024 // 4, dtype=torch.int32, device=self.device)
        return self._splitk

    def col_over_buffer(self, K: int) -> torch.Tensor:
        if self._col_over is None or self._col_over.numel() < K:
            self._col_over = torch.zeros(K, dtype=torch.uint8, device=self.device)
        return self._col_over

    def do_bench_cudagraph(self, fn):
        """Cache.py:26-38: warm up, capture `fn` into a CUDA graph, return the graph."""
        if torch.cuda.current_stream() == torch.cuda.default_stream():
            raise RuntimeError("Cannot capture graph in default stream. Please use side stream in benchmark code.")
        for _ in range(10):
            fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        torch.cuda.synchronize()
        return g


class MLPCache:
    """Cache.py:42-48."""

    def __init__(self, max_batch_size=4096, device="cuda"):
        self.device = device
        self.x_scale = torch.zeros((max_batch_size, 1), dtype=torch.float16, device=device)
        self.ind = None
        self.shape = None
        self.activation_outliers = None

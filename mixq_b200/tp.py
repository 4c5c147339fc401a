"""Column-/row-parallel sharding of the MixLinears of a Llama layer.

The reference has no collective code at all (its multi-GPU story is accelerate layer placement,
models/base.py:196-225); this is the Megatron-style split the north star asks for:
  * W_pack (q|k|v, by heads), gate_proj, up_proj: column-parallel — shard N; x is replicated, so every rank
    computes the same x_scale and the same outlier index set, and no collective is needed;
  * o_proj, down_proj: row-parallel — shard K; each rank quantises its K-slice with its own per-row scale and
    its local outlier columns, then ONE all-reduce(sum) of the fp16 [M, hidden] output.
Pure tensor slicing, device-agnostic (used on CPU by the gloo tests).
"""
from __future__ import annotations

import ctypes as C

import torch


def shard_rows(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Column-parallel Linear: weight [N,K] -> rows [rank*N/world, (rank+1)*N/world)."""
    n = t.shape[0]
    if n % world:
        raise ValueError(f"{n} rows do not divide by world_size {world}")
    s = n // world
    return t[rank * s:(rank + 1) * s]


def shard_cols(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Row-parallel Linear: weight [N,K] -> columns [rank*K/world, (rank+1)*K/world)."""
    k = t.shape[1]
    if k % world:
        raise ValueError(f"{k} columns do not divide by world_size {world}")
    s = k // world
    return t[:, rank * s:(rank + 1) * s]


def pack_qkv_shard(wq, wk, wv, rank: int, world: int) -> torch.Tensor:
    """models/llama.py:98-166 concatenates q/k/v along N; a rank takes its heads of each before concatenating,
    so that its W_pack output is [q_local | k_local | v_local]."""
    return torch.cat([shard_rows(wq, rank, world), shard_rows(wk, rank, world), shard_rows(wv, rank, world)], 0).contiguous()


def all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(t, group=group)
    return t


class _DeviceBuffer:
    """Zero-copy torch view of raw device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerExchange:
    """The exchange step of the row-parallel Linears over NVLink peer memory: ONE kernel per rank replaces
    ncclAllReduce(sum, fp16 [M, hidden]) + the decoder's residual add (include/mixq.h: mixq_allreduce_residual).

    Every rank owns two partial buffers (the row-parallel MixLinear writes straight into them, alternating) and eight flag
    words, allocated with cudaMalloc and mapped into the other ranks of the node through CUDA IPC handles exchanged over the
    process group.  Reads of the peers' partials travel over NVLink / NVSwitch inside the kernel; nothing goes through NCCL.
    """

    def __init__(self, rows: int, cols: int, rank: int, world: int, group=None, device="cuda", two_shot=None):
        from . import _lib
        import os
        import torch.distributed as dist
        if not (2 <= world <= 8):
            raise ValueError("PeerExchange needs 2..8 ranks on one node")
        if (rows * cols) % 8:
            raise ValueError("rows * cols must be a multiple of 8")
        if two_shot is None:
            # one-shot reads (world - 1) * n from the peers, two-shot moves 2 * (world - 1) / world * n and shakes hands twice
            env = os.environ.get("MIXQ_TP_TWO_SHOT")
            two_shot = (world >= 4) if env is None else env == "1"
        self.two_shot = bool(two_shot)
        self.lib = _lib.load()
        self._check = _lib.check
        self.rows, self.cols, self.rank, self.world, self.device = rows, cols, rank, world, device
        nbytes = rows * cols * 2
        sizes = [nbytes, nbytes, 256] + ([nbytes, nbytes] if self.two_shot else [])   # partial 0/1, flags, result 0/1
        self._own = []
        for size in sizes:
            p = C.c_void_p()
            self._check(self.lib.mixq_peer_alloc(size, C.byref(p)), "peer_alloc")
            self._own.append(p.value)
        handles = []
        for p in self._own:
            h = C.create_string_buffer(64)
            self._check(self.lib.mixq_ipc_get_handle(p, h), "ipc_get_handle")
            handles.append(h.raw)
        gathered = [None] * world
        dist.all_gather_object(gathered, handles, group=group)
        self._mapped = []      # peer mappings to close
        ptrs = [[0] * len(sizes) for _ in range(world)]
        for r in range(world):
            for j in range(len(sizes)):
                if r == rank:
                    ptrs[r][j] = self._own[j]
                else:
                    q = C.c_void_p()
                    self._check(self.lib.mixq_ipc_open_handle(gathered[r][j], C.byref(q)), f"ipc_open_handle(rank {r})")
                    ptrs[r][j] = q.value
                    self._mapped.append(q.value)
        self._state = torch.zeros(2, dtype=torch.int32, device=device)      # epoch, done
        self._args = _lib.AllReduceArgs()
        a = self._args
        for r in range(world):
            a.partial0[r], a.partial1[r], a.flags[r] = ptrs[r][0], ptrs[r][1], ptrs[r][2]
            if self.two_shot:
                a.result0[r], a.result1[r] = ptrs[r][3], ptrs[r][4]
        a.epoch = self._state.data_ptr()
        a.done = self._state.data_ptr() + 4
        a.world, a.rank, a.n = world, rank, rows * cols
        nbuf = 4 if self.two_shot else 2
        self._keep = [_DeviceBuffer(self._own[j if j < 2 else j + 1], (rows, cols), "<f2") for j in range(nbuf)]
        views = [torch.as_tensor(k, device=device) for k in self._keep]
        self.partials, self.results = views[:2], views[2:]
        self.buf = 0
        dist.barrier(group=group)       # nobody signals before everybody has mapped everything

    def next_partial(self) -> torch.Tensor:
        """fp16 [rows, cols] buffer the row-parallel Linear of THIS exchange must write its partial output into."""
        return self.partials[self.buf]

    def reduce(self, residual, out: torch.Tensor = None) -> torch.Tensor:
        """fp16(fp16(sum over ranks of the partials just written) + residual); flips to the other buffer.  One-shot writes
        `out` (allocated when None); two-shot returns this rank's result buffer of the exchange (valid until the exchange
        after the next one)."""
        a = self._args
        a.buf = self.buf
        a.residual = 0 if residual is None else residual.data_ptr()
        if self.two_shot:
            out = self.results[self.buf]
        elif out is None:
            out = torch.empty((self.rows, self.cols), dtype=torch.float16, device=self.device)
        a.out = out.data_ptr()
        self._check(self.lib.mixq_allreduce_residual(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                    "allreduce_residual")
        self.buf ^= 1
        return out

    def close(self):
        torch.cuda.synchronize()
        for p in self._mapped:
            self.lib.mixq_ipc_close_handle(p)
        self._mapped = []
        self.partials, self.results, self._keep = [], [], []
        for p in self._own:
            self.lib.mixq_peer_free(p)
        self._own = []

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


class MulticastExchange:
    """The exchange step through the NVSwitch: NVLink-SHARP multicast (multimem.ld_reduce / multimem.st / multimem.red) in ONE
    kernel of this library per rank (include/mixq.h: mixq_allreduce_multicast).  Same interface as PeerExchange.

    Plumbing only comes from torch: `torch.distributed._symmetric_memory` allocates one symmetric buffer per rank and hands
    back this rank's pointer and the multicast address; the data plane is this library's kernel.  Raises RuntimeError when
    the platform offers no multicast mapping (callers then fall back to PeerExchange).
    """

    two_shot = True   # reduce-scatter + all-gather through the switch

    def __init__(self, rows: int, cols: int, rank: int, world: int, group=None, device="cuda"):
        from . import _lib
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        if not (2 <= world <= 8):
            raise ValueError("MulticastExchange needs 2..8 ranks on one node")
        n = rows * cols
        if n % (8 * world):
            raise ValueError("rows * cols must be a multiple of 8 * world")
        self.lib, self._check = _lib.load(), _lib.check
        self.rows, self.cols, self.rank, self.world, self.device = rows, cols, rank, world, device
        nbytes = n * 2
        total = 4 * nbytes + 256           # partial 0/1, result 0/1, flag words
        grp = dist.group.WORLD if group is None else group
        try:
            symm.enable_symm_mem_for_group(grp.group_name)
        except Exception:
            pass
        self._buf = symm.empty(total // 2, dtype=torch.float16, device=torch.device(device) if not isinstance(device, torch.device) else device)
        self._buf.zero_()
        torch.cuda.synchronize()
        self._hdl = symm.rendezvous(self._buf, grp)
        mc = int(getattr(self._hdl, "multicast_ptr", 0) or 0)
        if mc == 0:
            raise RuntimeError("no NVLS multicast mapping on this platform")
        self._state = torch.zeros(2, dtype=torch.int32, device=device)      # epoch, done
        a = self._args = _lib.McAllReduceArgs()
        a.mc, a.local = mc, self._buf.data_ptr()
        a.partial_off[0], a.partial_off[1] = 0, nbytes
        a.result_off[0], a.result_off[1] = 2 * nbytes, 3 * nbytes
        a.flags_off = 4 * nbytes
        a.epoch = self._state.data_ptr()
        a.done = self._state.data_ptr() + 4
        a.n, a.world, a.rank = n, world, rank
        views = [self._buf[j * n:(j + 1) * n].view(rows, cols) for j in range(4)]
        self.partials, self.results = views[:2], views[2:]
        self.buf = 0
        dist.barrier(group=group)       # nobody signals before everybody has mapped everything

    def next_partial(self) -> torch.Tensor:
        return self.partials[self.buf]

    def reduce(self, residual, out: torch.Tensor = None) -> torch.Tensor:
        a = self._args
        a.buf = self.buf
        a.residual = 0 if residual is None else residual.data_ptr()
        self._check(self.lib.mixq_allreduce_multicast(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                    "allreduce_multicast")
        out = self.results[self.buf]
        self.buf ^= 1
        return out

    def close(self):
        torch.cuda.synchronize()
        self.partials, self.results = [], []
        self._hdl = None
        self._buf = None


def _symmetric_alloc(total_bytes: int, rank: int, world: int, group=None, device="cuda"):
    """One zero-initialised allocation of `total_bytes` per rank, mapped into every rank of the node.
    Returns (local_ptr, [ptr of rank r's copy as mapped here], multicast_ptr or 0, keepalive, close()).
    torch.distributed._symmetric_memory when it works (it also yields an NVLS multicast address on NVSwitch systems),
    else cudaMalloc + CUDA-IPC handles exchanged over the process group.  Plumbing only."""
    import torch.distributed as dist
    from . import _lib
    lib = _lib.load()
    try:
        import torch.distributed._symmetric_memory as symm
        grp = dist.group.WORLD if group is None else group
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        buf = symm.empty((total_bytes + 1) // 2, dtype=torch.float16, device=dev)
        buf.zero_()
        torch.cuda.synchronize()
        hdl = symm.rendezvous(buf, grp)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        assert ptrs[rank] == buf.data_ptr()
        return buf.data_ptr(), ptrs, mc, (buf, hdl), (lambda: None)
    except Exception as e:
        import sys
        if rank == 0:
            print(f"mixq: torch symmetric memory unavailable ({type(e).__name__}: {e}); using CUDA-IPC peer mappings", file=sys.stderr)
    p = C.c_void_p()
    _lib.check(lib.mixq_peer_alloc(total_bytes, C.byref(p)), "peer_alloc")
    h = C.create_string_buffer(64)
    _lib.check(lib.mixq_ipc_get_handle(p, h), "ipc_get_handle")
    gathered = [None] * world
    dist.all_gather_object(gathered, h.raw, group=group)
    ptrs, mapped = [0] * world, []
    for r in range(world):
        if r == rank:
            ptrs[r] = p.value
        else:
            q = C.c_void_p()
            _lib.check(lib.mixq_ipc_open_handle(gathered[r], C.byref(q)), f"ipc_open_handle(rank {r})")
            ptrs[r] = q.value
            mapped.append(q.value)

    def close():
        for q in mapped:
            lib.mixq_ipc_close_handle(q)
        lib.mixq_peer_free(p.value)
    return p.value, ptrs, 0, None, close


class PushExchange:
    """The FUSED row-parallel exchange (include/mixq.h: mixq_linear_args.y_peer + mixq_exchange_finish_poll[_quant]).

    Reduce-scatter half: the row-parallel MixLinear's epilogue warps store column slice j of this rank's partial straight into
    rank j's receive slot over NVLink (`push_targets()` hands the slot pointers to MixLinear_GEMM.forward(..., push=)), so the
    transfer rides under the GEMM's own tail.  All-gather half: `reduce(residual)` launches the small finish kernel — local fp32
    reduction of this rank's slice in rank order (+ residual, a separate fp16 rounding), the slice stored into every rank's
    result buffer (peer stores, or multimem.st through the NVSwitch), the other slices collected.
    sync="poll" (default): no handshake at all — slots and result buffers hold a sentinel until data lands, readers spin on the
    data itself and re-arm what they consume; `reduce(..., quant=)` also runs the next Linear's activation prologue on the
    assembled rows.  sync="flags": the release/acquire flag protocol (two handshakes, or one with one_shot).
    Bit-identical on all ranks and equal to oracle.mixq_oracle.tp_exchange (rank-order fp32 sum).
    The returned result buffer is valid until the NEXT reduce() has run (which re-arms it in the polling form)."""

    fused = True
    two_shot = True

    def __init__(self, rows: int, cols: int, rank: int, world: int, group=None, device="cuda", multicast: bool = True,
                 one_shot=None, sync=None):
        from . import _lib
        import os
        import torch.distributed as dist
        # sync: "poll" — no flags, the data is its own signal (mixq_exchange_finish_poll; buffers are armed with the fp16
        # sentinel 0xFFFF and re-armed by their consumer); "flags" — the release/acquire handshake (mixq_exchange_finish).
        self.sync = sync or os.environ.get("MIXQ_TP_SYNC", "poll")
        if self.sync not in ("poll", "flags"):
            raise ValueError("sync must be 'poll' or 'flags'")
        if not (2 <= world <= 8):
            raise ValueError("PushExchange needs 2..8 ranks on one node")
        if cols % world or (cols // world) % 128:
            raise ValueError("hidden size / world must be a multiple of 128 (a GEMM tile may not straddle two ranks' slices)")
        if one_shot is None:
            # one-shot: every rank pushes its WHOLE partial to every rank ((world-1) n fp16 over NVLink) and reduces all of it;
            # two-phase: reduce-scatter push + broadcast of the reduced slice, 2 (world-1)/world n.  Measured with the polling
            # form (profiles/r02_exchange_anatomy_tp*_poll.json, what one exchange adds to the row-parallel o_proj): 2 ranks
            # 8.4 us two-phase vs 11.4 one-shot; 8 ranks 18.9 vs 65.9 -> two-phase everywhere.  (With the flag protocol one-shot
            # won at 2 ranks: it saves a handshake.)
            env = os.environ.get("MIXQ_TP_ONE_SHOT")
            one_shot = (world == 2 and self.sync == "flags") if env is None else env == "1"
        self.one_shot = bool(one_shot)
        self.two_shot = not self.one_shot
        self.lib, self._check = _lib.load(), _lib.check
        self.rows, self.cols, self.rank, self.world, self.device = rows, cols, rank, world, device
        self.ns = cols // world
        nb = rows * cols * 2
        rb = nb * (world if self.one_shot else 1)          # receive area per exchange buffer
        local, ptrs, mc, self._keep, self._close = _symmetric_alloc(2 * rb + 2 * nb + 256, rank, world, group, device)
        # broadcast of the reduced slice: multimem.st through the switch (one store, (w-1)/w of the egress saved) or plain
        # stores to every peer.  Measured (us added per exchange, peer stores vs multicast): 2 ranks 8.4 vs 11.2 (the own copy
        # stays local), 4 ranks 13.5 vs 14.7, 8 ranks 19.5 vs 18.9.  MIXQ_TP_BCAST=mc|peer overrides.
        bcast = os.environ.get("MIXQ_TP_BCAST", "peer" if world <= 4 else "mc")
        if not multicast or bcast == "peer":
            mc = 0
        self.multicast = mc != 0 and not self.one_shot
        self._local, self._ptrs, self._mc = local, ptrs, mc
        self._state = torch.zeros(2, dtype=torch.int32, device=device)      # epoch, done
        self._fin, self._targets = [], []
        res0, fl = 2 * rb, 2 * rb + 2 * nb
        slot = nb if self.one_shot else rows * self.ns * 2
        for b in range(2):
            if self.sync == "poll":
                a = _lib.ExchangePollArgs()
                a.reset = local + res0 + (b ^ 1) * nb
            else:
                a = _lib.ExchangeFinishArgs()
                for r in range(world):
                    a.flags[r] = ptrs[r] + fl
                a.mc_flags = (mc + fl) if mc else 0
                a.epoch = self._state.data_ptr()
                a.done = self._state.data_ptr() + 4
            a.recv = local + b * rb
            for r in range(world):
                a.result[r] = ptrs[r] + res0 + b * nb
            a.mc_result = (mc + res0 + b * nb) if self.multicast else 0
            a.M, a.N, a.world, a.rank = rows, cols, world, rank
            a.one_shot = 1 if self.one_shot else 0
            self._fin.append(a)
            self._targets.append([ptrs[j] + b * rb + rank * slot for j in range(world)])
        self._views = [_DeviceBuffer(local + res0 + b * nb, (rows, cols), "<f2") for b in range(2)]
        self.results = [torch.as_tensor(v, device=device) for v in self._views]
        if self.sync == "poll":
            # arm the receive slots and both result buffers with the sentinel
            self._arm = _DeviceBuffer(local, ((2 * rb + 2 * nb) // 2,), "<i2")
            torch.as_tensor(self._arm, device=device).fill_(-1)
        self.buf = 0
        torch.cuda.synchronize()
        dist.barrier(group=group)       # nobody pushes or signals before everybody has mapped everything

    def push_targets(self):
        """(slot pointers, slice width, broadcast count) for MixLinear_GEMM.forward(..., push=) of THIS exchange: two-phase —
        column slice j of the partial goes to pointer j; one-shot — the whole partial goes to every pointer."""
        if self.one_shot:
            return self._targets[self.buf], 0, self.world
        return self._targets[self.buf], self.ns, 0

    can_quantize = property(lambda self: self.sync == "poll")

    def reduce(self, residual, out: torch.Tensor = None, quant=None) -> torch.Tensor:
        """quant = (norm_weight, eps, lin, cache): also run lin's activation prologue (RMSNorm -> outlier gather -> row scale ->
        quantise, fused/norm.py:24-33) on the exchanged rows inside the finish kernel and leave q_xcache / x_scale /
        activation_outliers in `cache`: lin (W_pack, or gate_proj of the SwiGLU pair) then runs forward_quantized / in its
        fused call mode without an activation prologue.  Polling form only."""
        a = self._fin[self.buf]
        a.residual = 0 if residual is None else residual.data_ptr()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if quant is not None:
            if self.sync != "poll":
                raise ValueError("the quantising finish needs sync='poll'")
            norm_w, eps, lin, cache = quant
            n = lin._n_ind
            ao, q_x = cache.ao_buffer(n), cache.q_x_buffer(self.rows, self.cols)
            self._check(self.lib.mixq_exchange_finish_poll_quant(C.byref(a), norm_w.data_ptr(), float(eps), lin._ind_buf.data_ptr(), n,
                                                                 ao.data_ptr(), ao.shape[1], q_x.data_ptr(), cache.x_scale.data_ptr(),
                                                                 lin.bit, st), "exchange_finish_poll_quant")
            cache.q_xcache, cache.activation_outliers, cache.ind = q_x, ao[:self.rows, :n], lin.ind
        else:
            fn = self.lib.mixq_exchange_finish_poll if self.sync == "poll" else self.lib.mixq_exchange_finish
            self._check(fn(C.byref(a), st), "exchange_finish")
        out = self.results[self.buf]
        self.buf ^= 1
        return out

    def close(self):
        torch.cuda.synchronize()
        self.results, self._views, self._arm = [], [], None
        self._close()
        self._close = lambda: None
        self._keep = None


def vocab_parallel_argmax(local_logits: torch.Tensor, rank: int, world: int, group=None) -> torch.Tensor:
    """Next-token ids [B] when every rank holds the logits of vocab / world consecutive vocabulary rows (vocab-parallel lm_head):
    each rank takes its local maximum, the (value, global index) pairs are all-gathered — 8 bytes per row and rank instead of the
    logits — and the best pair wins; ties go to the lowest vocabulary index, like torch.argmax over the gathered row.
    Works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    val, idx = torch.max(local_logits.float(), dim=-1)
    mine = torch.stack((val, (idx + rank * local_logits.shape[-1]).float()), dim=-1).contiguous()     # ids < 2^24: exact in fp32
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    allc = torch.stack(parts, dim=0)                    # [world, B, 2]
    best = torch.argmax(allc[..., 0], dim=0)            # first rank holding the maximum = lowest vocabulary index among ties
    return allc[..., 1].gather(0, best.unsqueeze(0)).squeeze(0).long()


def make_exchange(rows: int, cols: int, rank: int, world: int, group=None, device="cuda", kind: str = "auto"):
    """kind: "push" (reduce-scatter fused into the GEMM epilogue + finish kernel; the default), "push-nomc" (the same without
    the multicast broadcast), "multicast" (NVLS all-reduce kernel), "peer" (CUDA-IPC peer-memory all-reduce kernel),
    "auto" = push."""
    if kind in ("auto", "push", "push-nomc"):
        try:
            return PushExchange(rows, cols, rank, world, group=group, device=device, multicast=(kind != "push-nomc"))
        except ValueError:          # slice width not a multiple of the GEMM tile: every rank falls back alike
            if kind != "auto":
                raise
        kind = "multicast"
    if kind in ("multicast", "auto"):
        try:
            return MulticastExchange(rows, cols, rank, world, group=group, device=device)
        except Exception as e:      # every rank fails the same way: the decision is collective-consistent
            if kind == "multicast":
                raise
            import sys
            if rank == 0:
                print(f"mixq: NVLS multicast exchange unavailable ({type(e).__name__}: {e}); using the peer-memory exchange", file=sys.stderr)
    return PeerExchange(rows, cols, rank, world, group=group, device=device)

"""CPU, world_size 2 over gloo: the column-/row-parallel split of a Llama layer (mixq_b200/tp.py) — the product's
sharding helpers + torch.distributed plumbing, with the oracle standing in for the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

CFG = dict(hidden=256, inter=512, heads=4, kv_heads=2, head_dim=64, theta=10000.0, eps=1e-5)


def _weights(seed=0):
    g = torch.Generator().manual_seed(seed)
    H, I, D = CFG["hidden"], CFG["inter"], CFG["head_dim"]
    r = lambda n, k: (torch.randn(n, k, generator=g) * (0.5 / k ** 0.5)).half()
    ln1, ln2 = torch.ones(H).half(), torch.ones(H).half()
    ln1[[7, 100]] = 20
    ln2[[31, 200]] = 20
    return dict(ln1=ln1, ln2=ln2, wq=r(CFG["heads"] * D, H), wk=r(CFG["kv_heads"] * D, H), wv=r(CFG["kv_heads"] * D, H),
                wo=r(H, CFG["heads"] * D), wg=r(I, H), wu=r(I, H), wd=r(H, I),
                h=torch.randn(8, H, generator=g).half())


def _run_layer(w, rank, world, allreduce):
    from mixq_b200 import tp
    from oracle import llama_oracle as LO, mixq_oracle as O
    cache = O.MixLibCacheOracle(32, 6, 8)
    n = lambda t: t.contiguous().numpy()
    layer = LO.LlamaLayerOracle(n(w["ln1"]), n(w["ln2"]), n(tp.pack_qkv_shard(w["wq"], w["wk"], w["wv"], rank, world)),
                                n(tp.shard_cols(w["wo"], rank, world)), n(tp.shard_rows(w["wg"], rank, world)),
                                n(tp.shard_rows(w["wu"], rank, world)), n(tp.shard_cols(w["wd"], rank, world)), cache)
    cfg = dict(heads=CFG["heads"] // world, kv_heads=CFG["kv_heads"] // world, head_dim=CFG["head_dim"],
               theta=CFG["theta"], eps=CFG["eps"])
    h = n(w["h"]).copy()
    outs = []
    for _ in range(3):   # two discovery calls + one steady-state call
        outs.append(LO.decode_step(h.copy(), [layer], cache, cfg, allreduce))
    return outs, layer


def _worker(rank, world, port, q, peer_semantics=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mixq_b200 import tp
    from oracle import mixq_oracle as O

    def allreduce(a):
        if peer_semantics:
            # what the peer-memory exchange kernel does (mixq_allreduce_residual): every rank reads ALL partials and sums
            # them in fp32 in rank order, one rounding to fp16; the residual add stays a separate fp16 op (decode_step)
            parts = [torch.empty_like(torch.from_numpy(a)) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(a)))
            return O.tp_exchange([t.numpy() for t in parts])
        t = torch.from_numpy(a.astype(np.float32))   # fp16 partial sums, reduced; NCCL does this natively in fp16
        tp.all_reduce_sum(t)
        return t.numpy().astype(np.float16)
    outs, layer = _run_layer(_weights(), rank, world, allreduce)
    q.put((rank, [o.copy() for o in outs], layer.W_pack.ind.copy(), layer.up.ind.copy(), layer.o_proj.ind.copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("peer_semantics", [False, True])
def test_tp2_matches_single_rank(peer_semantics):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, peer_semantics)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, outs, ind_qkv, ind_up, ind_o = q.get(timeout=120)
        got[rank] = (outs, ind_qkv, ind_up, ind_o)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_outs, ref_layer = _run_layer(_weights(), 0, 1, None)
    # replicated input => identical outlier sets on every rank for the column-parallel Linears (bit-exact)
    for r in range(world):
        assert np.array_equal(got[r][1], ref_layer.W_pack.ind)
        assert np.array_equal(got[r][2], ref_layer.up.ind)
    assert sorted(ref_layer.W_pack.ind.tolist()) == [7, 100]
    # after the all-reduce every rank holds the same hidden state, within the stated 1e-2 of the 1-GPU oracle
    for t in range(3):
        a0, a1 = got[0][0][t].astype(np.float32), got[1][0][t].astype(np.float32)
        assert np.array_equal(a0, a1)
        ref = ref_outs[t].astype(np.float32)
        rel = np.linalg.norm(a0 - ref) / np.linalg.norm(ref)
        assert rel <= 1e-2, rel
    # union of the per-rank o_proj outlier columns (local ids + K-shard offset) == single-GPU set (absolute threshold)
    k_loc = CFG["heads"] * CFG["head_dim"] // world
    union = sorted(set(got[0][3].tolist()) | {c + k_loc for c in got[1][3].tolist()})
    assert union == sorted(ref_layer.o_proj.ind.tolist())


def test_shard_helpers():
    from mixq_b200 import tp
    w = torch.arange(24).reshape(4, 6)
    assert tp.shard_rows(w, 1, 2).tolist() == w[2:].tolist()
    assert tp.shard_cols(w, 0, 3).tolist() == w[:, :2].tolist()
    with pytest.raises(ValueError):
        tp.shard_rows(w, 0, 3)
    q, k, v = torch.zeros(4, 2), torch.ones(2, 2), 2 * torch.ones(2, 2)
    assert tp.pack_qkv_shard(q, k, v, 1, 2)[:, 0].tolist() == [0, 0, 1, 2]


def test_tp_exchange_oracle_properties():
    """oracle.tp_exchange (the arithmetic of allreduce_residual_kernel): exact for two ranks, rank-order deterministic, and the
    residual is a SEPARATE fp16 op (two roundings)."""
    from oracle import mixq_oracle as O
    rng = np.random.default_rng(0)
    parts = [rng.standard_normal((16, 64)).astype(np.float16) for _ in range(8)]
    res = rng.standard_normal((16, 64)).astype(np.float16)
    two = O.tp_exchange(parts[:2])
    assert np.array_equal(two, (parts[0].astype(np.float64) + parts[1].astype(np.float64)).astype(np.float16))
    assert np.array_equal(O.tp_exchange(parts[:2]), O.tp_exchange(parts[1::-1]))        # a + b == b + a in fp32
    y8 = O.tp_exchange(parts)
    ref8 = np.zeros((16, 64), np.float32)
    for p_ in parts:
        ref8 = ref8 + p_.astype(np.float32)
    assert np.array_equal(y8, ref8.astype(np.float16))
    exact = sum(p_.astype(np.float64) for p_ in parts)
    assert np.abs(y8.astype(np.float64) - exact).max() <= np.abs(exact).max() * 2 ** -10   # within fp16 rounding of the true sum
    with_res = O.tp_exchange(parts, res)
    assert np.array_equal(with_res, (y8.astype(np.float32) + res.astype(np.float32)).astype(np.float16))


def _argmax_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mixq_b200 import tp
    g = torch.Generator().manual_seed(3)
    full = torch.randn(16, 64, generator=g).half()
    full[0, 5] = full[0, 40] = 9.0            # a tie across the two shards: the lower index must win
    full[1, 33] = full[1, 34] = 9.0           # a tie inside one shard
    full[2, :] = -3.0                         # all equal: index 0
    v = full.shape[1] // world
    got = tp.vocab_parallel_argmax(full[:, rank * v:(rank + 1) * v], rank, world)
    q.put((rank, got.tolist(), torch.argmax(full.float(), dim=-1).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_vocab_parallel_argmax_two_ranks():
    """The next-token pick of the vocab-parallel lm_head (mixq_b200/tp.py: vocab_parallel_argmax) == torch.argmax over the full
    logits row on every rank, ties included."""
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_argmax_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, want in res:
        assert got == want, (rank, got, want)
    assert res[0][1][:3] == [5, 33, 0]

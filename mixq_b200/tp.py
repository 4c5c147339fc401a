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

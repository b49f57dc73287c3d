"""Multi-GPU partitioning of the attention core: batch x KV-head-group units, output all-gather only.

The reference is single-device; this is the north star's scale-out: every (sequence b, KV group g)
unit is independent — query heads [g*H/G, (g+1)*H/G) only ever read K/V head g — so the B x G grid is split across
the ranks with no exchange before or during the kernels.  The only collective is an all-gather of the bf16 outputs
(`torch.distributed`, NCCL over NVLink on the GPU box, gloo in the CPU tests), optionally chunked so the gather of
chunk c overlaps the kernel of chunk c+1 on a side stream.

Batch-first partitioning keeps the gathered tensor contiguous along dim 0; KV groups are split only when there are
fewer sequences than ranks.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    """Half-open ranges of sequences and KV groups owned by one rank."""
    b0: int
    b1: int
    g0: int
    g1: int

    @property
    def empty(self) -> bool:
        return self.b1 <= self.b0 or self.g1 <= self.g0


def partition(B: int, G: int, world: int, rank: int) -> Shard:
    """Split the B x G unit grid over `world` ranks.

    world <= B : contiguous, as-even-as-possible batch slices (all groups each).
    world >  B : `world // B` ranks share one sequence and split its KV groups (needs world % B == 0 and
                 G % (world // B) == 0 — true for every BASELINE config: B=8 / G=8 on 8 GPUs, B=1 / G=8 on 2-8 GPUs).
    """
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if B <= 0 or G <= 0:
        return Shard(0, 0, 0, 0)
    if world <= B:
        base, rem = divmod(B, world)
        b0 = rank * base + min(rank, rem)
        b1 = b0 + base + (1 if rank < rem else 0)
        return Shard(b0, b1, 0, G)
    if world % B != 0 or G % (world // B) != 0:
        raise ValueError(f"cannot split B={B} sequences x G={G} KV groups over {world} ranks evenly")
    per = world // B
    gper = G // per
    b = rank // per
    gi = rank % per
    return Shard(b, b + 1, gi * gper, (gi + 1) * gper)


def shard_qkv(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, s: Shard):
    """Views of the rank's units: q [b, T, heads of its groups, hd], k/v [b, T, its groups, hd]."""
    H, G = q.size(2), k.size(2)
    hpg = H // G
    return (q[s.b0:s.b1, :, s.g0 * hpg:s.g1 * hpg], k[s.b0:s.b1, :, s.g0:s.g1], v[s.b0:s.b1, :, s.g0:s.g1])


def gather_outputs(o_local: torch.Tensor, B: int, H: int, G: int, group: Optional[dist.ProcessGroup] = None
                   ) -> torch.Tensor:
    """All-gather the per-rank outputs into the full [B, Tq, H, hd] tensor on every rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return o_local
    Tq, hd = o_local.size(1), o_local.size(3)
    shards = [partition(B, G, world, r) for r in range(world)]
    hpg = H // G
    even_batch = world <= B and B % world == 0
    if even_batch:
        out = torch.empty((B, Tq, H, hd), dtype=o_local.dtype, device=o_local.device)
        dist.all_gather_into_tensor(out, o_local.contiguous(), group=group)
        return out
    # uneven batch slices or group-split: gather padded-equal pieces, then place them
    max_b = max(s.b1 - s.b0 for s in shards)
    max_h = max((s.g1 - s.g0) * hpg for s in shards)
    piece = torch.zeros((max_b, Tq, max_h, hd), dtype=o_local.dtype, device=o_local.device)
    piece[: o_local.size(0), :, : o_local.size(2)] = o_local
    pieces: List[torch.Tensor] = [torch.empty_like(piece) for _ in range(world)]
    dist.all_gather(pieces, piece, group=group)
    out = torch.empty((B, Tq, H, hd), dtype=o_local.dtype, device=o_local.device)
    for s, p in zip(shards, pieces):
        if s.empty:
            continue
        out[s.b0:s.b1, :, s.g0 * hpg:s.g1 * hpg] = p[: s.b1 - s.b0, :, : (s.g1 - s.g0) * hpg]
    return out


def sharded_attention(core: Callable[..., torch.Tensor], q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *,
                      q_valid: Optional[torch.Tensor] = None, k_valid: Optional[torch.Tensor] = None,
                      group: Optional[dist.ProcessGroup] = None, gather: bool = True, **core_kwargs) -> torch.Tensor:
    """Run `core` (e.g. `ops.gqa_swa_prefill` partial) on this rank's units of the FULL q/k/v and all-gather.

    `core(q, k, v, q_valid, k_valid, **core_kwargs) -> o`.  With gather=False the local output shard is returned
    (decode keeps its cache and outputs sharded for the whole generation).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B, H, G = q.size(0), q.size(2), k.size(2)
    s = partition(B, G, world, rank)
    ql, kl, vl = shard_qkv(q, k, v, s)
    qv = None if q_valid is None else q_valid[s.b0:s.b1]
    kv = None if k_valid is None else k_valid[s.b0:s.b1]
    o_local = core(ql, kl, vl, qv, kv, **core_kwargs)
    if not gather or world == 1:
        return o_local
    return gather_outputs(o_local, B, H, G, group)

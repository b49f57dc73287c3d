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


_comm_streams = {}


def _comm_stream(device: torch.device):
    """One side stream per device for the output gathers (so they overlap the next chunk's kernel)."""
    key = (device.type, device.index)
    if key not in _comm_streams:
        _comm_streams[key] = torch.cuda.Stream(device=device)
    return _comm_streams[key]


class PeerGather:
    """Gathered-output buffer in symmetric memory: every rank holds the full [B, Tq, H, hd] tensor and can write into
    its peers' copies directly (NVLink peer stores issued by the copy engines).

    Why not NCCL: the prefill kernel is persistent — one CTA per SM holding ~224 KB of shared memory — so an NCCL
    kernel enqueued on a side stream cannot become resident until the attention kernel has drained; a chunked NCCL
    all-gather therefore does not overlap (measured on 2 x B200: 10.5 ms vs 10.9 ms un-chunked for cfg5).  Copy-engine
    peer writes need no SM, so the transfer of piece c really runs under the kernel of piece c+1.

    Built on `torch.distributed._symmetric_memory` (one rendezvous per buffer; reuse the object across calls).
    """

    def __init__(self, B: int, Tq: int, H: int, hd: int, dtype: torch.dtype, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.shape = (B, Tq, H, hd)
        self.local = symm_mem.empty(self.shape, dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.local, group=self.group)
        self.peers = [self.local if r == self.rank else self.handle.get_buffer(r, self.shape, dtype)
                      for r in range(self.world)]
        self.streams = [torch.cuda.Stream(device=device) for _ in range(self.world)]

    def begin(self) -> None:
        """Every rank is done reading the previous contents (call in stream order before the first put)."""
        self.handle.barrier()

    def put(self, index, piece: torch.Tensor) -> None:
        """Write `piece` (ready on the current stream) into out[index] of every rank, on side streams."""
        ready = torch.cuda.Event()
        ready.record()
        # every rank starts with a different peer (own copy last), so no destination is hit by all ranks at once
        for i in range(1, self.world + 1):
            r = (self.rank + i) % self.world
            st = self.streams[r]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                self.peers[r][index].copy_(piece, non_blocking=True)
            piece.record_stream(st)

    def finish(self) -> torch.Tensor:
        """All pieces of all ranks have landed everywhere; returns this rank's full tensor."""
        cur = torch.cuda.current_stream()
        for st in self.streams:
            cur.wait_stream(st)
        self.handle.barrier()
        return self.local


class FusedGather:
    """Attention + output gather in ONE kernel: the prefill kernel's epilogue stores every finished O tile into the
    gathered [B, Tq, H, hd] tensor of every rank (TMA tile stores to peer memory over NVLink), so there is no second
    pass over O, no collective kernel competing for SMs with the persistent attention kernel, and the transfer
    overlaps the math tile by tile (`vats_attn_prefill_gather`).

    The gathered tensor lives in symmetric memory (`torch.distributed._symmetric_memory`, one rendezvous per buffer;
    reuse the object across calls).  `run` brackets the launch with the two device-side barriers the protocol needs:
    nobody still reads the previous contents when the peers start writing; everybody's tiles have landed before
    anybody reads.
    """

    def __init__(self, B: int, Tq: int, H: int, hd: int, G: int, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.B, self.H, self.G = B, H, G
        self.shape = (B, Tq, H, hd)
        self.local = symm_mem.empty(self.shape, dtype=torch.bfloat16, device=device)
        self.handle = symm_mem.rendezvous(self.local, group=self.group)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.shard = partition(B, G, self.world, self.rank)

    def run(self, ql: torch.Tensor, kl: torch.Tensor, vl: torch.Tensor, scale: float, causal: bool, left: int,
            right: int, q_valid: Optional[torch.Tensor] = None, k_valid: Optional[torch.Tensor] = None,
            logit_bound: float = 0.0) -> torch.Tensor:
        """ql / kl / vl: this rank's units (`shard_qkv`).  Returns the gathered tensor (this rank's copy)."""
        from . import ops
        s = self.shard
        hpg = self.H // self.G
        self.handle.barrier()        # every rank is done with the previous contents
        ops.gqa_swa_prefill_gather(ql, kl, vl, self.local, self.peer_ptrs, self.rank, s.b0, s.g0 * hpg, q_valid, k_valid,
                                   scale, causal, left, right, logit_bound)
        self.handle.barrier()        # every rank's tiles have landed everywhere
        return self.local


def local_attention_gather(core: Callable[..., torch.Tensor], ql: torch.Tensor, kl: torch.Tensor, vl: torch.Tensor,
                           B: int, H: int, G: int, *, q_valid: Optional[torch.Tensor] = None,
                           k_valid: Optional[torch.Tensor] = None, chunks: int = 1,
                           group: Optional[dist.ProcessGroup] = None, peer: Optional["PeerGather"] = None,
                           **core_kwargs) -> torch.Tensor:
    """Attention on this rank's LOCAL units (as returned by `shard_qkv`) + all-gather into the full [B,Tq,H,hd].

    With chunks > 1 (and an even batch split) the local work is cut into pieces whose output slices are contiguous in
    the gathered tensor, and the gather of piece c runs on a side stream while the kernel of piece c+1 computes:
      * >= 2 local sequences: pieces are batch slices;
      * one local sequence and causal attention: pieces are query-token ranges [t0, t1) attending keys [0, t1 + Tk-Tq)
        (bottom-right alignment keeps the mask identical), each a contiguous slice of that sequence's output rows.
    Anything else falls back to one kernel + one gather.  With `peer` (a PeerGather of the output shape) the pieces
    travel as copy-engine peer writes instead of NCCL all-gathers, which is what actually overlaps with the persistent
    attention kernel.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return core(ql, kl, vl, q_valid, k_valid, **core_kwargs)
    rank = dist.get_rank(group)
    nb, Tq, Tk = ql.size(0), ql.size(1), kl.size(1)
    even_batch = world <= B and B % world == 0 and ql.size(2) == H
    causal = bool(core_kwargs.get("causal", False))
    mode = None
    if chunks > 1 and even_batch:
        if nb >= 2:
            mode = "batch"
            chunks = min(chunks, nb)
        elif nb == 1 and causal and Tq >= 2 * chunks:
            mode = "tokens"
    if peer is not None and even_batch and mode is None:
        mode, chunks = "batch", 1
    if mode is None:
        return gather_outputs(core(ql, kl, vl, q_valid, k_valid, **core_kwargs), B, H, G, group)

    hd = ql.size(3)
    on_gpu = ql.is_cuda
    if peer is not None:
        peer.begin()
        out = peer.local
    else:
        out = torch.empty((B, Tq, H, hd), dtype=ql.dtype, device=ql.device)
    comm = _comm_stream(ql.device) if on_gpu else None
    bounds = [(c * (nb if mode == "batch" else Tq)) // chunks for c in range(chunks + 1)]
    for c in range(chunks):
        lo, hi = bounds[c], bounds[c + 1]
        if hi <= lo:
            continue
        if mode == "batch":
            o_c = core(ql[lo:hi], kl[lo:hi], vl[lo:hi], None if q_valid is None else q_valid[lo:hi],
                       None if k_valid is None else k_valid[lo:hi], **core_kwargs)
            index = lambda r: (slice(r * nb + lo, r * nb + hi),)
        else:
            kend = hi + (Tk - Tq)
            o_c = core(ql[:, lo:hi], kl[:, :kend], vl[:, :kend], None if q_valid is None else q_valid[:, lo:hi],
                       None if k_valid is None else k_valid[:, :kend], **core_kwargs)
            index = lambda r: (slice(r, r + 1), slice(lo, hi))
        o_c = o_c.contiguous()
        if peer is not None:
            peer.put(index(rank), o_c)
            continue
        dst = [out[index(r)] for r in range(world)]
        if on_gpu:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                dist.all_gather(dst, o_c, group=group)
            o_c.record_stream(comm)
        else:
            dist.all_gather(dst, o_c, group=group)
    if peer is not None:
        return peer.finish()
    if on_gpu:
        torch.cuda.current_stream().wait_stream(comm)
    return out


def sharded_attention(core: Callable[..., torch.Tensor], q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *,
                      q_valid: Optional[torch.Tensor] = None, k_valid: Optional[torch.Tensor] = None,
                      group: Optional[dist.ProcessGroup] = None, gather: bool = True, chunks: int = 1,
                      **core_kwargs) -> torch.Tensor:
    """Run `core` (e.g. `ops.gqa_swa_prefill` partial) on this rank's units of the FULL q/k/v and all-gather.

    `core(q, k, v, q_valid, k_valid, **core_kwargs) -> o`.  With gather=False the local output shard is returned
    (decode keeps its cache and outputs sharded for the whole generation).  chunks > 1 overlaps the gather with the
    kernels (see `local_attention_gather`).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B, H, G = q.size(0), q.size(2), k.size(2)
    s = partition(B, G, world, rank)
    ql, kl, vl = shard_qkv(q, k, v, s)
    qv = None if q_valid is None else q_valid[s.b0:s.b1]
    kv = None if k_valid is None else k_valid[s.b0:s.b1]
    if not gather or world == 1:
        return core(ql, kl, vl, qv, kv, **core_kwargs)
    return local_attention_gather(core, ql, kl, vl, B, H, G, q_valid=qv, k_valid=kv, chunks=chunks, group=group,
                                  **core_kwargs)

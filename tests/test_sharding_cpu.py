"""CPU, world_size 2 (gloo): the batch x KV-group partitioner and the output all-gather reproduce the unsharded
result.  The compute inside is the oracle (test infrastructure) — only the host-side sharding logic is under test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vats_multimodal_lm_b200 import sharding


def test_partition_covers_grid_exactly_once():
    for B, G, world in [(8, 8, 8), (8, 8, 4), (8, 8, 2), (1, 8, 8), (1, 8, 2), (2, 8, 8), (64, 8, 8), (5, 4, 2),
                        (7, 3, 3), (3, 8, 1)]:
        seen = torch.zeros(B, G, dtype=torch.int32)
        for r in range(world):
            s = sharding.partition(B, G, world, r)
            seen[s.b0:s.b1, s.g0:s.g1] += 1
        assert torch.all(seen == 1), (B, G, world)
    with pytest.raises(ValueError):
        sharding.partition(3, 8, 8, 0)   # 8 ranks over 3 sequences: not an even group split


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, G, H, results):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import mask_predicate, sdpa_explicit
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        T, hd = 12, 8
        q = torch.randn(B, T, H, hd, generator=g)
        k = torch.randn(B, T, G, hd, generator=g)
        v = torch.randn(B, T, G, hd, generator=g)
        kvalid = torch.rand(B, T, generator=g) > 0.2
        kvalid[:, 0] = True

        def core(q_, k_, v_, qv, kv, scale, causal, left, right):
            m = mask_predicate(q_.size(0), q_.size(1), k_.size(1), causal, left, right, qv, kv)
            return sdpa_explicit(q_, k_, v_, m, scale)

        kw = dict(scale=0.3, causal=True, left=5, right=0)
        full = core(q, k, v, None, kvalid, **kw)
        out = sharding.sharded_attention(core, q, k, v, k_valid=kvalid, **kw)
        ok = torch.allclose(out, full, atol=1e-6)
        local = sharding.sharded_attention(core, q, k, v, k_valid=kvalid, gather=False, **kw)
        s = sharding.partition(B, G, world, rank)
        hpg = H // G
        ok = ok and torch.allclose(local, full[s.b0:s.b1, :, s.g0 * hpg:s.g1 * hpg], atol=1e-6)
        # chunked (overlappable) gather: batch pieces when a rank owns >= 2 sequences, token pieces for one causal
        # sequence per rank, plain gather otherwise — always the same values
        for chunks in (2, 3):
            out_c = sharding.sharded_attention(core, q, k, v, k_valid=kvalid, chunks=chunks, **kw)
            ok = ok and torch.allclose(out_c, full, atol=1e-6)
        # chunked prefill against a longer key range (Tq < Tk, bottom-right aligned)
        full2 = core(q[:, 4:], k, v, None, kvalid, **kw)
        out2 = sharding.sharded_attention(core, q[:, 4:], k, v, k_valid=kvalid, chunks=2, **kw)
        ok = ok and torch.allclose(out2, full2, atol=1e-6)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,G,H", [(4, 2, 4), (3, 2, 6), (1, 4, 8), (2, 2, 4)])
def test_sharded_equals_unsharded_world2(B, G, H):
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, G, H, results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)

"""Two-GPU check of the sharded path on real devices (skipped on a one-GPU box): NCCL all-gather, chunked gathers and
copy-engine peer writes into symmetric memory all reproduce the single-GPU result bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, results):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vats_multimodal_lm_b200 import ops, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ok = True
        for (B, T, H, G, hd, causal, left) in [(4, 300, 8, 2, 64, True, 100), (2, 700, 4, 2, 128, True, 256),
                                               (6, 196, 4, 2, 72, False, -1), (1, 900, 8, 2, 128, True, 300)]:
            g = torch.Generator().manual_seed(11)
            q = torch.nn.functional.normalize(torch.randn(B, T, H, hd, generator=g), dim=-1).bfloat16().to(dev)
            k = torch.nn.functional.normalize(torch.randn(B, T, G, hd, generator=g), dim=-1).bfloat16().to(dev)
            v = torch.randn(B, T, G, hd, generator=g).bfloat16().to(dev)
            scale = hd ** -0.5

            def core(q_, k_, v_, qv, kv, causal=False):
                return ops.gqa_swa_prefill(q_, k_, v_, qv, kv, scale, causal, left, 0 if causal else -1, 0)

            full = core(q, k, v, None, None, causal=causal)
            plain = sharding.sharded_attention(core, q, k, v, causal=causal)
            chunked = sharding.sharded_attention(core, q, k, v, chunks=3, causal=causal)
            s = sharding.partition(B, G, world, rank)
            ql, kl, vl = sharding.shard_qkv(q, k, v, s)
            pg = sharding.PeerGather(B, T, H, hd, torch.bfloat16, dev)
            for _ in range(2):   # twice: the buffer is reused across calls
                peer = sharding.local_attention_gather(core, ql, kl, vl, B, H, G, chunks=3, causal=causal, peer=pg)
            fused = None
            if hd % 8 == 0:   # the fused gather needs a TMA-addressable output
                fg = sharding.FusedGather(B, T, H, hd, G, dev)
                for _ in range(2):
                    fused = fg.run(ql, kl, vl, scale, causal, left, 0 if causal else -1).clone()
            torch.cuda.synchronize()
            if fused is not None:
                # same tile kernel, same per-unit arithmetic: bit-identical; (<= 256 keys: `full` ran on the resident-K/V
                # kernel, the fused gather on the tile kernel — compare with the tolerance there)
                same = torch.equal(fused, full) if T > 256 else torch.allclose(fused.float(), full.float(), atol=2e-2, rtol=0)
                ok = ok and same
                if not same:
                    print(f"rank {rank}: fused gather mismatch for {(B, T, H, G, hd)}: "
                          f"{(fused.float() - full.float()).abs().max().item():.3e}")
            # token-chunked pieces see a shorter key range: same mask, same tiles for the rows they own, but a
            # different tile count can change the summation order of the online softmax -> compare with a tolerance
            for name, t in (("plain", plain), ("chunked", chunked), ("peer", peer)):
                same = torch.allclose(t.float(), full.float(), atol=2e-2, rtol=0)
                ok = ok and same
                if not same:
                    print(f"rank {rank}: {name} mismatch for {(B, T, H, G, hd)}: "
                          f"{(t.float() - full.float()).abs().max().item():.3e}")
            ok = ok and torch.equal(plain, full)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_gathers_match_single_gpu_world2():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_process_two_devices():
    """One thread launching on cuda:0 and then on cuda:1: the opt-in shared-memory attribute of every large-smem kernel
    (tile, resident-K/V, short-sequence, decode) is per device — a cache that is not keyed on the device makes the
    second device's launches fail with 'invalid argument'."""
    import math

    from conftest import make_qkv
    from gpu_util import check_close, oracle_prefill, run_prefill
    from oracle import decode_explicit
    from vats_multimodal_lm_b200 import ops

    cases = [((2, 300, 300, 4, 2, 64), True, 100, 0),       # tile kernel
             ((80, 196, 196, 4, 2, 72), False, -1, -1),     # resident-K/V kernel
             ((40, 8, 8, 8, 2, 66), False, -1, -1)]         # short-sequence kernel
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        for shape, causal, left, right in cases:
            N, Tq, Tk, H, G, hd = shape
            q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=N)
            scale = 1.0 / math.sqrt(hd)
            o = run_prefill(q, k, v, scale, causal, left, right, device=dev)
            torch.cuda.synchronize(dev)
            assert o.device == torch.device(dev)
            check_close(o, oracle_prefill(q, k, v, scale, causal, left, right), f"{dev} {shape}")
        B, S, H, G, hd = 3, 700, 8, 2, 128
        g = torch.Generator().manual_seed(5)
        kc = torch.nn.functional.normalize(torch.randn(B, S, G, hd, generator=g), dim=-1).bfloat16()
        vc = torch.randn(B, S, G, hd, generator=g).bfloat16()
        qd = torch.nn.functional.normalize(torch.randn(B, H, hd, generator=g), dim=-1).bfloat16()
        lens = torch.tensor([700, 1, 333], dtype=torch.int32)
        o = ops.gqa_swa_decode(qd.to(dev), kc.to(dev), vc.to(dev), lens.to(dev), hd ** -0.5, 256)
        torch.cuda.synchronize(dev)
        check_close(o, decode_explicit(qd, kc, vc, lens, hd ** -0.5, 256), f"{dev} decode")

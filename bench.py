#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the GQA + sliding-window attention hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extra]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[1], the config the metric is quoted on that fits one GPU):
    LLM KV-cache decode — single-query GQA+SWA over cached K/V, batch 64, context 8192, window 4096,
    H=32, G=8, hd=128, bf16.  A "step" is one decode call over the whole batch.
    metric = decode HBM GB/s = algorithmic bytes (un-expanded K/V window + q + o, each touched once) / time.
At N GPUs every rank owns its own batch of 64 sequences and its own cache (weak scaling, no data-path collective;
the cache stays sharded for the whole generation — SURVEY.md §8e).

The other BASELINE configs (prefill cfg1/cfg5, ViT cfg3/cfg4) are measured in the same run and reported under
"other_workloads" (TFLOP/s, GB/s, fraction of the bounding roofline); cfg5 at N>1 is sharded by batch x KV group with
an NCCL all-gather of the outputs, reported compute-only and compute+gather.

`--impl reference` times the reference's CPU path (oracle/cpu_baseline.py, a port: /root/reference does not exist on
the GPU box) with all host threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

LOG2E = 1.4426950408889634

# ------------------------------------------------------------------------------------------------ workloads
CFG2 = dict(name="cfg2_decode", B=64, S=8192, H=32, G=8, hd=128, left=4096)


def decode_bytes(c, B=None):
    B = c["B"] if B is None else B
    keys = min(c["S"], c["left"] + 1)
    return 2 * B * c["G"] * keys * c["hd"] * 2 + 2 * B * c["H"] * c["hd"] * 2


def decode_flops(c, B=None):
    B = c["B"] if B is None else B
    keys = min(c["S"], c["left"] + 1)
    return 4 * B * c["H"] * c["hd"] * keys


def prefill_pairs(T, causal, left):
    """allowed (i, j) pairs per head for Tq == Tk == T."""
    if not causal:
        return T * T
    if left < 0 or left >= T:
        return T * (T + 1) // 2
    w = left + 1
    return w * (w + 1) // 2 + (T - w) * w


PREFILL_CFGS = [
    # name, N, T, H, G, hd, causal, left, bound
    dict(name="cfg1_llm_prefill_T4096", N=1, T=4096, H=24, G=8, hd=60, causal=True, left=384, bound="tensor"),
    dict(name="cfg3_vit2d", N=256, T=196, H=16, G=8, hd=72, causal=False, left=-1, bound="hbm"),
    dict(name="cfg4a_vit3d_spatial", N=512, T=196, H=32, G=8, hd=66, causal=False, left=-1, bound="hbm"),
    dict(name="cfg4b_vit3d_temporal", N=12544, T=8, H=32, G=8, hd=66, causal=False, left=-1, bound="hbm"),
    dict(name="cfg5_long_prefill", N=8, T=32768, H=32, G=8, hd=128, causal=True, left=4096, bound="tensor"),
]


def prefill_flops(c, N=None):
    N = c["N"] if N is None else N
    return 4 * N * c["H"] * c["hd"] * prefill_pairs(c["T"], c["causal"], c["left"])


def prefill_bytes(c, N=None):
    N = c["N"] if N is None else N
    return 2 * N * c["T"] * c["hd"] * (2 * c["H"] + 2 * c["G"])


# ------------------------------------------------------------------------------------------------ helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def gen_unit_bf16(shape, seed, device, normalize):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
    if normalize:
        x = torch.nn.functional.normalize(x, dim=-1)
    return x.to(torch.bfloat16)


def barrier_sync(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x: float, world: int) -> float:
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """CPU port of the reference path on a bounded sample of cfg2 (and nothing of ours on the timed path)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle.cpu_baseline import reference_decode_cpu
    c = CFG2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = 8  # bounded sample: 8 of the 64 sequences (the expanded bf16 cache for all 64 is 4.3 GB)
    g = torch.Generator().manual_seed(1234 + 2)
    kc = torch.nn.functional.normalize(torch.randn(Bs, c["S"], c["G"], c["hd"], generator=g), dim=-1).bfloat16()
    vc = torch.randn(Bs, c["S"], c["G"], c["hd"], generator=g).bfloat16()
    q = torch.nn.functional.normalize(torch.randn(Bs, c["H"], c["hd"], generator=g), dim=-1).bfloat16()
    scale = c["hd"] ** -0.5
    fn = lambda: reference_decode_cpu(q, kc, vc, c["S"], scale, c["left"])
    for _ in range(max(1, args.warmup)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    gbs = decode_bytes(c, Bs) / (dt / args.steps) / 1e9
    sample = f"{Bs} of {c['B']} sequences per step (same context/window/heads), bf16, torch CPU SDPA with expanded heads"
    line = {
        "impl": "reference", "metric": "decode_hbm_gbps", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cfg2 LLM KV-cache decode: B=64 ctx=8192 window=4096 H=32 G=8 hd=128 bf16",
                   "sample": sample},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm
def bench_decode(args, world, peaks):
    from vats_multimodal_lm_b200 import _ffi, ops
    c = CFG2
    dev = torch.device("cuda", torch.cuda.current_device())
    rank = dist_env()[0]
    B, S, H, G, hd, left = c["B"], c["S"], c["H"], c["G"], c["hd"], c["left"]
    if args.cache_layout == "bhsd":
        # head-major storage [B, G, S, hd] exposed with the reference's [B, S, G, hd] shape (a strided view): every
        # (sequence, KV group) is one contiguous 1 MB run of keys — the layout the drop-in KVCache allocates
        kc = gen_unit_bf16((B, G, S, hd), 1234 + 2 + 100 * rank, dev, True).permute(0, 2, 1, 3)
        vc = gen_unit_bf16((B, G, S, hd), 2234 + 2 + 100 * rank, dev, False).permute(0, 2, 1, 3)
    else:
        kc = gen_unit_bf16((B, S, G, hd), 1234 + 2 + 100 * rank, dev, True)
        vc = gen_unit_bf16((B, S, G, hd), 2234 + 2 + 100 * rank, dev, False)
    q = gen_unit_bf16((B, H, hd), 3234 + 2 + 100 * rank, dev, True)
    lens = torch.full((B,), S, dtype=torch.int32, device=dev)
    scale = hd ** -0.5
    step = lambda: ops.gqa_swa_decode(q, kc, vc, lens, scale, left)

    for _ in range(max(3, args.warmup)):
        step()
    launches_per_step = None
    step()
    launches_per_step = _ffi.last_launch_count()

    # ---- device-resident timing: K steps bracketed by barrier + synchronize, per-step CUDA events inside
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(torch.cuda.current_device())
    barrier_sync(world)
    sampler.start()
    t0 = time.perf_counter()
    for a, b in evs:
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    barrier_sync(world)
    wall = max_over_ranks(t1 - t0, world)
    per_launch_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = statistics.mean(per_launch_ms)
    dev_ms_max = max_over_ranks(dev_ms, world)
    nbytes = decode_bytes(c)
    value = world * nbytes / (wall / args.steps) / 1e9
    achieved = nbytes / (dev_ms * 1e-3) / 1e9

    # ---- end to end through the public ops with HOST buffers: H2D of the step's projected q and new k/v (pinned),
    #      fused qk-norm + RoPE + cache append (vats::decode_prepare), decode, D2H of the result, every step
    qh = q.cpu().pin_memory()
    knh = kc[:, S - 1].cpu().pin_memory()
    vnh = vc[:, S - 1].cpu().pin_memory()
    lens_h = lens.cpu().pin_memory()
    oh = torch.empty((B, H, hd), dtype=torch.bfloat16).pin_memory()

    pos = torch.arange(S, dtype=torch.float32, device=dev)
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.float32, device=dev) / hd))
    cos_t, sin_t = torch.cos(torch.outer(pos, inv_freq)), torch.sin(torch.outer(pos, inv_freq))

    def e2e_step():
        qd = qh.to(dev, non_blocking=True)
        kn = knh.to(dev, non_blocking=True)
        vn = vnh.to(dev, non_blocking=True)
        ld = lens_h.to(dev, non_blocking=True)
        # L2-normalise + rotate q and k at position seq_len-1, append k/v there (one launch), then attend
        qr = ops.decode_prepare(qd, kn, vn, kc, vc, ld, cos_t, sin_t, True, 1e-6)
        o = ops.gqa_swa_decode(qr, kc, vc, ld, scale, left)
        oh.copy_(o, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(3):
        e2e_step()
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    t1 = time.perf_counter()
    barrier_sync(world)
    e2e_wall = max_over_ranks(t1 - t0, world)
    e2e_eager = world * nbytes / (e2e_wall / args.steps) / 1e9

    # the same step captured once in a CUDA graph (pinned H2D copies, the two kernels, the D2H copy) and replayed:
    # one launch per step instead of six plus the Python op overhead
    e2e_graph = None
    graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            def body():
                qd = qh.to(dev, non_blocking=True)
                kn = knh.to(dev, non_blocking=True)
                vn = vnh.to(dev, non_blocking=True)
                ld = lens_h.to(dev, non_blocking=True)
                qr = ops.decode_prepare(qd, kn, vn, kc, vc, ld, cos_t, sin_t, True, 1e-6)
                o = ops.gqa_swa_decode(qr, kc, vc, ld, scale, left)
                oh.copy_(o, non_blocking=True)
            for _ in range(2):
                body()
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ref_o = oh.clone()
        for _ in range(3):
            graph.replay()
            torch.cuda.synchronize()
        if not torch.equal(ref_o, oh):
            raise RuntimeError("graph replay does not reproduce the eager step")
    except Exception as e:   # the eager number stands
        graph = None
        sys.stderr.write(f"bench: CUDA-graph e2e skipped ({type(e).__name__}: {e})\n")
    # every rank must take the same branch (the timed loop contains barriers)
    if max_over_ranks(0.0 if graph is not None else 1.0, world) == 0.0:
        barrier_sync(world)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            graph.replay()
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier_sync(world)
        e2e_graph = world * nbytes / (max_over_ranks(t1 - t0, world) / args.steps) / 1e9
    clocks = sampler.stop()   # sampled over the device-timed region and the end-to-end region (both under load)
    e2e_value = e2e_graph if e2e_graph is not None else e2e_eager
    h2d = qh.numel() * 2 + knh.numel() * 2 + vnh.numel() * 2 + lens_h.numel() * 4
    d2h = oh.numel() * 2

    traffic = None
    prof = os.path.join(ROOT, "profiles", "decode_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    return dict(
        value=value, ms_per_step=wall / args.steps * 1e3, dev_ms=dev_ms_max, clocks=clocks,
        launches=launches_per_step * args.steps,
        e2e=dict(value=e2e_value, unit="GB/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 mode="cuda_graph_replay" if e2e_graph is not None else "eager", eager_value=e2e_eager,
                 note="KV cache is resident state (it never leaves HBM between steps); the step's inputs are the new "
                      "token's projected q/k/v: H2D, vats::decode_prepare (qk-norm + RoPE + cache append, one "
                      "launch), vats::gqa_swa_decode, D2H"),
        roofline=dict(bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s",
                      frac=achieved / peaks["hbm_gbs"], traffic=traffic,
                      peak_source=f"MEASURED_PEAKS.json hbm_gbs ({peaks['source']})",
                      frac_of_8tbs_spec=achieved / 8000.0,
                      kernel="decode_mma_kernel<128,4,32> (one launch per step: TMA ring + mma.sync consumers + flush warp)",
                      algorithmic_bytes_per_launch=nbytes,
                      launch_ms_mean=dev_ms, launch_ms_min=min(per_launch_ms)),
    )


def bench_prefill_cfg(c, peaks, steps, warmup, world, rank, shard: bool, layout: str = "dense"):
    """One prefill-class workload. With shard=True the N sequences (x KV groups) are split over the ranks and the
    outputs all-gathered (cfg5 / ViT); returns per-rank-max timings."""
    from vats_multimodal_lm_b200 import ops, sharding
    dev = torch.device("cuda", torch.cuda.current_device())
    N, T, H, G, hd = c["N"], c["T"], c["H"], c["G"], c["hd"]
    sh = sharding.partition(N, G, world, rank) if shard else sharding.Shard(0, N, 0, G)
    nb, ng = sh.b1 - sh.b0, sh.g1 - sh.g0
    hpg = H // G
    q = gen_unit_bf16((nb, T, ng * hpg, hd), 1234 + rank, dev, True)
    k = gen_unit_bf16((nb, T, ng, hd), 2234 + rank, dev, True)
    v = gen_unit_bf16((nb, T, ng, hd), 3234 + rank, dev, False)
    if layout == "module" and hd % 8 != 0:
        # the layout the drop-in modules produce for head dims TMA cannot address (modules/_common.py): same logical
        # [N,T,heads,hd] tensors, head stride rounded up to 8 elements
        from vats_multimodal_lm_b200.modules._common import _to_kernel_layout
        q, k, v = _to_kernel_layout(q), _to_kernel_layout(k), _to_kernel_layout(v)
    scale = hd ** -0.5
    step = lambda: ops.gqa_swa_prefill(q, k, v, None, None, scale, c["causal"], c["left"], 0 if c["causal"] else -1, 0)
    flush = None
    if prefill_bytes(c) / max(world if shard else 1, 1) < 2.5e8:   # working set could sit in the 126 MB L2
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        step()
    evs = []
    barrier_sync(world)
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        o = step()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
    ms = max_over_ranks(ms, world)
    res = dict(ms_compute=ms, l2="flushed between iterations" if flush is not None else "inputs larger than L2")
    if shard and world > 1:
        # compute + NCCL all-gather of the outputs (the only collective of the path)
        barrier_sync(world)
        t0 = time.perf_counter()
        for _ in range(steps):
            o = step()
            full = sharding.gather_outputs(o, N, H, G)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        res["ms_compute_plus_allgather"] = max_over_ranks((t1 - t0) / steps * 1e3, world)
        del full
        # the same, cut into 4 pieces whose gathers run on a side stream under the next piece's kernel
        core = lambda q_, k_, v_, qv_, kv_, causal=False: ops.gqa_swa_prefill(
            q_, k_, v_, qv_, kv_, scale, causal, c["left"], 0 if causal else -1, 0)
        for _ in range(2):
            full = sharding.local_attention_gather(core, q, k, v, N, H, G, chunks=4, causal=c["causal"])
        barrier_sync(world)
        t0 = time.perf_counter()
        for _ in range(steps):
            full = sharding.local_attention_gather(core, q, k, v, N, H, G, chunks=4, causal=c["causal"])
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        res["ms_compute_plus_allgather_overlapped"] = max_over_ranks((t1 - t0) / steps * 1e3, world)
        del full
        # ... and with copy-engine peer writes into symmetric memory instead of NCCL (needs no SM: really overlaps)
        pg, err = None, None
        try:
            pg = sharding.PeerGather(N, T, H, hd, torch.bfloat16, dev)
        except Exception as e:
            err = f"{type(e).__name__}: {e}"
        if max_over_ranks(0.0 if pg is not None else 1.0, world) == 0.0:   # every rank takes the same branch
            for _ in range(2):
                full = sharding.local_attention_gather(core, q, k, v, N, H, G, chunks=4, causal=c["causal"], peer=pg)
            barrier_sync(world)
            t0 = time.perf_counter()
            for _ in range(steps):
                full = sharding.local_attention_gather(core, q, k, v, N, H, G, chunks=4, causal=c["causal"], peer=pg)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            res["ms_compute_plus_peer_gather_overlapped"] = max_over_ranks((t1 - t0) / steps * 1e3, world)
            del full
        else:
            res["peer_gather_error"] = err or "symmetric memory unavailable on another rank"
        del pg
    units = world if (not shard) else 1  # unsharded workloads are replicated per rank (weak)
    fl = prefill_flops(c) * units
    by = prefill_bytes(c) * units
    res.update(tflops=fl / (ms * 1e-3) / 1e12, gbs=by / (ms * 1e-3) / 1e9, flops=fl, bytes=by,
               tokens_per_s=c["N"] * c["T"] * units / (ms * 1e-3))
    if c["bound"] == "tensor":
        res["roofline"] = dict(bound="tensor", achieved=res["tflops"] / max(world, 1) if shard else res["tflops"] / units,
                               peak=peaks["bf16_tflops"], unit="TFLOP/s")
    else:
        res["roofline"] = dict(bound="hbm", achieved=res["gbs"] / max(world, 1) if shard else res["gbs"] / units,
                               peak=peaks["hbm_gbs"], unit="GB/s")
    res["roofline"]["frac"] = res["roofline"]["achieved"] / res["roofline"]["peak"]
    res["roofline"]["note"] = "per-GPU achieved vs measured per-GPU peak"
    del q, k, v
    torch.cuda.empty_cache()
    return res


def cpu_baseline_decode():
    """Oracle-side port of the reference CPU path on a bounded sample, rank 0 / N=1 only."""
    from oracle.cpu_baseline import reference_decode_cpu, time_callable
    c = CFG2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = 8
    g = torch.Generator().manual_seed(1234 + 2)
    kc = torch.nn.functional.normalize(torch.randn(Bs, c["S"], c["G"], c["hd"], generator=g), dim=-1).bfloat16()
    vc = torch.randn(Bs, c["S"], c["G"], c["hd"], generator=g).bfloat16()
    q = torch.nn.functional.normalize(torch.randn(Bs, c["H"], c["hd"], generator=g), dim=-1).bfloat16()
    calls, dt = time_callable(lambda: reference_decode_cpu(q, kc, vc, c["S"], c["hd"] ** -0.5, c["left"]), 10.0, 200)
    gbs = decode_bytes(c, Bs) * calls / dt / 1e9
    return dict(value=gbs, unit="GB/s", cores=cores, kind="port",
                sample=f"{Bs} of 64 sequences x {calls} calls in {dt:.1f} s, bf16, torch CPU SDPA with heads expanded "
                       f"as the reference does (oracle/cpu_baseline.py)")


def run_ours(args):
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the attention kernels have no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = load_peaks()
    r = bench_decode(args, world, peaks)

    other = {}
    if not args.no_extra:
        for c in PREFILL_CFGS:
            shard = world > 1 and c["name"] in ("cfg5_long_prefill", "cfg3_vit2d", "cfg4a_vit3d_spatial",
                                               "cfg4b_vit3d_temporal")
            steps = max(2, min(args.steps, 5 if c["name"] == "cfg5_long_prefill" else 10))
            try:
                other[c["name"]] = bench_prefill_cfg(c, peaks, steps, 3, world, rank, shard)
                other[c["name"]]["sharded"] = shard
                other[c["name"]]["layout"] = "dense [N,T,heads,hd]"
                if c["hd"] % 8 != 0 and c["T"] > 32:   # (the modules do not pad sequences of <= 32 keys)
                    r2 = bench_prefill_cfg(c, peaks, steps, 3, world, rank, shard, layout="module")
                    r2["sharded"] = shard
                    r2["layout"] = "as produced by the drop-in modules: head stride padded to 8 elements (TMA-addressable)"
                    other[c["name"] + "_module_layout"] = r2
            except Exception as e:  # a secondary workload must not take the headline down
                other[c["name"]] = {"error": f"{type(e).__name__}: {e}"}
    cpu = None
    if rank == 0 and world == 1:
        try:
            cpu = cpu_baseline_decode()
        except Exception as e:
            cpu = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {
            "metric": "decode_hbm_gbps", "value": r["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": "cfg2 LLM KV-cache decode: single-query GQA+SWA, B=64 per GPU, ctx=8192, window=4096, "
                            "H=32, G=8, hd=128, bf16 (BASELINE.json configs[1])",
                "bytes_per_step_per_gpu": decode_bytes(CFG2), "flops_per_step_per_gpu": decode_flops(CFG2),
                "l2": "inputs larger than L2 (1.07 GB of K/V streamed per step vs 126 MB L2)",
                "parallelism": f"batch-sharded x{world}, caches stay sharded, no collective",
                "cache_layout": args.cache_layout,
            },
            "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["launches"], "roofline": r["roofline"],
            "device_ms_per_step": r["dev_ms"],
            "cpu_baseline": cpu, "other_workloads": other, "peaks": peaks,
        }
        emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--cache-layout", default="bshd", choices=["bshd", "bhsd"],
                    help="storage order of the KV cache behind its [B,S,G,hd] shape")
    args = ap.parse_args()
    # Exactly one line on stdout: libraries write banners there (NCCL prints its version at the first communicator),
    # so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

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

The other BASELINE configs (prefill cfg1 at T = 32 / 384 / 4096, cfg5, ViT cfg3 / cfg4a / cfg4b, and the decode of the
default LLM geometry, hd 60) are measured in the same run and reported under "other_workloads": time over >= 20
iterations, TFLOP/s, GB/s, fraction of the bounding roofline, clocks sampled during the workload's own timed region,
max-abs / relative-L2 error of sampled output rows against the fp32 oracle, and (N=1) the reference's CPU path on
the host cores (core-only fp32 SDPA with the explicit mask; cfg5 on a bounded slice, extrapolated and labelled so).
cfg3 also carries an end-to-end number (pinned host q/k/v -> H2D -> kernel -> D2H).  At N>1 cfg5 and the ViT
workloads are sharded by batch x KV group and the outputs gathered four ways (NCCL all-gather, chunked NCCL, copy-engine
peer writes, and the gather FUSED into the kernel epilogue as TMA stores to peer memory); every variant's gathered
tensor is checked against a locally recomputed remote slice ("sharded_parity_max_abs").

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


CFG2_MEDIUM = dict(name="cfg2_medium_hd60_decode", B=64, S=8192, H=24, G=8, hd=60, left=4096)

PREFILL_CFGS = [
    # name, N, T, H, G, hd, causal, left, bound
    dict(name="cfg1_llm_prefill_T32", N=1, T=32, H=24, G=8, hd=60, causal=True, left=384, bound="latency"),
    dict(name="cfg1_llm_prefill_T384", N=1, T=384, H=24, G=8, hd=60, causal=True, left=384, bound="latency"),
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
    # (the new token's q, k, v arrive as ONE host buffer [B, H + 2G, hd] — the layout of the reference's fused w_qkv
    #  projection, src/optimized_attention.py:437-461 — so a step is two H2D copies: that buffer and the lengths)
    G = c["G"]
    qkv_h = torch.cat([q.cpu(), kc[:, S - 1].cpu(), vc[:, S - 1].cpu()], dim=1).pin_memory()
    lens_h = lens.cpu().pin_memory()
    oh = torch.empty((B, H, hd), dtype=torch.bfloat16).pin_memory()

    pos = torch.arange(S, dtype=torch.float32, device=dev)
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.float32, device=dev) / hd))
    cos_t, sin_t = torch.cos(torch.outer(pos, inv_freq)), torch.sin(torch.outer(pos, inv_freq))

    def e2e_step():
        qkv = qkv_h.to(dev, non_blocking=True)
        qd, kn, vn = qkv[:, :H], qkv[:, H:H + G], qkv[:, H + G:]
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
                qkv = qkv_h.to(dev, non_blocking=True)
                qd, kn, vn = qkv[:, :H], qkv[:, H:H + G], qkv[:, H + G:]
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
    h2d = qkv_h.numel() * 2 + lens_h.numel() * 4
    d2h = oh.numel() * 2

    # DRAM traffic of the dominant kernel cannot be measured without a profiler: it comes from the committed ncu
    # `--set full` capture of this very workload (dram__bytes_read.sum + dram__bytes_write.sum of one launch), and the
    # line names the capture it was read from; null if the file is missing
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "decode_traffic.json")
    if os.path.exists(prof):
        try:
            tj = json.load(open(prof))
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
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
                      frac=achieved / peaks["hbm_gbs"], traffic=traffic, traffic_source=traffic_src,
                      peak_source=f"MEASURED_PEAKS.json hbm_gbs ({peaks['source']})",
                      frac_of_8tbs_spec=achieved / 8000.0,
                      kernel="decode_mma_kernel<128,4,32> (one launch per step: TMA ring + mma.sync consumers + flush warp)",
                      algorithmic_bytes_per_launch=nbytes,
                      launch_ms_mean=dev_ms, launch_ms_min=min(per_launch_ms)),
    )


def _pad_heads(x):
    """The layout the drop-in modules produce for head dims TMA cannot address: head stride rounded up to 8."""
    from vats_multimodal_lm_b200.modules._common import _to_kernel_layout
    return _to_kernel_layout(x)


def _gen_prefill_inputs(c, nb, ng, rank_seed, dev, layout):
    hpg = c["H"] // c["G"]
    q = gen_unit_bf16((nb, c["T"], ng * hpg, c["hd"]), 1234 + rank_seed, dev, True)
    k = gen_unit_bf16((nb, c["T"], ng, c["hd"]), 2234 + rank_seed, dev, True)
    v = gen_unit_bf16((nb, c["T"], ng, c["hd"]), 3234 + rank_seed, dev, False)
    if layout == "module" and c["hd"] % 8 != 0:
        q, k, v = _pad_heads(q), _pad_heads(k), _pad_heads(v)
    return q, k, v


def sampled_error(c, q, k, v, o, n_rows=48):
    """max-abs and relative-L2 error of a sample of output rows (first and last sequence, rows spread over the
    sequence incl. both ends) against the fp32 oracle on the same bf16 inputs."""
    from oracle import sdpa_rows
    T = c["T"]
    rows = torch.unique(torch.cat([torch.linspace(0, T - 1, min(n_rows, T)).long(), torch.tensor([0, T - 1])]))
    worst_abs, num, den = 0.0, 0.0, 0.0
    for n in sorted({0, q.size(0) - 1}):
        qc, kc, vc = q[n:n + 1].float().cpu(), k[n:n + 1].float().cpu(), v[n:n + 1].float().cpu()
        ref = sdpa_rows(qc, kc, vc, rows, c["hd"] ** -0.5, c["causal"], c["left"], 0 if c["causal"] else -1)
        got = o[n:n + 1, rows.to(o.device)].float().cpu()
        worst_abs = max(worst_abs, (got - ref).abs().max().item())
        num += (got - ref).pow(2).sum().item()
        den += ref.pow(2).sum().item()
    return dict(max_abs=worst_abs, rel_l2=math.sqrt(num / max(den, 1e-30)), rows_checked=int(rows.numel()) * 2,
                tolerance="max_abs <= 2e-2, rel_l2 <= 1e-2 (bf16 operands/outputs, fp32 accumulation, vs fp32 oracle)")


def cpu_baseline_prefill(c):
    """The reference's CPU path for one prefill-class workload: core-only fp32 SDPA with the explicit mask and expanded
    K/V heads (oracle/cpu_baseline.py), all host threads, on a bounded sample of the workload."""
    from oracle.cpu_baseline import reference_core_cpu, time_callable
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    N, T, H, G, hd = c["N"], c["T"], c["H"], c["G"], c["hd"]
    g = torch.Generator().manual_seed(77)
    scale = hd ** -0.5
    if c["name"].startswith("cfg5"):
        # one sequence's LAST 2048 queries against the 6144 keys they can see (same band geometry); scaled by allowed pairs
        Tq, Tk = 2048, 2048 + c["left"]
        q = torch.nn.functional.normalize(torch.randn(1, Tq, H, hd, generator=g), dim=-1)
        k = torch.nn.functional.normalize(torch.randn(1, Tk, G, hd, generator=g), dim=-1)
        v = torch.randn(1, Tk, G, hd, generator=g)
        calls, dt = time_callable(lambda: reference_core_cpu(q, k, v, scale, True, c["left"], 0), 4.0, 20)
        pairs = Tq * (c["left"] + 1)
        fl = 4 * H * hd * pairs
        tfl = fl * calls / dt / 1e12
        return dict(value=tfl, unit="TFLOP/s", cores=cores, kind="port",
                    ms_extrapolated_full_workload=prefill_flops(c) / (tfl * 1e12) * 1e3,
                    sample=f"EXTRAPOLATED: B=1, last {Tq} queries x {Tk} keys of the band (of 8 x 32768), fp32, "
                           f"{calls} calls in {dt:.1f} s, scaled by allowed (query, key) pairs")
    Ns = N
    while Ns > 1 and 4 * Ns * T * T * H * hd > 6e10:   # keep one call under ~0.3 s
        Ns //= 2
    q = torch.nn.functional.normalize(torch.randn(Ns, T, H, hd, generator=g), dim=-1)
    k = torch.nn.functional.normalize(torch.randn(Ns, T, G, hd, generator=g), dim=-1)
    v = torch.randn(Ns, T, G, hd, generator=g)
    calls, dt = time_callable(lambda: reference_core_cpu(q, k, v, scale, c["causal"], c["left"], 0 if c["causal"] else -1),
                              2.0, 200)
    ms = dt / calls * 1e3 * (N / Ns)
    fl = prefill_flops(c)
    return dict(value=fl / (ms * 1e-3) / 1e12, unit="TFLOP/s", cores=cores, kind="port", ms_full_workload=ms,
                sample=f"{Ns} of {N} sequences per call, {calls} calls in {dt:.1f} s, fp32 (the reference's dtype), torch "
                       f"CPU SDPA with the explicit mask and K/V expanded to H heads (oracle/cpu_baseline.py)")


def bench_prefill_cfg(c, peaks, steps, warmup, world, rank, shard: bool, layout: str = "dense", cpu: bool = False,
                      e2e: bool = False):
    """One prefill-class workload: >= 20 timed iterations with per-launch CUDA events, clocks sampled during them,
    sampled-row error against the oracle.  With shard=True the N sequences (x KV groups) are split over the ranks and
    the outputs gathered (cfg5 / ViT); timings are the max over ranks."""
    from vats_multimodal_lm_b200 import _ffi, ops, sharding
    dev = torch.device("cuda", torch.cuda.current_device())
    N, T, H, G, hd = c["N"], c["T"], c["H"], c["G"], c["hd"]
    sh = sharding.partition(N, G, world, rank) if shard else sharding.Shard(0, N, 0, G)
    nb, ng = sh.b1 - sh.b0, sh.g1 - sh.g0
    hpg = H // G
    q, k, v = _gen_prefill_inputs(c, nb, ng, rank, dev, layout)
    scale = hd ** -0.5
    right = 0 if c["causal"] else -1
    step = lambda: ops.gqa_swa_prefill(q, k, v, None, None, scale, c["causal"], c["left"], right, 0)
    flush = None
    if prefill_bytes(c) / max(world if shard else 1, 1) < 2.5e8:   # working set could sit in the 126 MB L2
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        o = step()
    kernel = _ffi.last_kernel()
    launches = _ffi.last_launch_count()
    evs = []
    sampler = ClockSampler(torch.cuda.current_device())
    barrier_sync(world)
    sampler.start()
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        o = step()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    per = [a.elapsed_time(b) for a, b in evs]
    # tiny workloads finish before nvidia-smi takes a sample: keep the GPU on the workload for ~0.6 s more
    if sum(per) < 600.0:
        t_end = time.perf_counter() + 0.6
        while time.perf_counter() < t_end:
            for _ in range(20):
                step()
            torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = max_over_ranks(statistics.mean(per), world)
    res = dict(ms_compute=ms, ms_min=min(per), iterations=steps, kernel=kernel, launches_per_call=launches,
               clocks=clocks, l2="flushed between iterations" if flush is not None else "inputs larger than L2")
    try:
        res["error_vs_oracle"] = sampled_error(c, q, k, v, o)
    except Exception as e:
        res["error_vs_oracle"] = {"error": f"{type(e).__name__}: {e}"}
    # The same workload the way the drop-in modules call it behind qk-norm (unit-norm q, k: logit_bound = 1.0, the
    # kernels skip the row maximum).  Reported beside the exact-maximum time above, never instead of it.
    bstep = lambda: ops.gqa_swa_prefill(q, k, v, None, None, scale, c["causal"], c["left"], right, 0, 1.0)
    for _ in range(warmup):
        ob = bstep()
    bevs = []
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ob = bstep()
        b.record()
        bevs.append((a, b))
    torch.cuda.synchronize()
    bper = [a.elapsed_time(b) for a, b in bevs]
    res["qk_norm_logit_bound"] = dict(ms_compute=max_over_ranks(statistics.mean(bper), world), ms_min=min(bper),
                                      logit_bound=1.0, max_abs_vs_exact_path=(ob.float() - o.float()).abs().max().item(),
                                      note="modules/_common.py QK_NORM_LOGIT_BOUND: q, k are l2-normalised by qk-norm")
    try:
        res["qk_norm_logit_bound"]["error_vs_oracle"] = sampled_error(c, q, k, v, ob)
    except Exception as e:
        res["qk_norm_logit_bound"]["error_vs_oracle"] = {"error": f"{type(e).__name__}: {e}"}
    del ob

    if e2e and world == 1:
        # end to end through the public op with HOST buffers: pinned q/k/v -> H2D, kernel, D2H of o, every step
        qh, kh, vh = (t.contiguous().cpu().pin_memory() for t in (q, k, v))
        oh = torch.empty((nb, T, ng * hpg, hd), dtype=torch.bfloat16).pin_memory()

        def e2e_step():
            qd, kd, vd = qh.to(dev, non_blocking=True), kh.to(dev, non_blocking=True), vh.to(dev, non_blocking=True)
            oh.copy_(ops.gqa_swa_prefill(qd, kd, vd, None, None, scale, c["causal"], c["left"], right, 0), non_blocking=True)
            torch.cuda.synchronize()
        for _ in range(2):
            e2e_step()
        t0 = time.perf_counter()
        n_e2e = 10
        for _ in range(n_e2e):
            e2e_step()
        dt = (time.perf_counter() - t0) / n_e2e
        h2d = (qh.numel() + kh.numel() + vh.numel()) * 2
        d2h = oh.numel() * 2
        res["e2e"] = dict(ms=dt * 1e3, gbs=prefill_bytes(c) / dt / 1e9, tflops=prefill_flops(c) / dt / 1e12,
                          h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                          note="pinned host q/k/v -> H2D -> vats::gqa_swa_prefill -> D2H of o; PCIe-bound "
                               f"({(h2d + d2h) / dt / 1e9:.1f} GB/s over the link)")
        del qh, kh, vh, oh

    if shard and world > 1:
        res.update(_bench_gathers(c, q, k, v, step, scale, right, world, rank, steps, dev))

    units = world if (not shard) else 1  # unsharded workloads are replicated per rank (weak)
    fl = prefill_flops(c) * units
    by = prefill_bytes(c) * units
    res.update(tflops=fl / (ms * 1e-3) / 1e12, gbs=by / (ms * 1e-3) / 1e9, flops=fl, bytes=by,
               tokens_per_s=c["N"] * c["T"] * units / (ms * 1e-3))
    per_gpu = max(world, 1) if shard else units
    if c["bound"] == "hbm":
        res["roofline"] = dict(bound="hbm", achieved=res["gbs"] / per_gpu, peak=peaks["hbm_gbs"], unit="GB/s")
    else:
        res["roofline"] = dict(bound="tensor", achieved=res["tflops"] / per_gpu, peak=peaks["bf16_tflops"], unit="TFLOP/s")
        if c["bound"] == "latency":
            res["roofline"]["note_bound"] = ("batch-1 prompt: a few work items on 148 SMs — launch / latency bound; the "
                                             "tensor-peak fraction is reported for completeness only")
    res["roofline"]["frac"] = res["roofline"]["achieved"] / res["roofline"]["peak"]
    res["roofline"]["note"] = "per-GPU achieved vs measured per-GPU peak (burst bf16 peak: the kernel is timed alone)"
    if cpu and rank == 0 and world == 1:
        try:
            res["cpu_baseline"] = cpu_baseline_prefill(c)
        except Exception as e:
            res["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    del q, k, v, o
    torch.cuda.empty_cache()
    return res


def _bench_gathers(c, q, k, v, step, scale, right, world, rank, steps, dev):
    """compute + gather of the sharded outputs, four ways; each variant's gathered tensor is verified: every rank
    regenerates the inputs of ANOTHER rank's shard, recomputes that slice locally and compares it with what arrived."""
    from vats_multimodal_lm_b200 import ops, sharding
    N, T, H, G, hd = c["N"], c["T"], c["H"], c["G"], c["hd"]
    hpg = H // G
    res = {}
    other = (rank + 1) % world
    so = sharding.partition(N, G, world, other)
    qo, ko, vo = _gen_prefill_inputs(c, so.b1 - so.b0, so.g1 - so.g0, other, dev, "dense")
    expect = ops.gqa_swa_prefill(qo, ko, vo, None, None, scale, c["causal"], c["left"], right, 0)
    del qo, ko, vo
    worst = [0.0]

    def check(full, name):
        got = full[so.b0:so.b1, :, so.g0 * hpg:so.g1 * hpg]
        d = (got.float() - expect.float()).abs().max().item()
        res[f"parity_{name}_max_abs"] = max_over_ranks(d, world)
        worst[0] = max(worst[0], res[f"parity_{name}_max_abs"])

    def timed(fn, name):
        for _ in range(2):
            full = fn()
        torch.cuda.synchronize()
        check(full, name)
        barrier_sync(world)
        t0 = time.perf_counter()
        for _ in range(steps):
            full = fn()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        del full
        return max_over_ranks((t1 - t0) / steps * 1e3, world)

    core = lambda q_, k_, v_, qv_, kv_, causal=False: ops.gqa_swa_prefill(
        q_, k_, v_, qv_, kv_, scale, causal, c["left"], 0 if causal else -1, 0)
    res["ms_compute_plus_allgather"] = timed(lambda: sharding.gather_outputs(step(), N, H, G), "allgather")
    res["ms_compute_plus_allgather_overlapped"] = timed(
        lambda: sharding.local_attention_gather(core, q, k, v, N, H, G, chunks=4, causal=c["causal"]), "chunked")
    pg, err = None, None
    try:
        pg = sharding.PeerGather(N, T, H, hd, torch.bfloat16, dev)
    except Exception as e:
        err = f"{type(e).__name__}: {e}"
    if max_over_ranks(0.0 if pg is not None else 1.0, world) == 0.0:   # every rank takes the same branch
        res["ms_compute_plus_peer_gather_overlapped"] = timed(
            lambda: sharding.local_attention_gather(core, q, k, v, N, H, G, chunks=4, causal=c["causal"], peer=pg), "peer")
    else:
        res["peer_gather_error"] = err or "symmetric memory unavailable on another rank"
    del pg
    fg, err = None, None
    if hd % 8 == 0:
        try:
            fg = sharding.FusedGather(N, T, H, hd, G, dev)
        except Exception as e:
            err = f"{type(e).__name__}: {e}"
        if max_over_ranks(0.0 if fg is not None else 1.0, world) == 0.0:
            ms = timed(lambda: fg.run(q, k, v, scale, c["causal"], c["left"], right), "fused")
            res["ms_compute_plus_fused_gather"] = ms
            out_bytes = 2 * N * T * H * hd
            res["fused_gather_nvlink_ingress_gbs_per_gpu"] = out_bytes * (world - 1) / world / (ms * 1e-3) / 1e9
        else:
            res["fused_gather_error"] = err or "symmetric memory unavailable on another rank"
        del fg
    else:
        res["fused_gather_error"] = "head_dim not a multiple of 8: output not TMA-addressable"
    res["sharded_parity_max_abs"] = worst[0]
    return res


def bench_decode_secondary(c, peaks, steps, world, rank):
    """Decode of the default LLM geometry (hd 60): cache with the head stride the drop-in KVCache allocates (64)."""
    from vats_multimodal_lm_b200 import _ffi, ops
    from oracle import decode_explicit
    dev = torch.device("cuda", torch.cuda.current_device())
    B, S, H, G, hd, left = c["B"], c["S"], c["H"], c["G"], c["hd"], c["left"]
    hs = (hd + 7) // 8 * 8
    kc = gen_unit_bf16((B, S, G, hs), 4234 + rank, dev, True)[..., :hd]
    vc = gen_unit_bf16((B, S, G, hs), 5234 + rank, dev, False)[..., :hd]
    q = gen_unit_bf16((B, H, hd), 6234 + rank, dev, True)
    lens = torch.full((B,), S, dtype=torch.int32, device=dev)
    scale = hd ** -0.5
    step = lambda: ops.gqa_swa_decode(q, kc, vc, lens, scale, left)
    for _ in range(5):
        o = step()
    kernel = _ffi.last_kernel()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sampler = ClockSampler(torch.cuda.current_device())
    barrier_sync(world)
    sampler.start()
    for a, b in evs:
        a.record()
        o = step()
        b.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter() + 0.5
    while time.perf_counter() < t_end:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = max_over_ranks(statistics.mean(a.elapsed_time(b) for a, b in evs), world)
    nbytes = decode_bytes(c)
    sl = slice(0, 2)
    ref = decode_explicit(q[sl].float().cpu(), kc[sl].float().cpu(), vc[sl].float().cpu(), lens[sl].cpu(), scale, left)
    d = o[sl].float().cpu() - ref
    return dict(ms_compute=ms, iterations=steps, kernel=kernel, clocks=clocks, gbs=nbytes / (ms * 1e-3) / 1e9,
                bytes=nbytes, cache_head_stride=hs,
                error_vs_oracle=dict(max_abs=d.abs().max().item(), rel_l2=(d.norm() / ref.norm()).item(),
                                     sequences_checked=2),
                roofline=dict(bound="hbm", achieved=nbytes / (ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s",
                              frac=nbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                              note="algorithmic bytes use the UNPADDED head dim (60); the cache rows are 64 wide"),
                l2="inputs larger than L2")


TRAIN_CFG = dict(name="cfg1_llm_train_step_T2048", N=8, T=2048, H=24, G=8, hd=60, causal=True, left=384)


def bench_backward(c, peaks, steps, rank):
    """Forward + backward of the core through autograd (reference training loop
    training/transformers/nlp/loops/training_loop.py:54-65): vats::gqa_swa_prefill with requires_grad inputs in the
    modules' layout, dO ~ N(0,1).  Forward and backward timed separately with CUDA events; gradients of a few sampled
    rows are checked against torch.autograd on the fp32 oracle."""
    from vats_multimodal_lm_b200 import _ffi, ops
    from oracle import mask_predicate, sdpa_explicit
    dev = torch.device("cuda", torch.cuda.current_device())
    N, T, H, G, hd = c["N"], c["T"], c["H"], c["G"], c["hd"]
    q, k, v = _gen_prefill_inputs(c, N, G, rank, dev, "module")
    q, k, v = q.detach().requires_grad_(True), k.detach().requires_grad_(True), v.detach().requires_grad_(True)
    do = gen_unit_bf16((N, T, H, hd), 9234 + rank, dev, False)
    scale = hd ** -0.5
    fwd = lambda: ops.gqa_swa_prefill(q, k, v, None, None, scale, c["causal"], c["left"], 0, 0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3):
        o = fwd()
        o.backward(do)
        q.grad = k.grad = v.grad = None
    tf, tb = [], []
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    for _ in range(steps):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        o = fwd()
        e[1].record()
        o.backward(do)
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1]))
        tb.append(e[1].elapsed_time(e[2]))
        if _ < steps - 1:
            q.grad = k.grad = v.grad = None
    kernel = _ffi.last_kernel()
    clocks = sampler.stop()
    # gradient parity on one sequence, fp32 oracle + torch.autograd
    qc, kc, vc = (t[:1].detach().float().cpu().requires_grad_(True) for t in (q, k, v))
    ref = sdpa_explicit(qc, kc, vc, mask_predicate(1, T, T, c["causal"], c["left"], 0), scale)
    ref.backward(do[:1].float().cpu())
    errs = {}
    for name, g, r in (("dq", q.grad, qc.grad), ("dk", k.grad, kc.grad), ("dv", v.grad, vc.grad)):
        d = g[:1].float().cpu() - r
        errs[name] = dict(max_abs=d.abs().max().item(), rel_l2=(d.norm() / r.norm()).item())
    ms_f, ms_b = statistics.mean(tf), statistics.mean(tb)
    fl = prefill_flops(c)
    return dict(ms_forward=ms_f, ms_backward=ms_b, iterations=steps, kernel_last=kernel, clocks=clocks,
                tflops_forward=fl / (ms_f * 1e-3) / 1e12, tflops_backward=2.5 * fl / (ms_b * 1e-3) / 1e12,
                flops_forward=fl, flops_backward=2.5 * fl, tokens_per_s=N * T / ((ms_f + ms_b) * 1e-3),
                grad_error_vs_autograd_of_oracle=errs, sequences_checked=1,
                l2="flushed between iterations", layout="module layout (head stride padded to 8 elements)",
                note="backward = recompute of the row statistics + dQ kernel + dK/dV kernel (mma.sync, deterministic); "
                     "2.5x the forward's algorithmic FLOPs (5 matmuls vs 2), allowed pairs only; timings include the "
                     "autograd dispatch")


MODULE_LAYERS = [
    # name, builder, input shape, forward kwargs — the drop-in attention modules at the BASELINE geometries, bf16 weights
    dict(name="layer_cfg1_llm_attention_T4096", kind="llm", d_model=1440, H=24, G=8, shape=(1, 4096, 1440)),
    dict(name="layer_cfg3_vit2d_spatial_attention", kind="vit2d", d_model=1152, H=16, G=8, shape=(256, 196, 1152)),
    dict(name="layer_cfg4_vit3d_spatiotemporal_attention", kind="vit3d", d_model=2112, H=32, G=8, shape=(64, 8, 196, 2112)),
]


def bench_module_layer(c, steps, rank):
    """One forward of a drop-in attention module (projection GEMMs in torch + producers + core + output projection),
    inference mode, twice: with the fused producer kernels (qk-norm + RoPE + bf16 cast + kernel layout in one launch,
    SURVEY §8f ranks 1-2) and with the PyTorch element-wise passes the modules use under autograd.  Same weights, same
    input; the outputs of the two paths are compared."""
    import vats_multimodal_lm_b200 as vl
    from vats_multimodal_lm_b200 import _ffi
    from vats_multimodal_lm_b200.modules import llm as L, vit2d as V2, vit3d as V3
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(4321 + rank)
    hd = c["d_model"] // c["H"]
    if c["kind"] == "llm":
        m = vl.Attention(c["d_model"], c["H"], c["G"], 10000.0, hd ** -0.5)
        call = lambda x: m(x, 384, 0, True, None, None, None, False, False, True)[0]
    elif c["kind"] == "vit2d":
        m = vl.SpatialAttention(c["d_model"], c["H"], c["G"], 10000.0, 224, 16, hd ** -0.5, False, False, True)
        call = lambda x: m(x, False, True, -1, -1)
    else:
        m = vl.SpatioTemporalAttention(c["d_model"], c["H"], c["G"], 10000.0, (2, 16, 16))
        call = lambda x: m(x, (c["shape"][1], 14, 14), False, True, None, None)
    m = m.to(dev).bfloat16().eval()
    x = (torch.randn(c["shape"], device=dev) * 0.5).bfloat16()
    mods = (L, V2, V3)
    saved = [mod._on_gpu for mod in mods]

    def timed(fused):
        for mod, orig in zip(mods, saved):
            mod._on_gpu = orig if fused else (lambda t: False)
        try:
            with torch.no_grad():
                for _ in range(3):
                    y = call(x)
                torch.cuda.synchronize()
                kernel = _ffi.last_kernel()
                evs = []
                for _ in range(steps):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    y = call(x)
                    b.record()
                    evs.append((a, b))
                torch.cuda.synchronize()
            return statistics.mean(a.elapsed_time(b) for a, b in evs), y, kernel
        finally:
            for mod, orig in zip(mods, saved):
                mod._on_gpu = orig

    ms_fused, y_fused, kernel = timed(True)
    ms_plain, y_plain, _ = timed(False)
    d = (y_fused.float() - y_plain.float())
    return dict(ms_layer_fused_producers=ms_fused, ms_layer_pytorch_producers=ms_plain, iterations=steps,
                core_kernel=kernel, tokens=int(x.numel() // c["d_model"]),
                tokens_per_s=int(x.numel() // c["d_model"]) / (ms_fused * 1e-3),
                fused_vs_pytorch_path=dict(max_abs=d.abs().max().item(),
                                           rel_l2=(d.norm() / y_plain.float().norm()).item()),
                note="whole module forward incl. the torch projection GEMMs (bf16 weights); the two timings differ only "
                     "in how q / k / v get from the projection to the core: one producer launch vs ~12 element-wise "
                     "passes + cast / layout copies")


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
        steps = max(20, min(args.steps, 30))
        sharded_names = ("cfg5_long_prefill", "cfg3_vit2d", "cfg4a_vit3d_spatial", "cfg4b_vit3d_temporal")
        for c in PREFILL_CFGS:
            shard = world > 1 and c["name"] in sharded_names
            try:
                other[c["name"]] = bench_prefill_cfg(c, peaks, steps, 3, world, rank, shard, cpu=True,
                                                     e2e=(c["name"] == "cfg3_vit2d"))
                other[c["name"]]["sharded"] = shard
                other[c["name"]]["layout"] = "dense [N,T,heads,hd]"
                if c["hd"] % 8 != 0 and c["T"] > 32 and world == 1:   # (the modules do not pad sequences of <= 32 keys)
                    r2 = bench_prefill_cfg(c, peaks, steps, 3, world, rank, False, layout="module")
                    r2["sharded"] = False
                    r2["layout"] = "as produced by the drop-in modules: head stride padded to 8 elements (TMA-addressable)"
                    other[c["name"] + "_module_layout"] = r2
            except Exception as e:  # a secondary workload must not take the headline down
                other[c["name"]] = {"error": f"{type(e).__name__}: {e}"}
        try:
            other[CFG2_MEDIUM["name"]] = bench_decode_secondary(CFG2_MEDIUM, peaks, max(20, min(args.steps, 100)), world, rank)
        except Exception as e:
            other[CFG2_MEDIUM["name"]] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1:
            try:
                other[TRAIN_CFG["name"]] = bench_backward(TRAIN_CFG, peaks, 20, rank)
            except Exception as e:
                other[TRAIN_CFG["name"]] = {"error": f"{type(e).__name__}: {e}"}
            # whole-module forwards run in a child process: whatever happens there cannot take the headline line down
            torch.cuda.empty_cache()
            try:
                cp = subprocess.run([sys.executable, os.path.abspath(__file__), "--layers-only"], capture_output=True,
                                    text=True, timeout=420)
                layers = json.loads(cp.stdout.strip().splitlines()[-1])
            except Exception as e:
                layers = {lc["name"]: {"error": f"{type(e).__name__}: {e}"} for lc in MODULE_LAYERS}
            other.update(layers)
    cpu = None
    if rank == 0 and world == 1:
        try:
            cpu = cpu_baseline_decode()
        except Exception as e:
            cpu = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        parity = [w.get("sharded_parity_max_abs") for w in other.values() if isinstance(w, dict) and
                  w.get("sharded_parity_max_abs") is not None]
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
        if parity:
            line["sharded_parity_max_abs"] = max(parity)
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
    ap.add_argument("--layers-only", action="store_true", help="(internal) time the whole-module forwards, print a JSON dict")
    ap.add_argument("--cache-layout", default="bshd", choices=["bshd", "bhsd"],
                    help="storage order of the KV cache behind its [B,S,G,hd] shape")
    args = ap.parse_args()
    # Exactly one line on stdout: libraries write banners there (NCCL prints its version at the first communicator),
    # so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.layers_only:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        res = {}
        for lc in MODULE_LAYERS:
            try:
                res[lc["name"]] = bench_module_layer(lc, 10, 0)
            except Exception as e:
                res[lc["name"]] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
        emit(res)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

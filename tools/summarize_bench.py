import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(f"decode: value={d['value']:.0f} GB/s achieved={r['achieved']:.0f} ({100*r['frac']:.1f}% of measured) launch_ms mean={r['launch_ms_mean']:.4f} min={r['launch_ms_min']:.4f} e2e={d['e2e']['value']:.0f} launches={d['gpu_launches']}")
for k, v in d.get("other_workloads", {}).items():
    if "error" in v:
        print(k, "ERROR", v["error"]); continue
    if "ms_layer_fused_producers" in v:
        print(f"{k:26s} fused {v['ms_layer_fused_producers']:.4f} ms  pytorch producers {v['ms_layer_pytorch_producers']:.4f} ms  diff {v['fused_vs_pytorch_path']}")
        continue
    if "ms_forward" in v:
        print(f"{k:26s} fwd {v['ms_forward']:.4f} ms ({v['tflops_forward']:.1f} TFLOP/s)  bwd {v['ms_backward']:.4f} ms ({v['tflops_backward']:.1f} TFLOP/s)")
        continue
    rf = v["roofline"]
    extra = f" +allgather {v['ms_compute_plus_allgather']:.3f} ms" if "ms_compute_plus_allgather" in v else ""
    if "ms_compute_plus_allgather_overlapped" in v:
        extra += f" (overlapped {v['ms_compute_plus_allgather_overlapped']:.3f} ms)"
    if "ms_compute_plus_peer_gather_overlapped" in v:
        extra += f" (peer writes {v['ms_compute_plus_peer_gather_overlapped']:.3f} ms)"
    if "peer_gather_error" in v:
        extra += " peer_gather_error=" + v["peer_gather_error"][:80]
    print(f"{k:26s} {v['ms_compute']:.4f} ms  {v.get('tflops', 0.0):8.1f} TFLOP/s {v['gbs']:8.1f} GB/s  {rf['bound']} frac={100*rf['frac']:.1f}%{extra}")
if d.get("cpu_baseline"): print("cpu_baseline", d["cpu_baseline"])
print("clocks", d.get("clocks"))

"""GPU: the LLM drop-in under the reference generator's exact call sequence
(src/transformers/nlp/inference/generate.py:89-127, model built as src/transformers/nlp/model.py:148-154 builds it):

    1. prefill   model(ids [B,T], padding_mask [B,T], use_cache=True)   -> cache filled
    2. step 0    model(ids [B,T], padding_mask,       use_cache=False)  -> uncached re-forward, cache untouched
    3. step s>0  model(last [B,1], padding_mask [B,1] = unfinished, use_cache=True)

with a `KVCache(num_heads = H)` exactly as the unmodified call sites construct it.  Every cached single-token step must
end in the TMA decode kernel (checked through vats_attn_last_kernel), also for head_dim 60 (cache head stride 64), and
its output must match an fp32 oracle that recomputes the step from the full token history (projections of the module's
own weights, qk-norm, RoPE at the true positions, explicit mask).  tests/test_reference_callsites_cpu.py runs the same
sequence through the unmodified reference generator on CPU stand-ins.
"""
import pytest
import torch

from gpu_util import check_close
import vats_multimodal_lm_b200 as vl
from vats_multimodal_lm_b200 import _ffi, ops
from oracle import decode_explicit, mask_predicate, prefill_prepare_explicit, rope_tables, sdpa_explicit

pytestmark = pytest.mark.gpu
MOD_MAX_ABS, MOD_REL_L2 = 3e-2, 1.5e-2


def _close(out, ref, what):
    out, ref = out.float().cpu(), ref.float()
    assert torch.isfinite(out).all(), what
    max_abs = (out - ref).abs().max().item()
    rel = (out - ref).norm().item() / max(ref.norm().item(), 1e-12)
    assert max_abs <= MOD_MAX_ABS and rel <= MOD_REL_L2, f"{what}: max_abs={max_abs:.3e} rel_l2={rel:.3e}"


def _oracle_layer(attn, x_hist, q_rows, causal, left, q_valid, theta):
    """fp32 restatement of one Attention layer over the whole history x_hist [B,P,d]; returns rows `q_rows` of the output.
    Projections -> qk-norm -> RoPE at positions 0..P-1 -> masked softmax(QK^T)V (bottom-right aligned) -> w_o."""
    sd = {k: v.detach().float().cpu() for k, v in attn.state_dict().items()}
    B, P, _ = x_hist.shape
    H, G, hd = attn.num_heads, attn.query_groups, attn.head_dim
    qkv = x_hist.float() @ sd["w_qkv.weight"].T
    q, k, v = torch.split(qkv, [H * hd, G * hd, G * hd], dim=-1)
    cos, sin = rope_tables(hd, theta, P)
    q, k, v = prefill_prepare_explicit(q.view(B, P, H, hd), k.view(B, P, G, hd), v.view(B, P, G, hd), cos, sin, 0, True)
    # the kernels see bf16-rounded q, k, v
    q, k, v = (t.bfloat16().float() for t in (q, k, v))
    q = q[:, q_rows]
    m = mask_predicate(B, q.size(1), P, causal, left, 0 if causal else -1, q_valid, None)
    o = sdpa_explicit(q, k, v, m, attn.softmax_scale)
    return o.reshape(B, q.size(1), H * hd) @ sd["w_o.weight"].T


@pytest.mark.parametrize("d_model,H,G,left", [(240, 4, 2, 6), (512, 4, 2, -1), (256, 8, 2, 9)])   # hd 60, 128, 32
def test_generate_py_call_sequence_reaches_decode_kernel(d_model, H, G, left):
    torch.manual_seed(d_model)
    dev = "cuda"
    hd, theta, S_max = d_model // H, 10000.0, 64
    attn = vl.Attention(d_model, H, G, theta, hd ** -0.5).to(dev)
    cache = vl.KVCache(max_batch_size=8, max_seq_len=S_max, num_heads=H, head_dim=hd, num_layers=3)   # H heads!
    sibling = vl.KVCache(max_batch_size=8, max_seq_len=S_max, num_heads=H, head_dim=hd, num_layers=3)  # generate.py:27
    layer_idx = 1
    B, T, new = 3, 11, 9
    x0 = torch.randn(B, T, d_model)
    pm = torch.ones(B, T, dtype=torch.bool)
    pm[1, 7:] = False                                       # padded prompt
    sibling.reset()
    sibling.initialize(B)                                   # generate.py:93-94 (its own object; must stay harmless)

    # 1. prefill with cache + mask
    o, c = attn(x0.to(dev), left, 0, True, pm.to(dev), cache, layer_idx, True)
    assert cache.kv_heads == G and cache.layer_seq_len(layer_idx) == T and c["k"].shape == (B, T, G, hd)
    ref = _oracle_layer(attn, x0, slice(0, T), True, left, pm, theta)
    _close(o, ref, "cached prefill")
    # 2. uncached re-forward leaves the cache alone
    o2, c2 = attn(x0.to(dev), left, 0, True, pm.to(dev), cache, layer_idx, False)
    assert c2 is None and cache.layer_seq_len(layer_idx) == T
    _close(o2, ref, "uncached re-forward")
    # 3. cached single-token steps with the [B,1] mask of unfinished sequences
    hist = x0
    unfinished = torch.ones(B, dtype=torch.bool)
    for s in range(1, new):
        if s == 4:
            unfinished[0] = False                           # sequence 0 hit EOS
        x = torch.randn(B, 1, d_model)
        hist = torch.cat([hist, x], 1)
        last_attention = unfinished[:, None].clone()
        o, c = attn(x.to(dev), left, 0, True, last_attention.to(dev), cache, layer_idx, True)
        assert _ffi.last_kernel() == "decode_mma", f"step {s} ended in {_ffi.last_kernel()}"
        assert cache.layer_seq_len(layer_idx) == T + s and c["k"].shape == (B, 1, G, hd)
        P = hist.size(1)
        ref = _oracle_layer(attn, hist, slice(P - 1, P), True, left, last_attention, theta)
        _close(o, ref, f"cached step {s}")
        assert (o[~unfinished.to(dev)] == 0).all() or attn.w_o.bias is not None   # finished rows: zero attention output
    # the cache holds every token's k (also those of finished sequences), at the padded head stride
    k_all = cache.cache[layer_idx]["k"]
    assert k_all.stride(2) == (hd + 7) // 8 * 8 and k_all.shape == (B, S_max, G, hd)
    assert (k_all[:, :T + new - 1].float().abs().sum(-1) > 0).all()
    # generate.py:240 resets ITS cache: the model's sibling must start the next call empty
    sibling.reset()
    assert cache.current_seq_len is None
    o3, _ = attn(x0.to(dev), left, 0, True, pm.to(dev), cache, layer_idx, True)
    _close(o3, _oracle_layer(attn, x0, slice(0, T), True, left, pm, theta), "prefill after reset")


@pytest.mark.parametrize("hd,stride", [(60, 64), (48, 48), (72, 72), (66, 72), (96, 96), (120, 128), (8, 8), (128, 128)])
def test_decode_mma_serves_every_tma_addressable_head_dim(hd, stride):
    """decode_mma_kernel tiles of 16 / 32 / 64 / 128 columns with the real head dim below the tile width (TMA zero-fills
    the rest): hd 60 in a 64-element head stride is the default LLM's cache."""
    g = torch.Generator().manual_seed(hd)
    B, S, H, G, left = 5, 700, 12, 3, 300
    buf_k = torch.randn(B, S, G, stride, generator=g).bfloat16()
    buf_v = torch.randn(B, S, G, stride, generator=g).bfloat16()
    buf_v[:, 650:] = float("nan")                              # stale rows past every sequence end must not leak
    q = torch.nn.functional.normalize(torch.randn(B, H, hd, generator=g), dim=-1).bfloat16()
    lens = torch.tensor([650, 1, 333, 0, 64], dtype=torch.int32)
    dk, dv = buf_k.cuda()[..., :hd], buf_v.cuda()[..., :hd]
    o = ops.gqa_swa_decode(q.cuda(), dk, dv, lens.cuda(), hd ** -0.5, left)
    assert _ffi.last_kernel() == "decode_mma"
    ref = decode_explicit(q, torch.nan_to_num(buf_k[..., :hd]), torch.nan_to_num(buf_v[..., :hd]), lens, hd ** -0.5, left)
    check_close(o, ref, f"decode hd={hd}")
    assert (o[3] == 0).all()


def test_decode_dense_hd60_cache_falls_back_to_the_cuda_core_kernel():
    g = torch.Generator().manual_seed(3)
    B, S, H, G, hd = 2, 300, 6, 2, 60
    k = torch.randn(B, S, G, hd, generator=g).bfloat16()
    v = torch.randn(B, S, G, hd, generator=g).bfloat16()
    q = torch.nn.functional.normalize(torch.randn(B, H, hd, generator=g), dim=-1).bfloat16()
    lens = torch.tensor([300, 17], dtype=torch.int32)
    o = ops.gqa_swa_decode(q.cuda(), k.cuda(), v.cuda(), lens.cuda(), hd ** -0.5, -1)
    assert _ffi.last_kernel() == "decode_split"
    check_close(o, decode_explicit(q, k, v, lens, hd ** -0.5, -1), "dense hd60 decode")
    # alternating between the two kernels on one workspace key must not disturb the split counters
    k64 = torch.zeros(B, S, G, 64, dtype=torch.bfloat16)
    k64[..., :hd] = k
    v64 = torch.zeros(B, S, G, 64, dtype=torch.bfloat16)
    v64[..., :hd] = v
    for _ in range(3):
        o1 = ops.gqa_swa_decode(q.cuda(), k64.cuda()[..., :hd], v64.cuda()[..., :hd], lens.cuda(), hd ** -0.5, -1)
        assert _ffi.last_kernel() == "decode_mma"
        o2 = ops.gqa_swa_decode(q.cuda(), k.cuda(), v.cuda(), lens.cuda(), hd ** -0.5, -1)
        check_close(o1, decode_explicit(q, k, v, lens, hd ** -0.5, -1), "padded hd60 decode")
        check_close(o2, decode_explicit(q, k, v, lens, hd ** -0.5, -1), "dense hd60 decode")


def test_decode_inside_cuda_graph_uses_graph_owned_workspace():
    g = torch.Generator().manual_seed(5)
    B, S, H, G, hd = 4, 4096, 8, 2, 128
    k = torch.randn(B, S, G, hd, generator=g).bfloat16().cuda()
    v = torch.randn(B, S, G, hd, generator=g).bfloat16().cuda()
    q = torch.nn.functional.normalize(torch.randn(B, H, hd, generator=g), dim=-1).bfloat16().cuda()
    lens = torch.full((B,), S, dtype=torch.int32).cuda()
    eager = ops.gqa_swa_decode(q, k, v, lens, hd ** -0.5, 2048)
    n_cached = len(ops._DECODE_WS)
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.gqa_swa_decode(q, k, v, lens, hd ** -0.5, 2048)
        with torch.cuda.graph(graph, stream=s):
            out = ops.gqa_swa_decode(q, k, v, lens, hd ** -0.5, 2048)
    torch.cuda.current_stream().wait_stream(s)
    ops.reset_decode_workspaces()                      # nothing the graph uses lives in the cache
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
    assert n_cached >= 1


def test_bf16_cast_module_runs_fused_paths():
    """rope tables stay fp32 after model.to(bfloat16), so the fused producers accept them."""
    torch.manual_seed(0)
    attn = vl.Attention(256, 4, 2, 10000.0, 0.125).to("cuda").to(torch.bfloat16)
    cache = vl.KVCache(4, 32, 4, 64, 1)
    x = torch.randn(2, 5, 256, device="cuda", dtype=torch.bfloat16)
    o, _ = attn(x, -1, 0, True, None, cache, 0, True)
    o1, _ = attn(x[:, :1], -1, 0, True, None, cache, 0, True)
    assert o.dtype == torch.bfloat16 and torch.isfinite(o.float()).all() and torch.isfinite(o1.float()).all()
    assert _ffi.last_kernel() == "decode_mma"

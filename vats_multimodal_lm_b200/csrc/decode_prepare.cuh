// decode_prepare.cuh — the pre-core step of one cached decode token, fused: qk L2-norm + RoPE + bf16 rounding + KV-cache
// append (decode part).
//
// Reference, per new token at position p (src/optimized_attention.py:463-474):
//     q, k = F.normalize(q, eps=1e-6), F.normalize(k, eps=1e-6)      utils/attention_utils.py:80-102
//     q, k = rope(q), rope(k)                                         src/optimized_attention.py:97-143 (interleaved pairs)
//     cache.update(k, v)                                              src/optimized_attention.py:224-257 (intended contract)
// Four to six elementwise launches plus the cache write in PyTorch; here one launch: a warp per (sequence, head row),
// rows = H query heads + G key heads + G value heads.  fp32 arithmetic, one rounding to bf16 at the end.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace vats {

struct PrepareParams {
  const void* q_in;   // [B, H, hd]   (bf16 or fp32)
  const void* k_in;   // [B, G, hd]
  const void* v_in;   // [B, G, hd]
  int in_fp32;
  __nv_bfloat16* q_out;    // [B, H, hd]
  __nv_bfloat16* k_cache;  // [B, S_max, G, hd]
  __nv_bfloat16* v_cache;
  const int32_t* seq_lens; // [B] length including the new token; position = seq_lens[b] - 1
  const float* cos_table;  // [positions, hd/2] fp32 or NULL (no RoPE)
  const float* sin_table;
  int B, H, G, hd, S_max;
  long long qi_b, qi_h, ki_b, ki_h, vi_b, vi_h, qo_b, qo_h;
  long long ks_b, ks_t, ks_h, vs_b, vs_t, vs_h;
  int qk_norm;
  float eps;
};

__device__ __forceinline__ float prepare_load(const void* base, long long idx, int fp32) {
  return fp32 ? reinterpret_cast<const float*>(base)[idx] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}

constexpr int kPrepareWarps = 8;
constexpr int kPrepareMaxPairs = 4;   // per lane: hd <= 256

__global__ void __launch_bounds__(kPrepareWarps * 32) decode_prepare_kernel(const PrepareParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_seq = p.H + 2 * p.G;
  const long long row = (long long)blockIdx.x * kPrepareWarps + warp;
  if (row >= (long long)p.B * rows_per_seq) return;
  const int b = (int)(row / rows_per_seq);
  const int rr = (int)(row % rows_per_seq);
  const int L = p.seq_lens[b];
  if (L <= 0 || L > p.S_max) return;   // nothing to append (or no room: the host validates lengths it can see)
  const int pos = L - 1;

  // which tensor / head
  const void* src;
  long long src_off;
  __nv_bfloat16* dst;
  bool rotate;
  if (rr < p.H) {
    src = p.q_in; src_off = (long long)b * p.qi_b + (long long)rr * p.qi_h;
    dst = p.q_out + (long long)b * p.qo_b + (long long)rr * p.qo_h;
    rotate = true;
  } else if (rr < p.H + p.G) {
    const int g = rr - p.H;
    src = p.k_in; src_off = (long long)b * p.ki_b + (long long)g * p.ki_h;
    dst = p.k_cache + (long long)b * p.ks_b + (long long)pos * p.ks_t + (long long)g * p.ks_h;
    rotate = true;
  } else {
    const int g = rr - p.H - p.G;
    src = p.v_in; src_off = (long long)b * p.vi_b + (long long)g * p.vi_h;
    dst = p.v_cache + (long long)b * p.vs_b + (long long)pos * p.vs_t + (long long)g * p.vs_h;
    rotate = false;   // values are stored as they are
  }

  // lane owns element pairs (2i, 2i+1), i = lane, lane + 32, ...
  const int half = (p.hd + 1) >> 1;
  float x0[kPrepareMaxPairs], x1[kPrepareMaxPairs];
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < kPrepareMaxPairs; ++u) {
    const int i = lane + 32 * u;
    x0[u] = 0.f;
    x1[u] = 0.f;
    if (i < half) {
      x0[u] = prepare_load(src, src_off + 2 * i, p.in_fp32);
      if (2 * i + 1 < p.hd) x1[u] = prepare_load(src, src_off + 2 * i + 1, p.in_fp32);
      ss += x0[u] * x0[u] + x1[u] * x1[u];
    }
  }
  float scale = 1.f;
  if (rotate && p.qk_norm) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    scale = 1.f / fmaxf(sqrtf(ss), p.eps);   // F.normalize: x / max(||x||, eps)
  }
  const bool rope = rotate && p.cos_table != nullptr;
#pragma unroll
  for (int u = 0; u < kPrepareMaxPairs; ++u) {
    const int i = lane + 32 * u;
    if (i < half) {
      float a = x0[u] * scale, c = x1[u] * scale;
      if (rope) {
        const float cs = p.cos_table[(long long)pos * (p.hd >> 1) + i];
        const float sn = p.sin_table[(long long)pos * (p.hd >> 1) + i];
        const float ra = a * cs - c * sn;
        const float rc = a * sn + c * cs;
        a = ra;
        c = rc;
      }
      dst[2 * i] = __float2bfloat16(a);
      if (2 * i + 1 < p.hd) dst[2 * i + 1] = __float2bfloat16(c);
    }
  }
}

}  // namespace vats

// ---------------------------------------------------------------------------------------------------------------------
// The same producers for a whole prefill chunk (prefill half, 1-D RoPE): q, k [N, T, heads, hd] are
// L2-normalised and rotated at position pos0 + t, v is passed through; everything is rounded to bf16 once and written
// with a caller-chosen head stride (the TMA-addressable layout: head stride rounded up to 8 elements), replacing the
// normalise / rotate / cast / pad passes of the PyTorch path with one launch.
namespace vats {

struct PrefillPrepareParams {
  const void* q_in;   // [N, T, H, hd]  (bf16 or fp32)
  const void* k_in;   // [N, T, G, hd]
  const void* v_in;
  int in_fp32;
  __nv_bfloat16* q_out;   // [N, T, H, hd] with strides (qo_n, qo_t, qo_h)
  __nv_bfloat16* k_out;
  __nv_bfloat16* v_out;
  const float* cos_table;  // [>= pos0 + T, hd/2] or NULL
  const float* sin_table;
  int N, T, H, G, hd, pos0;
  long long qi_n, qi_t, qi_h, ki_n, ki_t, ki_h, vi_n, vi_t, vi_h;
  long long qo_n, qo_t, qo_h, ko_n, ko_t, ko_h, vo_n, vo_t, vo_h;
  int qk_norm;
  float eps;
};

__global__ void __launch_bounds__(kPrepareWarps * 32) prefill_prepare_kernel(const PrefillPrepareParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_tok = p.H + 2 * p.G;
  const long long total = (long long)p.N * p.T * rows_per_tok;
  const long long warps = (long long)gridDim.x * kPrepareWarps;
  const int half = (p.hd + 1) >> 1;
  for (long long row = (long long)blockIdx.x * kPrepareWarps + warp; row < total; row += warps) {
    const long long nt = row / rows_per_tok;
    const int rr = (int)(row - nt * rows_per_tok);
    const long long n = nt / p.T;
    const int t = (int)(nt - n * p.T);
    const void* src;
    long long src_off;
    __nv_bfloat16* dst;
    bool rotate;
    if (rr < p.H) {
      src = p.q_in; src_off = n * p.qi_n + (long long)t * p.qi_t + (long long)rr * p.qi_h;
      dst = p.q_out + n * p.qo_n + (long long)t * p.qo_t + (long long)rr * p.qo_h;
      rotate = true;
    } else if (rr < p.H + p.G) {
      const int g = rr - p.H;
      src = p.k_in; src_off = n * p.ki_n + (long long)t * p.ki_t + (long long)g * p.ki_h;
      dst = p.k_out + n * p.ko_n + (long long)t * p.ko_t + (long long)g * p.ko_h;
      rotate = true;
    } else {
      const int g = rr - p.H - p.G;
      src = p.v_in; src_off = n * p.vi_n + (long long)t * p.vi_t + (long long)g * p.vi_h;
      dst = p.v_out + n * p.vo_n + (long long)t * p.vo_t + (long long)g * p.vo_h;
      rotate = false;
    }
    float x0[kPrepareMaxPairs], x1[kPrepareMaxPairs];
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < kPrepareMaxPairs; ++u) {
      const int i = lane + 32 * u;
      x0[u] = 0.f;
      x1[u] = 0.f;
      if (i < half) {
        x0[u] = prepare_load(src, src_off + 2 * i, p.in_fp32);
        if (2 * i + 1 < p.hd) x1[u] = prepare_load(src, src_off + 2 * i + 1, p.in_fp32);
        ss += x0[u] * x0[u] + x1[u] * x1[u];
      }
    }
    float scale = 1.f;
    if (rotate && p.qk_norm) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      scale = 1.f / fmaxf(sqrtf(ss), p.eps);
    }
    const bool rope = rotate && p.cos_table != nullptr;
    const long long pos = (long long)p.pos0 + t;
#pragma unroll
    for (int u = 0; u < kPrepareMaxPairs; ++u) {
      const int i = lane + 32 * u;
      if (i < half) {
        float a = x0[u] * scale, c = x1[u] * scale;
        if (rope) {
          const float cs = p.cos_table[pos * (p.hd >> 1) + i];
          const float sn = p.sin_table[pos * (p.hd >> 1) + i];
          const float ra = a * cs - c * sn;
          const float rc = a * sn + c * cs;
          a = ra;
          c = rc;
        }
        dst[2 * i] = __float2bfloat16(a);
        if (2 * i + 1 < p.hd) dst[2 * i + 1] = __float2bfloat16(c);
      }
    }
  }
}

}  // namespace vats

// ---------------------------------------------------------------------------------------------------------------------
// The producers of the ViT passes (SURVEY §8f rank 1, 2-D axial and 3-D RoPE): every rotary variant of the reference
// (vit_2d/optimized_attention.py:128-172 — four blocks (x1, x2, y1, y2); vit_3d/rope_3d.py:97-219 — interleaved pairs
// inside the h / w blocks or the t block) is, per token and column,
//        out[c] = xn[c] * cos[tok][c] + xn[partner[c]] * sin[tok][c]
// with a fixed column permutation `partner` and signed per-token tables — so ONE kernel serves them all: L2-normalise
// q and k, rotate by table, pass v through, round to bf16 once and write the kernels' layout (head stride rounded up to
// 8).  Sequences are indexed as n = n_outer * Ni + n_inner with separate input strides, so the ViT-3D temporal pass reads
// its [B, T, S, heads, hd] projections in place and writes [B*S, T, heads, hd]: the reference's
// `x.transpose(1, 2).contiguous()` (vit_3d/optimized_attention.py:474-479) never happens (§8f rank 2).
namespace vats {

struct PrepareTableParams {
  const void* q_in;   // logical [No, Ni, T, H, hd], strides (q_no, q_ni, q_t, q_h), hd contiguous (bf16 or fp32)
  const void* k_in;   // [No, Ni, T, G, hd]
  const void* v_in;
  int in_fp32;
  __nv_bfloat16* q_out;   // dense sequences [No * Ni, T, heads, hd_stride]: strides (qo_n, qo_t, qo_h)
  __nv_bfloat16* k_out;
  __nv_bfloat16* v_out;
  const float* cos_table;   // [T, hd]
  const float* sin_table;   // [T, hd] (signed)
  const int* partner;       // [hd]
  int No, Ni, T, H, G, hd;
  long long q_no, q_ni, q_t, q_h, k_no, k_ni, k_t, k_h, v_no, v_ni, v_t, v_h;
  long long qo_n, qo_t, qo_h, ko_n, ko_t, ko_h, vo_n, vo_t, vo_h;
  int qk_norm;
  float eps;
};

constexpr int kPrepareTableMaxHd = 256;

__global__ void __launch_bounds__(kPrepareWarps * 32) prefill_prepare_table_kernel(const PrepareTableParams p) {
  __shared__ float rows[kPrepareWarps][kPrepareTableMaxHd];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_tok = p.H + 2 * p.G;
  const long long total = (long long)p.No * p.Ni * p.T * rows_per_tok;
  const long long warps = (long long)gridDim.x * kPrepareWarps;
  float* row = rows[warp];
  for (long long r = (long long)blockIdx.x * kPrepareWarps + warp; r < total; r += warps) {
    const long long nt = r / rows_per_tok;
    const int rr = (int)(r - nt * rows_per_tok);
    const long long n = nt / p.T;
    const int t = (int)(nt - n * p.T);
    const long long no = n / p.Ni;
    const int ni = (int)(n - no * p.Ni);
    const void* src;
    long long src_off;
    __nv_bfloat16* dst;
    bool rotate;
    if (rr < p.H) {
      src = p.q_in; src_off = no * p.q_no + (long long)ni * p.q_ni + (long long)t * p.q_t + (long long)rr * p.q_h;
      dst = p.q_out + n * p.qo_n + (long long)t * p.qo_t + (long long)rr * p.qo_h;
      rotate = true;
    } else if (rr < p.H + p.G) {
      const int g = rr - p.H;
      src = p.k_in; src_off = no * p.k_no + (long long)ni * p.k_ni + (long long)t * p.k_t + (long long)g * p.k_h;
      dst = p.k_out + n * p.ko_n + (long long)t * p.ko_t + (long long)g * p.ko_h;
      rotate = true;
    } else {
      const int g = rr - p.H - p.G;
      src = p.v_in; src_off = no * p.v_no + (long long)ni * p.v_ni + (long long)t * p.v_t + (long long)g * p.v_h;
      dst = p.v_out + n * p.vo_n + (long long)t * p.vo_t + (long long)g * p.vo_h;
      rotate = false;
    }
    float ss = 0.f;
    for (int c = lane; c < p.hd; c += 32) {
      const float x = prepare_load(src, src_off + c, p.in_fp32);
      row[c] = x;
      ss += x * x;
    }
    float scale = 1.f;
    if (rotate && p.qk_norm) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      scale = 1.f / fmaxf(sqrtf(ss), p.eps);
    }
    __syncwarp();
    const bool rope = rotate && p.cos_table != nullptr;
    for (int c = lane; c < p.hd; c += 32) {
      float x = row[c] * scale;
      if (rope) {
        const float y = row[p.partner[c]] * scale;
        x = x * p.cos_table[(long long)t * p.hd + c] + y * p.sin_table[(long long)t * p.hd + c];
      }
      dst[c] = __float2bfloat16(x);
    }
    __syncwarp();
  }
}

}  // namespace vats


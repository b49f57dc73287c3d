// decode.cuh — KV-cache single-query GQA + sliding-window attention (split-K), HBM-bound.
//
// Replaces the intended "1 query x cached K/V" step of the reference
// (src/optimized_attention.py:508-516 cache branch + 709-714 SDPA; cache layout KVCache, :169-287).
//
// Design:
//   * one CTA = (split s, sequence b x KV group g, batch of <= HPG query heads of that group); the H/G query heads of a
//     group share one pass over the group's K/V rows, so the cache is read exactly once;
//   * the window [lo, L) of a sequence is cut into chunks of CH keys; split s handles chunk s (split-K);
//   * K and V rows are streamed with 128-bit (or 64/32-bit for odd head dims) ld.global.nc.L1::no_allocate,
//     U rows in flight per lane; a key row is spread over LPK lanes, 32/LPK keys per warp step;
//   * two passes per chunk: (1) scores q.k for all keys of the chunk -> smem, chunk max / exp2 / sum with warp
//     shuffles; (2) acc += p * v.  Partials (m, l, acc[hd]) go to the fp32 workspace; a small combine kernel merges
//     the splits (skipped when there is a single split).
//   No tensor cores: 4 flop/byte, the FP32 pipe has ~2x headroom over the HBM stream.
#pragma once
#include "ptx.cuh"

namespace vats {

struct DecodeParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* o;
  const int32_t* seq_lens;
  int B, H, G, hd, S_max;
  int hpg;              // H / G
  int head_batches;     // ceil(hpg / HPG)
  long long qs_b, qs_h; // strides in elements
  long long ks_b, ks_t, ks_h;
  long long vs_b, vs_t, vs_h;
  long long os_b, os_h;
  float scale_log2;     // scale * log2(e)
  int left;
  int chunk;            // keys per split
  int num_splits;
  float* ws_acc;        // [B*H, num_splits, hd]
  float* ws_ml;         // [B*H, num_splits, 2]
};

constexpr int kDecodeThreads = 128;
constexpr int kDecodeWarps = kDecodeThreads / 32;
constexpr int kDecodeMaxChunk = 512;

template <int VEC>
struct VecLoad;
template <>
struct VecLoad<8> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, uint32_t (&w)[4]) {
    uint4 r = ptx::ldg_nc_v4(p);
    w[0] = r.x; w[1] = r.y; w[2] = r.z; w[3] = r.w;
  }
};
template <>
struct VecLoad<4> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, uint32_t (&w)[2]) {
    uint2 r = ptx::ldg_nc_v2(p);
    w[0] = r.x; w[1] = r.y;
  }
};
template <>
struct VecLoad<2> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, uint32_t (&w)[1]) { w[0] = ptx::ldg_nc_u32(p); }
};

// Reduce N (= W) values across the W lanes of a lane group so that lane l ends with the total of value l.
// Recursive halving: W-1 shuffles for W values (a butterfly all-reduce would need W*log2(W)).
template <int W>
__device__ __forceinline__ float reduce_scatter(float (&v)[W], int lane_in_group) {
#pragma unroll
  for (int s = W / 2; s >= 1; s >>= 1) {
    const bool upper = (lane_in_group & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <int W>
__device__ __forceinline__ float group_allreduce_sum(float x) {
#pragma unroll
  for (int s = W / 2; s >= 1; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
  return x;
}

// VEC  elements per lane load (8 = 128-bit); LPK lanes per key row; CPL chunks per lane; HPG heads per CTA (padded
// to HPGP = power of two for the smem score rows); U = key rows in flight per lane.
template <int VEC, int LPK, int CPL, int HPG, int U>
__global__ void __launch_bounds__(kDecodeThreads) decode_split_kernel(const DecodeParams p) {
  constexpr int KPW = 32 / LPK;        // keys per warp step
  constexpr int EPL = VEC * CPL;       // elements per lane
  constexpr int WPV = VEC / 2;         // 32-bit words per vector
  constexpr int HPGP = HPG <= 1 ? 1 : (HPG <= 2 ? 2 : (HPG <= 4 ? 4 : 8));
  constexpr int KPI = kDecodeWarps * U * KPW;  // keys per CTA iteration

  __shared__ __align__(16) float s_scores[kDecodeMaxChunk * HPGP];
  __shared__ float s_m[HPGP], s_l[HPGP];
  extern __shared__ __align__(16) float s_acc[];  // [kDecodeWarps][HPG][hd] cross-warp reduction

  const int split = blockIdx.x;
  const int b = blockIdx.y / p.G;
  const int g = blockIdx.y % p.G;
  const int hb = blockIdx.z;
  const int h0 = g * p.hpg + hb * HPG;                     // first query head of this CTA
  const int nh = min(HPG, p.hpg - hb * HPG);               // active heads
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPK;                              // key slot within the warp step
  const int ll = lane % LPK;                               // lane within the key row

  const int L = p.seq_lens[b];
  int lo = 0;
  if (p.left >= 0) lo = max(0, L - 1 - p.left);
  const int ks = lo + split * p.chunk;                     // first key of this split
  const int ke = min(L, ks + p.chunk);                     // one past the last key
  const int nkeys = max(0, ke - ks);

  // ---- q slice of this lane, fp32, pre-multiplied by scale*log2(e)
  float qf[HPG][EPL];
#pragma unroll
  for (int h = 0; h < HPG; ++h) {
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int e0 = (ll + c * LPK) * VEC;
      uint32_t w[WPV];
#pragma unroll
      for (int i = 0; i < WPV; ++i) w[i] = 0u;
      if (h < nh && e0 < p.hd) VecLoad<VEC>::ld(p.q + b * p.qs_b + (long long)(h0 + h) * p.qs_h + e0, w);
#pragma unroll
      for (int i = 0; i < WPV; ++i) {
        qf[h][c * VEC + 2 * i] = ptx::bf16lo(w[i]) * p.scale_log2;
        qf[h][c * VEC + 2 * i + 1] = ptx::bf16hi(w[i]) * p.scale_log2;
      }
    }
  }

  const __nv_bfloat16* kbase = p.k + b * p.ks_b + (long long)g * p.ks_h;
  const __nv_bfloat16* vbase = p.v + b * p.vs_b + (long long)g * p.vs_h;

  // ---- pass 1: scores for every key of the chunk
  for (int it = 0; it * KPI < nkeys; ++it) {
    uint32_t kw[U][CPL][WPV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kl = it * KPI + (warp * U + u) * KPW + sub;  // key index local to the chunk
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int e0 = (ll + c * LPK) * VEC;
#pragma unroll
        for (int i = 0; i < WPV; ++i) kw[u][c][i] = 0u;
        if (kl < nkeys && e0 < p.hd) VecLoad<VEC>::ld(kbase + (long long)(ks + kl) * p.ks_t + e0, kw[u][c]);
      }
    }
    float part[U][HPGP];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float kf[EPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c)
#pragma unroll
        for (int i = 0; i < WPV; ++i) {
          kf[c * VEC + 2 * i] = ptx::bf16lo(kw[u][c][i]);
          kf[c * VEC + 2 * i + 1] = ptx::bf16hi(kw[u][c][i]);
        }
#pragma unroll
      for (int h = 0; h < HPGP; ++h) {
        float a = 0.f;
        if (h < HPG) {
#pragma unroll
          for (int e = 0; e < EPL; ++e) a = fmaf(qf[h][e], kf[e], a);
        }
        part[u][h] = a;
      }
    }
    if constexpr ((U * HPGP) % LPK == 0) {
      // transposed reduction: groups of LPK values; lane ll ends up owning value ll of each group
      constexpr int NG = (U * HPGP) / LPK;
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        float vals[LPK];
#pragma unroll
        for (int i = 0; i < LPK; ++i) {
          const int flat = gi * LPK + i;
          vals[i] = part[flat / HPGP][flat % HPGP];
        }
        const float tot = reduce_scatter<LPK>(vals, ll);
        const int flat = gi * LPK + ll;
        const int u = flat / HPGP, h = flat % HPGP;
        const int kl = it * KPI + (warp * U + u) * KPW + sub;
        if (kl < nkeys) s_scores[kl * HPGP + h] = tot;
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int kl = it * KPI + (warp * U + u) * KPW + sub;
#pragma unroll
        for (int h = 0; h < HPGP; ++h) {
          const float tot = group_allreduce_sum<LPK>(part[u][h]);
          if (ll == 0 && kl < nkeys) s_scores[kl * HPGP + h] = tot;
        }
      }
    }
  }
  __syncthreads();

  // ---- chunk softmax statistics per head: m = max, p = exp2(s - m) written back, l = sum p
  for (int h = warp; h < HPG; h += kDecodeWarps) {
    float m = -INFINITY;
    for (int kl = lane; kl < nkeys; kl += 32) m = fmaxf(m, s_scores[kl * HPGP + h]);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    const float mref = (m == -INFINITY) ? 0.f : m;
    float l = 0.f;
    for (int kl = lane; kl < nkeys; kl += 32) {
      const float pv = ptx::ex2(s_scores[kl * HPGP + h] - mref);
      s_scores[kl * HPGP + h] = pv;
      l += pv;
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) l += __shfl_xor_sync(0xffffffffu, l, s);
    if (lane == 0) {
      s_m[h] = m;
      s_l[h] = l;
    }
  }
  __syncthreads();

  // ---- pass 2: acc[h][e] += p[key][h] * v[key][e]
  float acc[HPG][EPL];
#pragma unroll
  for (int h = 0; h < HPG; ++h)
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[h][e] = 0.f;

  for (int it = 0; it * KPI < nkeys; ++it) {
    uint32_t vw[U][CPL][WPV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kl = it * KPI + (warp * U + u) * KPW + sub;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int e0 = (ll + c * LPK) * VEC;
#pragma unroll
        for (int i = 0; i < WPV; ++i) vw[u][c][i] = 0u;
        if (kl < nkeys && e0 < p.hd) VecLoad<VEC>::ld(vbase + (long long)(ks + kl) * p.vs_t + e0, vw[u][c]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kl = it * KPI + (warp * U + u) * KPW + sub;
      float pr[HPGP];
      if (kl < nkeys) {
        if constexpr (HPGP == 4) {
          const float4 t = *reinterpret_cast<const float4*>(&s_scores[kl * 4]);
          pr[0] = t.x; pr[1] = t.y; pr[2] = t.z; pr[3] = t.w;
        } else if constexpr (HPGP == 2) {
          const float2 t = *reinterpret_cast<const float2*>(&s_scores[kl * 2]);
          pr[0] = t.x; pr[1] = t.y;
        } else {
#pragma unroll
          for (int h = 0; h < HPGP; ++h) pr[h] = s_scores[kl * HPGP + h];
        }
      } else {
#pragma unroll
        for (int h = 0; h < HPGP; ++h) pr[h] = 0.f;
      }
      float vf[EPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c)
#pragma unroll
        for (int i = 0; i < WPV; ++i) {
          vf[c * VEC + 2 * i] = ptx::bf16lo(vw[u][c][i]);
          vf[c * VEC + 2 * i + 1] = ptx::bf16hi(vw[u][c][i]);
        }
#pragma unroll
      for (int h = 0; h < HPG; ++h)
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[h][e] = fmaf(pr[h], vf[e], acc[h][e]);
    }
  }

  // ---- reduce acc over the key slots of a warp, then over warps through smem
#pragma unroll
  for (int h = 0; h < HPG; ++h)
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      float x = acc[h][e];
#pragma unroll
      for (int s = 16; s >= LPK; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
      acc[h][e] = x;
    }
  if (sub == 0) {
#pragma unroll
    for (int h = 0; h < HPG; ++h)
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int e0 = (ll + c * LPK) * VEC;
        if (e0 < p.hd) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) s_acc[(warp * HPG + h) * p.hd + e0 + i] = acc[h][c * VEC + i];
        }
      }
  }
  __syncthreads();

  for (int idx = threadIdx.x; idx < nh * p.hd; idx += kDecodeThreads) {
    const int h = idx / p.hd, e = idx % p.hd;
    float x = 0.f;
#pragma unroll
    for (int w = 0; w < kDecodeWarps; ++w) x += s_acc[(w * HPG + h) * p.hd + e];
    const int hq = h0 + h;
    if (p.num_splits == 1) {
      const float l = s_l[h];
      const float y = l > 0.f ? x / l : 0.f;
      p.o[b * p.os_b + (long long)hq * p.os_h + e] = __float2bfloat16(y);
    } else {
      p.ws_acc[((long long)(b * p.H + hq) * p.num_splits + split) * p.hd + e] = x;
    }
  }
  if (p.num_splits > 1 && threadIdx.x < nh) {
    const int hq = h0 + threadIdx.x;
    float* ml = p.ws_ml + ((long long)(b * p.H + hq) * p.num_splits + split) * 2;
    ml[0] = s_m[threadIdx.x];
    ml[1] = s_l[threadIdx.x];
  }
}

// Merge the split-K partials: one CTA per (b, h).
__global__ void __launch_bounds__(128) decode_combine_kernel(const DecodeParams p) {
  const int bh = blockIdx.x;
  const int b = bh / p.H, h = bh % p.H;
  const float* ml = p.ws_ml + (long long)bh * p.num_splits * 2;
  float M = -INFINITY;
  for (int s = 0; s < p.num_splits; ++s) M = fmaxf(M, ml[2 * s]);
  const float Mref = (M == -INFINITY) ? 0.f : M;
  float Lsum = 0.f;
  for (int s = 0; s < p.num_splits; ++s) {
    const float m = ml[2 * s];
    if (m != -INFINITY) Lsum += ptx::ex2(m - Mref) * ml[2 * s + 1];
  }
  const float inv = Lsum > 0.f ? 1.f / Lsum : 0.f;
  const float* acc = p.ws_acc + (long long)bh * p.num_splits * p.hd;
  for (int e = threadIdx.x; e < p.hd; e += blockDim.x) {
    float x = 0.f;
    for (int s = 0; s < p.num_splits; ++s) {
      const float m = ml[2 * s];
      if (m != -INFINITY) x += ptx::ex2(m - Mref) * acc[(long long)s * p.hd + e];
    }
    p.o[b * p.os_b + (long long)h * p.os_h + e] = __float2bfloat16(x * inv);
  }
}

}  // namespace vats

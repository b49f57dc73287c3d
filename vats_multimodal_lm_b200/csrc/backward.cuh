// backward.cuh — dQ / dK / dV of the GQA + sliding-window attention core.
//
// The reference trains through the same modules (training/transformers/nlp/loops/training_loop.py:54-65,
// src/transformers/nlp/model.py:281-294 re-runs forward under `checkpoint`, tests/transformers/nlp/attention_tests.py
// `test_gradients`), so the drop-in op needs a backward.  torch differentiates F.scaled_dot_product_attention
// (src/optimized_attention.py:709-714) for the reference; here:
//
//   P = softmax(scale * Q K^T | mask)        D_i = sum_c dO_ic * O_ic
//   dV = P^T dO        dP = dO V^T        dS = P o (dP - D)        dQ = scale * dS K        dK = scale * dS^T Q
//   (GQA: dK / dV of a KV head sum over the H/G query heads that read it.)
//
// Two kernels, both bf16 operands / fp32 accumulation on mma.sync.m16n8k16 with operands staged in shared memory
// (ldmatrix), deterministic (no atomics):
//   attn_bwd_dq_kernel   one CTA per (sequence, query head, 64 query rows): ONE pass over the KV blocks the mask allows
//                        — the row statistics are recomputed online (the forward kernels do not store a
//                        log-sum-exp) and dQ accumulates against the running maximum, rescaled when it grows —
//                        then writes lse and D for the other kernel;
//   attn_bwd_dkv_kernel  one CTA per (sequence, KV head, 64 keys): loops over the query heads of the group and the
//                        query blocks, works on the TRANSPOSED tiles (S^T = K Q^T, so P^T / dS^T come out directly in
//                        the fragment layout the dV / dK products need) and accumulates dK, dV.
// The mask predicate is the one of mask.cuh (bit-exact with the forward kernels), including q_valid / k_valid and
// fully masked rows (zero gradient).  This is a correct, tensor-core backward — not yet a tcgen05 one.
#pragma once
#include "decode_mma.cuh"   // ldmatrix / mma.sync wrappers
#include "mask.cuh"
#include "prefill_simt.cuh" // PrefillParams
#include "prefill_tc.cuh"   // cp.async helpers

namespace vats {

#ifndef VATS_BWD_DQ_FAST   // A/B knob: unmasked fast path of the dQ kernel's score tile
#define VATS_BWD_DQ_FAST 1
#endif
#ifndef VATS_BWD_OCC   // A/B knob: 1 = ask ptxas for 3 CTAs per SM up to head dim 64 (2 up to 96)
#define VATS_BWD_OCC 1
#endif
constexpr int kBwdThreads = 128;
constexpr int bwd_min_blocks(int ks) { return VATS_BWD_OCC ? (ks <= 4 ? 3 : (ks <= 6 ? 2 : 1)) : 1; }
// (asking for 4 CTAs per SM in the dQ kernel — its 4-tile shared memory would allow it up to hd 64 — caps it at 128
// registers, 124 bytes of spills, and measured the same 0.84 ms on the training step)
constexpr int kBwdBM = 64;   // query rows per block
constexpr int kBwdBN = 64;   // keys per block

struct BwdParams {
  PrefillParams a;               // q, k, v, o (forward output), strides, mask, scale_log2
  const __nv_bfloat16* dout;     // [N, Tq, H, hd], strides dos_*
  long long dos_n, dos_t, dos_h;
  __nv_bfloat16* dq;             // [N, Tq, H, hd] dense
  __nv_bfloat16* dk;             // [N, Tk, G, hd] dense
  __nv_bfloat16* dv;
  float* lse;                    // [N, H, Tq] scaled-log2 log-sum-exp (+inf for rows without any allowed key)
  float* dsum;                   // [N, H, Tq] D_i
  float scale;
  int hd_pad;                    // head dim rounded up to 16
  int vec16;                     // 1: q, k, v, o, dout have 16-byte aligned bases and strides: tiles are staged with
                                 // 16-byte cp.async copies (else 4-byte)
};

// dQ kernel: 4 tiles (Q / dO / O, then two K / V pairs) + D; dK/dV kernel: 6 tiles (K, V, two Q / dO pairs) + two lse / D pairs
__host__ __device__ inline size_t bwd_smem_bytes(int hd_pad) {
  return (size_t)6 * kBwdBM * (hd_pad + 8) * 2 + 4 * kBwdBM * sizeof(float);
}
__host__ __device__ inline size_t bwd_dq_smem_bytes(int hd_pad) {
  return (size_t)4 * kBwdBM * (hd_pad + 8) * 2 + kBwdBM * sizeof(float);
}

// rows [row0, row0 + 64) x hd of a row-strided bf16 matrix -> smem tile [64][KS * 16 + 8], asynchronously: cp.async copies
// (16-byte when the source allows, else 4-byte), zero-filled past `rows` / hd.  The caller commits and waits.
// (The first version staged tiles with a loop of dependent 32-bit load / store pairs: ~11 K clk per tile, every tile,
// with nothing else running in the CTA — the dK/dV kernel spent 37 K clk per (query block, key block) pair.)
template <int KS>
__device__ __forceinline__ void bwd_tile_async(__nv_bfloat16* dst, const __nv_bfloat16* src, long long stride, int row0,
                                               int rows, int hd, int vec16) {
  constexpr int pitch = KS * 16 + 8;
  constexpr int cpr = KS * 2;            // 16-byte chunks per row
  constexpr int segs = (cpr + 7) / 8;    // 128-byte segments per row
  // Eight consecutive threads take the eight 16-byte chunks of one 128-byte row segment (one line per row and warp
  // instruction, conflict-free shared-memory writes); a thread's rows are tid / 8 + 16 m, so it computes one address
  // per tile and steps it by 16 rows.  (A division of the chunk index per copy made the address arithmetic a fifth of
  // all executed instructions; two threads per row with interleaved chunks quadrupled the L2 requests.)
  const int rb = (int)threadIdx.x >> 3, c0 = (int)threadIdx.x & 7;
  const __nv_bfloat16* g = src + (long long)(row0 + rb) * stride + c0 * 8;
  const uint32_t d = ptx::smem_u32(dst + rb * pitch + c0 * 8);
#pragma unroll
  for (int m = 0; m < kBwdBM / 16; ++m) {
    const bool live = row0 + rb + 16 * m < rows;
#pragma unroll
    for (int sg = 0; sg < segs; ++sg) {
      const int c = c0 + 8 * sg;
      if (cpr % 8 == 0 || c < cpr) {
        if (vec16) {
          int bytes = (hd - c * 8) * 2;
          bytes = !live ? 0 : (bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d + (m * 16 * pitch + sg * 64) * 2),
                       "l"(bytes ? g + (long long)(16 * m) * stride + sg * 64 : src), "r"(bytes)
                       : "memory");
        } else {
#pragma unroll
          for (int w = 0; w < 4; ++w) {   // the same chunk as four 32-bit copies (rows only 4-byte aligned)
            const int bytes = (live && c * 8 + 2 * w < hd) ? 4 : 0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d + (m * 16 * pitch + sg * 64) * 2 + w * 4),
                         "l"(bytes ? g + (long long)(16 * m) * stride + sg * 64 + 2 * w : src), "r"(bytes)
                         : "memory");
          }
        }
      }
    }
  }
}

// A fragment (16 x 16) of a row-major smem tile at (row0, k0)
__device__ __forceinline__ void bwd_ldsm_a(const __nv_bfloat16* tile, int pitch, int row0, int k0, uint32_t (&a)[4]) {
  const int lane = threadIdx.x & 31;
  const uint32_t addr = ptx::smem_u32(tile + (row0 + (lane & 15)) * pitch + k0 + (lane >> 4) * 8);
  ldmatrix_x4(addr, a[0], a[1], a[2], a[3]);
}
// B fragments of two n-tiles (16 n x 16 k) from a tile stored [n][k]: (b[0], b[1]) = n-tile 0, (b[2], b[3]) = n-tile 1
__device__ __forceinline__ void bwd_ldsm_b(const __nv_bfloat16* tile, int pitch, int n0, int k0, uint32_t (&b)[4]) {
  const int lane = threadIdx.x & 31;
  const uint32_t addr = ptx::smem_u32(tile + (n0 + (lane & 7) + (lane >> 4) * 8) * pitch + k0 + ((lane >> 3) & 1) * 8);
  ldmatrix_x4(addr, b[0], b[1], b[2], b[3]);
}
// ... from a tile stored [k][n] (transposed on the fly)
__device__ __forceinline__ void bwd_ldsm_bt(const __nv_bfloat16* tile, int pitch, int k0, int n0, uint32_t (&b)[4]) {
  const int lane = threadIdx.x & 31;
  const uint32_t addr = ptx::smem_u32(tile + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * pitch + n0 + (lane >> 4) * 8);
  ldmatrix_x4_trans(addr, b[0], b[1], b[2], b[3]);
}

// key range [lo, hi] (inclusive, clamped to [0, Tk)) a query row may attend, geometry only; hi < lo = nothing
__device__ __forceinline__ void bwd_row_range(const MaskParams& mp, int i, int* lo, int* hi) {
  long long l = key_lo(mp, i), h = key_hi(mp, i);
  if (l < 0) l = 0;
  if (h > (long long)mp.Tk - 1) h = (long long)mp.Tk - 1;
  *lo = l > 0x3fffffffLL ? 0x3fffffff : (int)l;
  *hi = h < -1 ? -1 : (int)h;
}

// ---------------------------------------------------------------------------------------------------- dQ (+ lse, D)
template <int KS>   // KS = hd_pad / 16
__global__ void __launch_bounds__(kBwdThreads, bwd_min_blocks(KS)) attn_bwd_dq_kernel(const BwdParams P) {
  using namespace ptx;
  extern __shared__ __align__(16) unsigned char bwd_smem[];
  const PrefillParams& a = P.a;
  constexpr int pitch = KS * 16 + 8;   // == P.hd_pad + 8: compile-time, so fragment addresses fold into constants
  // four tile buffers: Q / dO / O while the A fragments and D are taken, then K (pass 1) or K / V pairs (pass 2),
  // double-buffered: the next KV block is on its way (cp.async) while the current one is computed on
  auto sT = [&](int i) { return reinterpret_cast<__nv_bfloat16*>(bwd_smem) + i * kBwdBM * pitch; };
  __nv_bfloat16* sQ = sT(0);
  __nv_bfloat16* sdO = sT(1);
  float* sD = reinterpret_cast<float*>(reinterpret_cast<__nv_bfloat16*>(bwd_smem) + 4 * kBwdBM * pitch);

  const int q0 = blockIdx.x * kBwdBM, h = blockIdx.y, n = blockIdx.z;
  const int g = h / a.hpg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
  const __nv_bfloat16* qp = a.q + n * a.qs_n + (long long)h * a.qs_h;
  const __nv_bfloat16* op = a.o + n * a.os_n + (long long)h * a.os_h;
  const __nv_bfloat16* dop = P.dout + n * P.dos_n + (long long)h * P.dos_h;
  const __nv_bfloat16* kp = a.k + n * a.ks_n + (long long)g * a.ks_h;
  const __nv_bfloat16* vp = a.v + n * a.vs_n + (long long)g * a.vs_h;

  // ---- Q, dO (and O, parked in the K buffer) -> smem;  D_i = sum_c dO_ic * O_ic
  bwd_tile_async<KS>(sQ, qp, a.qs_t, q0, a.Tq, a.hd, P.vec16);
  bwd_tile_async<KS>(sdO, dop, P.dos_t, q0, a.Tq, a.hd, P.vec16);
  bwd_tile_async<KS>(sT(2), op, a.os_t, q0, a.Tq, a.hd, P.vec16);
  cpasync_commit();
  cpasync_wait<0>();
  __syncthreads();
  {
    const int r = threadIdx.x >> 1, half = threadIdx.x & 1;
    float acc = 0.f;
    for (int c = half; c < P.hd_pad; c += 2) acc += __bfloat162float(sdO[r * pitch + c]) * __bfloat162float(sT(2)[r * pitch + c]);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (half == 0) sD[r] = acc;
  }
  __syncthreads();

  // this thread's two rows: r0 = warp * 16 + gq, r1 = r0 + 8
  const int rl0 = warp * 16 + gq, rl1 = rl0 + 8;
  const int i0 = q0 + rl0, i1 = q0 + rl1;
  int lo0 = 0, hi0 = -1, lo1 = 0, hi1 = -1;
  if (i0 < a.Tq && (a.q_valid == nullptr || a.q_valid[(long long)n * a.Tq + i0])) bwd_row_range(a.mask, i0, &lo0, &hi0);
  if (i1 < a.Tq && (a.q_valid == nullptr || a.q_valid[(long long)n * a.Tq + i1])) bwd_row_range(a.mask, i1, &lo1, &hi1);
  const float D0 = sD[rl0], D1 = sD[rl1];

  uint32_t qa[KS][4], da[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    bwd_ldsm_a(sQ, pitch, warp * 16, ks * 16, qa[ks]);
    bwd_ldsm_a(sdO, pitch, warp * 16, ks * 16, da[ks]);
  }

  int t_first, t_last;
  tile_range(a.mask, q0, kBwdBM, kBwdBN, &t_first, &t_last);

  // all 64 rows of the block are live and no key mask: blocks the geometry allows entirely need no predicate
  const bool plain_rows = a.q_valid == nullptr && a.k_valid == nullptr && q0 + kBwdBM <= a.Tq;
  // S[16 x 64] of this warp for the KV block in sK, scaled to log2 units and masked (-inf)
  auto scores = [&](const __nv_bfloat16* sK, int k0, float (&s)[8][4]) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        bwd_ldsm_b(sK, pitch, np * 16, ks * 16, b);
        mma_bf16_16816(s[np * 2], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b[0], b[1]);
        mma_bf16_16816(s[np * 2 + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b[2], b[3]);
      }
    }
    if (VATS_BWD_DQ_FAST && plain_rows && tile_is_full(a.mask, k0 / kBwdBN, q0, kBwdBM, kBwdBN)) {   // every pair allowed
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] *= a.scale_log2;
      }
      return;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = k0 + nt * 8 + tq * 2 + (e & 1);
        const bool top = e < 2;
        bool ok = top ? (j >= lo0 && j <= hi0) : (j >= lo1 && j <= hi1);
        if (ok && a.k_valid != nullptr) ok = a.k_valid[(long long)n * a.Tk + j] != 0;
        s[nt][e] = ok ? s[nt][e] * a.scale_log2 : -INFINITY;
      }
    }
  };

  // ---- ONE pass over the KV blocks the mask allows: the row statistics are accumulated online, as in the forward,
  //      and so is dQ:   dQ_i = (1 / l_i) * sum_j exp2(s_ij - m_i) (dP_ij - D_i) K_j   with the running maximum in
  //      place of m_i and the accumulator rescaled whenever it grows (exact: a common factor per row).  The first
  //      version made a pass for (m, l) and a second one for dQ: a quarter more matrix products, twice the exponentials
  //      and the K tiles fetched twice.
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float dq[KS * 2][4];
#pragma unroll
  for (int nt = 0; nt < KS * 2; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
  __syncthreads();   // every warp holds its Q / dO fragments and D: the four buffers are free
  if (t_first <= t_last) {
    bwd_tile_async<KS>(sT(0), kp, a.ks_t, t_first * kBwdBN, a.Tk, a.hd, P.vec16);
    bwd_tile_async<KS>(sT(1), vp, a.vs_t, t_first * kBwdBN, a.Tk, a.hd, P.vec16);
  }
  cpasync_commit();
  for (int t = t_first, it = 0; t <= t_last; ++t, ++it) {
    cpasync_wait<0>();
    __syncthreads();   // block t has landed for everyone, and everyone is done with block t - 1
    if (t < t_last) {
      bwd_tile_async<KS>(sT(2 * ((it + 1) & 1)), kp, a.ks_t, (t + 1) * kBwdBN, a.Tk, a.hd, P.vec16);
      bwd_tile_async<KS>(sT(2 * ((it + 1) & 1) + 1), vp, a.vs_t, (t + 1) * kBwdBN, a.Tk, a.hd, P.vec16);
    }
    cpasync_commit();
    const __nv_bfloat16* sK = sT(2 * (it & 1));
    const __nv_bfloat16* sV = sT(2 * (it & 1) + 1);
    float s[8][4];
    scores(sK, t * kBwdBN, s);
    // ---- running maximum / sum, rescale of the accumulator
    float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      t0 = fmaxf(t0, fmaxf(s[nt][0], s[nt][1]));
      t1 = fmaxf(t1, fmaxf(s[nt][2], s[nt][3]));
    }
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
    const float n0 = fmaxf(m0, t0), n1 = fmaxf(m1, t1);
    const float r0 = n0 == -INFINITY ? 0.f : n0, r1 = n1 == -INFINITY ? 0.f : n1;
    const float c0 = m0 == -INFINITY ? 0.f : ex2(m0 - r0), c1 = m1 == -INFINITY ? 0.f : ex2(m1 - r1);
    m0 = n0;
    m1 = n1;
    if (__any_sync(0xffffffffu, c0 != 1.f || c1 != 1.f)) {
#pragma unroll
      for (int nt = 0; nt < KS * 2; ++nt) {
        dq[nt][0] *= c0;
        dq[nt][1] *= c0;
        dq[nt][2] *= c1;
        dq[nt][3] *= c1;
      }
    }
    // ---- dP = dO V^T
    float dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        bwd_ldsm_b(sV, pitch, np * 16, ks * 16, b);
        mma_bf16_16816(dp[np * 2], da[ks][0], da[ks][1], da[ks][2], da[ks][3], b[0], b[1]);
        mma_bf16_16816(dp[np * 2 + 1], da[ks][0], da[ks][1], da[ks][2], da[ks][3], b[2], b[3]);
      }
    }
    // ---- p = exp2(s - m) (un-normalised), row sums, dS = p o (dP - D) in A-fragment form (one k-step per n-tile pair)
    float a0 = 0.f, a1 = 0.f;
    uint32_t dsa[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float d[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int nt = 2 * j + u;
        const float p0 = ex2(s[nt][0] - r0), p1 = ex2(s[nt][1] - r0), p2 = ex2(s[nt][2] - r1), p3 = ex2(s[nt][3] - r1);
        a0 += p0 + p1;
        a1 += p2 + p3;
        d[u][0] = p0 * (dp[nt][0] - D0);
        d[u][1] = p1 * (dp[nt][1] - D0);
        d[u][2] = p2 * (dp[nt][2] - D1);
        d[u][3] = p3 * (dp[nt][3] - D1);
      }
      dsa[j][0] = pack_bf16x2(d[0][0], d[0][1]);
      dsa[j][1] = pack_bf16x2(d[0][2], d[0][3]);
      dsa[j][2] = pack_bf16x2(d[1][0], d[1][1]);
      dsa[j][3] = pack_bf16x2(d[1][2], d[1][3]);
    }
    l0 = l0 * c0 + a0;
    l1 = l1 * c1 + a1;
    // ---- dQ += dS K
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int np = 0; np < KS; ++np) {
        uint32_t b[4];
        bwd_ldsm_bt(sK, pitch, j * 16, np * 16, b);
        mma_bf16_16816(dq[np * 2], dsa[j][0], dsa[j][1], dsa[j][2], dsa[j][3], b[0], b[1]);
        mma_bf16_16816(dq[np * 2 + 1], dsa[j][0], dsa[j][1], dsa[j][2], dsa[j][3], b[2], b[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // scaled-log2 log-sum-exp for the dK/dV kernel; rows without any allowed key get +inf, so that exp2(s - lse) = 0
  const float lse0 = (l0 > 0.f) ? m0 + __log2f(l0) : INFINITY;
  const float lse1 = (l1 > 0.f) ? m1 + __log2f(l1) : INFINITY;
  if (tq == 0) {
    if (i0 < a.Tq) {
      P.lse[((long long)n * a.H + h) * a.Tq + i0] = lse0;
      P.dsum[((long long)n * a.H + h) * a.Tq + i0] = D0;
    }
    if (i1 < a.Tq) {
      P.lse[((long long)n * a.H + h) * a.Tq + i1] = lse1;
      P.dsum[((long long)n * a.H + h) * a.Tq + i1] = D1;
    }
  }
  const float inv0 = l0 > 0.f ? P.scale / l0 : 0.f, inv1 = l1 > 0.f ? P.scale / l1 : 0.f;
  // ---- dQ * scale -> bf16 (dense [N, Tq, H, hd])
  __nv_bfloat16* dqp = P.dq + (((long long)n * a.Tq) * a.H + h) * a.hd;
#pragma unroll
  for (int nt = 0; nt < KS * 2; ++nt) {
    const int col = nt * 8 + tq * 2;
    if (col < a.hd) {
      if (i0 < a.Tq)
        *reinterpret_cast<uint32_t*>(dqp + (long long)i0 * a.H * a.hd + col) = pack_bf16x2(dq[nt][0] * inv0, dq[nt][1] * inv0);
      if (i1 < a.Tq)
        *reinterpret_cast<uint32_t*>(dqp + (long long)i1 * a.H * a.hd + col) = pack_bf16x2(dq[nt][2] * inv1, dq[nt][3] * inv1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------- dK, dV
template <int KS>
__global__ void __launch_bounds__(kBwdThreads, bwd_min_blocks(KS)) attn_bwd_dkv_kernel(const BwdParams P) {
  using namespace ptx;
  extern __shared__ __align__(16) unsigned char bwd_smem[];
  const PrefillParams& a = P.a;
  constexpr int pitch = KS * 16 + 8;   // == P.hd_pad + 8: compile-time, so fragment addresses fold into constants
  // six tile buffers: K, V of this block (resident) and two Q / dO pairs — the next (head, query block) pair is on its
  // way (cp.async) while the current one is computed on; its lse / D rows travel through registers
  __nv_bfloat16* sT0 = reinterpret_cast<__nv_bfloat16*>(bwd_smem);
  __nv_bfloat16* sK = sT0;
  __nv_bfloat16* sV = sT0 + kBwdBM * pitch;
  float* sStat = reinterpret_cast<float*>(sT0 + 6 * kBwdBM * pitch);   // [2 buffers][lse | D][64]

  const int k0 = blockIdx.x * kBwdBN, g = blockIdx.y, n = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
  const __nv_bfloat16* kp = a.k + n * a.ks_n + (long long)g * a.ks_h;
  const __nv_bfloat16* vp = a.v + n * a.vs_n + (long long)g * a.vs_h;
  bwd_tile_async<KS>(sK, kp, a.ks_t, k0, a.Tk, a.hd, P.vec16);
  bwd_tile_async<KS>(sV, vp, a.vs_t, k0, a.Tk, a.hd, P.vec16);

  // this thread's two key rows
  const int j0 = k0 + warp * 16 + gq, j1 = j0 + 8;
  const bool kv0 = j0 < a.Tk && (a.k_valid == nullptr || a.k_valid[(long long)n * a.Tk + j0]);
  const bool kv1 = j1 < a.Tk && (a.k_valid == nullptr || a.k_valid[(long long)n * a.Tk + j1]);
  const long long off = (long long)a.Tk - a.Tq;

  float dk[KS * 2][4], dv[KS * 2][4];
#pragma unroll
  for (int nt = 0; nt < KS * 2; ++nt) {
    dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
    dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
  }

  const int q_blocks = (a.Tq + kBwdBM - 1) / kBwdBM;
  // the (head of the group, query block) pairs with at least one allowed (query, key) pair against this key block, in
  // order; `advance` moves to the next one (every thread walks the same sequence)
  auto pair_allowed = [&](int qb) -> bool {
    int t_first, t_last;
    tile_range(a.mask, qb * kBwdBM, kBwdBM, kBwdBN, &t_first, &t_last);
    return (int)blockIdx.x >= t_first && (int)blockIdx.x <= t_last;
  };
  // the query blocks that can see this key block do not depend on the head: their span [qb_lo, qb_hi] is found once
  // (a band mask leaves most of the q_blocks x hpg candidates out — walking all of them for every head was a fifth of
  // the kernel's instructions on a 384-key window)
  int qb_lo = q_blocks, qb_hi = -1;
  for (int qb = 0; qb < q_blocks; ++qb) {
    if (pair_allowed(qb)) {
      if (qb < qb_lo) qb_lo = qb;
      qb_hi = qb;
    }
  }
  auto advance = [&](int& hh, int& qb) -> bool {
    if (qb_hi < qb_lo) return false;
    for (;;) {
      if (qb < qb_lo) {
        qb = qb_lo;
      } else if (++qb > qb_hi) {
        qb = qb_lo;
        if (++hh >= a.hpg) return false;
      }
      if (pair_allowed(qb)) return true;
    }
  };
  auto stage_pair = [&](int hh, int qb, int buf) {
    const int h = g * a.hpg + hh;
    bwd_tile_async<KS>(sT0 + (2 + 2 * buf) * kBwdBM * pitch, a.q + n * a.qs_n + (long long)h * a.qs_h, a.qs_t,
                       qb * kBwdBM, a.Tq, a.hd, P.vec16);
    bwd_tile_async<KS>(sT0 + (3 + 2 * buf) * kBwdBM * pitch, P.dout + n * P.dos_n + (long long)h * P.dos_h, P.dos_t,
                       qb * kBwdBM, a.Tq, a.hd, P.vec16);
  };
  auto load_stats = [&](int hh, int qb, float* lse, float* dsum) {   // threads < 64: row threadIdx.x of the query block
    const int h = g * a.hpg + hh;
    const int i = qb * kBwdBM + (int)threadIdx.x;
    const bool live = i < a.Tq && (a.q_valid == nullptr || a.q_valid[(long long)n * a.Tq + i]);
    *lse = live ? P.lse[((long long)n * a.H + h) * a.Tq + i] : INFINITY;
    *dsum = live ? P.dsum[((long long)n * a.H + h) * a.Tq + i] : 0.f;
  };
  int hh = 0, qb = -1;
  bool have = advance(hh, qb);
  if (have) {
    stage_pair(hh, qb, 0);
    if (threadIdx.x < kBwdBM) load_stats(hh, qb, &sStat[threadIdx.x], &sStat[kBwdBM + threadIdx.x]);
  }
  cpasync_commit();
  for (int it = 0; have; ++it) {
    int nh = hh, nq = qb;
    const bool more = advance(nh, nq);
    cpasync_wait<0>();
    __syncthreads();   // this pair (and K, V) has landed for everyone, and everyone is done with the previous pair
    float lse_n = INFINITY, dsum_n = 0.f;
    if (more) {
      stage_pair(nh, nq, (it + 1) & 1);
      if (threadIdx.x < kBwdBM) load_stats(nh, nq, &lse_n, &dsum_n);
    }
    cpasync_commit();
    const __nv_bfloat16* sQ = sT0 + (2 + 2 * (it & 1)) * kBwdBM * pitch;
    const __nv_bfloat16* sdO = sT0 + (3 + 2 * (it & 1)) * kBwdBM * pitch;
    const float* sLse = sStat + (it & 1) * 2 * kBwdBM;
    const float* sD = sLse + kBwdBM;
    const int q0 = qb * kBwdBM;
    {

      // ---- S^T = K Q^T for this warp's 16 keys x 64 queries, then P^T (rows = keys j0 / j1, columns = queries)
      float st[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ka[4];
        bwd_ldsm_a(sK, pitch, warp * 16, ks * 16, ka);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bq[4];
          bwd_ldsm_b(sQ, pitch, np * 16, ks * 16, bq);
          mma_bf16_16816(st[np * 2], ka[0], ka[1], ka[2], ka[3], bq[0], bq[1]);
          mma_bf16_16816(st[np * 2 + 1], ka[0], ka[1], ka[2], ka[3], bq[2], bq[3]);
        }
      }
      // allowed query columns (local index il) of this thread's two key rows, from the predicate solved for i:
      //   causal: j <= i + off  <=>  il >= j - off - q0;   right: il >= j - right - off - q0;   left: il <= j + left - off - q0
      int ilo0 = 0, ihi0 = kBwdBM - 1, ilo1 = 0, ihi1 = kBwdBM - 1;
      if (!tile_is_full(a.mask, (int)blockIdx.x, q0, kBwdBM, kBwdBN)) {
        const long long base = off + (long long)q0;
        auto clampi = [](long long x) { return x < -1 ? -1 : (x > kBwdBM ? kBwdBM : (int)x); };
        long long l0 = 0, l1 = 0, h0 = kBwdBM - 1, h1 = kBwdBM - 1;
        if (a.mask.causal) { l0 = (long long)j0 - base; l1 = (long long)j1 - base; }
        if (a.mask.right >= 0) {
          const long long r0 = (long long)j0 - a.mask.right - base, r1 = (long long)j1 - a.mask.right - base;
          l0 = (a.mask.causal && l0 > r0) ? l0 : r0;
          l1 = (a.mask.causal && l1 > r1) ? l1 : r1;
        }
        if (a.mask.left >= 0) { h0 = (long long)j0 + a.mask.left - base; h1 = (long long)j1 + a.mask.left - base; }
        ilo0 = clampi(l0 < 0 ? 0 : l0); ilo1 = clampi(l1 < 0 ? 0 : l1);
        ihi0 = clampi(h0 > kBwdBM - 1 ? kBwdBM - 1 : h0); ihi1 = clampi(h1 > kBwdBM - 1 ? kBwdBM - 1 : h1);
      }
      if (!kv0) ihi0 = -1;   // key row past the sequence or masked by k_valid
      if (!kv1) ihi1 = -1;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int il = nt * 8 + tq * 2 + (e & 1);      // query column within the block
          const bool ok = e < 2 ? (il >= ilo0 && il <= ihi0) : (il >= ilo1 && il <= ihi1);
          st[nt][e] = ok ? ex2(st[nt][e] * a.scale_log2 - sLse[il]) : 0.f;   // lse = +inf for dead / padded rows
        }
      }
      // ---- dV += P^T dO  (P^T straight into A-fragment form; contraction over the 64 queries)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(st[2 * jj][0], st[2 * jj][1]);
        pa[1] = pack_bf16x2(st[2 * jj][2], st[2 * jj][3]);
        pa[2] = pack_bf16x2(st[2 * jj + 1][0], st[2 * jj + 1][1]);
        pa[3] = pack_bf16x2(st[2 * jj + 1][2], st[2 * jj + 1][3]);
#pragma unroll
        for (int np = 0; np < KS; ++np) {
          uint32_t bd[4];
          bwd_ldsm_bt(sdO, pitch, jj * 16, np * 16, bd);
          mma_bf16_16816(dv[np * 2], pa[0], pa[1], pa[2], pa[3], bd[0], bd[1]);
          mma_bf16_16816(dv[np * 2 + 1], pa[0], pa[1], pa[2], pa[3], bd[2], bd[3]);
        }
      }
      // ---- dP^T = V dO^T,  dS^T = P^T o (dP^T - D)
      float dpt[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t va[4];
        bwd_ldsm_a(sV, pitch, warp * 16, ks * 16, va);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bd[4];
          bwd_ldsm_b(sdO, pitch, np * 16, ks * 16, bd);
          mma_bf16_16816(dpt[np * 2], va[0], va[1], va[2], va[3], bd[0], bd[1]);
          mma_bf16_16816(dpt[np * 2 + 1], va[0], va[1], va[2], va[3], bd[2], bd[3]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int il = nt * 8 + tq * 2 + (e & 1);
          st[nt][e] *= dpt[nt][e] - sD[il];
        }
      }
      // ---- dK += dS^T Q
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint32_t dsa[4];
        dsa[0] = pack_bf16x2(st[2 * jj][0], st[2 * jj][1]);
        dsa[1] = pack_bf16x2(st[2 * jj][2], st[2 * jj][3]);
        dsa[2] = pack_bf16x2(st[2 * jj + 1][0], st[2 * jj + 1][1]);
        dsa[3] = pack_bf16x2(st[2 * jj + 1][2], st[2 * jj + 1][3]);
#pragma unroll
        for (int np = 0; np < KS; ++np) {
          uint32_t bq[4];
          bwd_ldsm_bt(sQ, pitch, jj * 16, np * 16, bq);
          mma_bf16_16816(dk[np * 2], dsa[0], dsa[1], dsa[2], dsa[3], bq[0], bq[1]);
          mma_bf16_16816(dk[np * 2 + 1], dsa[0], dsa[1], dsa[2], dsa[3], bq[2], bq[3]);
        }
      }
    }
    if (more && threadIdx.x < kBwdBM) {   // (the other stat buffer was last read two pairs ago)
      sStat[((it + 1) & 1) * 2 * kBwdBM + threadIdx.x] = lse_n;
      sStat[((it + 1) & 1) * 2 * kBwdBM + kBwdBM + threadIdx.x] = dsum_n;
    }
    hh = nh;
    qb = nq;
    have = more;
  }
  cpasync_wait<0>();   // (a block without any pair still staged its K / V)
  // ---- bf16 outputs (dense [N, Tk, G, hd])
  __nv_bfloat16* dkp = P.dk + (((long long)n * a.Tk) * a.G + g) * a.hd;
  __nv_bfloat16* dvp = P.dv + (((long long)n * a.Tk) * a.G + g) * a.hd;
#pragma unroll
  for (int nt = 0; nt < KS * 2; ++nt) {
    const int col = nt * 8 + tq * 2;
    if (col < a.hd) {
      if (j0 < a.Tk) {
        *reinterpret_cast<uint32_t*>(dkp + (long long)j0 * a.G * a.hd + col) = pack_bf16x2(dk[nt][0] * P.scale, dk[nt][1] * P.scale);
        *reinterpret_cast<uint32_t*>(dvp + (long long)j0 * a.G * a.hd + col) = pack_bf16x2(dv[nt][0], dv[nt][1]);
      }
      if (j1 < a.Tk) {
        *reinterpret_cast<uint32_t*>(dkp + (long long)j1 * a.G * a.hd + col) = pack_bf16x2(dk[nt][2] * P.scale, dk[nt][3] * P.scale);
        *reinterpret_cast<uint32_t*>(dvp + (long long)j1 * a.G * a.hd + col) = pack_bf16x2(dv[nt][2], dv[nt][3]);
      }
    }
  }
}

}  // namespace vats

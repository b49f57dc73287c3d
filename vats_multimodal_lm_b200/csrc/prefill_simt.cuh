// prefill_simt.cuh — CUDA-core GQA + window attention for tiny or irregular sequences.
//
// Serves the shapes where tensor-core tiles (128 x 128) would be >90 % padding or the layout is not TMA-legal:
// the ViT-3D temporal attention (12 544 sequences of 8 tokens, reference vit_3d/optimized_attention.py:393-430),
// unit-test geometries such as hd = 6 (reference vit_3d/optimized_attention.py:744), odd strides.  Those cases
// are HBM-bound (6 flop/byte for the temporal pass), so a warp-per-row flash-style kernel is the right tool.
//
// One CTA = (sequence n, KV group g, block of ROWS_PER_CTA "rows"); a row is a (query token, query head of the
// group) pair, so all H/G query heads of a group share the K/V tile staged in shared memory (GQA reuse).
// Each warp owns RPW rows.  Per KV tile of 32 keys: lane j computes the logit of key j for each row (q from smem
// broadcast, K row from smem), softmax statistics by warp shuffles (online, fp32), then the lanes switch to owning
// head_dim columns for acc += p_j * V[j] with p_j broadcast by shuffle.
#pragma once
#include "mask.cuh"
#include "ptx.cuh"

namespace vats {

struct PrefillParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* o;
  const uint8_t* q_valid;
  const uint8_t* k_valid;
  int N, Tq, Tk, H, G, hd;
  int hpg;
  long long qs_n, qs_t, qs_h;
  long long ks_n, ks_t, ks_h;
  long long vs_n, vs_t, vs_h;
  long long os_n, os_t, os_h;
  float scale_log2;  // scale * log2(e)
  MaskParams mask;
};

constexpr int kSimtWarps = 4;
constexpr int kSimtRPW = 8;                          // rows per warp
constexpr int kSimtRows = kSimtWarps * kSimtRPW;     // rows per CTA
constexpr int kSimtTileN = 32;                       // keys per tile
constexpr int kSimtMaxHd = 256;

// CPLN = ceil(hd / 32): head_dim columns owned by each lane in the PV phase.
template <int CPLN>
__global__ void __launch_bounds__(kSimtWarps * 32) prefill_simt_kernel(const PrefillParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int hd = p.hd;
  const int pitch = hd + 1;  // fp32 words per staged row; odd pitch => conflict-free column walks
  float* sK = reinterpret_cast<float*>(smem_raw);         // [32][pitch]
  float* sV = sK + kSimtTileN * pitch;                    // [32][pitch]
  float* sQ = sV + kSimtTileN * pitch;                    // [kSimtRows][pitch]
  __shared__ uint8_t sKvalid[kSimtTileN];

  const int n = blockIdx.z;
  const int g = blockIdx.y;
  const int row0 = blockIdx.x * kSimtRows;                // first row of this CTA
  const int total_rows = p.Tq * p.hpg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // row r -> (token = r / hpg, head-in-group = r % hpg): consecutive rows share a token => tight key range
  const int tok_first = row0 / p.hpg;
  int tok_last = (min(row0 + kSimtRows, total_rows) - 1) / p.hpg;

  // ---- stage the q rows (scaled) in smem as fp32
  for (int idx = threadIdx.x; idx < kSimtRows * hd; idx += blockDim.x) {
    const int r = idx / hd, e = idx % hd;
    const int row = row0 + r;
    float x = 0.f;
    if (row < total_rows) {
      const int tok = row / p.hpg, hh = row % p.hpg;
      x = __bfloat162float(p.q[n * p.qs_n + (long long)tok * p.qs_t + (long long)(g * p.hpg + hh) * p.qs_h + e]) *
          p.scale_log2;
    }
    sQ[r * pitch + e] = x;
  }

  float m_run[kSimtRPW], l_run[kSimtRPW], acc[kSimtRPW][CPLN];
#pragma unroll
  for (int r = 0; r < kSimtRPW; ++r) {
    m_run[r] = -INFINITY;
    l_run[r] = 0.f;
#pragma unroll
    for (int c = 0; c < CPLN; ++c) acc[r][c] = 0.f;
  }

  int t_first, t_last;
  tile_range(p.mask, tok_first, tok_last - tok_first + 1, kSimtTileN, &t_first, &t_last);

  for (int t = t_first; t <= t_last; ++t) {
    const int k0 = t * kSimtTileN;
    __syncthreads();  // previous tile fully consumed (also orders the sQ fill before first use)
    for (int idx = threadIdx.x; idx < kSimtTileN * hd; idx += blockDim.x) {
      const int j = idx / hd, e = idx % hd;
      const int key = k0 + j;
      float kx = 0.f, vx = 0.f;
      if (key < p.Tk) {
        kx = __bfloat162float(p.k[n * p.ks_n + (long long)key * p.ks_t + (long long)g * p.ks_h + e]);
        vx = __bfloat162float(p.v[n * p.vs_n + (long long)key * p.vs_t + (long long)g * p.vs_h + e]);
      }
      sK[j * pitch + e] = kx;
      sV[j * pitch + e] = vx;
    }
    if (threadIdx.x < kSimtTileN) {
      const int key = k0 + threadIdx.x;
      uint8_t ok = key < p.Tk ? 1 : 0;
      if (ok && p.k_valid) ok = p.k_valid[(long long)n * p.Tk + key] ? 1 : 0;
      sKvalid[threadIdx.x] = ok;
    }
    __syncthreads();

    const int key = k0 + lane;
    const bool key_ok = sKvalid[lane] != 0;
#pragma unroll
    for (int r = 0; r < kSimtRPW; ++r) {
      const int rl = warp * kSimtRPW + r;
      const int row = row0 + rl;
      if (row >= total_rows) continue;  // warp-uniform
      const int tok = row / p.hpg;
      // logit of (row, key = lane)
      float s = 0.f;
      const float* qrow = sQ + rl * pitch;
      const float* krow = sK + lane * pitch;
      for (int e = 0; e < hd; ++e) s = fmaf(qrow[e], krow[e], s);
      const bool ok = key_ok && allowed_geom(p.mask, tok, key);
      s = ok ? s : -INFINITY;
      // online softmax over the 32 keys of the tile: warp-level max / sum
      float mt = s;
#pragma unroll
      for (int x = 16; x >= 1; x >>= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, x));
      const float m_new = fmaxf(m_run[r], mt);
      const float mref = (m_new == -INFINITY) ? 0.f : m_new;
      const float pj = ptx::ex2(s - mref);                       // 0 for masked keys
      const float corr = (m_run[r] == -INFINITY) ? 0.f : ptx::ex2(m_run[r] - mref);
      float lt = pj;
#pragma unroll
      for (int x = 16; x >= 1; x >>= 1) lt += __shfl_xor_sync(0xffffffffu, lt, x);
      l_run[r] = l_run[r] * corr + lt;
      m_run[r] = m_new;
#pragma unroll
      for (int c = 0; c < CPLN; ++c) acc[r][c] *= corr;
      // acc += p_j * V[j]; lanes own columns lane, lane+32, ...
      const int jn = min(kSimtTileN, p.Tk - k0);
      for (int j = 0; j < jn; ++j) {
        const float pb = __shfl_sync(0xffffffffu, pj, j);
        const float* vrow = sV + j * pitch;
#pragma unroll
        for (int c = 0; c < CPLN; ++c) {
          const int e = lane + 32 * c;
          if (e < hd) acc[r][c] = fmaf(pb, vrow[e], acc[r][c]);
        }
      }
    }
  }

  // ---- epilogue
#pragma unroll
  for (int r = 0; r < kSimtRPW; ++r) {
    const int row = row0 + warp * kSimtRPW + r;
    if (row >= total_rows) continue;
    const int tok = row / p.hpg, hh = row % p.hpg;
    bool qok = true;
    if (p.q_valid) qok = p.q_valid[(long long)n * p.Tq + tok] != 0;
    const float inv = (qok && l_run[r] > 0.f) ? 1.f / l_run[r] : 0.f;
    __nv_bfloat16* orow = p.o + n * p.os_n + (long long)tok * p.os_t + (long long)(g * p.hpg + hh) * p.os_h;
#pragma unroll
    for (int c = 0; c < CPLN; ++c) {
      const int e = lane + 32 * c;
      if (e < hd) orow[e] = __float2bfloat16(acc[r][c] * inv);
    }
  }
}

}  // namespace vats

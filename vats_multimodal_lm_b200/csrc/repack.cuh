// repack.cuh — copy [N, T, heads, hd] bf16 tensors whose rows are only 4-byte aligned (head_dim 60, 66, ... or strided
// views) into a TMA-addressable layout: head stride rounded up to 8 elements, token / sequence strides dense.
//
// The tcgen05 kernel can stage such tensors itself (cp.async variant), but its 96 loader threads issue one 4-byte
// copy per element pair and become the limit (cfg4a: 2.0 ms).  One streaming pass at HBM speed plus the TMA-fed kernel
// is faster (cfg4a: 0.20 + 0.43 ms with repack_chunk_kernel and the resident-K/V kernel), which is also what the
// drop-in modules do when they cast to bf16.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include "prefill_tc.cuh"   // tc_fastdiv

namespace vats {

struct RepackTensor {
  const __nv_bfloat16* src;
  __nv_bfloat16* dst;
  long long s_n, s_t, s_h;   // source strides (elements)
  int T, heads;
  long long rows;            // N * T * heads
  unsigned div_heads[2], div_T[2];   // magic-number division by heads / T (chunk kernel)
};

struct RepackParams {
  RepackTensor t[3];
  int hd2;      // head_dim / 2 (32-bit words per row)
  int hd_pad;   // destination head stride (elements, multiple of 8)
  unsigned div_cpr[2];   // magic-number division by hd_pad / 8 (16-byte chunks per destination row)
};

constexpr int kRepackWarps = 8;
constexpr int kRepackRows = 8;   // rows in flight per warp (loads of all of them are issued before the first store)

// One warp per group of kRepackRows consecutive rows (sequence, token, head); blockIdx.y selects the tensor.
// head_dim <= 128: at most two 32-bit words per lane and row.
__global__ void __launch_bounds__(kRepackWarps * 32) repack_kernel(const RepackParams p) {
  const RepackTensor& x = p.t[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const long long groups = (x.rows + kRepackRows - 1) / kRepackRows;
  const long long warps = (long long)gridDim.x * kRepackWarps;
  for (long long g = (long long)blockIdx.x * kRepackWarps + (threadIdx.x >> 5); g < groups; g += warps) {
    const long long row0 = g * kRepackRows;
    long long nt = row0 / x.heads;
    int h = (int)(row0 - nt * x.heads);
    long long n = nt / x.T;
    int t = (int)(nt - n * x.T);
    uint32_t v0[kRepackRows], v1[kRepackRows];
#pragma unroll
    for (int r = 0; r < kRepackRows; ++r) {
      v0[r] = 0u;
      v1[r] = 0u;
      if (row0 + r < x.rows) {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(x.src + n * x.s_n + (long long)t * x.s_t + (long long)h * x.s_h);
        if (lane < p.hd2) v0[r] = __ldg(s + lane);
        if (lane + 32 < p.hd2) v1[r] = __ldg(s + lane + 32);
      }
      if (++h == x.heads) {
        h = 0;
        if (++t == x.T) {
          t = 0;
          ++n;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRepackRows; ++r) {
      if (row0 + r < x.rows) {
        uint32_t* d = reinterpret_cast<uint32_t*>(x.dst + (row0 + r) * p.hd_pad);
        if (lane < p.hd2) d[lane] = v0[r];
        if (lane + 32 < p.hd2) d[lane + 32] = v1[r];
      }
    }
  }
}

// The same copy with one thread per 16-byte DESTINATION chunk (all tensors of the launch < 2^31 chunks): four 32-bit
// loads (source rows are only 4-byte aligned; neighbouring lanes read neighbouring words, so the sectors are shared
// in L1), one 128-bit store, the pad columns written as zeros.  The row-per-warp kernel above spends a warp
// instruction on the 33rd word of a 132-byte row and recomputes the (n, t, h) walk per row: ncu had it ALU-bound at
// 3.7 TB/s of combined traffic (profiles/r02s_ncu_full_repack_cfg4a_raw.csv, 0.38 ms on cfg4a); this one is a plain
// streaming copy.  kRepackChunkUnroll chunks per thread are loaded before the first store.
constexpr int kRepackChunkThreads = 256;
constexpr int kRepackChunkUnroll = 4;

__global__ void __launch_bounds__(kRepackChunkThreads) repack_chunk_kernel(const RepackParams p) {
  const RepackTensor& x = p.t[blockIdx.y];
  const unsigned cpr = (unsigned)p.hd_pad >> 3;
  const unsigned total = (unsigned)x.rows * cpr;
  const unsigned stride = gridDim.x * kRepackChunkThreads;
  const unsigned hd2 = (unsigned)p.hd2;
  for (unsigned c0 = blockIdx.x * kRepackChunkThreads + threadIdx.x; c0 < total; c0 += stride * kRepackChunkUnroll) {
    uint4 val[kRepackChunkUnroll];
#pragma unroll
    for (int u = 0; u < kRepackChunkUnroll; ++u) {
      const unsigned c = c0 + (unsigned)u * stride;
      val[u] = make_uint4(0u, 0u, 0u, 0u);
      if (c < total) {
        unsigned row, j, nt, h, n, t;
        tc_fastdiv(c, p.div_cpr, cpr, &row, &j);
        tc_fastdiv(row, x.div_heads, (unsigned)x.heads, &nt, &h);
        tc_fastdiv(nt, x.div_T, (unsigned)x.T, &n, &t);
        const uint32_t* s = reinterpret_cast<const uint32_t*>(x.src + (long long)n * x.s_n + (long long)t * x.s_t +
                                                              (long long)h * x.s_h) + 4u * j;
        const unsigned left = hd2 - 4u * j;   // words of this row from the chunk's first on (>= 1)
        val[u].x = __ldg(s);
        if (left > 1u) val[u].y = __ldg(s + 1);
        if (left > 2u) val[u].z = __ldg(s + 2);
        if (left > 3u) val[u].w = __ldg(s + 3);
      }
    }
#pragma unroll
    for (int u = 0; u < kRepackChunkUnroll; ++u) {
      const unsigned c = c0 + (unsigned)u * stride;
      if (c < total) reinterpret_cast<uint4*>(x.dst)[c] = val[u];
    }
  }
}

}  // namespace vats

// repack.cuh — copy [N, T, heads, hd] bf16 tensors whose rows are only 4-byte aligned (head_dim 60, 66, ... or strided
// views) into a TMA-addressable layout: head stride rounded up to 8 elements, token / sequence strides dense.
//
// The tcgen05 kernel can stage such tensors itself (cp.async variant), but its 96 loader threads issue one 4-byte
// copy per element pair and become the limit (cfg4a: 2.5 ms).  One streaming pass at HBM speed plus the TMA-fed kernel
// is faster (cfg4a: 0.25 + 1.1 ms), which is also what the drop-in modules do when they cast to bf16.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace vats {

struct RepackTensor {
  const __nv_bfloat16* src;
  __nv_bfloat16* dst;
  long long s_n, s_t, s_h;   // source strides (elements)
  int T, heads;
  long long rows;            // N * T * heads
};

struct RepackParams {
  RepackTensor t[3];
  int hd2;      // head_dim / 2 (32-bit words per row)
  int hd_pad;   // destination head stride (elements, multiple of 8)
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

}  // namespace vats

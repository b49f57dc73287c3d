// mask.cuh — the mask predicate and the KV tile-range arithmetic, shared by every kernel and by the host.
//
// Integer-only, __host__ __device__, so the very same code is exercised by the CPU test-suite
// (vats_attn_debug_tile_range) and by the kernels.  Predicate = SURVEY.md §8a-0:
//   causal / right:=0      reference src/optimized_attention.py:519-520, 632-634
//   (left,right) window    reference src/optimized_attention.py:634, vit_2d/optimized_attention.py:337,
//                          vit_3d/optimized_attention.py:162 (flash-attn semantics: keys in [i-left, i+right])
//   q_valid (query rows)   reference src/optimized_attention.py:673-675
//   k_valid (keys)         reference vit_3d/optimized_attention.py:276-277
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VATS_HD __host__ __device__ __forceinline__
#else
#define VATS_HD inline
#endif

namespace vats {

struct MaskParams {
  int Tq, Tk;
  int causal;       // 0/1
  int left, right;  // <0 = unlimited
};

// Lower / upper key bound (inclusive) for query row i, before clamping to [0,Tk) and before k_valid.
// Uses 64-bit to stay exact for any int32 inputs.
VATS_HD long long key_lo(const MaskParams& p, int i) {
  const long long off = (long long)p.Tk - (long long)p.Tq;
  return p.left < 0 ? 0LL : (long long)i + off - (long long)p.left;
}
VATS_HD long long key_hi(const MaskParams& p, int i) {
  const long long off = (long long)p.Tk - (long long)p.Tq;
  long long hi = (long long)p.Tk - 1;
  if (p.causal) {
    const long long c = (long long)i + off;
    hi = c < hi ? c : hi;
  }
  if (p.right >= 0) {
    const long long r = (long long)i + off + (long long)p.right;
    hi = r < hi ? r : hi;
  }
  return hi;
}

// Geometric part of the predicate (everything except the q_valid / k_valid byte masks).
VATS_HD bool allowed_geom(const MaskParams& p, int i, int j) {
  const long long off = (long long)p.Tk - (long long)p.Tq;
  const long long ii = (long long)i + off;
  if (p.causal && (long long)j > ii) return false;
  if (p.left >= 0 && (long long)j < ii - (long long)p.left) return false;
  if (p.right >= 0 && (long long)j > ii + (long long)p.right) return false;
  return true;
}

// Tile range for the query block [q0, q0+block_m) ∩ [0,Tq): first/last KV tile (inclusive) holding at least one
// key that some row of the block may attend (geometry only).  first > last ⇒ nothing to visit.
VATS_HD void tile_range(const MaskParams& p, int q0, int block_m, int block_n, int* first, int* last) {
  int q_last = q0 + block_m - 1;
  if (q_last > p.Tq - 1) q_last = p.Tq - 1;
  if (q_last < q0 || p.Tk <= 0) {
    *first = 0;
    *last = -1;
    return;
  }
  long long lo = key_lo(p, q0);      // smallest lower bound is at the first row
  long long hi = key_hi(p, q_last);  // largest upper bound is at the last row
  if (lo < 0) lo = 0;
  if (hi > (long long)p.Tk - 1) hi = (long long)p.Tk - 1;
  if (hi < lo) {
    *first = 0;
    *last = -1;
    return;
  }
  // lo, hi are inside [0, Tk) here: 32-bit division (a shift when block_n is a compile-time power of two; a 64-bit
  // division costs a GPU thread hundreds of cycles)
  *first = (int)lo / block_n;
  *last = (int)hi / block_n;
}

// True when every (row, key) of the block × tile rectangle is allowed by the geometry and lies inside [0,Tk):
// such tiles skip the per-element predicate (k_valid, if present, is still applied by the caller).
VATS_HD bool tile_is_full(const MaskParams& p, int tile, int q0, int block_m, int block_n) {
  int q_last = q0 + block_m - 1;
  if (q_last > p.Tq - 1) q_last = p.Tq - 1;
  const long long k_first = (long long)tile * block_n;
  const long long k_last = k_first + block_n - 1;
  if (k_last > (long long)p.Tk - 1) return false;
  // every row must allow k_first..k_last: the tightest lower bound is at q_last, the tightest upper at q0
  if (key_lo(p, q_last) > k_first) return false;
  if (key_hi(p, q0) < k_last) return false;
  return true;
}

}  // namespace vats

// prefill_mid.cuh — GQA attention for sequences of 33..256 keys (ViT-2D / ViT-3D spatial passes, text encoders,
// cross-attention contexts) on tcgen05 / TMEM / TMA.
//
// Same contract as prefill_tc.cuh (reference vit_2d/optimized_attention.py:348-423, vit_3d/optimized_attention.py:185-348,
// src/optimized_attention.py:657-723 for short prompts); a different shape of the work: with at most 256 keys the whole
// K / V of a (sequence, KV group) is ONE tile, so
//   * K and V of the group are fetched once and stay in shared memory for ALL of the group's query heads and all of its
//     query blocks (prefill_tc re-fetches them per head pair and per 128-token block);
//   * S = Q.K^T is a single UMMA with N = keys rounded up to 16 (<= 256) and the softmax is single-pass and exact — no
//     online rescaling, no running maximum, one S -> P round trip per tile;
//   * the 128 rows of a tile are (token, head) pairs: the H/G query heads of the group are packed into the M dimension
//     (TMA box {64, heads, 128/heads}), so 196 tokens x 4 heads are 7 tiles instead of 8 and the tile's rows are
//     contiguous in q and o;
//   * two tile slots ping-pong: while one warpgroup runs its exponentials (the MUFU pipe is the co-critical resource at
//     these shapes: 128 x 208 ex2 per tile = 1 664 cycles per SM sub-partition) the other slot's MMAs, TMEM traffic
//     and write-out proceed.
//
// Persistent CTAs (one per SM) walk the (sequence, KV group) items round-robin.
//   warps 0-3 / 4-7  softmax warpgroup of slot 0 / 1: thread r owns tile row r (TMEM lane r)
//   warps 8, 9, 10   TMA producers (one lane each): K ring, V ring, Q tiles
//                    (kLdg instantiation: 96 loader threads stage K, Q, V with cp.async — rows TMA cannot address,
//                    e.g. dense head_dim 66: 132-byte rows)
//   warp 11          MMA issuer: S(f) = Q.K^T (SS), O(f) = P.V (TS, P from TMEM), in the order S(0) S(1) PV(0) S(2) PV(1) ...
// TMEM: slot s owns columns [256 s, 256 s + 256): S in [0, n_pad); P (bf16) is written over S[0, n_pad/2) chunk by
// chunk behind the softmax's second read pass; O accumulates in [o_off, o_off + hd_pad) with o_off = ceil16(n_pad/2),
// i.e. inside the (by then consumed) upper half of S.
#pragma once
#include "mask.cuh"
#include "prefill_tc.cuh"  // PrefillParams, tc_fastdiv, cp.async helpers, setmaxnreg, kTcOStageBytes

namespace vats {

constexpr int kMidThreads = 384;
constexpr int kMidLoaderThreads = 96;
constexpr int kMidMaxKv = 4;        // deepest K / V ring
constexpr int kMidSlotCols = 256;   // TMEM columns per tile slot
constexpr int kMidQRegionBytes = 128 * 128;

struct MidParams {
  PrefillParams a;
  int hd_pad;          // head dim rounded up to 16 (MMA K of S, MMA N of P.V)
  int regions;         // ceil(hd_pad / 64) 128-byte swizzle regions per row
  int n_pad;           // keys rounded up to 16 (MMA N of S, K extent of P.V), <= 256
  int o_off;           // TMEM column of O inside a slot
  int pack, pack_shift;  // query heads packed into one tile (power of two <= 32), its log2
  int tok_per_tile;    // 128 >> pack_shift
  int q_tiles;         // ceil(Tq / tok_per_tile)
  int head_sets;       // hpg / pack
  int tiles_per_item;  // q_tiles * head_sets
  int num_items;       // N * G
  int nkv;             // K / V ring depth
  int ldg_vec;         // kLdg: 32-bit words per cp.async copy (1 or 2)
  int o_stage;         // 1 = TMA tile stores, 2 = coalesced 32-bit stores from the staging tile
  int simple_mask;     // 1: no band, no q_valid / k_valid — only the columns >= Tk are masked
  unsigned div_g[2], div_qt[2];
};

struct MidBarriers {
  uint64_t q_full[2], q_empty[2], s_full[2], p_full[2], o_full[2], o_empty[2];
  uint64_t k_full[kMidMaxKv], k_empty[kMidMaxKv], v_full[kMidMaxKv], v_empty[kMidMaxKv];
  uint32_t kbits[8][8];   // per softmax warp: k_valid of the current item as bit words (generic-mask path)
  uint32_t tmem_base;
  uint32_t pad;
};

__host__ __device__ inline size_t mid_smem_bytes(int regions, int n_pad, int nkv) {
  return (size_t)2 * regions * kMidQRegionBytes + (size_t)2 * nkv * regions * n_pad * 128 + 8 * kTcOStageBytes + 1024 +
         sizeof(MidBarriers);
}

struct MidCursor {
  int item, tt;   // global item index (sequence x KV group), tile within the item
};
struct MidTile {
  int n, g, q0, head0;
};
__device__ __forceinline__ void mid_advance(const MidParams& P, MidCursor& c, int steps) {
  c.tt += steps;
  while (c.tt >= P.tiles_per_item) {
    c.tt -= P.tiles_per_item;
    c.item += (int)gridDim.x;
  }
}
__device__ __forceinline__ MidTile mid_decode(const MidParams& P, const MidCursor& c) {
  MidTile t;
  unsigned n, g, hs, qt;
  tc_fastdiv((unsigned)c.item, P.div_g, (unsigned)P.a.G, &n, &g);
  tc_fastdiv((unsigned)c.tt, P.div_qt, (unsigned)P.q_tiles, &hs, &qt);
  t.n = (int)n;
  t.g = (int)g;
  t.q0 = (int)qt * P.tok_per_tile;
  t.head0 = (int)g * P.a.hpg + (int)hs * P.pack;
  return t;
}

// 32-bit word of allowed columns [base, base + 32) for a row whose allowed keys are [lo, hi]
__device__ __forceinline__ uint32_t mid_range_word(int lo, int hi, int base) {
  const int l = lo - base, h = hi - base;
  const uint32_t ml = l <= 0 ? 0xffffffffu : (l >= 32 ? 0u : 0xffffffffu << l);
  const uint32_t mh = h >= 31 ? 0xffffffffu : (h < 0 ? 0u : 0xffffffffu >> (31 - h));
  return ml & mh;
}

__device__ __forceinline__ float mid_max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// NC columns of S (NC = 64, 32 or 16) starting at column c0 of this thread's row -> registers
template <int NC>
__device__ __forceinline__ void mid_ld_cols(uint32_t taddr, uint32_t* v) {
  if (NC == 64) {
    ptx::tmem_ld_32x32b_x32(taddr, v);
    ptx::tmem_ld_32x32b_x32(taddr + 32, v + 32);
  } else if (NC == 32) {
    ptx::tmem_ld_32x32b_x32(taddr, v);
  } else {
    ptx::tmem_ld_32x32b_x16(taddr, v);
  }
  ptx::tmem_ld_wait();
}

// Mask the NC values of one row in place (-inf where the key is not allowed).
template <int NC>
__device__ __forceinline__ void mid_apply_mask(uint32_t* v, int c0, int lo, int hi, const uint32_t* kb, bool use_kb) {
#pragma unroll
  for (int w = 0; w < (NC + 31) / 32; ++w) {
    uint32_t bits = mid_range_word(lo, hi, c0 + 32 * w);
    if (use_kb) bits &= kb[(c0 >> 5) + w] >> (c0 & 16);   // (c0 & 16 != 0 only for the 16-column tail block)
    if (NC == 16) bits |= 0xffff0000u;
    if (bits != 0xffffffffu) {
#pragma unroll
      for (int i = 0; i < (NC < 32 ? NC : 32); ++i)
        if (!((bits >> i) & 1u)) v[32 * w + i] = 0xff800000u;
    }
  }
}

// pass 1: running row maximum (raw logits) over one block of columns
template <int NC>
__device__ __forceinline__ float mid_pass_max(uint32_t tS, int c0, int lo, int hi, const uint32_t* kb, bool use_kb,
                                              float m) {
  uint32_t v[NC];
  mid_ld_cols<NC>(tS + (uint32_t)c0, v);
  mid_apply_mask<NC>(v, c0, lo, hi, kb, use_kb);
  float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < NC; i += 8) {
    m0 = mid_max3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    m1 = mid_max3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    m2 = mid_max3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
    m3 = mid_max3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// pass 2: p = exp2(s * scale_log2 - m), row sum, bf16 pack, P written over S[c0/2, c0/2 + NC/2)
template <int NC>
__device__ __forceinline__ float mid_pass_exp(uint32_t tS, int c0, int lo, int hi, const uint32_t* kb, bool use_kb,
                                              float scale_log2, float neg_m) {
  using namespace ptx;
  uint32_t v[NC];
  mid_ld_cols<NC>(tS + (uint32_t)c0, v);
  mid_apply_mask<NC>(v, c0, lo, hi, kb, use_kb);
  const float2 sc2 = make_float2(scale_log2, scale_log2);
  const float2 nm2 = make_float2(neg_m, neg_m);
  float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
  uint32_t pk[NC / 2];
#pragma unroll
  for (int c = 0; c < NC; c += 4) {
    const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), sc2, nm2);
    const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])), sc2, nm2);
    const float2 p0 = make_float2(ex2(x0.x), ex2(x0.y));
    const float2 p1 = make_float2(ex2(x1.x), ex2(x1.y));
    sum_a = __fadd2_rn(sum_a, p0);
    sum_b = __fadd2_rn(sum_b, p1);
    pk[c >> 1] = pack_bf16x2(p0.x, p0.y);
    pk[(c >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
  }
  const uint32_t tP = tS + (uint32_t)(c0 >> 1);
  if (NC == 64) {
    tmem_st_32x32b_x32(tP, pk);
  } else if (NC == 32) {
    tmem_st_32x32b_x16(tP, pk);
  } else {
    tmem_st_32x32b_x8(tP, pk);
  }
  return (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
}

// Stage `rows_pad` rows x hd of bf16 into 128B-swizzled regions (the layout a SWIZZLE_128B TMA box {64, ..., rows}
// produces) with asynchronous 4- / 8-byte copies.  RowPtr(r) gives the global address of row r or nullptr for a row
// that must be zero-filled; columns [hd, hd_pad) are zero-filled too.
template <int NT, int VW, typename RowPtr>
__device__ __forceinline__ void mid_cpasync_rows(uint32_t dst, RowPtr row_ptr, const __nv_bfloat16* any_valid,
                                                 int rows_pad, int hd, int hd_pad, uint32_t region_bytes, int tid) {
  const int upr = hd_pad / (2 * VW);
  const int hu = hd / (2 * VW);
  const int dq = NT / upr, dr = NT % upr;
  int r = tid / upr, u = tid - r * upr;
  while (r < rows_pad) {
    const __nv_bfloat16* row = row_ptr(r);
    const bool ok = row != nullptr && u < hu;
    const __nv_bfloat16* g = ok ? row + 2 * VW * u : any_valid;
    const uint32_t w = (uint32_t)(u * VW);
    const uint32_t wi = w & 31u;
    const uint32_t off = (w >> 5) * region_bytes + (uint32_t)r * 128u + (((wi >> 2) ^ ((uint32_t)r & 7u)) << 4) + ((wi & 3u) << 2);
    const uint32_t nbytes = ok ? 4u * VW : 0u;
    if (VW == 2)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + off), "l"(g), "r"(nbytes) : "memory");
    else
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + off), "l"(g), "r"(nbytes) : "memory");
    u += dr;
    r += dq;
    if (u >= upr) {
      u -= upr;
      ++r;
    }
  }
}

template <bool kLdg>
__global__ void __launch_bounds__(kMidThreads, 1)
prefill_mid_kernel(const MidParams P, const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                   const __grid_constant__ CUtensorMap tmap_o) {
  using namespace ptx;
  extern __shared__ unsigned char smem_raw[];
  const PrefillParams& a = P.a;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t q_tile = (uint32_t)P.regions * kMidQRegionBytes;
  const uint32_t kv_region = (uint32_t)P.n_pad * 128u;
  const uint32_t kv_tile = (uint32_t)P.regions * kv_region;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + 2 * q_tile;
  const uint32_t sV = sK + (uint32_t)P.nkv * kv_tile;
  const uint32_t sO = sV + (uint32_t)P.nkv * kv_tile;
  MidBarriers* bars = reinterpret_cast<MidBarriers*>(smem_raw + (base - raw) + (size_t)2 * q_tile +
                                                     (size_t)2 * P.nkv * kv_tile + 8 * kTcOStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    if (!kLdg) {
      prefetch_tmap(&tmap_q);
      prefetch_tmap(&tmap_k);
      prefetch_tmap(&tmap_v);
    }
    if (P.o_stage == 1) prefetch_tmap(&tmap_o);
    const uint32_t load_arrivals = kLdg ? kMidLoaderThreads : 1;
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->q_full[s]), load_arrivals);
      mbar_init(smem_u32(&bars->q_empty[s]), 1);
      mbar_init(smem_u32(&bars->s_full[s]), 1);
      mbar_init(smem_u32(&bars->p_full[s]), 128);
      mbar_init(smem_u32(&bars->o_full[s]), 1);
      mbar_init(smem_u32(&bars->o_empty[s]), 128);
    }
    for (int s = 0; s < kMidMaxKv; ++s) {
      mbar_init(smem_u32(&bars->k_full[s]), load_arrivals);
      mbar_init(smem_u32(&bars->k_empty[s]), 1);
      mbar_init(smem_u32(&bars->v_full[s]), load_arrivals);
      mbar_init(smem_u32(&bars->v_empty[s]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 11) {
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp >= 8) {
    setmaxnreg_dec<80>();
    if (warp <= 10) {
      if (!kLdg) {
        // ------------------------------------------------------------------ TMA producers
        if (lane == 0) {
          if (warp == 10) {
            MidCursor c{(int)blockIdx.x, 0};
            for (uint32_t f = 0; c.item < P.num_items; ++f) {
              const MidTile t = mid_decode(P, c);
              const uint32_t s = f & 1u, u = f >> 1;
              mbar_wait(smem_u32(&bars->q_empty[s]), (u & 1u) ^ 1u, 0x300u);
              const uint32_t bar = smem_u32(&bars->q_full[s]);
              mbar_expect_tx(bar, q_tile);
              for (int r = 0; r < P.regions; ++r)
                tma_load_4d(sQ + s * q_tile + (uint32_t)r * kMidQRegionBytes, &tmap_q, bar, 64 * r, t.head0, t.q0, t.n);
              mid_advance(P, c, 1);
            }
          } else {
            const bool is_k = warp == 8;
            const CUtensorMap* tm = is_k ? &tmap_k : &tmap_v;
            uint64_t* full = is_k ? bars->k_full : bars->v_full;
            uint64_t* empty = is_k ? bars->k_empty : bars->v_empty;
            const uint32_t ring = is_k ? sK : sV;
            int slot = 0;
            uint32_t ph = 0u;
            for (int item = (int)blockIdx.x; item < P.num_items; item += (int)gridDim.x) {
              unsigned n, g;
              tc_fastdiv((unsigned)item, P.div_g, (unsigned)a.G, &n, &g);
              mbar_wait(smem_u32(&empty[slot]), ph ^ 1u, 0x310u);
              const uint32_t bar = smem_u32(&full[slot]);
              mbar_expect_tx(bar, kv_tile);
              for (int r = 0; r < P.regions; ++r)
                tma_load_4d(ring + (uint32_t)slot * kv_tile + (uint32_t)r * kv_region, tm, bar, 64 * r, (int)g, 0, (int)n);
              if (++slot == P.nkv) { slot = 0; ph ^= 1u; }
            }
          }
        }
      } else {
        // ------------------------------------------------------------------ cp.async loaders (96 threads), in the
        // order the MMA warp consumes: K, Q(0), Q(1), V, Q(2), ... per item.  A tile is signalled only after the
        // copies of the next one were issued, so two tiles are in flight per thread.
        const int ltid = (int)threadIdx.x - 8 * 32;
        uint32_t pending = 0u;
        // A tile's arrival is deferred until the next tile's copies are in flight — but never across a wait that may
        // depend on it (with a one-deep K ring the next item's k_empty needs S of the tile still pending here).
        auto wait_flush = [&](uint32_t bar, uint32_t parity, uint32_t tag) {
          if (mbar_try_wait(bar, parity)) return;
          if (pending != 0u) {
            cpasync_wait<0>();
            fence_proxy_async_smem();
            mbar_arrive(pending);
            pending = 0u;
          }
          mbar_wait(bar, parity, tag);
        };
        auto finish = [&](uint32_t full_bar) {
          cpasync_commit();
          if (pending != 0u) {
            cpasync_wait<1>();
            fence_proxy_async_smem();
            mbar_arrive(pending);
          }
          pending = full_bar;
        };
        auto stage_kv = [&](uint32_t dst, const __nv_bfloat16* src, long long stride_t) {
          auto rp = [&](int r) -> const __nv_bfloat16* { return r < a.Tk ? src + (long long)r * stride_t : nullptr; };
          if (P.ldg_vec == 2)
            mid_cpasync_rows<kMidLoaderThreads, 2>(dst, rp, src, P.n_pad, a.hd, P.hd_pad, kv_region, ltid);
          else
            mid_cpasync_rows<kMidLoaderThreads, 1>(dst, rp, src, P.n_pad, a.hd, P.hd_pad, kv_region, ltid);
        };
        int ks = 0, vs = 0;
        uint32_t kph = 0u, vph = 0u, f = 0u;
        for (int item = (int)blockIdx.x; item < P.num_items; item += (int)gridDim.x) {
          unsigned n, g;
          tc_fastdiv((unsigned)item, P.div_g, (unsigned)a.G, &n, &g);
          MidCursor c{item, 0};
          for (int tt = 0; tt < P.tiles_per_item; ++tt, ++f) {
            if (tt == 0) {
              wait_flush(smem_u32(&bars->k_empty[ks]), kph ^ 1u, 0x320u);
              stage_kv(sK + (uint32_t)ks * kv_tile, a.k + (long long)n * a.ks_n + (long long)g * a.ks_h, a.ks_t);
              finish(smem_u32(&bars->k_full[ks]));
              if (++ks == P.nkv) { ks = 0; kph ^= 1u; }
            }
            {
              c.tt = tt;
              const MidTile t = mid_decode(P, c);
              const uint32_t s = f & 1u, u = f >> 1;
              wait_flush(smem_u32(&bars->q_empty[s]), (u & 1u) ^ 1u, 0x321u);
              const __nv_bfloat16* qb = a.q + (long long)t.n * a.qs_n;
              auto rp = [&](int r) -> const __nv_bfloat16* {
                const int tok = t.q0 + (r >> P.pack_shift);
                return tok < a.Tq ? qb + (long long)tok * a.qs_t + (long long)(t.head0 + (r & (P.pack - 1))) * a.qs_h
                                  : nullptr;
              };
              if (P.ldg_vec == 2)
                mid_cpasync_rows<kMidLoaderThreads, 2>(sQ + s * q_tile, rp, a.q, 128, a.hd, P.hd_pad, kMidQRegionBytes, ltid);
              else
                mid_cpasync_rows<kMidLoaderThreads, 1>(sQ + s * q_tile, rp, a.q, 128, a.hd, P.hd_pad, kMidQRegionBytes, ltid);
              finish(smem_u32(&bars->q_full[s]));
            }
            // V is first needed by P.V(0), one softmax after S(0): it goes behind the item's first two Q tiles
            if (tt == (P.tiles_per_item > 1 ? 1 : 0)) {
              wait_flush(smem_u32(&bars->v_empty[vs]), vph ^ 1u, 0x322u);
              stage_kv(sV + (uint32_t)vs * kv_tile, a.v + (long long)n * a.vs_n + (long long)g * a.vs_h, a.vs_t);
              finish(smem_u32(&bars->v_full[vs]));
              if (++vs == P.nkv) { vs = 0; vph ^= 1u; }
            }
          }
        }
        if (pending != 0u) {
          cpasync_wait<0>();
          fence_proxy_async_smem();
          mbar_arrive(pending);
        }
      }
    } else {
      // -------------------------------------------------------------------- MMA issuer (all lanes convergent)
      const uint32_t leader = elect_one() ? 1u : 0u;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc_s = make_idesc_bf16(128, P.n_pad, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, P.hd_pad, 0, 1);
      const int ksteps_s = P.hd_pad / 16;
      const int ksteps_o = P.n_pad / 16;
      const uint32_t hi_sw = smem_desc_hi_sw128(1024);
      const uint32_t q_lo[2] = {smem_desc_lo(sQ, 16), smem_desc_lo(sQ + q_tile, 16)};
      const uint32_t k_lo_base = smem_desc_lo(sK, 16);
      const uint32_t v_lo_base = smem_desc_lo(sV, kv_region);
      const uint32_t kv_step = kv_tile >> 4;
      const uint32_t rq = kMidQRegionBytes >> 4, rk = kv_region >> 4;

      const int tpi = P.tiles_per_item;
      const int my_items = (P.num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int total = my_items * tpi;
      int s_tt = 0, ks = 0, pv_tt = 0, vs = 0;
      uint32_t kph = 0u, vph = 0u;
      for (int f = 0; f < total + 2; ++f) {
        if (f >= 2) {
          // ---- O(f-2) = P.V
          const uint32_t s = (uint32_t)f & 1u, u = (uint32_t)(f - 2) >> 1;
          mbar_wait(smem_u32(&bars->p_full[s]), u & 1u, 0x100u);
          if (pv_tt == 0) mbar_wait(smem_u32(&bars->v_full[vs]), vph, 0x101u);
          tc_fence_after();
          const uint32_t tP = tmem_u + s * kMidSlotCols;
          const uint32_t v_lo = v_lo_base + (uint32_t)vs * kv_step;
          uint32_t acc = 0u;
          for (int k = 0; k < ksteps_o; ++k) {
            mma_ts_lohi(tP + (uint32_t)P.o_off, tP + (uint32_t)k * 8u, v_lo + (uint32_t)k * (2048u >> 4), hi_sw, idesc_o, acc,
                        leader);
            acc = 1u;
          }
          tc_commit_pred(smem_u32(&bars->o_full[s]), leader);
          if (++pv_tt == tpi) {
            pv_tt = 0;
            tc_commit_pred(smem_u32(&bars->v_empty[vs]), leader);
            if (++vs == P.nkv) { vs = 0; vph ^= 1u; }
          }
        }
        if (f < total) {
          // ---- S(f) = Q.K^T
          const uint32_t s = (uint32_t)f & 1u, u = (uint32_t)f >> 1;
          mbar_wait(smem_u32(&bars->q_full[s]), u & 1u, 0x110u);
          if (s_tt == 0) mbar_wait(smem_u32(&bars->k_full[ks]), kph, 0x111u);
          if (u > 0) mbar_wait(smem_u32(&bars->o_empty[s]), (u - 1u) & 1u, 0x112u);   // the slot's previous O was read out
          tc_fence_after();
          uint32_t ql = q_lo[s], kl = k_lo_base + (uint32_t)ks * kv_step;
          uint32_t acc = 0u;
          for (int k = 0; k < ksteps_s; ++k) {
            mma_ss_lohi(tmem_u + s * kMidSlotCols, ql, hi_sw, kl, hi_sw, idesc_s, acc, leader);
            acc = 1u;
            if ((k & 3) == 3) {
              ql += rq - 6u;
              kl += rk - 6u;
            } else {
              ql += 2u;
              kl += 2u;
            }
          }
          tc_commit_pred(smem_u32(&bars->s_full[s]), leader);
          tc_commit_pred(smem_u32(&bars->q_empty[s]), leader);
          if (++s_tt == tpi) {
            s_tt = 0;
            tc_commit_pred(smem_u32(&bars->k_empty[ks]), leader);
            if (++ks == P.nkv) { ks = 0; kph ^= 1u; }
          }
        }
      }
    }
  } else {
    // ====================================================================== softmax warpgroups (slot = warp / 4)
    setmaxnreg_inc<208>();
    const uint32_t slot = (uint32_t)warp >> 2;
    const int wq = warp & 3;
    const int r = (int)threadIdx.x & 127;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t tS = tmem + lane_base + slot * kMidSlotCols;
    const uint32_t tO = tS + (uint32_t)P.o_off;
    const uint32_t stage = sO + (uint32_t)warp * kTcOStageBytes;
    uint32_t* kb = bars->kbits[warp];
    const bool use_kb = a.k_valid != nullptr;

    MidCursor c{(int)blockIdx.x, 0};
    mid_advance(P, c, (int)slot);
    for (uint32_t u = 0; c.item < P.num_items; ++u, mid_advance(P, c, 2)) {
      const MidTile t = mid_decode(P, c);
      const int tok = t.q0 + (r >> P.pack_shift);
      const int head = t.head0 + (r & (P.pack - 1));
      const int tok_w = t.q0 + ((wq * 32) >> P.pack_shift);   // first token of this warp's 32 rows
      const bool warp_live = tok_w < a.Tq;

      int lo = 0, hi = a.Tk - 1;
      if (!P.simple_mask && tok < a.Tq) {
        const long long l = key_lo(a.mask, tok), h = key_hi(a.mask, tok);
        lo = l < 0 ? 0 : (l > 256 ? 256 : (int)l);
        hi = h < -1 ? -1 : (int)h;   // key_hi is already <= Tk - 1
      }
      if (use_kb && warp_live) {
        for (int w = 0; w < (P.n_pad + 31) / 32; ++w) {
          const int key = w * 32 + lane;
          const bool ok = key < a.Tk && a.k_valid[(long long)t.n * a.Tk + key] != 0;
          const uint32_t bits = __ballot_sync(0xffffffffu, ok);
          if (lane == 0) kb[w] = bits;
        }
        __syncwarp();
      }

      mbar_wait(smem_u32(&bars->s_full[slot]), u & 1u, 0x200u);
      tc_fence_after();
      float l_sum = 0.f;
      if (warp_live) {
        // ---- pass 1: exact row maximum
        float m = -INFINITY;
        int c0 = 0;
        for (; c0 + 64 <= P.n_pad; c0 += 64) m = mid_pass_max<64>(tS, c0, lo, hi, kb, use_kb, m);
        if (P.n_pad & 32) {
          m = mid_pass_max<32>(tS, c0, lo, hi, kb, use_kb, m);
          c0 += 32;
        }
        if (P.n_pad & 16) m = mid_pass_max<16>(tS, c0, lo, hi, kb, use_kb, m);
        // ---- pass 2: exponentials, row sum, P over S
        const float neg_m = (m == -INFINITY) ? 0.f : -m * a.scale_log2;
        c0 = 0;
        for (; c0 + 64 <= P.n_pad; c0 += 64) l_sum += mid_pass_exp<64>(tS, c0, lo, hi, kb, use_kb, a.scale_log2, neg_m);
        if (P.n_pad & 32) {
          l_sum += mid_pass_exp<32>(tS, c0, lo, hi, kb, use_kb, a.scale_log2, neg_m);
          c0 += 32;
        }
        if (P.n_pad & 16) l_sum += mid_pass_exp<16>(tS, c0, lo, hi, kb, use_kb, a.scale_log2, neg_m);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->p_full[slot]));

      // ---- epilogue: O / l -> bf16 -> staging tile -> global
      mbar_wait(smem_u32(&bars->o_full[slot]), u & 1u, 0x210u);
      tc_fence_after();
      uint32_t acc[128];
      if (warp_live) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i * 16 < P.hd_pad) tmem_ld_32x32b_x16(tO + i * 16, acc + i * 16);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->o_empty[slot]));
      if (warp_live) {
        bool qok = tok < a.Tq;
        if (qok && a.q_valid != nullptr) qok = a.q_valid[(long long)t.n * a.Tq + tok] != 0;
        const float inv = (qok && l_sum > 0.f) ? 1.f / l_sum : 0.f;
#pragma unroll
        for (int cbi = 0; cbi < 2; ++cbi) {
          const int cb = cbi * 64;
          if (cb >= P.hd_pad) continue;
          if (P.o_stage == 1 && lane == 0) bulk_wait_group_read0();
          __syncwarp();
#pragma unroll
          for (int uu = 0; uu < 8; ++uu) {
            if (cb + uu * 8 < P.hd_pad) {
              const uint32_t x = pack_bf16x2(__uint_as_float(acc[cb + 8 * uu + 0]) * inv, __uint_as_float(acc[cb + 8 * uu + 1]) * inv);
              const uint32_t y = pack_bf16x2(__uint_as_float(acc[cb + 8 * uu + 2]) * inv, __uint_as_float(acc[cb + 8 * uu + 3]) * inv);
              const uint32_t z = pack_bf16x2(__uint_as_float(acc[cb + 8 * uu + 4]) * inv, __uint_as_float(acc[cb + 8 * uu + 5]) * inv);
              const uint32_t w = pack_bf16x2(__uint_as_float(acc[cb + 8 * uu + 6]) * inv, __uint_as_float(acc[cb + 8 * uu + 7]) * inv);
              const uint32_t dst = stage + (uint32_t)lane * 128u + (((uint32_t)uu ^ ((uint32_t)lane & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
            }
          }
          if (P.o_stage == 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmap_o, stage, cb, t.head0, tok_w, t.n);
              bulk_commit_group();
            }
          } else {
            __syncwarp();
            const int wv = (a.hd - cb >= 64 ? 64 : a.hd - cb) >> 1;   // valid 32-bit words per row in this chunk
            if (wv > 0) {
              const int dq = 32 / wv, dr = 32 % wv;
              int row = lane / wv, w = lane - row * wv;
              while (row < 32) {
                const int rt = tok_w + (row >> P.pack_shift);
                if (rt < a.Tq) {
                  const uint32_t src = stage + (uint32_t)row * 128u + ((((uint32_t)w >> 2) ^ ((uint32_t)row & 7u)) << 4) +
                                       (((uint32_t)w & 3u) << 2);
                  uint32_t val;
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(val) : "r"(src) : "memory");
                  uint32_t* orow = reinterpret_cast<uint32_t*>(a.o + (long long)t.n * a.os_n + (long long)rt * a.os_t +
                                                               (long long)(t.head0 + (row & (P.pack - 1))) * a.os_h + cb);
                  orow[w] = val;
                }
                w += dr;
                row += dq;
                if (w >= wv) {
                  w -= wv;
                  ++row;
                }
              }
            }
            __syncwarp();
          }
        }
      }
      (void)head;
    }
  }

  // ---- teardown
  if (P.o_stage == 1 && warp < 8 && lane == 0) bulk_wait_group0();
  tc_fence_before();
  __syncthreads();
  if (warp == 11) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace vats

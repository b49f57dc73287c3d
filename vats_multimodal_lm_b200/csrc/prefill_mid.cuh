// prefill_mid.cuh — GQA attention for sequences of 33..256 keys (ViT-2D / ViT-3D spatial passes, text encoders,
// cross-attention contexts) on tcgen05 / TMEM / TMA.
//
// Same contract as prefill_tc.cuh (reference vit_2d/optimized_attention.py:348-423, vit_3d/optimized_attention.py:185-348,
// src/optimized_attention.py:657-723 for short prompts); a different shape of the work: with at most 256 keys the whole
// K / V of a (sequence, KV group) is ONE tile, so
//   * K and V of the group are fetched once and stay in shared memory for ALL of the group's query heads and all of its
//     query blocks (prefill_tc re-fetches them per head pair and per 128-token block);
//   * S = Q.K^T is a single UMMA with N = keys rounded up to 16 (<= 256) and the softmax is exact over the whole row — no
//     online rescaling, no running maximum, one S -> P round trip per tile;
//   * the 128 rows of a tile are (token, head) pairs: the H/G query heads of the group are packed into the M dimension
//     (TMA box {64, heads, 128/heads}), so 196 tokens x 4 heads are 7 tiles instead of 8 and the tile's rows are
//     contiguous in q and o;
//   * the two softmax warpgroups each own ONE of the two TMEM tile slots (every second tile) and run half a tile apart:
//     a thread streams its row through registers 32 columns at a time (double-buffered tcgen05.ld — a 128-column read
//     costs ~50 clk, tools/micro/tmem_bw.cu), one pass for the row maximum and one for the exponentials (the first
//     is skipped under a logit bound), and the dead time of one warpgroup (tile bookkeeping, barrier round trips,
//     TMEM latency) hides behind the exponentials of the other — the MUFU pipe is the resource they share
//     (128 x 208 ex2 per tile = 1664 clk per SM sub-partition);
//   * the write-out has its own warpgroup, so P.V, the TMEM read of O and the TMA store of tile f overlap the
//     softmax of tiles f+1, f+2;
// History, measured on B200 (cfg3): (1) one warpgroup per slot doing softmax AND its own epilogue: 0.218 ms; (2) a row
// split between two threads of two warpgroups working on the SAME tile, logits held in 128 registers, row maximum
// exchanged through shared memory: 0.173 ms — both warpgroups stalled at the same points of every tile (and a
// mbarrier.try_wait used as a poll suspended them for >1000 clk per tile: try_wait blocks, test_wait does not).
//
// Persistent CTAs (one per SM, 16 warps) walk the (sequence, KV group) items round-robin.
//   warps 0-3 / 4-7   softmax warpgroups: warpgroup h owns TMEM slot h (tiles f = h, h+2, ...), thread r tile row r
//   warps 8-11        epilogue warpgroup: O from TMEM, 1/l, bf16, staging tile, TMA store (or coalesced stores)
//   warps 12, 13, 14  TMA producers (one lane each): K ring, V ring, Q tiles
//                     (kLdg instantiation: 96 loader threads stage K, Q, V with cp.async — rows TMA cannot address,
//                     e.g. dense head_dim 66: 132-byte rows)
//   warp 15           MMA issuer: S(f) = Q.K^T (SS), O(f) = P.V (TS, P from TMEM), in the order S(0) S(1) PV(0) S(2) PV(1) ...
// TMEM: slot s holds S in n_pad columns; P (bf16) is written over S[0, n_pad/2) behind the reads.  When
// 2 * n_pad + hd_pad <= 512 (ViT shapes: 2 * 208 + 80) there is ONE O accumulator behind the two slots: S(f+2) then
// only waits for P.V(f) to have consumed P(f) — in-order on the tensor pipe — and not for the epilogue to drain O(f),
// which takes the epilogue out of the S -> softmax -> P.V -> S loop.  Otherwise
// (256 keys x hd 128) O accumulates inside the slot, in the consumed upper half of S, at o_off = ceil16(n_pad/2).
#pragma once
#include "mask.cuh"
#include "prefill_tc.cuh"  // PrefillParams, tc_fastdiv, cp.async helpers, setmaxnreg, kTcOStageBytes

namespace vats {

constexpr int kMidThreads = 512;
constexpr int kMidLoaderThreads = 96;
constexpr int kMidMaxKv = 4;        // deepest K / V ring
constexpr int kMidSlotCols = 256;   // TMEM columns per tile slot
constexpr int kMidQRegionBytes = 128 * 128;

struct MidParams {
  PrefillParams a;
  int hd_pad;          // head dim rounded up to 16 (MMA K of S, MMA N of P.V)
  int regions;         // ceil(hd_pad / 64) 128-byte swizzle regions per row
  int n_pad;           // keys rounded up to 16 (MMA N of S, K extent of P.V), <= 256
  int slot_cols;       // TMEM columns between the two tile slots
  int o_shared;        // 1: ONE O accumulator behind both S slots (2 * n_pad + hd_pad <= 512): a slot is free for the
                       //    next S as soon as its P was consumed; 0: O lives inside each slot's consumed S columns
  int o_off;           // TMEM column of O: absolute (o_shared) or relative to the slot
  int bounded;         // 1: the caller guarantees |q.k| <= logit bound (qk-norm): the row maximum is not computed,
                       //    p = exp2(s * scale_log2 - bound_log2) — softmax is shift-invariant, so the result is the same
  float bound_log2;    // logit bound * scale * log2(e)
  int o_bufs;          // staging tiles per epilogue warp (2 = the TMA read of one chunk overlaps staging the next)
  int pack, pack_shift;  // query heads packed into one tile (power of two <= 32), its log2
  int tok_per_tile;    // 128 >> pack_shift
  int q_tiles;         // ceil(Tq / tok_per_tile)
  int head_sets;       // hpg / pack
  int tiles_per_item;  // q_tiles * head_sets
  int num_items;       // N * G
  int nkv;             // K / V ring depth
  int ldg_vec;         // kLdg: 32-bit words per cp.async copy (1 or 2)
  int o_stage;         // 1 = 128B-swizzled staging + TMA tile stores (one per 64 columns); 2 = the same staging, written
                       // out with coalesced 32-bit stores; 3 = dense row staging + ONE untiled TMA store per warp
                       // (O dense within a token: the warp's 32 rows are 32/pack runs of pack*hd elements)
  int simple_mask;     // 1: no band, no q_valid / k_valid — only the columns >= Tk are masked
  unsigned div_g[2], div_qt[2];
  unsigned long long* trace;   // debug timeline (VATS_ENABLE_TRACE builds): block 0 appends (tag, clock) records
  int trace_cap;
};

#if defined(VATS_ENABLE_TRACE)
struct MidTracer {
  unsigned long long* base;
  int n, cap;
  __device__ __forceinline__ MidTracer(const MidParams& P, int role)
      : base(P.trace != nullptr && blockIdx.x == 0 ? P.trace + (size_t)role * 2 * P.trace_cap : nullptr), n(0),
        cap(P.trace_cap) {}
  __device__ __forceinline__ void operator()(unsigned tag) {
    if (base != nullptr && n < cap) {
      base[2 * n] = tag;
      base[2 * n + 1] = (unsigned long long)clock64();
      ++n;
    }
  }
};
#else
struct MidTracer {
  __device__ __forceinline__ MidTracer(const MidParams&, int) {}
  __device__ __forceinline__ void operator()(unsigned) {}
};
#endif

struct MidBarriers {
  uint64_t q_full[2], q_empty[2], s_full[2], p_full[2], o_full[2], o_empty[2];
  uint64_t k_full[kMidMaxKv], k_empty[kMidMaxKv], v_full[kMidMaxKv], v_empty[kMidMaxKv];
  float row_sum[2][128];      // [slot][row]: row sums of P, read by the epilogue
  uint32_t kbits[8][8];       // per softmax warp: k_valid of the current item as bit words (generic-mask path)
  uint32_t tmem_base;
  uint32_t pad;
};

__host__ __device__ inline size_t mid_smem_bytes(int regions, int n_pad, int nkv, int o_bufs) {
  return (size_t)2 * regions * kMidQRegionBytes + (size_t)2 * nkv * regions * n_pad * 128 +
         (size_t)4 * o_bufs * kTcOStageBytes + 1024 + sizeof(MidBarriers);
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

// 16-bit word of allowed columns [base, base + 16) for a row whose allowed keys are [lo, hi]
__device__ __forceinline__ uint32_t mid_range16(int lo, int hi, int base) {
  const int l = lo - base, h = hi - base;
  const uint32_t ml = l <= 0 ? 0xffffu : (l >= 16 ? 0u : (0xffffu << l) & 0xffffu);
  const uint32_t mh = h >= 15 ? 0xffffu : (h < 0 ? 0u : 0xffffu >> (15 - h));
  return ml & mh;
}

__device__ __forceinline__ void mid_named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
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

// kLdg: Q / K / V staged with cp.async instead of TMA.  kSimple: no band, no q_valid / k_valid — the only masked
// columns are those past the end of the sequence (the ViT passes); the generic per-row predicate is then not even
// compiled in (the softmax warps are instruction-fetch-bound: every instruction and branch removed from the per-tile
// path counts).
template <bool kLdg, bool kSimple>
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
                                                     (size_t)2 * P.nkv * kv_tile + (size_t)4 * P.o_bufs * kTcOStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 12 && lane == 0) {
    if (!kLdg) {
      prefetch_tmap(&tmap_q);
      prefetch_tmap(&tmap_k);
      prefetch_tmap(&tmap_v);
    }
    if (P.o_stage == 1 || P.o_stage == 3) prefetch_tmap(&tmap_o);
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
  if (warp == 15) {
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp >= 12) {
    setmaxnreg_dec<64>();
    if (warp <= 14) {
      if (!kLdg) {
        // ------------------------------------------------------------------ TMA producers
        if (lane == 0) {
          MidTracer trace(P, 0);
          if (warp == 14) {
            MidCursor c{(int)blockIdx.x, 0};
            for (uint32_t f = 0; c.item < P.num_items; ++f) {
              const MidTile t = mid_decode(P, c);
              const uint32_t s = f & 1u, u = f >> 1;
              mbar_wait_relaxed(smem_u32(&bars->q_empty[s]), (u & 1u) ^ 1u);
              const uint32_t bar = smem_u32(&bars->q_full[s]);
              mbar_expect_tx(bar, q_tile);
              for (int r = 0; r < P.regions; ++r)
                tma_load_4d(sQ + s * q_tile + (uint32_t)r * kMidQRegionBytes, &tmap_q, bar, 64 * r, t.head0, t.q0, t.n);
              trace(0x300u + (f & 15u));
              mid_advance(P, c, 1);
            }
          } else {
            const bool is_k = warp == 12;
            const CUtensorMap* tm = is_k ? &tmap_k : &tmap_v;
            uint64_t* full = is_k ? bars->k_full : bars->v_full;
            uint64_t* empty = is_k ? bars->k_empty : bars->v_empty;
            const uint32_t ring = is_k ? sK : sV;
            int slot = 0;
            uint32_t ph = 0u;
            for (int item = (int)blockIdx.x; item < P.num_items; item += (int)gridDim.x) {
              unsigned n, g;
              tc_fastdiv((unsigned)item, P.div_g, (unsigned)a.G, &n, &g);
              mbar_wait_relaxed(smem_u32(&empty[slot]), ph ^ 1u);
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
        const int ltid = (int)threadIdx.x - 12 * 32;
        uint32_t pending = 0u;
        // A tile's arrival is deferred until the next tile's copies are in flight — but never across a wait that may
        // depend on it (with a one-deep K ring the next item's k_empty needs S of the tile still pending here).
        auto wait_flush = [&](uint32_t bar, uint32_t parity, uint32_t tag) {
          if (mbar_test_wait(bar, parity)) return;
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
      MidTracer trace(P, 1);
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
          if (leader) trace(0x100u + ((uint32_t)f & 15u));
          mbar_wait_relaxed(smem_u32(&bars->p_full[s]), u & 1u, 20);
          if (pv_tt == 0) mbar_wait(smem_u32(&bars->v_full[vs]), vph, 0x101u);
          if (P.o_shared && f >= 3)   // the one accumulator: the epilogue must have read O(f-3) out
            mbar_wait(smem_u32(&bars->o_empty[s ^ 1u]), ((uint32_t)(f - 3) >> 1) & 1u, 0x102u);
          if (leader) trace(0x110u + ((uint32_t)f & 15u));
          tc_fence_after();
          const uint32_t tP = tmem_u + s * (uint32_t)P.slot_cols;
          const uint32_t tD = P.o_shared ? tmem_u + (uint32_t)P.o_off : tP + (uint32_t)P.o_off;
          const uint32_t v_lo = v_lo_base + (uint32_t)vs * kv_step;
          uint32_t acc = 0u;
          for (int k = 0; k < ksteps_o; ++k) {
            mma_ts_lohi(tD, tP + (uint32_t)k * 8u, v_lo + (uint32_t)k * (2048u >> 4), hi_sw, idesc_o, acc, leader);
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
          if (leader) trace(0x120u + ((uint32_t)f & 15u));
          mbar_wait(smem_u32(&bars->q_full[s]), u & 1u, 0x110u);
          if (s_tt == 0) mbar_wait(smem_u32(&bars->k_full[ks]), kph, 0x111u);
          if (leader) trace(0x130u + ((uint32_t)f & 15u));
          if (!P.o_shared && u > 0)   // O inside the slot: its previous O must have been read out
            mbar_wait(smem_u32(&bars->o_empty[s]), (u - 1u) & 1u, 0x112u);
          if (leader) trace(0x140u + ((uint32_t)f & 15u));
          tc_fence_after();
          uint32_t ql = q_lo[s], kl = k_lo_base + (uint32_t)ks * kv_step;
          uint32_t acc = 0u;
          for (int k = 0; k < ksteps_s; ++k) {
            mma_ss_lohi(tmem_u + s * (uint32_t)P.slot_cols, ql, hi_sw, kl, hi_sw, idesc_s, acc, leader);
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
  } else if (warp >= 8) {
    // ====================================================================== epilogue warpgroup (warps 8-11)
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t stage0 = sO + (uint32_t)(wq * P.o_bufs) * kTcOStageBytes;
    MidTracer trace(P, 3);
    uint32_t nstore = 0u;   // staging tiles used so far (alternates the two buffers)
    MidCursor c{(int)blockIdx.x, 0};
    for (uint32_t f = 0; c.item < P.num_items; ++f, mid_advance(P, c, 1)) {
      const MidTile t = mid_decode(P, c);
      const uint32_t slot = f & 1u, u = f >> 1;
      const int tok = t.q0 + (r >> P.pack_shift);
      const int tok_w = t.q0 + ((wq * 32) >> P.pack_shift);   // first token of this warp's 32 rows
      const bool warp_live = tok_w < a.Tq;
      const uint32_t tO = tmem + lane_base + (P.o_shared ? 0u : slot * (uint32_t)P.slot_cols) + (uint32_t)P.o_off;
      bool qok = tok < a.Tq;
      if (qok && a.q_valid != nullptr) qok = a.q_valid[(long long)t.n * a.Tq + tok] != 0;

      if (r == 0) trace(0x240u + (f & 15u));
      mbar_wait_relaxed(smem_u32(&bars->o_full[slot]), u & 1u, 40);
      if (r == 0) trace(0x250u + (f & 15u));
      tc_fence_after();
      const float l_sum = bars->row_sum[slot][r];
      const float inv = (qok && l_sum > 0.f) ? 1.f / l_sum : 0.f;
      if (P.o_stage == 3) {
        // ---- dense row staging: lane's row at stage + lane * hd * 2 (exactly the box {pack*hd/2 words, 32/pack tokens}
        //      of the untiled tensor map), both 64-column halves, then ONE TMA store for the warp's 32 rows
        const uint32_t stage = stage0 + (P.o_bufs == 2 && a.hd <= 64 ? (nstore & 1u) * kTcOStageBytes : 0u);
        const uint32_t row_bytes = (uint32_t)a.hd * 2u;
        const bool vec16 = (row_bytes & 15u) == 0u;
        const int npieces = (P.hd_pad + 31) >> 5;   // 32 accumulator columns at a time (the epilogue warps run on 96 registers)
        if (warp_live) {
          if (lane == 0) {   // the store that last read this staging tile must be done with it
            if (P.o_bufs == 2 && a.hd <= 64)
              bulk_wait_group_read<1>();
            else
              bulk_wait_group_read0();
          }
          __syncwarp();
        }
#pragma unroll 1
        for (int pi = 0; pi < 4; ++pi) {
          if (pi < npieces) {
            const int cb = pi * 32;
            if (warp_live) {
              // (both halves are always loaded: the slot's 256 columns exist, columns >= hd are never stored)
              uint32_t acc[32];
              tmem_ld_32x32b_x16(tO + cb, acc);
              tmem_ld_32x32b_x16(tO + cb + 16, acc + 16);
              tmem_ld_wait();
              if (pi == npieces - 1) {   // the last columns of O are in registers: the slot may take the next S
                tc_fence_before();
                mbar_arrive(smem_u32(&bars->o_empty[slot]));
              }
              const uint32_t dst_row = stage + (uint32_t)lane * row_bytes + (uint32_t)cb * 2u;
#pragma unroll
              for (int uu = 0; uu < 4; ++uu) {
                const int col = cb + uu * 8;
                const uint32_t x = pack_bf16x2(__uint_as_float(acc[8 * uu + 0]) * inv, __uint_as_float(acc[8 * uu + 1]) * inv);
                const uint32_t y = pack_bf16x2(__uint_as_float(acc[8 * uu + 2]) * inv, __uint_as_float(acc[8 * uu + 3]) * inv);
                const uint32_t z = pack_bf16x2(__uint_as_float(acc[8 * uu + 4]) * inv, __uint_as_float(acc[8 * uu + 5]) * inv);
                const uint32_t w = pack_bf16x2(__uint_as_float(acc[8 * uu + 6]) * inv, __uint_as_float(acc[8 * uu + 7]) * inv);
                const uint32_t dst = dst_row + (uint32_t)uu * 16u;
                if (vec16) {
                  if (col < a.hd)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
                } else {   // rows only 4-byte aligned (hd 66, 60): word stores, conflict-free for an odd word pitch
                  if (col < a.hd) asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(x) : "memory");
                  if (col + 2 < a.hd) asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + 4u), "r"(y) : "memory");
                  if (col + 4 < a.hd) asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + 8u), "r"(z) : "memory");
                  if (col + 6 < a.hd) asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + 12u), "r"(w) : "memory");
                }
              }
            } else if (pi == npieces - 1) {
              tc_fence_before();
              mbar_arrive(smem_u32(&bars->o_empty[slot]));
            }
          }
        }
        if (warp_live) {
          ++nstore;
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmap_o, stage, (t.head0 * a.hd) >> 1, tok_w, t.n);
            bulk_commit_group();
          }
        }
      } else {
#pragma unroll
      for (int cbi = 0; cbi < 2; ++cbi) {
        const int cb = cbi * 64;
        if (cb >= P.hd_pad) continue;
        uint32_t acc[64];
        if (warp_live) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (cb + i * 16 < P.hd_pad) tmem_ld_32x32b_x16(tO + cb + i * 16, acc + i * 16);
          tmem_ld_wait();
        }
        if (cb + 64 >= P.hd_pad) {   // the last columns of O are in registers: the slot may take the next S
          tc_fence_before();
          mbar_arrive(smem_u32(&bars->o_empty[slot]));
        }
        if (!warp_live) continue;
        const uint32_t stage = stage0 + (P.o_bufs == 2 ? (nstore & 1u) * kTcOStageBytes : 0u);
        ++nstore;
        if (P.o_stage == 1 && lane == 0) {   // the TMA store that last read this staging tile must be done with it
          if (P.o_bufs == 2)
            bulk_wait_group_read<1>();
          else
            bulk_wait_group_read0();
        }
        __syncwarp();
#pragma unroll
        for (int uu = 0; uu < 8; ++uu) {
          if (cb + uu * 8 < P.hd_pad) {
            const uint32_t x = pack_bf16x2(__uint_as_float(acc[8 * uu + 0]) * inv, __uint_as_float(acc[8 * uu + 1]) * inv);
            const uint32_t y = pack_bf16x2(__uint_as_float(acc[8 * uu + 2]) * inv, __uint_as_float(acc[8 * uu + 3]) * inv);
            const uint32_t z = pack_bf16x2(__uint_as_float(acc[8 * uu + 4]) * inv, __uint_as_float(acc[8 * uu + 5]) * inv);
            const uint32_t w = pack_bf16x2(__uint_as_float(acc[8 * uu + 6]) * inv, __uint_as_float(acc[8 * uu + 7]) * inv);
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
      if (r == 0) trace(0x260u + (f & 15u));
    }
    if ((P.o_stage == 1 || P.o_stage == 3) && lane == 0) bulk_wait_group0();   // staged O tiles must be out before the CTA's smem goes away
  } else {
    // ====================================================================== softmax warpgroups (warps 0-3, 4-7)
    // Warpgroup h owns TMEM slot h, i.e. every second tile of this CTA: the two run half a tile apart, so the
    // per-tile bookkeeping, barrier round trips and TMEM latencies of one hide behind the exponentials of the other
    // (the MUFU pipe is the resource they share).  Thread r owns tile row r (TMEM lane r) and streams it through
    // registers 32 columns at a time, double-buffered: exact mode makes one pass for the row maximum and one for
    // the exponentials (a TMEM read costs ~50 clk per 128 columns — tools/micro/tmem_bw.cu), bounded mode
    // (qk-norm) the second pass only.  P (bf16) overwrites the slot's S columns in place, behind the reads.
    setmaxnreg_inc<152>();
    const int h = warp >> 2;
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t tS = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)h * (uint32_t)P.slot_cols;
    uint32_t* kb = bars->kbits[warp];
    const bool use_kb = !kSimple && a.k_valid != nullptr;
    const int nch = (P.n_pad + 31) >> 5;
    const bool tail16 = (P.n_pad & 16) != 0;   // the last chunk holds 16 columns (its upper half is never stored)
    const uint32_t s_bar = smem_u32(&bars->s_full[h]);
    const uint32_t p_bar = smem_u32(&bars->p_full[h]);
    MidTracer trace(P, 2);
    MidCursor c{(int)blockIdx.x, 0};
    mid_advance(P, c, h);
    for (uint32_t j = 0; c.item < P.num_items; ++j, mid_advance(P, c, 2)) {
      const MidTile t = mid_decode(P, c);
      const int tok = t.q0 + (r >> P.pack_shift);
      const bool warp_live = (t.q0 + ((wq * 32) >> P.pack_shift)) < a.Tq;
      int lo = 0, hi = a.Tk - 1;
      if (!kSimple && tok < a.Tq) {
        const long long l = key_lo(a.mask, tok), hh = key_hi(a.mask, tok);
        lo = l < 0 ? 0 : (l > 256 ? 256 : (int)l);
        hi = hh < -1 ? -1 : (int)hh;   // key_hi is already <= Tk - 1
      }
      if (use_kb && warp_live) {
        for (int w = 0; w < nch; ++w) {
          const int key = w * 32 + lane;
          const bool ok = key < a.Tk && a.k_valid[(long long)t.n * a.Tk + key] != 0;
          const uint32_t bits = __ballot_sync(0xffffffffu, ok);
          if (lane == 0) kb[w] = bits;
        }
        __syncwarp();
      }
      // allowed columns of chunk ci for this row, applied to the chunk in registers
      auto mask_chunk = [&](uint32_t (&x)[32], int ci) {
        uint32_t bits = 0xffffffffu;
        if (kSimple) {
          const int over = ci * 32 + 32 - a.Tk;   // columns of the chunk past the end of the sequence
          if (over > 0) bits = over >= 32 ? 0u : (0xffffffffu >> over);
        } else {
          bits = mid_range_word(lo, hi, ci * 32);
          if (use_kb) bits &= kb[ci];
        }
        if (bits != 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (!((bits >> i) & 1u)) x[i] = 0xff800000u;
        }
      };

      if (r == 0 && h == 0) trace(0x200u + (j & 15u));
      mbar_wait(s_bar, j & 1u, 0x200u);
      tc_fence_after();
      if (r == 0 && h == 0) trace(0x210u + (j & 15u));

      float l_sum = 0.f;
      if (warp_live) {
        uint32_t xa[32], xb[32];
        float neg_m = -P.bound_log2;
        if (!P.bounded) {
          // ---- pass 1: row maximum
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
          auto max_chunk = [&](uint32_t (&x)[32], int ci) {
            mask_chunk(x, ci);
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              m0 = mid_max3(m0, __uint_as_float(x[i + 0]), __uint_as_float(x[i + 1]));
              m1 = mid_max3(m1, __uint_as_float(x[i + 2]), __uint_as_float(x[i + 3]));
              m2 = mid_max3(m2, __uint_as_float(x[i + 4]), __uint_as_float(x[i + 5]));
              m3 = mid_max3(m3, __uint_as_float(x[i + 6]), __uint_as_float(x[i + 7]));
            }
          };
          tmem_ld_32x32b_x32(tS, xa);
          tmem_ld_wait();
#pragma unroll 1
          for (int ci = 0; ci < nch; ci += 2) {
            if (ci + 1 < nch) tmem_ld_32x32b_x32(tS + (uint32_t)(ci + 1) * 32u, xb);
            max_chunk(xa, ci);
            tmem_ld_wait();
            if (ci + 1 < nch) {
              if (ci + 2 < nch) tmem_ld_32x32b_x32(tS + (uint32_t)(ci + 2) * 32u, xa);
              max_chunk(xb, ci + 1);
              tmem_ld_wait();
            }
          }
          const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          neg_m = (m == -INFINITY) ? 0.f : -m * a.scale_log2;
          if (r == 0 && h == 0) trace(0x220u + (j & 15u));
        }
        // ---- pass 2: p = exp2(s * scale_log2 - m), row sum, P -> TMEM
        const float2 sc2 = make_float2(a.scale_log2, a.scale_log2);
        const float2 nm2 = make_float2(neg_m, neg_m);
        float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
        auto exp_chunk = [&](uint32_t (&x)[32], int ci) {
          mask_chunk(x, ci);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(x[i]), __uint_as_float(x[i + 1])), sc2, nm2);
            const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(x[i + 2]), __uint_as_float(x[i + 3])), sc2, nm2);
            const float2 p0 = make_float2(ex2(x0.x), ex2(x0.y));
            const float2 p1 = make_float2(ex2(x1.x), ex2(x1.y));
            sum_a = __fadd2_rn(sum_a, p0);
            sum_b = __fadd2_rn(sum_b, p1);
            pk[i >> 1] = pack_bf16x2(p0.x, p0.y);
            pk[(i >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
          }
          tmem_st_32x32b_x8(tS + (uint32_t)ci * 16u, pk);
          if (!(tail16 && ci == nch - 1)) tmem_st_32x32b_x8(tS + (uint32_t)ci * 16u + 8u, pk + 8);
        };
        tmem_ld_32x32b_x32(tS, xa);
        tmem_ld_wait();
#pragma unroll 1
        for (int ci = 0; ci < nch; ci += 2) {
          if (ci + 1 < nch) tmem_ld_32x32b_x32(tS + (uint32_t)(ci + 1) * 32u, xb);
          exp_chunk(xa, ci);
          tmem_ld_wait();
          if (ci + 1 < nch) {
            if (ci + 2 < nch) tmem_ld_32x32b_x32(tS + (uint32_t)(ci + 2) * 32u, xa);
            exp_chunk(xb, ci + 1);
            tmem_ld_wait();
          }
        }
        l_sum = (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
        tmem_st_wait();
      }
      // the epilogue reads row_sum of the slot's previous tile before it releases o_empty: do not overwrite it earlier
      // (short rows: this softmax can finish before the epilogue got to tile f-2)
      if (j > 0) mbar_wait(smem_u32(&bars->o_empty[h]), (j - 1u) & 1u, 0x202u);
      bars->row_sum[h][r] = l_sum;
      tc_fence_before();
      mbar_arrive(p_bar);
      if (r == 0 && h == 0) trace(0x230u + (j & 15u));
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 15) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace vats

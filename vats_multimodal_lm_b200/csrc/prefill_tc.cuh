// prefill_tc.cuh — GQA + sliding-window attention on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
//
// Replaces the attention core of the reference: the SDPA call and mask build of
// src/optimized_attention.py:657-723 (LLM), vit_2d/optimized_attention.py:348-423 and
// vit_3d/optimized_attention.py:185-348, with the window semantics of the (dead) FA2 call at
// src/optimized_attention.py:628-635, and `extend_kv_heads` (utils/attention_utils.py:7-27) folded into indexing.
//
// One CTA owns one KV group g of one sequence n, one block of 128 query tokens, and a PAIR of query heads of that
// group (M-tile 0 and M-tile 1, 128 rows each).  Both tiles consume the same K/V tiles from shared memory
// (GQA head-group reuse) and ping-pong on the tensor core so the softmax of one overlaps the MMAs of the other.
//
//   warps 0-3   softmax warpgroup of tile 0: thread r owns S row r (TMEM lane r) — row max / row sum are in-thread
//   warps 4-7   softmax warpgroup of tile 1
//   warp  8     TMA producer: Q tiles once, then K_j, V_j into mbarrier-guarded rings (128B-swizzled boxes)
//   warps 8-10  (only for head dims TMA cannot address, e.g. 60 or 66: rows are 4-byte, not 16-byte, aligned)
//               LDG staging loaders: coalesced 32-bit loads, stored into the same 128B-swizzled layout
//   warp  11    MMA issuer (one elected lane): S = Q·K^T (SS, both K-major), O += P·V (TS: P from TMEM, V MN-major)
//   Registers are re-balanced with setmaxnreg: 80 for warps 8-11, 208 for the softmax warps.
//
//   TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P (bf16) aliases the first 64 columns of S.
//
// KV tiles outside the (causal, window) band of the query block are skipped entirely (mask.cuh: tile_range); only
// tiles cut by the band, by the end of the sequence or by k_valid run the per-element predicate.
// Softmax is online with lazy rescaling: O is rescaled only when the running maximum grows by more than 2^8.
#pragma once
#include "mask.cuh"
#include "prefill_simt.cuh"  // PrefillParams
#include "ptx.cuh"

namespace vats {

constexpr int kTcBlockM = 128;
constexpr int kTcBlockN = 128;
constexpr int kTcThreads = 384;
constexpr int kTcLoaderThreads = 96;  // warps 8-10 in LDG staging mode
constexpr int kTcRegionBytes = 128 * 128;  // 128 rows x 64 bf16, one 128B-swizzled box
constexpr int kTcMaxStages = 4;
constexpr float kTcRescaleThreshold = 8.0f;  // log2 units

struct TcParams {
  PrefillParams a;
  int hd_pad;        // head dim rounded up to a multiple of 16 (MMA K of QK^T, MMA N of PV)
  int regions;       // ceil(hd_pad / 64) swizzle regions per tile
  int q_blocks;      // ceil(Tq / 128)
  int pairs;         // ceil(hpg / 2) head pairs per KV group
  int nk, nv;        // ring depths
  int q_ldg, kv_ldg; // 1: the tensor is not TMA-addressable (row starts only 4-byte aligned): LDG staging instead
  int o_vec16;       // 1: O rows may be written with 16-byte stores
  int exp_poly;      // 1: on unmasked tiles half of the exponentials run on the FMA pipe (degree-3 polynomial)
};

struct TcSmemBarriers {
  uint64_t q_full[2];
  uint64_t q_fixed[2];
  uint64_t k_full[kTcMaxStages], k_empty[kTcMaxStages];
  uint64_t v_full[kTcMaxStages], v_empty[kTcMaxStages];
  uint64_t s_full[2];
  uint64_t p_full[2];
  uint64_t o_full;
  uint32_t tmem_base;
  uint32_t pad;
};

__host__ __device__ inline size_t tc_smem_bytes(int regions, int nk, int nv) {
  return (size_t)(2 + nk + nv) * regions * kTcRegionBytes + 1024 /*alignment slack*/ + sizeof(TcSmemBarriers);
}

// Stage rows [row0, row0+128) x hd of a row-strided bf16 matrix into a 128B-swizzled K-major tile (the layout a
// (64 x 128) SWIZZLE_128B TMA box would produce), with 32-bit loads.  Rows >= T and columns [hd, hd_pad) are zero.
template <int NT>
__device__ __forceinline__ void ldg_stage_tile(unsigned char* dst, const __nv_bfloat16* src, long long stride_t,
                                               int row0, int T, int hd, int hd_pad, int tid) {
  const int wpr = hd_pad >> 1;  // 32-bit words per staged row
  const int hw = hd >> 1;       // valid words per row
  // word index idx = r * wpr + w walks tid, tid + NT, ...; (r, w) advance incrementally (no per-word division)
  const int dq = NT / wpr, dr = NT % wpr;
  int r_l = tid / wpr, w_l = tid - r_l * wpr;  // load cursor
  int r_s = r_l, w_s = w_l;                    // store cursor
  constexpr int UNR = 8;
  while (r_s < 128) {
    uint32_t val[UNR];
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
      val[i] = 0u;
      if (r_l < 128 && (row0 + r_l) < T && w_l < hw)
        val[i] = ptx::ldg_nc_u32(src + (long long)(row0 + r_l) * stride_t + 2 * w_l);
      w_l += dr;
      r_l += dq;
      if (w_l >= wpr) {
        w_l -= wpr;
        ++r_l;
      }
    }
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
      if (r_s < 128) {
        const uint32_t r = (uint32_t)r_s, w = (uint32_t)w_s;
        const uint32_t wi = w & 31u;
        const uint32_t off = (w >> 5) * kTcRegionBytes + r * 128u + (((wi >> 2) ^ (r & 7u)) << 4) + ((wi & 3u) << 2);
        *reinterpret_cast<uint32_t*>(dst + off) = val[i];
      }
      w_s += dr;
      r_s += dq;
      if (w_s >= wpr) {
        w_s -= wpr;
        ++r_s;
      }
    }
  }
}

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

__global__ void __launch_bounds__(kTcThreads, 1)
prefill_tc_kernel(const TcParams P, const __grid_constant__ CUtensorMap tmap_q,
                  const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v) {
  using namespace ptx;
  extern __shared__ unsigned char smem_raw[];
  const PrefillParams& a = P.a;

  // ---- carve shared memory (1024-byte aligned for the 128B swizzle atoms)
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t tile_bytes = (uint32_t)P.regions * kTcRegionBytes;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + 2 * tile_bytes;
  const uint32_t sV = sK + (uint32_t)P.nk * tile_bytes;
  TcSmemBarriers* bars =
      reinterpret_cast<TcSmemBarriers*>(smem_raw + (base - raw) + (size_t)(2 + P.nk + P.nv) * tile_bytes);

  // ---- work decode: q block fastest, then (group, pair), then sequence
  int bid = blockIdx.x;
  const int qb = (P.q_blocks - 1) - (bid % P.q_blocks);  // heavy (late) causal blocks first
  bid /= P.q_blocks;
  const int pair = bid % P.pairs;
  bid /= P.pairs;
  const int g = bid % a.G;
  const int n = bid / a.G;
  const int q0 = qb * kTcBlockM;
  const int hh0 = pair * 2;  // head-in-group of tile 0
  const bool active1 = (hh0 + 1) < a.hpg;
  const int head0 = g * a.hpg + hh0;

  int t_first, t_last;
  tile_range(a.mask, q0, kTcBlockM, kTcBlockN, &t_first, &t_last);
  const int n_tiles = t_last - t_first + 1;  // <= 0: nothing to attend

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup
  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bars->q_full[t]), 1);
      mbar_init(smem_u32(&bars->q_fixed[t]), 128);
      mbar_init(smem_u32(&bars->s_full[t]), 1);
      mbar_init(smem_u32(&bars->p_full[t]), 128);
    }
    const uint32_t kv_arrivals = P.kv_ldg ? kTcLoaderThreads : 1;
    for (int s = 0; s < kTcMaxStages; ++s) {
      mbar_init(smem_u32(&bars->k_full[s]), kv_arrivals);
      mbar_init(smem_u32(&bars->k_empty[s]), 1);
      mbar_init(smem_u32(&bars->v_full[s]), kv_arrivals);
      mbar_init(smem_u32(&bars->v_empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->o_full), 1);
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
    // =========================================================== warpgroup 2: producers (warps 8-10) + MMA issuer (11)
    setmaxnreg_dec<80>();
  if (warp <= 10) {
    if (warp == 8 && lane == 0 && n_tiles > 0 && !P.q_ldg) {
      for (int t = 0; t < 2; ++t) {
        if (t == 1 && !active1) break;
        const uint32_t bar = smem_u32(&bars->q_full[t]);
        mbar_expect_tx(bar, tile_bytes);
        for (int c = 0; c < P.regions; ++c)
          tma_load_4d(sQ + t * tile_bytes + c * kTcRegionBytes, &tmap_q, bar, 64 * c, head0 + t, q0, n);
      }
    }
    if (!P.kv_ldg) {
      // ---- TMA: one elected lane feeds the K and V rings
      if (warp == 8 && lane == 0) {
        for (int j = 0; j < n_tiles; ++j) {
          const int k0 = (t_first + j) * kTcBlockN;
          {
            const int s = j % P.nk;
            mbar_wait(smem_u32(&bars->k_empty[s]), ((uint32_t)(j / P.nk) & 1u) ^ 1u);
            const uint32_t bar = smem_u32(&bars->k_full[s]);
            mbar_expect_tx(bar, tile_bytes);
            for (int c = 0; c < P.regions; ++c)
              tma_load_4d(sK + s * tile_bytes + c * kTcRegionBytes, &tmap_k, bar, 64 * c, g, k0, n);
          }
          {
            const int s = j % P.nv;
            mbar_wait(smem_u32(&bars->v_empty[s]), ((uint32_t)(j / P.nv) & 1u) ^ 1u);
            const uint32_t bar = smem_u32(&bars->v_full[s]);
            mbar_expect_tx(bar, tile_bytes);
            for (int c = 0; c < P.regions; ++c)
              tma_load_4d(sV + s * tile_bytes + c * kTcRegionBytes, &tmap_v, bar, 64 * c, g, k0, n);
          }
        }
      }
    } else {
      // ---- LDG staging: 96 threads copy each K / V tile into the swizzled layout
      const int ltid = threadIdx.x - 8 * 32;
      const __nv_bfloat16* kbase = a.k + n * a.ks_n + (long long)g * a.ks_h;
      const __nv_bfloat16* vbase = a.v + n * a.vs_n + (long long)g * a.vs_h;
      unsigned char* sK_g = smem_raw + (sK - raw);
      unsigned char* sV_g = smem_raw + (sV - raw);
      for (int j = 0; j < n_tiles; ++j) {
        const int k0 = (t_first + j) * kTcBlockN;
        {
          const int s = j % P.nk;
          mbar_wait(smem_u32(&bars->k_empty[s]), ((uint32_t)(j / P.nk) & 1u) ^ 1u);
          ldg_stage_tile<kTcLoaderThreads>(sK_g + (size_t)s * tile_bytes, kbase, a.ks_t, k0, a.Tk, a.hd, P.hd_pad, ltid);
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(&bars->k_full[s]));
        }
        {
          const int s = j % P.nv;
          mbar_wait(smem_u32(&bars->v_empty[s]), ((uint32_t)(j / P.nv) & 1u) ^ 1u);
          ldg_stage_tile<kTcLoaderThreads>(sV_g + (size_t)s * tile_bytes, vbase, a.vs_t, k0, a.Tk, a.hd, P.hd_pad, ltid);
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(&bars->v_full[s]));
        }
      }
    }
  } else {
    // =========================================================== MMA issuer
    if (lane == 0 && n_tiles > 0) {
      const uint32_t idesc_s = make_idesc_bf16(kTcBlockM, kTcBlockN, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(kTcBlockM, P.hd_pad, 0, 1);
      const int ksteps = P.hd_pad / 16;
      const int ntile_heads = active1 ? 2 : 1;

      // The issuing thread is a single lane running a dependent instruction stream: every instruction costs ~5
      // cycles of latency, so the loop is kept to a handful of integer ops per MMA — descriptor high words are
      // constants, low words advance by fixed increments, ring slots / phases are tracked incrementally.
      const uint32_t hi_k = smem_desc_hi_sw128(1024);   // K-major tiles (Q, K): SBO = 1024 B
      const uint32_t hi_v = hi_k;                        // MN-major V: SBO = 1024 B, LBO (in lo) = region stride
      const uint32_t q_lo0 = smem_desc_lo(sQ, 16);
      const uint32_t q_lo1 = smem_desc_lo(sQ + tile_bytes, 16);
      const uint32_t k_lo_base = smem_desc_lo(sK, 16);
      const uint32_t v_lo_base = smem_desc_lo(sV, kTcRegionBytes);
      const uint32_t tile_step = tile_bytes >> 4;        // descriptor units (16 B) per ring slot
      const uint32_t tS0 = tmem, tS1 = tmem + 128, tO0 = tmem + 256, tO1 = tmem + 384;

      // K-step ks reads 32 bytes further along the 128-byte swizzled row; after 4 steps it moves to the next
      // 64-element region.  Offsets are in descriptor units (16 B).  (A fully unrolled variant with immediate
      // offsets measured no better on B200.)
      auto issue_s = [&](uint32_t q_lo, uint32_t k_lo, uint32_t d_tmem) {
        constexpr uint32_t R = kTcRegionBytes >> 4;
        uint32_t acc = 0u;
        for (int ks = 0; ks < ksteps; ++ks) {
          mma_ss_lohi(d_tmem, q_lo, hi_k, k_lo, hi_k, idesc_s, acc);
          acc = 1u;
          const uint32_t step = ((ks & 3) == 3) ? (R - 6u) : 2u;
          q_lo += step;
          k_lo += step;
        }
      };
      auto issue_pv = [&](uint32_t p_tmem, uint32_t v_lo, uint32_t d_tmem, uint32_t acc) {
#pragma unroll
        for (int ks = 0; ks < kTcBlockN / 16; ++ks) {
          mma_ts_lohi(d_tmem, p_tmem + ks * 8, v_lo + ks * (2048 >> 4), hi_v, idesc_o, acc);
          acc = 1u;
        }
      };

      for (int t = 0; t < ntile_heads; ++t) {
        if (P.q_ldg) mbar_wait(smem_u32(&bars->q_fixed[t]), 0);
        else mbar_wait(smem_u32(&bars->q_full[t]), 0);
      }
      mbar_wait(smem_u32(&bars->k_full[0]), 0);
      tc_fence_after();
      issue_s(q_lo0, k_lo_base, tS0);
      tc_commit(smem_u32(&bars->s_full[0]));
      if (active1) {
        issue_s(q_lo1, k_lo_base, tS1);
        tc_commit(smem_u32(&bars->s_full[1]));
      }
      tc_commit(smem_u32(&bars->k_empty[0]));

      int ksn = (P.nk > 1) ? 1 : 0;          // ring slot of K_{j+1}
      uint32_t kph = (P.nk > 1) ? 0u : 1u;   // its phase parity
      int vs = 0;                            // ring slot of V_j
      uint32_t vph = 0u;
      const uint32_t p_bar0 = smem_u32(&bars->p_full[0]), p_bar1 = smem_u32(&bars->p_full[1]);
      const uint32_t s_bar0 = smem_u32(&bars->s_full[0]), s_bar1 = smem_u32(&bars->s_full[1]);
      for (int j = 0; j < n_tiles; ++j) {
        const bool has_next = (j + 1) < n_tiles;
        const uint32_t jp = (uint32_t)j & 1u;
        const uint32_t v_lo = v_lo_base + (uint32_t)vs * tile_step;
        const uint32_t k_lo = k_lo_base + (uint32_t)ksn * tile_step;
        mbar_wait(smem_u32(&bars->v_full[vs]), vph);
        // ---- tile 0
        mbar_wait(p_bar0, jp);
        tc_fence_after();
        issue_pv(tS0, v_lo, tO0, j > 0 ? 1u : 0u);
        if (has_next) {
          mbar_wait(smem_u32(&bars->k_full[ksn]), kph);
          tc_fence_after();
          issue_s(q_lo0, k_lo, tS0);
          tc_commit(s_bar0);
        }
        // ---- tile 1
        if (active1) {
          mbar_wait(p_bar1, jp);
          tc_fence_after();
          issue_pv(tS1, v_lo, tO1, j > 0 ? 1u : 0u);
          if (has_next) {
            issue_s(q_lo1, k_lo, tS1);
            tc_commit(s_bar1);
          }
        }
        tc_commit(smem_u32(&bars->v_empty[vs]));
        if (has_next) tc_commit(smem_u32(&bars->k_empty[ksn]));
        if (++vs == P.nv) { vs = 0; vph ^= 1u; }
        if (++ksn == P.nk) { ksn = 0; kph ^= 1u; }
      }
      tc_commit(smem_u32(&bars->o_full));
    }
  }
  } else {
    // =========================================================== softmax warpgroups (tile t = warp / 4)
    setmaxnreg_inc<208>();
    const int t = warp >> 2;
    const int r = threadIdx.x & 127;           // row within the tile == TMEM lane
    const int tok = q0 + r;
    const int head = head0 + t;
    const bool tile_active = (t == 0) || active1;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_base + (uint32_t)t * 128;
    const uint32_t tO = tmem + lane_base + 256 + (uint32_t)t * 128;

    float l_run = 0.f;
    float m_used = -INFINITY;  // reference maximum in scaled-log2 units; -inf = not set yet

    if (tile_active && n_tiles > 0) {
      if (P.q_ldg) {
        // the 128 threads of this warpgroup stage their own Q tile (once per CTA)
        const __nv_bfloat16* qbase = a.q + n * a.qs_n + (long long)head * a.qs_h;
        ldg_stage_tile<128>(smem_raw + (sQ - raw) + (size_t)t * tile_bytes, qbase, a.qs_t, q0, a.Tq, a.hd, P.hd_pad, r);
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&bars->q_fixed[t]));
      }

      for (int j = 0; j < n_tiles; ++j) {
        const int tile = t_first + j;
        const int k0 = tile * kTcBlockN;
        mbar_wait(smem_u32(&bars->s_full[t]), (uint32_t)j & 1u);
        tc_fence_after();
        uint32_t sr[128];
        tmem_ld_32x32b_x32(tS + 0, sr + 0);
        tmem_ld_32x32b_x32(tS + 32, sr + 32);
        tmem_ld_32x32b_x32(tS + 64, sr + 64);
        tmem_ld_32x32b_x32(tS + 96, sr + 96);
        tmem_ld_wait();

        // ---- predicate (edge tiles only)
        const bool full = tile_is_full(a.mask, tile, q0, kTcBlockM, kTcBlockN);
        if (!full || a.k_valid != nullptr) {
          uint32_t kbits[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
          if (a.k_valid != nullptr) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const int key = k0 + w * 32 + lane;
              const bool ok = key < a.Tk && a.k_valid[(long long)n * a.Tk + key] != 0;
              kbits[w] = __ballot_sync(0xffffffffu, ok);
            }
          }
          long long lo = key_lo(a.mask, tok) - k0;
          long long hi = key_hi(a.mask, tok) - k0;
          const int lo_c = lo < 0 ? 0 : (lo > 128 ? 128 : (int)lo);
          const int hi_c = hi < -1 ? -1 : (hi > 127 ? 127 : (int)hi);
#pragma unroll
          for (int c = 0; c < 128; ++c) {
            const bool ok = (c >= lo_c) && (c <= hi_c) && ((kbits[c >> 5] >> (c & 31)) & 1u);
            if (!ok) sr[c] = 0xff800000u;  // -inf
          }
        }

        // ---- row max of this tile (raw logits), in-thread
        float mt0 = __uint_as_float(sr[0]), mt1 = __uint_as_float(sr[1]), mt2 = __uint_as_float(sr[2]),
              mt3 = __uint_as_float(sr[3]);
#pragma unroll
        for (int c = 4; c < 128; c += 4) {
          mt0 = fmaxf(mt0, __uint_as_float(sr[c]));
          mt1 = fmaxf(mt1, __uint_as_float(sr[c + 1]));
          mt2 = fmaxf(mt2, __uint_as_float(sr[c + 2]));
          mt3 = fmaxf(mt3, __uint_as_float(sr[c + 3]));
        }
        const float mt = fmaxf(fmaxf(mt0, mt1), fmaxf(mt2, mt3)) * a.scale_log2;  // scaled-log2 units (scale > 0)

        // ---- lazy rescale of the running state
        float factor = 1.f;
        if (m_used == -INFINITY) {
          m_used = mt;  // may stay -inf; O and l are still zero, nothing to rescale
        } else if (mt > m_used + kTcRescaleThreshold) {
          factor = ex2(m_used - mt);
          m_used = mt;
        }
        if (j > 0 && __any_sync(0xffffffffu, factor != 1.f)) {
          l_run *= factor;
          for (int c = 0; c < P.hd_pad; c += 16) {
            uint32_t orr[16];
            tmem_ld_32x32b_x16(tO + c, orr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * factor);
            tmem_st_32x32b_x16(tO + c, orr);
          }
        }

        // ---- p = exp2(s*scale_log2 - m_used), row sum, pack to bf16, write P over S.
        //      The scale/subtract and the row sums run as packed f32x2 operations (FFMA2 / FADD2).  The MUFU pipe
        //      (16 ex2/clk/SM) needs as long for a 128x128 tile as the tensor core needs for its two MMAs, so on
        //      unmasked tiles every other pair of elements takes exp2 on the FMA pipe instead: round to nearest
        //      integer with the 1.5*2^23 trick, degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative
        //      error 7.5e-5, 50x below bf16 rounding), integer part added into the exponent field.  Masked tiles
        //      keep the MUFU path, where ex2(-inf) is exactly 0.
        const float mref = (m_used == -INFINITY) ? 0.f : m_used;
        const float2 sc2 = make_float2(a.scale_log2, a.scale_log2);
        const float2 nm2 = make_float2(-mref, -mref);
        float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
        uint32_t pk[64];
        if (P.exp_poly && full && a.k_valid == nullptr) {
          const float2 magic = make_float2(12582912.f, 12582912.f);
          const float2 nmagic = make_float2(-12582912.f, -12582912.f);
          const float2 neg1 = make_float2(-1.f, -1.f);
          const float2 c0 = make_float2(0.9999280571937561f, 0.9999280571937561f);
          const float2 c1 = make_float2(0.6932609677314758f, 0.6932609677314758f);
          const float2 c2 = make_float2(0.2426111400127411f, 0.2426111400127411f);
          const float2 c3 = make_float2(0.05517186224460602f, 0.05517186224460602f);
#pragma unroll
          for (int c = 0; c < 128; c += 4) {
            const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), sc2, nm2);
            float2 x1 = __ffma2_rn(make_float2(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])), sc2, nm2);
            const float2 p0 = make_float2(ex2(x0.x), ex2(x0.y));
            x1 = make_float2(fmaxf(x1.x, -126.f), fmaxf(x1.y, -126.f));
            const float2 r = __fadd2_rn(x1, magic);                       // integer part lands in the low mantissa bits
            const float2 fr = __ffma2_rn(__fadd2_rn(r, nmagic), neg1, x1); // x - round(x)  in [-0.5, 0.5]
            float2 q = __ffma2_rn(fr, c3, c2);
            q = __ffma2_rn(fr, q, c1);
            q = __ffma2_rn(fr, q, c0);
            const float2 p1 = make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(r.x) << 23)),
                                          __int_as_float(__float_as_int(q.y) + (__float_as_int(r.y) << 23)));
            sum_a = __fadd2_rn(sum_a, p0);
            sum_b = __fadd2_rn(sum_b, p1);
            pk[c >> 1] = pack_bf16x2(p0.x, p0.y);
            pk[(c >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 128; c += 4) {
            const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), sc2, nm2);
            const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])), sc2, nm2);
            const float2 p0 = make_float2(ex2(x0.x), ex2(x0.y));
            const float2 p1 = make_float2(ex2(x1.x), ex2(x1.y));
            sum_a = __fadd2_rn(sum_a, p0);
            sum_b = __fadd2_rn(sum_b, p1);
            pk[c >> 1] = pack_bf16x2(p0.x, p0.y);
            pk[(c >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
          }
        }
        l_run += (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
        tmem_st_32x32b_x32(tS + 0, pk + 0);
        tmem_st_32x32b_x32(tS + 32, pk + 32);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->p_full[t]));
      }
    }

    // ---- epilogue: O / l -> bf16 -> global
    // (tcgen05.ld is warp-collective: every lane of an active tile's warps runs the loads; only the global
    //  stores are predicated on the row being inside the sequence)
    if (tile_active) {
      const bool do_store = tok < a.Tq;
      bool qok = true;
      if (do_store && a.q_valid != nullptr) qok = a.q_valid[(long long)n * a.Tq + tok] != 0;
      const float inv = (qok && l_run > 0.f && n_tiles > 0) ? 1.f / l_run : 0.f;
      __nv_bfloat16* orow = a.o + n * a.os_n + (long long)tok * a.os_t + (long long)head * a.os_h;
      if (n_tiles > 0) {
        mbar_wait(smem_u32(&bars->o_full), 0);
        tc_fence_after();
      }
      for (int c = 0; c < P.hd_pad; c += 16) {
        uint32_t orr[16];
        if (n_tiles > 0) {
          tmem_ld_32x32b_x16(tO + c, orr);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) orr[i] = 0u;
        }
        __syncwarp();
        if (do_store) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          w[i] = pack_bf16x2(__uint_as_float(orr[2 * i]) * inv, __uint_as_float(orr[2 * i + 1]) * inv);
        if (P.o_vec16 && c + 16 <= a.hd) {
          *reinterpret_cast<uint4*>(orow + c) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(orow + c + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int e = c + 2 * i;
            if (e + 1 < a.hd) {
              if ((reinterpret_cast<uintptr_t>(orow + e) & 3u) == 0) {
                *reinterpret_cast<uint32_t*>(orow + e) = w[i];
              } else {
                reinterpret_cast<uint16_t*>(orow)[e] = (uint16_t)(w[i] & 0xffffu);
                reinterpret_cast<uint16_t*>(orow)[e + 1] = (uint16_t)(w[i] >> 16);
              }
            } else if (e < a.hd) {
              reinterpret_cast<uint16_t*>(orow)[e] = (uint16_t)(w[i] & 0xffffu);
            }
          }
        }
        }  // do_store
        __syncwarp();
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 11) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace vats

// prefill_tc64.cuh — the tile kernel of prefill_tc.cuh with 64-key KV steps and DOUBLE-BUFFERED S (long sequences).
//
// Same contract, same work items (sequence, KV group, 128-token query block, pair of query heads), same roles and
// register budgets as prefill_tc_kernel.  What changes is the dependency loop of a tile.  There, S and P of a tile share
// the tile's only 128 TMEM columns, so per 128-key step a softmax warpgroup runs
//     wait S(j) -> softmax -> P(j)  ...  [MMA warp: P.V(j), then S(j+1) over the same columns]  ...  wait S(j+1)
// and ncu's source view (profiles/r02w: cfg5) shows the softmax warps waiting for S in ~40 % of their samples: the tensor
// pipe is busy 58 % of the time and the MUFU pipe 59 %, neither can go up while each waits for the other.
// Here a tile's 128 S columns are TWO 64-key buffers.  S(j+1) is issued into the other buffer while the warpgroup is
// still in softmax(j), so the warpgroup goes from P(j) straight to S(j+1); the MMA warp answers P(j) with P.V(j) and
// S(j+2) into the buffer P(j) just left:
//     MMA warp:   S(0) S(1) | P.V(0) S(2) | P.V(1) S(3) | ...          (per tile; the two tiles interleave)
//     warpgroup:  softmax(0) softmax(1) softmax(2) ...                  (no tensor-core round trip in between)
// Costs: an N = 64 UMMA reads the same 4 KB of Q per k-step as an N = 128 one (S is shared-memory-bound in SS form), so
// the S work per key grows; twice as many barrier round trips, row-maximum reductions and rescale checks per key.
//
// TMEM (512 columns): S buffers (t, b) at [(2 t + b) * 64, +64); O0 [256, 384), O1 [384, 512); P (bf16) aliases the
// first 32 columns of its S buffer.  K / V tiles are 64 keys (one 8 KB box per 64-column region), rings up to 4 deep.
// TMA-addressable q / k / v only (the launcher keeps prefill_tc_kernel for everything else), no fused gather.
#pragma once
#include "prefill_tc.cuh"

namespace vats {

constexpr int kTc64BlockN = 64;
constexpr int kTc64RegionBytes = 64 * 128;   // K / V: 64 rows x 64 bf16, one 128B-swizzled box

struct Tc64Barriers {
  uint64_t q_full[2], q_empty[2];
  uint64_t k_full[kTcMaxStages], k_empty[kTcMaxStages];
  uint64_t v_full[kTcMaxStages], v_empty[kTcMaxStages];
  uint64_t s_full[2][2];   // [tile][buffer]
  uint64_t p_full[2][2];
  uint64_t o_full[2], o_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__host__ __device__ inline size_t tc64_smem_bytes(int regions, int nk, int nv, int o_stage) {
  return (size_t)2 * regions * kTcRegionBytes + (size_t)(nk + nv) * regions * kTc64RegionBytes +
         (o_stage ? 8 * kTcOStageBytes : 0) + 1024 /*alignment slack*/ + sizeof(Tc64Barriers);
}

__device__ __forceinline__ TcWork tc64_decode_work(const TcParams& P, int w) {
  const PrefillParams& a = P.a;
  TcWork k;
  unsigned rest, qbr, pair, g, n;
  tc_fastdiv((unsigned)w, P.div_qb, (unsigned)P.q_blocks, &rest, &qbr);
  const int qb = (P.q_blocks - 1) - (int)qbr;  // heavy (late) causal blocks first
  tc_fastdiv(rest, P.div_pairs, (unsigned)P.pairs, &rest, &pair);
  tc_fastdiv(rest, P.div_g, (unsigned)a.G, &n, &g);
  k.g = (int)g;
  k.n = (int)n;
  k.q0 = qb * kTcBlockM;
  const int hh0 = (int)pair * 2;
  k.active1 = (hh0 + 1) < a.hpg;
  k.head0 = k.g * a.hpg + hh0;
  if (P.no_band) {
    k.t_first = 0;
    k.n_tiles = (a.Tk + kTc64BlockN - 1) >> 6;
  } else {
    int t_last;
    tile_range(a.mask, k.q0, kTcBlockM, kTc64BlockN, &k.t_first, &t_last);
    k.n_tiles = t_last - k.t_first + 1;
  }
  return k;
}

__global__ void __launch_bounds__(kTcThreads, 1)
prefill_tc64_kernel(const TcParams P, const __grid_constant__ CUtensorMap tmap_q,
                    const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                    const __grid_constant__ CUtensorMap tmap_o) {
  using namespace ptx;
  extern __shared__ unsigned char smem_raw[];
  const PrefillParams& a = P.a;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t q_tile = (uint32_t)P.regions * kTcRegionBytes;
  const uint32_t kv_tile = (uint32_t)P.regions * kTc64RegionBytes;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + 2 * q_tile;
  const uint32_t sV = sK + (uint32_t)P.nk * kv_tile;
  const uint32_t sO = sV + (uint32_t)P.nv * kv_tile;   // 8 x kTcOStageBytes when P.o_stage
  Tc64Barriers* bars = reinterpret_cast<Tc64Barriers*>(smem_raw + (base - raw) + (size_t)2 * q_tile +
                                                       (size_t)(P.nk + P.nv) * kv_tile + (P.o_stage ? 8 * kTcOStageBytes : 0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    if (P.o_stage == 1) prefetch_tmap(&tmap_o);
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bars->q_full[t]), 1);
      mbar_init(smem_u32(&bars->q_empty[t]), 1);
      mbar_init(smem_u32(&bars->o_full[t]), 1);
      mbar_init(smem_u32(&bars->o_empty[t]), 128);
      for (int b = 0; b < 2; ++b) {
        mbar_init(smem_u32(&bars->s_full[t][b]), 1);
        mbar_init(smem_u32(&bars->p_full[t][b]), 128);
      }
    }
    for (int s = 0; s < kTcMaxStages; ++s) {
      mbar_init(smem_u32(&bars->k_full[s]), 1);
      mbar_init(smem_u32(&bars->k_empty[s]), 1);
      mbar_init(smem_u32(&bars->v_full[s]), 1);
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
      // ================================================================ TMA producers: warp 8 K ring, 9 V ring, 10 Q tiles
      if (lane == 0) {
        if (warp == 10) {
          uint32_t qn[2] = {0u, 0u};
          for (int w = blockIdx.x; w < P.num_work; w += gridDim.x) {
            const TcWork wk = tc64_decode_work(P, w);
            if (wk.n_tiles <= 0) continue;
            for (int t = 0; t < 2; ++t) {
              if (t == 1 && !wk.active1) break;
              VATS_TC_PRODUCER_WAIT(smem_u32(&bars->q_empty[t]), (qn[t] & 1u) ^ 1u);
              const uint32_t bar = smem_u32(&bars->q_full[t]);
              mbar_expect_tx(bar, q_tile);
              for (int c = 0; c < P.regions; ++c)
                tma_load_4d(sQ + t * q_tile + c * kTcRegionBytes, &tmap_q, bar, 64 * c, wk.head0 + t, wk.q0, wk.n);
              ++qn[t];
            }
          }
        } else {
          const bool is_k = warp == 8;
          const CUtensorMap* tmap = is_k ? &tmap_k : &tmap_v;
          uint64_t* full = is_k ? bars->k_full : bars->v_full;
          uint64_t* empty = is_k ? bars->k_empty : bars->v_empty;
          const uint32_t ring = is_k ? sK : sV;
          const int depth = is_k ? P.nk : P.nv;
          int slot = 0;
          uint32_t ph = 0u;
          for (int w = blockIdx.x; w < P.num_work; w += gridDim.x) {
            const TcWork wk = tc64_decode_work(P, w);
            for (int j = 0; j < wk.n_tiles; ++j) {
              const int k0 = (wk.t_first + j) * kTc64BlockN;
              VATS_TC_PRODUCER_WAIT(smem_u32(&empty[slot]), ph ^ 1u);
              const uint32_t bar = smem_u32(&full[slot]);
              mbar_expect_tx(bar, kv_tile);
              for (int c = 0; c < P.regions; ++c)
                tma_load_4d(ring + slot * kv_tile + c * kTc64RegionBytes, tmap, bar, 64 * c, wk.g, k0, wk.n);
              if (++slot == depth) { slot = 0; ph ^= 1u; }
            }
          }
        }
      }
    } else {
      // ================================================================ MMA issuer (all lanes convergent, one elected)
      const uint32_t leader = elect_one() ? 1u : 0u;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc_s = make_idesc_bf16(kTcBlockM, kTc64BlockN, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(kTcBlockM, P.hd_pad, 0, 1);
      const int ksteps = P.hd_pad / 16;
      const uint32_t hi_k = smem_desc_hi_sw128(1024);
      const uint32_t hi_v = hi_k;
      const uint32_t q_lo[2] = {smem_desc_lo(sQ, 16), smem_desc_lo(sQ + q_tile, 16)};
      const uint32_t k_lo_base = smem_desc_lo(sK, 16);
      const uint32_t v_lo_base = smem_desc_lo(sV, kTc64RegionBytes);
      const uint32_t kv_step = kv_tile >> 4;
      const uint32_t tO[2] = {tmem_u + 256, tmem_u + 384};

      auto issue_s = [&](uint32_t ql, uint32_t kl, uint32_t d_tmem) {
        constexpr uint32_t RQ = kTcRegionBytes >> 4, RK = kTc64RegionBytes >> 4;
        uint32_t acc = 0u;
        for (int ks = 0; ks < ksteps; ++ks) {
          mma_ss_lohi(d_tmem, ql, hi_k, kl, hi_k, idesc_s, acc, leader);
          acc = 1u;
          if ((ks & 3) == 3) {
            ql += RQ - 6u;
            kl += RK - 6u;
          } else {
            ql += 2u;
            kl += 2u;
          }
        }
      };
      auto issue_pv = [&](uint32_t p_tmem, uint32_t v_lo, uint32_t d_tmem, uint32_t acc) {
#pragma unroll
        for (int ks = 0; ks < kTc64BlockN / 16; ++ks) {
          mma_ts_lohi(d_tmem, p_tmem + ks * 8, v_lo + ks * (2048 >> 4), hi_v, idesc_o, acc, leader);
          acc = 1u;
        }
      };

      int ks = 0, vs = 0;
      uint32_t kph = 0u, vph = 0u;
      uint32_t pcbits = 0u;   // parity of the P phases consumed per (tile, buffer): bit 2 t + b (running across items)
      uint32_t qn[2] = {0u, 0u};
      for (int w = blockIdx.x; w < P.num_work; w += gridDim.x) {
        const TcWork wk = tc64_decode_work(P, w);
        if (wk.n_tiles <= 0) continue;
        const int nt = wk.active1 ? 2 : 1;
        const int n_steps = wk.n_tiles;
        mbar_wait(smem_u32(&bars->q_full[0]), qn[0] & 1u);
        if (wk.active1) mbar_wait(smem_u32(&bars->q_full[1]), qn[1] & 1u);
        // ---- prologue: S(0) and S(1) of both tiles into their two buffers
        for (int j = 0; j < 2 && j < n_steps; ++j) {
          mbar_wait(smem_u32(&bars->k_full[ks]), kph);
          tc_fence_after();
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (t >= nt) break;
            issue_s(q_lo[t], k_lo_base + (uint32_t)ks * kv_step, tmem_u + (uint32_t)(2 * t + j) * 64u);
            tc_commit_pred(smem_u32(&bars->s_full[t][j]), leader);
            if (j + 1 == n_steps) tc_commit_pred(smem_u32(&bars->q_empty[t]), leader);
          }
          tc_commit_pred(smem_u32(&bars->k_empty[ks]), leader);
          if (++ks == P.nk) { ks = 0; kph ^= 1u; }
        }
        // ---- steady state: P(j) -> P.V(j), then S(j + 2) into the buffer P(j) leaves
        for (int j = 0; j < n_steps; ++j) {
          const int b = j & 1;
          const bool more = (j + 2) < n_steps;
          mbar_wait(smem_u32(&bars->v_full[vs]), vph);
          if (more) mbar_wait(smem_u32(&bars->k_full[ks]), kph);
          const uint32_t v_lo = v_lo_base + (uint32_t)vs * kv_step;
          const uint32_t k_lo = k_lo_base + (uint32_t)ks * kv_step;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (t >= nt) break;
            mbar_wait(smem_u32(&bars->p_full[t][b]), (pcbits >> (2 * t + b)) & 1u);
            pcbits ^= 1u << (2 * t + b);
            if (j == 0) mbar_wait(smem_u32(&bars->o_empty[t]), (qn[t] & 1u) ^ 1u);   // previous item's O_t was read out
            tc_fence_after();
            const uint32_t tSb = tmem_u + (uint32_t)(2 * t + b) * 64u;
            issue_pv(tSb, v_lo, tO[t], j > 0 ? 1u : 0u);
            if (more) {
              issue_s(q_lo[t], k_lo, tSb);
              tc_commit_pred(smem_u32(&bars->s_full[t][b]), leader);
              if (j + 3 == n_steps) tc_commit_pred(smem_u32(&bars->q_empty[t]), leader);   // the item's last read of Q_t
            } else if (j + 2 == n_steps) {
              // no S follows P.V(n - 2): an extra ("drain") phase of this buffer's S barrier tells the warpgroup that
              // O_t is quiescent before it may rescale it in the last step
              tc_commit_pred(smem_u32(&bars->s_full[t][b]), leader);
            } else {
              tc_commit_pred(smem_u32(&bars->o_full[t]), leader);
            }
          }
          tc_commit_pred(smem_u32(&bars->v_empty[vs]), leader);
          if (++vs == P.nv) { vs = 0; vph ^= 1u; }
          if (more) {
            tc_commit_pred(smem_u32(&bars->k_empty[ks]), leader);
            if (++ks == P.nk) { ks = 0; kph ^= 1u; }
          }
        }
        ++qn[0];
        if (wk.active1) ++qn[1];
      }
    }
  } else {
    // ==================================================================== softmax warpgroups (tile t = warp / 4)
    setmaxnreg_inc<208>();
    const int t = warp >> 2;
    const int r = threadIdx.x & 127;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_base + (uint32_t)t * 128;
    const uint32_t tO = tmem + lane_base + 256 + (uint32_t)t * 128;
    uint32_t scbits = 0u;   // parity of the S phases consumed per buffer: bit b (running across items)
    uint32_t qn = 0u;

    for (int w = blockIdx.x; w < P.num_work; w += gridDim.x) {
      const TcWork wk = tc64_decode_work(P, w);
      const int n = wk.n, q0 = wk.q0, t_first = wk.t_first, n_tiles = wk.n_tiles;
      const int tok = q0 + r;
      const int head = wk.head0 + t;
      const bool tile_active = (t == 0) || wk.active1;

      float l_run = 0.f;
      float m_used = -INFINITY;
      int full_first, full_last;   // KV tiles every row of the block may attend entirely
      if (P.no_band) {
        full_first = 0;
        full_last = (a.Tk >> 6) - 1;
      } else {
        int q_last = q0 + kTcBlockM - 1;
        if (q_last > a.Tq - 1) q_last = a.Tq - 1;
        long long lo = key_lo(a.mask, q_last);
        long long hi = key_hi(a.mask, q0);
        if (lo < 0) lo = 0;
        full_first = lo > (long long)a.Tk ? (a.Tk >> 6) + 1 : ((int)lo + kTc64BlockN - 1) >> 6;
        full_last = hi < 0 ? -1 : (((int)hi + 1) >> 6) - 1;
      }

      if (tile_active && n_tiles > 0) {
        for (int j = 0; j < n_tiles; ++j) {
          const int b = j & 1;
          const int tile = t_first + j;
          const int k0 = tile * kTc64BlockN;
          const uint32_t tSb = tS + (uint32_t)b * 64u;
          mbar_wait(smem_u32(&bars->s_full[t][b]), (scbits >> b) & 1u);
          scbits ^= 1u << b;
          tc_fence_after();
          uint32_t sr[64];
          tmem_ld_32x32b_x32(tSb + 0, sr + 0);
          tmem_ld_32x32b_x32(tSb + 32, sr + 32);
          tmem_ld_wait();

          const bool full = tile >= full_first && tile <= full_last;
          if (!full || a.k_valid != nullptr) {
            uint32_t kbits[2] = {0xffffffffu, 0xffffffffu};
            if (a.k_valid != nullptr) {
#pragma unroll
              for (int ww = 0; ww < 2; ++ww) {
                const int key = k0 + ww * 32 + lane;
                const bool ok = key < a.Tk && a.k_valid[(long long)n * a.Tk + key] != 0;
                kbits[ww] = __ballot_sync(0xffffffffu, ok);
              }
            }
            long long lo = key_lo(a.mask, tok) - k0;
            long long hi = key_hi(a.mask, tok) - k0;
            const int lo_c = lo < 0 ? 0 : (lo > 64 ? 64 : (int)lo);
            const int hi_c = hi < -1 ? -1 : (hi > 63 ? 63 : (int)hi);
#pragma unroll
            for (int ww = 0; ww < 2; ++ww) {
              const int l = lo_c - 32 * ww, h = hi_c - 32 * ww;
              const uint32_t ml = l <= 0 ? 0xffffffffu : (l >= 32 ? 0u : 0xffffffffu << l);
              const uint32_t mh = h >= 31 ? 0xffffffffu : (h < 0 ? 0u : 0xffffffffu >> (31 - h));
              kbits[ww] &= ml & mh;
            }
#pragma unroll
            for (int c = 0; c < 64; ++c)
              if (!((kbits[c >> 5] >> (c & 31)) & 1u)) sr[c] = 0xff800000u;  // -inf
          }

          float mt;
          if (P.bounded) {
            mt = P.bound_log2;
          } else {
            float mx[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) mx[i] = __uint_as_float(sr[i]);
#pragma unroll
            for (int c = 8; c < 64; c += 8)
#pragma unroll
              for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], __uint_as_float(sr[c + i]));
            mt = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7]))) *
                 a.scale_log2;
          }
          // ---- lazy rescale of the running state.  O_t may only be touched once P.V(j - 1) is complete, and with S
          //      running ahead that is no longer implied by S(j): the (rare) rescale waits for the OTHER buffer's S
          //      barrier — S(j + 1), committed behind P.V(j - 1) — which the warpgroup would wait for next anyway.
          float factor = 1.f;
          if (m_used == -INFINITY) {
            m_used = mt;
          } else if (mt > m_used + kTcRescaleThreshold) {
            factor = ex2(m_used - mt);
            m_used = mt;
          }
          const bool last_step = j > 0 && j + 1 == n_tiles;
          if (last_step) {   // the drain phase (P.V(n - 2) complete) is always consumed, rescale or not
            mbar_wait(smem_u32(&bars->s_full[t][b ^ 1]), (scbits >> (b ^ 1)) & 1u);
            scbits ^= 1u << (b ^ 1);
            tc_fence_after();
          }
          if (j > 0 && __any_sync(0xffffffffu, factor != 1.f)) {
            if (!last_step) {
              // S(j + 1) was issued behind P.V(j - 1) on the in-order tensor pipe: its barrier covers both (the phase
              // is consumed by the next step)
              mbar_wait(smem_u32(&bars->s_full[t][b ^ 1]), (scbits >> (b ^ 1)) & 1u);
              tc_fence_after();
            }
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

          const float mref = (m_used == -INFINITY) ? 0.f : m_used;
          const float2 sc2 = make_float2(a.scale_log2, a.scale_log2);
          const float2 nm2 = make_float2(-mref, -mref);
          float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
          uint32_t pk[32];
#pragma unroll
          for (int c = 0; c < 64; c += 4) {
            const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), sc2, nm2);
            const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])), sc2, nm2);
            const float2 p0 = make_float2(ex2(x0.x), ex2(x0.y));
            const float2 p1 = make_float2(ex2(x1.x), ex2(x1.y));
            sum_a = __fadd2_rn(sum_a, p0);
            sum_b = __fadd2_rn(sum_b, p1);
            pk[c >> 1] = pack_bf16x2(p0.x, p0.y);
            pk[(c >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
          }
          l_run += (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
          tmem_st_32x32b_x32(tSb, pk);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(smem_u32(&bars->p_full[t][b]));
        }
      }

      // ---- epilogue (as in prefill_tc_kernel): O / l -> bf16 -> staging tile -> TMA tile store / coalesced stores
      if (tile_active) {
        const bool do_store = tok < a.Tq;
        bool qok = true;
        if (do_store && a.q_valid != nullptr) qok = a.q_valid[(long long)n * a.Tq + tok] != 0;
        const float inv = (qok && l_run > 0.f && n_tiles > 0) ? 1.f / l_run : 0.f;
        if (n_tiles > 0) {
          mbar_wait(smem_u32(&bars->o_full[t]), qn & 1u);
          tc_fence_after();
        }
        const uint32_t stage = sO + (uint32_t)warp * kTcOStageBytes;
        const int row_w = q0 + (warp & 3) * 32;
        uint32_t acc[128];
#pragma unroll
        for (int pc16 = 0; pc16 < 8; ++pc16) tmem_ld_32x32b_x16(tO + pc16 * 16, acc + pc16 * 16);
        tmem_ld_wait();
        if (n_tiles > 0) {
          tc_fence_before();
          mbar_arrive(smem_u32(&bars->o_empty[t]));
          ++qn;
        }
        const bool have = n_tiles > 0;
#pragma unroll
        for (int cbi = 0; cbi < 2; ++cbi) {
          const int cb = cbi * 64;
          if (cb >= P.hd_pad) continue;
          if (lane == 0) bulk_wait_group_read0();
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (cb + u * 8 < P.hd_pad) {
              uint32_t x = pack_bf16x2(__uint_as_float(acc[cb + 8 * u + 0]) * inv, __uint_as_float(acc[cb + 8 * u + 1]) * inv);
              uint32_t y = pack_bf16x2(__uint_as_float(acc[cb + 8 * u + 2]) * inv, __uint_as_float(acc[cb + 8 * u + 3]) * inv);
              uint32_t z = pack_bf16x2(__uint_as_float(acc[cb + 8 * u + 4]) * inv, __uint_as_float(acc[cb + 8 * u + 5]) * inv);
              uint32_t ww = pack_bf16x2(__uint_as_float(acc[cb + 8 * u + 6]) * inv, __uint_as_float(acc[cb + 8 * u + 7]) * inv);
              if (!have) x = y = z = ww = 0u;
              const uint32_t dst = stage + (uint32_t)lane * 128u + (((uint32_t)u ^ ((uint32_t)lane & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(x), "r"(y), "r"(z), "r"(ww) : "memory");
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row_w < a.Tq) {
            tma_store_4d(&tmap_o, stage, cb, head, row_w, n);
            bulk_commit_group();
          }
        }
        __syncwarp();
      }
    }
  }

  // ---- teardown
  if (warp < 8 && lane == 0) bulk_wait_group0();
  tc_fence_before();
  __syncthreads();
  if (warp == 11) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace vats

// prefill_tc.cuh — GQA + sliding-window attention on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
//
// Replaces the attention core of the reference: the SDPA call and mask build of
// src/optimized_attention.py:657-723 (LLM), vit_2d/optimized_attention.py:348-423 and
// vit_3d/optimized_attention.py:185-348, with the window semantics of the (dead) FA2 call at
// src/optimized_attention.py:628-635, and `extend_kv_heads` (utils/attention_utils.py:7-27) folded into indexing.
//
// Persistent CTAs (one per SM) walk the work items round-robin.  One item = one KV group g of one sequence n, one
// block of 128 query tokens, and a PAIR of query heads of that
// group (M-tile 0 and M-tile 1, 128 rows each).  Both tiles consume the same K/V tiles from shared memory
// (GQA head-group reuse) and ping-pong on the tensor core so the softmax of one overlaps the MMAs of the other.
//
//   warps 0-3   softmax warpgroup of tile 0: thread r owns S row r (TMEM lane r) — row max / row sum are in-thread
//   warps 4-7   softmax warpgroup of tile 1
//   warps 8-10  TMA producers (one lane each): K ring, V ring, Q tiles — mbarrier-guarded, 128B-swizzled boxes
//   warps 8-10  (kLdg instantiation: head dims TMA cannot address, e.g. 60 or 66 — rows 4- or 8-byte aligned)
//               staging loaders: cp.async copies into the same 128B-swizzled layout, two tiles in flight per thread
//   warp  11    MMA warp (all lanes convergent, one elected): S = Q·K^T (SS, both K-major), O += P·V (TS: P from
//               TMEM, V MN-major)
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

// TMA producer lanes wait for a free ring slot most of the time: sleeping between polls (mbar_wait_relaxed) keeps their
// poll loops off the issue ports of the softmax warps sharing the sub-partition.  -DVATS_TC_TIGHT_PRODUCERS restores the
// plain poll loop for A/B measurements.
#if defined(VATS_TC_TIGHT_PRODUCERS)
#define VATS_TC_PRODUCER_WAIT(bar, parity) mbar_wait((bar), (parity))
#else
#define VATS_TC_PRODUCER_WAIT(bar, parity) mbar_wait_relaxed((bar), (parity), 64)
#endif

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
  int ldg_vec;       // kLdg staging: 32-bit words per cp.async copy (2 when every q/k/v row is 8-byte aligned, else 1)
  int o_vec16;       // 1: O rows may be written with 16-byte stores
  int o_stage;       // O write-out: 0 = per-thread row stores; 1 = per-warp shared-memory staging + TMA tile stores
                     // (tmap_o is valid); 2 = the same staging, written out with coalesced 32-bit stores (rows only
                     // 4-byte aligned, e.g. head_dim 66)
  int num_work;      // N * G * pairs * q_blocks work items, walked round-robin by the persistent CTAs
  int no_band;       // 1: no causal mask and no window — every query block visits all ceil(Tk/128) KV tiles and only
                     // the sequence end cuts a tile (skips the 64-bit band arithmetic every role runs per item)
  unsigned div_qb[2], div_pairs[2], div_g[2];  // magic (multiplier, shift) pairs for division by q_blocks / pairs / G
  unsigned long long* trace;  // debug: block 0 appends (tag, clock64) pairs here (NULL = off); [0] = count
  int trace_cap;
  int order_softmax; // 1: the two softmax warpgroups take turns in the exponential phase (staggers the ping-pong)
  // fused output gather (vats_attn_prefill_gather): every staged O tile is stored into the gathered output of ALL
  // ranks (peer memory over NVLink) instead of the local output only
  int bounded;       // 1: the caller guarantees |q.k| <= logit bound (qk-norm): no row maximum, no rescaling —
  float bound_log2;  //    p = exp2(s * scale_log2 - bound_log2), bound_log2 = bound * scale * log2(e)
  int peers;         // 0 = plain launch (tmap_o); W = number of ranks whose gathered tensors receive the tiles
  int peer_first;    // first destination (each rank starts elsewhere, so no copy is hit by all ranks at once)
  int seq_off;       // this rank's first sequence / head inside the gathered [N_total, Tq, H_total, hd] tensor
  int head_off;
};

constexpr int kTcMaxPeers = 8;
struct TcPeerMaps {
  CUtensorMap m[kTcMaxPeers];   // tensor maps of the gathered output on rank 0..W-1 (box {64, 1, 32, 1}, 128B swizzle)
};

struct TcSmemBarriers {
  uint64_t q_full[2];
  uint64_t q_fixed[2];
  uint64_t k_full[kTcMaxStages], k_empty[kTcMaxStages];
  uint64_t v_full[kTcMaxStages], v_empty[kTcMaxStages];
  uint64_t q_empty[2];   // the item's last S MMA has read Q tile t: the next item's Q may land
  uint64_t s_full[2];
  uint64_t p_full[2];
  uint64_t o_full[2];
  uint64_t turn[2];      // turn[t]: warpgroup t may run its exponential phase (the other one finished its own)
  uint64_t o_empty[2];   // the softmax warpgroup has read O_t out of TMEM: the next item's P.V may overwrite it
  uint32_t tmem_base;
  uint32_t pad;
};

constexpr int kTcOStageBytes = 32 * 128;  // per softmax warp: 32 rows x 64 bf16, one 128B-swizzled TMA store box

__host__ __device__ inline size_t tc_smem_bytes(int regions, int nk, int nv, int o_stage) {
  return (size_t)(2 + nk + nv) * regions * kTcRegionBytes + (o_stage ? 8 * kTcOStageBytes : 0) +
         1024 /*alignment slack*/ + sizeof(TcSmemBarriers);
}

// Stage rows [row0, row0+128) x hd of a row-strided bf16 matrix into a 128B-swizzled K-major tile (the layout a
// (64 x 128) SWIZZLE_128B TMA box would produce) with asynchronous 4- or 8-byte copies (cp.async, LDGSTS): nothing is
// staged through registers, so a loader thread keeps a whole tile (and the next one) in flight.  Rows >= T and
// columns [hd, hd_pad) are zero-filled (src-size 0).  VW = 32-bit words per copy: 2 when rows are 8-byte aligned.
// The caller commits the group and, once it has landed, makes it visible to the async proxy before signalling.
template <int NT, int VW>
__device__ __forceinline__ void cpasync_stage_tile(uint32_t dst, const __nv_bfloat16* src, long long stride_t, int row0,
                                                   int T, int hd, int hd_pad, int tid) {
  const int upr = hd_pad / (2 * VW);  // copy units per staged row
  const int hu = hd / (2 * VW);       // valid units per row
  // unit index idx = r * upr + u walks tid, tid + NT, ...; (r, u) advance incrementally (no per-copy division)
  const int dq = NT / upr, dr = NT % upr;
  int r = tid / upr, u = tid - r * upr;
  while (r < 128) {
    const bool ok = (row0 + r) < T && u < hu;
    const __nv_bfloat16* g = ok ? src + (long long)(row0 + r) * stride_t + 2 * VW * u : src;
    const uint32_t w = (uint32_t)(u * VW);   // first 32-bit word of the unit within the row
    const uint32_t wi = w & 31u;
    const uint32_t off = (w >> 5) * kTcRegionBytes + (uint32_t)r * 128u + (((wi >> 2) ^ ((uint32_t)r & 7u)) << 4) + ((wi & 3u) << 2);
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
__device__ __forceinline__ void cpasync_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpasync_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// Debug timeline: one thread per role of block 0 appends (tag, SM clock) records to its own lane of the buffer
// (plain stores, no atomics, so the probe costs a few cycles): trace[role][i] = {tag, clock}, trace_cap records/role.
#if defined(VATS_ENABLE_TRACE)
struct TcTracer {
  unsigned long long* base;
  int n, cap;
  __device__ __forceinline__ TcTracer(const TcParams& P, int role)
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
struct TcTracer {  // production build: the probes compile to nothing (build with -DVATS_ENABLE_TRACE to record)
  __device__ __forceinline__ TcTracer(const TcParams&, int) {}
  __device__ __forceinline__ void operator()(unsigned) {}
};
#endif

// One work item: (sequence n, KV group g, head pair, 128-token query block).
struct TcWork {
  int n, g, q0, head0;
  int t_first, n_tiles;  // KV tile range (n_tiles <= 0: nothing to attend)
  bool active1;          // the pair's second head exists
};

// n / d for n < 2^31 with a host-computed (multiplier, shift) pair — an integer division costs ~125 cycles on the
// single-lane roles, and every role decodes every work item.
__host__ __device__ __forceinline__ void tc_fastdiv(unsigned n, const unsigned (&magic)[2], unsigned d, unsigned* q,
                                                    unsigned* r) {
#if defined(__CUDA_ARCH__)
  const unsigned quo = d != 1u ? __umulhi(n, magic[0]) >> magic[1] : n;
#else
  const unsigned quo = d != 1u ? (unsigned)(((unsigned long long)n * magic[0]) >> 32) >> magic[1] : n;
#endif
  *q = quo;
  *r = n - quo * d;
}
inline void tc_find_divisor(unsigned d, unsigned (&magic)[2]) {
  if (d <= 1u) {
    magic[0] = 0u;
    magic[1] = 0u;
    return;
  }
  unsigned lg = 0;
  while ((1ull << lg) < d) ++lg;
  const unsigned p = 31 + lg;
  magic[0] = (unsigned)(((1ull << p) + d - 1) / d);
  magic[1] = p - 32;
}

__device__ __forceinline__ TcWork tc_decode_work(const TcParams& P, int w) {
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
    k.n_tiles = (a.Tk + kTcBlockN - 1) >> 7;
  } else {
    int t_last;
    tile_range(a.mask, k.q0, kTcBlockM, kTcBlockN, &k.t_first, &t_last);
    k.n_tiles = t_last - k.t_first + 1;
  }
  return k;
}

// kLdg = false: Q / K / V arrive by TMA.  kLdg = true: at least one of them is not TMA-addressable (row starts only
// 4-byte aligned) and all three are staged with LDG loaders instead.  Two instantiations rather than a run-time
// switch: the per-item code of a short-sequence launch otherwise thrashes the instruction cache.
template <bool kLdg>
__global__ void __launch_bounds__(kTcThreads, 1)
prefill_tc_kernel(const TcParams P, const __grid_constant__ CUtensorMap tmap_q,
                  const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                  const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ TcPeerMaps peer_maps) {
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
  const uint32_t sO = sV + (uint32_t)P.nv * tile_bytes;   // 8 x kTcOStageBytes when P.o_stage
  TcSmemBarriers* bars = reinterpret_cast<TcSmemBarriers*>(smem_raw + (base - raw) + (size_t)(2 + P.nk + P.nv) * tile_bytes +
                                                           (P.o_stage ? 8 * kTcOStageBytes : 0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup
  if (warp == 8 && lane == 0) {
    if (!kLdg) {
      prefetch_tmap(&tmap_q);
      prefetch_tmap(&tmap_k);
      prefetch_tmap(&tmap_v);
    }
    if (P.o_stage == 1 && P.peers == 0) prefetch_tmap(&tmap_o);
    for (int rk = 0; rk < P.peers; ++rk) prefetch_tmap(&peer_maps.m[rk]);
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bars->q_full[t]), 1);
      mbar_init(smem_u32(&bars->q_fixed[t]), 128);
      mbar_init(smem_u32(&bars->q_empty[t]), 1);
      mbar_init(smem_u32(&bars->s_full[t]), 1);
      mbar_init(smem_u32(&bars->p_full[t]), 128);
      mbar_init(smem_u32(&bars->o_full[t]), 1);
      mbar_init(smem_u32(&bars->o_empty[t]), 128);
      mbar_init(smem_u32(&bars->turn[t]), 128);
    }
    const uint32_t kv_arrivals = kLdg ? kTcLoaderThreads : 1;
    for (int s = 0; s < kTcMaxStages; ++s) {
      mbar_init(smem_u32(&bars->k_full[s]), kv_arrivals);
      mbar_init(smem_u32(&bars->k_empty[s]), 1);
      mbar_init(smem_u32(&bars->v_full[s]), kv_arrivals);
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
    // =========================================================== warpgroup 2: producers (warps 8-10) + MMA issuer (11)
    setmaxnreg_dec<80>();
  if (warp <= 10) {
    if (!kLdg) {
      // ---- TMA: three issuing lanes — warp 8 feeds the K ring, warp 9 the V ring, warp 10 the Q tiles.  One thread
      //      needs ~300 cycles per box (tools/micro/tma_bw.cu), so a single producer would serialise the 8 boxes an
      //      item starts with; the rings run continuously across items.
      if (lane == 0) {
        TcTracer trace(P, 0);
        if (warp == 10) {
          {
            uint32_t qn[2] = {0u, 0u};      // items in which tile t took part so far
            for (int w = blockIdx.x; w < P.num_work; w += gridDim.x) {
              const TcWork wk = tc_decode_work(P, w);
              if (wk.n_tiles <= 0) continue;
              for (int t = 0; t < 2; ++t) {
                if (t == 1 && !wk.active1) break;
                VATS_TC_PRODUCER_WAIT(smem_u32(&bars->q_empty[t]), (qn[t] & 1u) ^ 1u);
                const uint32_t bar = smem_u32(&bars->q_full[t]);
                mbar_expect_tx(bar, tile_bytes);
                for (int c = 0; c < P.regions; ++c)
                  tma_load_4d(sQ + t * tile_bytes + c * kTcRegionBytes, &tmap_q, bar, 64 * c, wk.head0 + t, wk.q0, wk.n);
                ++qn[t];
              }
              trace(0x300);
            }
          }
        } else {
          // warp 8: K, warp 9: V
          const bool is_k = warp == 8;
          const CUtensorMap* tmap = is_k ? &tmap_k : &tmap_v;
          uint64_t* full = is_k ? bars->k_full : bars->v_full;
          uint64_t* empty = is_k ? bars->k_empty : bars->v_empty;
          const uint32_t ring = is_k ? sK : sV;
          const int depth = is_k ? P.nk : P.nv;
          int slot = 0;
          uint32_t ph = 0u;
          for (int w = blockIdx.x; w < P.num_work; w += gridDim.x) {
            const TcWork wk = tc_decode_work(P, w);
            for (int j = 0; j < wk.n_tiles; ++j) {
              const int k0 = (wk.t_first + j) * kTcBlockN;
              VATS_TC_PRODUCER_WAIT(smem_u32(&empty[slot]), ph ^ 1u);
              if (is_k) trace(0x310 + (j & 15));
              const uint32_t bar = smem_u32(&full[slot]);
              mbar_expect_tx(bar, tile_bytes);
              for (int c = 0; c < P.regions; ++c)
                tma_load_4d(ring + slot * tile_bytes + c * kTcRegionBytes, tmap, bar, 64 * c, wk.g, k0, wk.n);
              if (++slot == depth) { slot = 0; ph ^= 1u; }
            }
          }
        }
      }
    } else {
      // ---- cp.async staging: 96 threads copy each K / V tile into the swizzled layout.  A thread signals a tile
      //      (wait_group -> fence.proxy.async -> arrive) only after issuing the copies of the next one, so two tiles
      //      are in flight per thread.
      const int ltid = threadIdx.x - 8 * 32;
      int ks = 0, vs = 0;
      uint32_t kph = 0u, vph = 0u;
      uint32_t pending = 0u;   // full barrier of the tile whose copies were committed last (0 = none)
      auto stage = [&](uint32_t dst, const __nv_bfloat16* base_ptr, long long stride_t, int k0, uint32_t full_bar) {
        if (P.ldg_vec == 2)
          cpasync_stage_tile<kTcLoaderThreads, 2>(dst, base_ptr, stride_t, k0, a.Tk, a.hd, P.hd_pad, ltid);
        else
          cpasync_stage_tile<kTcLoaderThreads, 1>(dst, base_ptr, stride_t, k0, a.Tk, a.hd, P.hd_pad, ltid);
        cpasync_commit();
        if (pending != 0u) {
          cpasync_wait<1>();
          fence_proxy_async_smem();
          mbar_arrive(pending);
        }
        pending = full_bar;
      };
      for (int w = blockIdx.x; w < P.num_work; w += gridDim.x) {
        const TcWork wk = tc_decode_work(P, w);
        if (wk.n_tiles <= 0) continue;
        const __nv_bfloat16* kbase = a.k + wk.n * a.ks_n + (long long)wk.g * a.ks_h;
        const __nv_bfloat16* vbase = a.v + wk.n * a.vs_n + (long long)wk.g * a.vs_h;
        for (int j = 0; j < wk.n_tiles; ++j) {
          const int k0 = (wk.t_first + j) * kTcBlockN;
          mbar_wait(smem_u32(&bars->k_empty[ks]), kph ^ 1u);
          stage(sK + (uint32_t)ks * tile_bytes, kbase, a.ks_t, k0, smem_u32(&bars->k_full[ks]));
          if (++ks == P.nk) { ks = 0; kph ^= 1u; }
          mbar_wait(smem_u32(&bars->v_empty[vs]), vph ^ 1u);
          stage(sV + (uint32_t)vs * tile_bytes, vbase, a.vs_t, k0, smem_u32(&bars->v_full[vs]));
          if (++vs == P.nv) { vs = 0; vph ^= 1u; }
        }
      }
      if (pending != 0u) {
        cpasync_wait<0>();
        fence_proxy_async_smem();
        mbar_arrive(pending);
      }
    }
  } else {
    // =========================================================== MMA issuer
    {
      // every lane runs this block; `leader` marks the one lane whose tcgen05 instructions take effect
      const uint32_t leader = elect_one() ? 1u : 0u;
      TcTracer trace(P, 1);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);  // tell the compiler it is warp-uniform
      const uint32_t idesc_s = make_idesc_bf16(kTcBlockM, kTcBlockN, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(kTcBlockM, P.hd_pad, 0, 1);
      const int ksteps = P.hd_pad / 16;

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
      const uint32_t tS0 = tmem_u, tS1 = tmem_u + 128, tO0 = tmem_u + 256, tO1 = tmem_u + 384;

      // K-step ks reads 32 bytes further along the 128-byte swizzled row; after 4 steps it moves to the next
      // 64-element region.  Offsets are in descriptor units (16 B).  (A fully unrolled variant with immediate
      // offsets measured no better on B200.)
      auto issue_s = [&](uint32_t q_lo, uint32_t k_lo, uint32_t d_tmem) {
        constexpr uint32_t R = kTcRegionBytes >> 4;
        uint32_t acc = 0u;
        for (int ks = 0; ks < ksteps; ++ks) {
          mma_ss_lohi(d_tmem, q_lo, hi_k, k_lo, hi_k, idesc_s, acc, leader);
          acc = 1u;
          const uint32_t step = ((ks & 3) == 3) ? (R - 6u) : 2u;
          q_lo += step;
          k_lo += step;
        }
      };
      auto issue_pv = [&](uint32_t p_tmem, uint32_t v_lo, uint32_t d_tmem, uint32_t acc) {
#pragma unroll
        for (int ks = 0; ks < kTcBlockN / 16; ++ks) {
          mma_ts_lohi(d_tmem, p_tmem + ks * 8, v_lo + ks * (2048 >> 4), hi_v, idesc_o, acc, leader);
          acc = 1u;
        }
      };

      int ks = 0, vs = 0;                 // ring slot of the next K tile / of the current V tile
      uint32_t kph = 0u, vph = 0u;
      uint32_t pc[2] = {0u, 0u};          // P phases consumed per tile (running across items)
      uint32_t qn[2] = {0u, 0u};          // items in which tile t took part so far
      const uint32_t p_bar0 = smem_u32(&bars->p_full[0]), p_bar1 = smem_u32(&bars->p_full[1]);
      const uint32_t s_bar0 = smem_u32(&bars->s_full[0]), s_bar1 = smem_u32(&bars->s_full[1]);
      for (int w = blockIdx.x; w < P.num_work; w += gridDim.x) {
        const TcWork wk = tc_decode_work(P, w);
        if (wk.n_tiles <= 0) continue;
        const bool active1 = wk.active1;
        const int n_tiles = wk.n_tiles;
        // ---- Q tiles of this item, first K tile, S(0) for both heads
        mbar_wait(smem_u32(kLdg ? &bars->q_fixed[0] : &bars->q_full[0]), qn[0] & 1u);
        if (active1) mbar_wait(smem_u32(kLdg ? &bars->q_fixed[1] : &bars->q_full[1]), qn[1] & 1u);
        if (leader) trace(0x100);
        mbar_wait(smem_u32(&bars->k_full[ks]), kph);
        if (leader) trace(0x101);
        tc_fence_after();
        issue_s(q_lo0, k_lo_base + (uint32_t)ks * tile_step, tS0);
        tc_commit_pred(s_bar0, leader);
        if (n_tiles == 1) tc_commit_pred(smem_u32(&bars->q_empty[0]), leader);
        if (active1) {
          issue_s(q_lo1, k_lo_base + (uint32_t)ks * tile_step, tS1);
          tc_commit_pred(s_bar1, leader);
          if (n_tiles == 1) tc_commit_pred(smem_u32(&bars->q_empty[1]), leader);
        }
        tc_commit_pred(smem_u32(&bars->k_empty[ks]), leader);
        if (++ks == P.nk) { ks = 0; kph ^= 1u; }

        for (int j = 0; j < n_tiles; ++j) {
          const bool has_next = (j + 1) < n_tiles;
          const bool last_s = (j + 2) == n_tiles;   // the S issued in this iteration is the item's last read of Q
          const uint32_t v_lo = v_lo_base + (uint32_t)vs * tile_step;
          const uint32_t k_lo = k_lo_base + (uint32_t)ks * tile_step;
          mbar_wait(smem_u32(&bars->v_full[vs]), vph);
          if (leader) trace(0x110 + (j & 15));
          // ---- tile 0
          mbar_wait(p_bar0, pc[0] & 1u);
          if (leader) trace(0x120 + (j & 15));
          ++pc[0];
          if (j == 0) mbar_wait(smem_u32(&bars->o_empty[0]), (qn[0] & 1u) ^ 1u);  // previous item's O_0 was read out
          tc_fence_after();
          issue_pv(tS0, v_lo, tO0, j > 0 ? 1u : 0u);
          if (has_next) {
            mbar_wait(smem_u32(&bars->k_full[ks]), kph);
            tc_fence_after();
            issue_s(q_lo0, k_lo, tS0);
            tc_commit_pred(s_bar0, leader);
            if (last_s) tc_commit_pred(smem_u32(&bars->q_empty[0]), leader);
          } else {
            // tile 0 is complete: let its warpgroup start the epilogue now instead of behind tile 1's last P.V (it
            // then runs ahead of tile 1 by that much, which also takes the two exponential phases apart)
            tc_commit_pred(smem_u32(&bars->o_full[0]), leader);
          }
          // ---- tile 1
          if (active1) {
            mbar_wait(p_bar1, pc[1] & 1u);
            if (leader) trace(0x130 + (j & 15));
            ++pc[1];
            if (j == 0) mbar_wait(smem_u32(&bars->o_empty[1]), (qn[1] & 1u) ^ 1u);
            tc_fence_after();
            issue_pv(tS1, v_lo, tO1, j > 0 ? 1u : 0u);
            if (has_next) {
              issue_s(q_lo1, k_lo, tS1);
              tc_commit_pred(s_bar1, leader);
              if (last_s) tc_commit_pred(smem_u32(&bars->q_empty[1]), leader);
            }
          }
          tc_commit_pred(smem_u32(&bars->v_empty[vs]), leader);
          if (++vs == P.nv) { vs = 0; vph ^= 1u; }
          if (has_next) {
            tc_commit_pred(smem_u32(&bars->k_empty[ks]), leader);
            if (++ks == P.nk) { ks = 0; kph ^= 1u; }
          }
        }
        if (leader) trace(0x140);
        ++qn[0];
        if (active1) {
          tc_commit_pred(smem_u32(&bars->o_full[1]), leader);
          ++qn[1];
        }
      }
    }
  }
  } else {
    // =========================================================== softmax warpgroups (tile t = warp / 4)
    setmaxnreg_inc<208>();
    const int t = warp >> 2;
    const int r = threadIdx.x & 127;           // row within the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_base + (uint32_t)t * 128;
    const uint32_t tO = tmem + lane_base + 256 + (uint32_t)t * 128;
    TcTracer trace(P, 2 + t);
    uint32_t sc = 0u;   // S phases consumed by this tile (running across items)
    uint32_t tk = 0u;   // exponential phases run in turn with the other warpgroup so far
    uint32_t qn = 0u;   // items in which this tile took part so far

    for (int w = blockIdx.x; w < P.num_work; w += gridDim.x) {
    const TcWork wk = tc_decode_work(P, w);
    const int n = wk.n, q0 = wk.q0, t_first = wk.t_first, n_tiles = wk.n_tiles;
    const int tok = q0 + r;
    const int head = wk.head0 + t;
    const bool tile_active = (t == 0) || wk.active1;

    float l_run = 0.f;
    float m_used = -INFINITY;  // reference maximum in scaled-log2 units; -inf = not set yet
    // KV tiles [full_first, full_last] are allowed for every row of the block: no per-element predicate there
    int full_first, full_last;
    if (P.no_band) {
      full_first = 0;
      full_last = (a.Tk >> 7) - 1;
    } else {
      int q_last = q0 + kTcBlockM - 1;
      if (q_last > a.Tq - 1) q_last = a.Tq - 1;
      long long lo = key_lo(a.mask, q_last);   // tightest lower bound
      long long hi = key_hi(a.mask, q0);       // tightest upper bound (already <= Tk - 1)
      if (lo < 0) lo = 0;
      full_first = lo > (long long)a.Tk ? (a.Tk >> 7) + 1 : ((int)lo + kTcBlockN - 1) >> 7;
      full_last = hi < 0 ? -1 : (((int)hi + 1) >> 7) - 1;
    }

    if (tile_active && n_tiles > 0) {
      if (kLdg) {
        // the 128 threads of this warpgroup stage their own Q tile, once the previous item's MMAs are done with it
        mbar_wait(smem_u32(&bars->q_empty[t]), (qn & 1u) ^ 1u);
        const __nv_bfloat16* qbase = a.q + n * a.qs_n + (long long)head * a.qs_h;
        if (P.ldg_vec == 2)
          cpasync_stage_tile<128, 2>(sQ + (uint32_t)t * tile_bytes, qbase, a.qs_t, q0, a.Tq, a.hd, P.hd_pad, r);
        else
          cpasync_stage_tile<128, 1>(sQ + (uint32_t)t * tile_bytes, qbase, a.qs_t, q0, a.Tq, a.hd, P.hd_pad, r);
        cpasync_commit();
        cpasync_wait<0>();
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&bars->q_fixed[t]));
      }

      for (int j = 0; j < n_tiles; ++j) {
        const int tile = t_first + j;
        const int k0 = tile * kTcBlockN;
        if (r == 0) trace(0x200 + (j & 15));
        mbar_wait(smem_u32(&bars->s_full[t]), sc & 1u);
        if (r == 0) trace(0x210 + (j & 15));
        ++sc;
        tc_fence_after();
        uint32_t sr[128];
        tmem_ld_32x32b_x32(tS + 0, sr + 0);
        tmem_ld_32x32b_x32(tS + 32, sr + 32);
        tmem_ld_32x32b_x32(tS + 64, sr + 64);
        tmem_ld_32x32b_x32(tS + 96, sr + 96);
        tmem_ld_wait();
        if (r == 0) trace(0x250 + (j & 15));

        // ---- predicate (edge tiles only)
        const bool full = tile >= full_first && tile <= full_last;  // == tile_is_full(a.mask, tile, q0, 128, 128)
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
          // allowed columns = [lo_c, hi_c] & k_valid bits, as four 32-bit words: two instructions per element below
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int l = lo_c - 32 * w, h = hi_c - 32 * w;
            const uint32_t ml = l <= 0 ? 0xffffffffu : (l >= 32 ? 0u : 0xffffffffu << l);
            const uint32_t mh = h >= 31 ? 0xffffffffu : (h < 0 ? 0u : 0xffffffffu >> (31 - h));
            kbits[w] &= ml & mh;
          }
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (!((kbits[c >> 5] >> (c & 31)) & 1u)) sr[c] = 0xff800000u;  // -inf
        }

        // ---- row max of this tile (raw logits), in-thread.  With bounded logits (qk-norm: |q.k| <= bound) the bound
        //      itself is the reference point: softmax is shift-invariant, so the 128 FMNMX per row and every rescale
        //      of O disappear (m_used is set once and never moves).
        float mt;
        if (P.bounded) {
          mt = P.bound_log2;
        } else {
          float mx[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) mx[i] = __uint_as_float(sr[i]);
#pragma unroll
          for (int c = 8; c < 128; c += 8)
#pragma unroll
            for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], __uint_as_float(sr[c + i]));
          mt = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7]))) *
               a.scale_log2;  // scaled-log2 units (scale > 0)
        }

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

        // ---- the exponential phase is the MUFU-bound part: when both tiles are live the two warpgroups take turns
        //      in it, which staggers them by half a period — one tile's softmax then overlaps the other tile's MMAs
        //      instead of both softmaxes (and then both MMA batches) running in lock-step.
        if (r == 0) trace(0x260 + (j & 15));
        const bool ordered = P.order_softmax && wk.active1;
        if (ordered) mbar_wait(smem_u32(&bars->turn[t]), t == 0 ? ((tk & 1u) ^ 1u) : (tk & 1u));
        // ---- p = exp2(s*scale_log2 - m_used), row sum, pack to bf16, write P over S.
        //      The scale/subtract and the row sums run as packed f32x2 operations (FFMA2 / FADD2), the exponentials on
        //      the MUFU pipe (ex2(-inf) is exactly 0 for masked keys).  Two alternatives were built and measured on
        //      B200 — half of the exponentials as a degree-3 polynomial on the FMA pipe, and packed-fp16 ex2 — and
        //      both were slower: this phase is issue/latency-bound, not MUFU-bound.
        const float mref = (m_used == -INFINITY) ? 0.f : m_used;
        const float2 sc2 = make_float2(a.scale_log2, a.scale_log2);
        const float2 nm2 = make_float2(-mref, -mref);
        float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
        uint32_t pk[64];
        if (a.Tk - k0 <= 80) {
          // last KV tile of a sequence with at most 80 live keys (ViT, 196 tokens: 68): columns past the end of the
          // sequence are masked for every row, so their exponentials are exactly 0 — a second, shorter straight-line
          // version of the loop (per-group predication inside one loop destroyed its instruction-level parallelism:
          // cfg5 605 instead of 1 000 TFLOP/s)
#pragma unroll
          for (int c = 0; c < 80; c += 4) {
            const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), sc2, nm2);
            const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])), sc2, nm2);
            const float2 p0 = make_float2(ex2(x0.x), ex2(x0.y));
            const float2 p1 = make_float2(ex2(x1.x), ex2(x1.y));
            sum_a = __fadd2_rn(sum_a, p0);
            sum_b = __fadd2_rn(sum_b, p1);
            pk[c >> 1] = pack_bf16x2(p0.x, p0.y);
            pk[(c >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
          }
#pragma unroll
          for (int c = 40; c < 64; ++c) pk[c] = 0u;
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
        if (ordered) {
          mbar_arrive(smem_u32(&bars->turn[t ^ 1]));
          ++tk;
        }
        if (r == 0) trace(0x270 + (j & 15));
        tmem_st_32x32b_x32(tS + 0, pk + 0);
        tmem_st_32x32b_x32(tS + 32, pk + 32);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->p_full[t]));
        if (r == 0) trace(0x220 + (j & 15));
      }
    }

    // ---- epilogue: O / l -> bf16 -> global.  The accumulator is handed back (o_empty) as soon as its last chunk is
    //      out of TMEM, so the next item's MMAs overlap this item's write-out.
    //      (tcgen05.ld is warp-collective: every lane runs the loads; only the stores are predicated on the row.)
    if (tile_active) {
      const bool do_store = tok < a.Tq;
      bool qok = true;
      if (do_store && a.q_valid != nullptr) qok = a.q_valid[(long long)n * a.Tq + tok] != 0;
      const float inv = (qok && l_run > 0.f && n_tiles > 0) ? 1.f / l_run : 0.f;
      if (n_tiles > 0) {
        if (r == 0) trace(0x230);
        mbar_wait(smem_u32(&bars->o_full[t]), qn & 1u);
        if (r == 0) trace(0x231);
        tc_fence_after();
      }
      if (P.o_stage) {
        // Each warp transposes its 32 rows through a private 4 KB staging tile (64 columns at a time, written in the
        // 128B-swizzle pattern): rows are 2-4 KB apart in global memory, so per-thread row stores would cost the LSU
        // one transaction per 16 bytes.  The tile then leaves with one TMA tile store issued by lane 0 (TMA clips
        // rows >= Tq and columns >= hd) or, when O rows are only 4-byte aligned, with coalesced 32-bit stores.
        const uint32_t stage = sO + (uint32_t)warp * kTcOStageBytes;
        const int row_w = q0 + (warp & 3) * 32;  // first token of this warp's 32 rows
        // the whole accumulator row leaves TMEM in one batch of loads (one wait), and the accumulator is handed back
        // to the MMA warp before anything is written out
        uint32_t acc[128];
#pragma unroll
        for (int pc16 = 0; pc16 < 8; ++pc16) tmem_ld_32x32b_x16(tO + pc16 * 16, acc + pc16 * 16);   // (all 128 columns of
        tmem_ld_wait();                                       // the tile's O region exist; those >= hd_pad are never stored)
        if (n_tiles > 0) {
          tc_fence_before();
          mbar_arrive(smem_u32(&bars->o_empty[t]));
          ++qn;
        }
        const bool have = n_tiles > 0;   // an item with nothing to attend writes zeros (TMEM then holds stale values)
#pragma unroll
        for (int cbi = 0; cbi < 2; ++cbi) {
          const int cb = cbi * 64;
          if (cb >= P.hd_pad) continue;
          // the previous store of this warp must have finished reading the staging tile
          if (P.o_stage == 1 && lane == 0) bulk_wait_group_read0();
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (cb + u * 8 < P.hd_pad) {
              uint32_t x = pack_bf16x2(__uint_as_float(acc[cb + 8 * u + 0]) * inv, __uint_as_float(acc[cb + 8 * u + 1]) * inv);
              uint32_t y = pack_bf16x2(__uint_as_float(acc[cb + 8 * u + 2]) * inv, __uint_as_float(acc[cb + 8 * u + 3]) * inv);
              uint32_t z = pack_bf16x2(__uint_as_float(acc[cb + 8 * u + 4]) * inv, __uint_as_float(acc[cb + 8 * u + 5]) * inv);
              uint32_t w = pack_bf16x2(__uint_as_float(acc[cb + 8 * u + 6]) * inv, __uint_as_float(acc[cb + 8 * u + 7]) * inv);
              if (!have) x = y = z = w = 0u;
              const uint32_t dst = stage + (uint32_t)lane * 128u + (((uint32_t)u ^ ((uint32_t)lane & 7u)) << 4);
              // (columns >= hd_pad of the last chunk are never written; TMA clips everything >= hd)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
            }
          }
          if (P.o_stage == 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && row_w < a.Tq) {
              if (P.peers == 0) {
                tma_store_4d(&tmap_o, stage, cb, head, row_w, n);
              } else {
                // the gather, fused: the tile goes to every rank's copy of the gathered output while it is still in
                // shared memory — no second pass over O, no collective kernel, no SM taken from the attention
                int rk = P.peer_first;
                for (int i = 0; i < P.peers; ++i) {
                  tma_store_4d(&peer_maps.m[rk], stage, cb, head + P.head_off, row_w, n + P.seq_off);
                  if (++rk == P.peers) rk = 0;
                }
              }
              bulk_commit_group();
            }
          } else {
            __syncwarp();
            // word index f = row * wv + w walks lane, lane + 32, ...: consecutive lanes write consecutive words of a row
            const int wv = (a.hd - cb >= 64 ? 64 : a.hd - cb) >> 1;   // valid 32-bit words per row in this chunk
            if (wv > 0) {
              uint32_t* obase = reinterpret_cast<uint32_t*>(a.o + n * a.os_n + (long long)head * a.os_h + cb);
              const int dq = 32 / wv, dr = 32 % wv;
              int row = lane / wv, w = lane - row * wv;
              const int rows_ok = a.Tq - row_w;   // rows of this warp below the end of the sequence
              while (row < 32) {
                if (row < rows_ok) {
                  const uint32_t src = stage + (uint32_t)row * 128u + ((((uint32_t)w >> 2) ^ ((uint32_t)row & 7u)) << 4) +
                                       (((uint32_t)w & 3u) << 2);
                  uint32_t val;
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(val) : "r"(src) : "memory");
                  obase[(((long long)(row_w + row) * a.os_t) >> 1) + w] = val;
                }
                w += dr;
                row += dq;
                if (w >= wv) {
                  w -= wv;
                  ++row;
                }
              }
            }
            __syncwarp();   // the tile is rewritten by the next chunk
          }
        }
      } else {
        __nv_bfloat16* orow = a.o + n * a.os_n + (long long)tok * a.os_t + (long long)head * a.os_h;
#pragma unroll 1
        for (int cb = 0; cb < P.hd_pad; cb += 32) {  // 32 accumulator columns at a time (keeps the row out of local memory)
          uint32_t tmp[32];
          if (n_tiles > 0) {
            tmem_ld_32x32b_x16(tO + cb, tmp);
            if (cb + 16 < P.hd_pad) tmem_ld_32x32b_x16(tO + cb + 16, tmp + 16);
            tmem_ld_wait();
            if (cb + 32 >= P.hd_pad) {   // last chunk is out of TMEM: hand the accumulator back before storing
              tc_fence_before();
              mbar_arrive(smem_u32(&bars->o_empty[t]));
              ++qn;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) tmp[i] = 0u;
          }
          if (do_store) {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
              w[i] = pack_bf16x2(__uint_as_float(tmp[2 * i]) * inv, __uint_as_float(tmp[2 * i + 1]) * inv);
            if (P.o_vec16) {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (cb + u * 8 < a.hd)   // hd % 8 == 0 here: 16-byte units are all-in or all-out
                  *reinterpret_cast<uint4*>(orow + cb + u * 8) = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
            } else {
              const bool al4 = (reinterpret_cast<uintptr_t>(orow) & 3u) == 0;   // even hd/strides: pairs are 4-byte aligned
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int e = cb + 2 * i;
                if (e + 1 < a.hd && al4) {
                  *reinterpret_cast<uint32_t*>(orow + e) = w[i];
                } else {
                  if (e < a.hd) reinterpret_cast<uint16_t*>(orow)[e] = (uint16_t)(w[i] & 0xffffu);
                  if (e + 1 < a.hd) reinterpret_cast<uint16_t*>(orow)[e + 1] = (uint16_t)(w[i] >> 16);
                }
              }
            }
          }
        }
      }
      __syncwarp();
      if (r == 0) trace(0x240);
    }
    }  // work items
  }

  // ---- teardown
  if (P.o_stage == 1 && warp < 8 && lane == 0) bulk_wait_group0();  // staged O tiles must be out before the CTA's smem goes away
  tc_fence_before();
  __syncthreads();
  if (warp == 11) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace vats

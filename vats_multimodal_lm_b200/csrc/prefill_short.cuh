// prefill_short.cuh — GQA attention for very short sequences (Tk <= 32), e.g. the ViT-3D temporal pass:
// 12 544 sequences of 8 tokens, 32 heads / 8 KV groups, head_dim 66 (reference vit_3d/optimized_attention.py:393-430,
// reached through _grouped_query_attention :185-348).
//
// At 6 flop/byte the pass is HBM-bound: the whole job is to stream every sequence through an SM once, with enough
// bytes in flight, and to keep the arithmetic out of the way.  A 128 x 128 tcgen05 tile would be 94 % padding; a
// CUDA-core version of this kernel was measured shared-memory-bound (every K / V word re-read per query row: 69 % of
// the LSU wavefront peak at 0.37 ms).  So the contractions run on warp-level tensor-core MMAs (mma.sync m16n8k16,
// bf16 in / fp32 out), whose fragments read each staged word once per 16 rows.
//
//   Persistent CTAs (two per SM; 8 compute warps + 1 data-movement warp) walk the sequences round-robin over a ring
//   of shared-memory stages: while the warps work on sequence i, the Q / K / V blocks of the next one are landing.
//     kBulk = true   each of Q[n], K[n], V[n] is one contiguous, 16-byte aligned block: three 1-D bulk copies (TMA)
//                    issued by one thread, completion on an mbarrier; O[n] leaves as one bulk store when it is
//                    contiguous too.  The shared-memory image is the global one (row pitch = head_dim / 2 words).
//     kBulk = false  any 4-byte aligned row layout: one warp per row issues 4-byte cp.async copies (completion on the
//                    same mbarrier through cp.async.mbarrier.arrive), rows padded to an odd pitch.
//   One warp = one (KV group, block of 32 query rows of that group) unit; a row is a (token, head of the group) pair,
//   so the K / V fragments are shared by all H/G heads (GQA reuse in registers):
//     S = Q K^T    A fragments straight from the staged Q rows, B fragments from the staged K rows (32-bit loads)
//     softmax      exact, fp32, all keys in one pass; row max / sum over the 4 lanes that hold a row (shuffles);
//                  P is normalised, rounded to bf16 and re-used in registers as the A fragment of the second MMA
//     O = P V      needs V with two consecutive KEYS per 32-bit word: each warp first transposes its group's V into a
//                  private scratch, already in fragment order
//   The output rows overwrite the warp's own Q rows in shared memory and leave coalesced.
//
// Needs 4-byte aligned rows (even head_dim and strides); everything else goes to prefill_simt.cuh.
#pragma once
#include "mask.cuh"
#include "prefill_simt.cuh"  // PrefillParams
#include "ptx.cuh"

namespace vats {

constexpr int kShortMaxThreads = 288;   // 8 compute warps + the data-movement warp
constexpr int kShortMaxStages = 4;

struct ShortParams {
  PrefillParams a;
  int hd2;        // head_dim / 2: 32-bit words per row
  int pitch;      // words per staged row
  int q_rows;     // tq_chunk * H: query rows of a full work item
  int tq_chunk;   // query tokens per work item (== Tq unless the Q block of a sequence does not fit a stage)
  int chunks;     // ceil(Tq / tq_chunk): work items per sequence; item = sequence * chunks + chunk
  long long num_items;   // N * chunks
  int kv_rows;    // Tk * G
  int q_words;    // staged words per sequence and stage: Q/O block, K block, V block (multiples of 4)
  int kv_words;
  int vt_words;   // per compute warp: nss * 32 * KMAX / 2 words of transposed V (its group's B operand of P.V)
  int nss;        // ceil(hd2 / 16): 16-word (32-column) super-steps of head_dim
  int m_rows;     // tq_chunk * hpg: query rows per KV group in a full work item
  int m_blocks;   // ceil(m_rows / 32): 32-row blocks per group; units = G * m_blocks, one warp each
  int o_bulk;     // O[n] is one contiguous 16-byte aligned block: bulk store
  int stages;     // shared-memory ring depth (2..4) and how many sequences ahead the loads run (1 .. stages - 1)
  int dist;
  int no_mask;    // every (query, key) pair is allowed: skip the predicate
  unsigned div_H[2], div_G[2], div_hpg[2], div_mb[2], div_hd2[2], div_chunks[2];  // magic numbers (tc_fastdiv)
  unsigned long long* trace;   // debug (-DVATS_ENABLE_TRACE): thread 0 of block 0 appends (tag, clock64) pairs
  int trace_cap;
};

__device__ __forceinline__ void short_fastdiv(unsigned n, const unsigned (&magic)[2], unsigned d, unsigned* q, unsigned* r) {
  const unsigned quo = d != 1u ? __umulhi(n, magic[0]) >> magic[1] : n;
  *q = quo;
  *r = n - quo * d;
}

// ring of raw stages + one transposed-V scratch per compute warp + the mbarriers
__host__ __device__ inline size_t short_smem_bytes(int q_words, int kv_words, int vt_words, int stages, int cwarps) {
  return (size_t)stages * ((size_t)q_words + 2 * (size_t)kv_words) * 4 + (size_t)cwarps * vt_words * 4 + 64;
}

namespace ptx {
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
// the executing thread's earlier cp.async copies arrive on the barrier when they land (counts as one expected arrival)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_16816_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
}  // namespace ptx

// KMAX = keys per pass (8, 16 or 32 >= Tk)
template <int KMAX, bool kBulk>
__global__ void __launch_bounds__(kShortMaxThreads, 2) prefill_short_kernel(const ShortParams P) {
  using namespace ptx;
  constexpr int NT = KMAX / 8;          // n-tiles of S
  constexpr int KK = (KMAX + 15) / 16;  // k-steps of P.V
  constexpr int VW = KMAX / 2;          // words (key pairs) per Vt column
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PrefillParams& a = P.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const uint32_t stage_words = (uint32_t)P.q_words + 2u * (uint32_t)P.kv_words;
  uint32_t* smem_w = reinterpret_cast<uint32_t*>(smem_raw);
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int cwarps = nwarps - 1;          // compute warps; the last warp only moves data (bulk loads / stores)
  const bool io_lane = warp == cwarps && lane == 0;
  uint32_t* sVt = smem_w + (size_t)P.stages * stage_words + (size_t)(warp < cwarps ? warp : 0) * P.vt_words;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_w + (size_t)P.stages * stage_words + (size_t)cwarps * P.vt_words);

  if (tid == 0) {
    for (int s = 0; s < P.stages; ++s) mbar_init(smem_u32(&full[s]), kBulk ? 1 : nthreads);
    fence_mbar_init();
  }
  __syncthreads();

  // work item -> (sequence n, first query token t0, query tokens nt)
  auto item_of = [&](long long item, long long* n, int* t0, int* ntok) {
    long long seq = item;
    int c = 0;
    if (P.chunks > 1) {
      unsigned qq, rr;
      short_fastdiv((unsigned)item, P.div_chunks, (unsigned)P.chunks, &qq, &rr);
      seq = qq;
      c = (int)rr;
    }
    *n = seq;
    *t0 = c * P.tq_chunk;
    *ntok = min(P.tq_chunk, a.Tq - c * P.tq_chunk);
  };

  auto issue_load = [&](long long item, int s) {
    long long n;
    int t0, ntok;
    item_of(item, &n, &t0, &ntok);
    const int q_rows = ntok * a.H;
    const uint32_t sq = smem_u32(smem_w + (size_t)s * stage_words);
    const uint32_t sk = sq + (uint32_t)P.q_words * 4u, sv = sk + (uint32_t)P.kv_words * 4u;
    const uint32_t bar = smem_u32(&full[s]);
    if (kBulk) {
      if (io_lane) {
        const uint32_t qb = (uint32_t)q_rows * (uint32_t)P.hd2 * 4u, kb = (uint32_t)P.kv_rows * (uint32_t)P.hd2 * 4u;
        mbar_expect_tx(bar, qb + 2u * kb);
        bulk_load_1d(sq, a.q + n * a.qs_n + (long long)t0 * a.qs_t, qb, bar);
        bulk_load_1d(sk, a.k + n * a.ks_n, kb, bar);
        bulk_load_1d(sv, a.v + n * a.vs_n, kb, bar);
      }
    } else {
      // one warp per row; lanes walk the row's 32-bit words
      for (int row = warp; row < q_rows; row += nwarps) {
        unsigned t, h;
        short_fastdiv((unsigned)row, P.div_H, (unsigned)a.H, &t, &h);
        const __nv_bfloat16* src = a.q + n * a.qs_n + (long long)(t0 + (int)t) * a.qs_t + (long long)h * a.qs_h;
        for (int w = lane; w < P.hd2; w += 32) cp_async_4(sq + (uint32_t)(row * P.pitch + w) * 4u, src + 2 * w);
      }
      for (int row = warp; row < P.kv_rows; row += nwarps) {
        unsigned j, g;
        short_fastdiv((unsigned)row, P.div_G, (unsigned)a.G, &j, &g);
        const __nv_bfloat16* ksrc = a.k + n * a.ks_n + (long long)j * a.ks_t + (long long)g * a.ks_h;
        const __nv_bfloat16* vsrc = a.v + n * a.vs_n + (long long)j * a.vs_t + (long long)g * a.vs_h;
        for (int w = lane; w < P.hd2; w += 32) {
          cp_async_4(sk + (uint32_t)(row * P.pitch + w) * 4u, ksrc + 2 * w);
          cp_async_4(sv + (uint32_t)(row * P.pitch + w) * 4u, vsrc + 2 * w);
        }
      }
      cp_async_mbar_arrive_noinc(bar);
    }
  };

  // prologue: the first `dist` sequences of this CTA
  for (int d = 0; d < P.dist; ++d) {
    const long long i0 = (long long)blockIdx.x + (long long)d * gridDim.x;
    if (i0 < P.num_items) issue_load(i0, d);
  }
#if defined(VATS_ENABLE_TRACE)
  int trace_n = 0;
#define VATS_SHORT_TRACE(tag)                                                                   \
  if (P.trace != nullptr && blockIdx.x == 0 && tid == 0 && trace_n < P.trace_cap) {            \
    P.trace[2 * trace_n] = (tag);                                                               \
    P.trace[2 * trace_n + 1] = (unsigned long long)clock64();                                   \
    ++trace_n;                                                                                  \
  }
#else
#define VATS_SHORT_TRACE(tag)
#endif
  int it = 0, s = 0, sp = P.dist % P.stages;   // stage of this iteration / of the sequence `dist` ahead
  uint32_t ph = 0u;
  const int units = a.G * P.m_blocks;
  const uint32_t all_keys = a.Tk >= 32 ? 0xffffffffu : (1u << a.Tk) - 1u;
  for (long long item = blockIdx.x; item < P.num_items; item += gridDim.x, ++it) {
    long long n;
    int t0, ntok;
    item_of(item, &n, &t0, &ntok);
    const int q_rows = ntok * a.H;       // rows of this item (the last chunk of a sequence may be shorter)
    const int m_rows = ntok * a.hpg;
    // ---- prefetch the sequence `dist` ahead; its stage was last used by iteration it + dist - stages, whose bulk
    //      store must have finished reading shared memory: at most stages - dist - 1 younger stores may still be pending
    {
      const long long nn = item + (long long)P.dist * gridDim.x;
      if (nn < P.num_items) {
        if (kBulk && P.o_bulk && io_lane) {
          const int pend = P.stages - P.dist - 1;
          if (pend <= 0) bulk_wait_group_read<0>();
          else if (pend == 1) bulk_wait_group_read<1>();
          else bulk_wait_group_read<2>();
        }
        issue_load(nn, sp);
      }
      if (++sp == P.stages) sp = 0;
    }
    VATS_SHORT_TRACE(1)
    if (warp < cwarps || !kBulk) mbar_wait(smem_u32(&full[s]), ph);
    VATS_SHORT_TRACE(2)

    uint32_t* sQ = smem_w + (size_t)s * stage_words;
    const uint32_t* sK = sQ + P.q_words;
    const uint32_t* sV = sK + P.kv_words;

    // ---- one compute warp per (group g, 32-row block mb): rows r = mb*32 + mt*16 + {gq, gq+8}, r -> (token r / hpg,
    //      head r % hpg).  The staged rows of one group all start in the same four banks (row pitch = 33 words, token
    //      stride = 0 mod 32), so the fragments are NOT loaded in the canonical k order: within a 16-word super-step
    //      lane tq takes words 4*tq .. 4*tq+3 (two MMAs; the contraction does not care about the order of k as long
    //      as A and B agree), and the n-tiles of P.V interleave the same way.  That turns 8-way bank conflicts into
    //      2-way ones.
    for (int u = warp; u < units && warp < cwarps; u += cwarps) {
      unsigned g, mb;
      short_fastdiv((unsigned)u, P.div_mb, (unsigned)P.m_blocks, &g, &mb);

      // Vt[nt2][n = gq][key pair]: B operand of P.V for this group, in fragment order; keys >= Tk are zero.
      // Word w of a V row holds columns 2w, 2w+1; w = 16*A + 4*t + c  ->  n-tile nt2 = 4*A + c, n = 2*t + {0,1}.
      __syncwarp();
      for (int x = lane; x < P.hd2 * VW; x += 32) {
        const int jp = x % VW, w = x / VW;
        const int j0 = 2 * jp, j1 = 2 * jp + 1;
        const uint32_t v0 = j0 < a.Tk ? sV[(j0 * a.G + (int)g) * P.pitch + w] : 0u;
        const uint32_t v1 = j1 < a.Tk ? sV[(j1 * a.G + (int)g) * P.pitch + w] : 0u;
        const int nt2 = 4 * (w >> 4) + (w & 3), n0 = 2 * ((w & 15) >> 2);
        uint32_t* dst = sVt + (size_t)((nt2 * 8 + n0) * VW + jp);
        dst[0] = (v0 & 0xffffu) | (v1 << 16);
        dst[VW] = (v0 >> 16) | (v1 & 0xffff0000u);
      }
      __syncwarp();

      uint32_t* qrow[2][2];
      uint32_t allow[2][2];
      bool rvalid[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int r = (int)mb * 32 + mt * 16 + hf * 8 + gq;
          rvalid[mt][hf] = r < m_rows;
          unsigned tok, hh;
          short_fastdiv((unsigned)(rvalid[mt][hf] ? r : 0), P.div_hpg, (unsigned)a.hpg, &tok, &hh);
          qrow[mt][hf] = sQ + (size_t)((int)tok * a.H + (int)g * a.hpg + (int)hh) * P.pitch;
          uint32_t al = all_keys;
          if (!P.no_mask) {
            al = 0u;
#pragma unroll
            for (int j = 0; j < KMAX; ++j) {
              bool ok = j < a.Tk && allowed_geom(a.mask, t0 + (int)tok, j);
              if (ok && a.k_valid != nullptr) ok = a.k_valid[n * a.Tk + j] != 0;
              al |= ok ? (1u << j) : 0u;
            }
            if (a.q_valid != nullptr && a.q_valid[n * a.Tq + t0 + tok] == 0) al = 0u;
          }
          allow[mt][hf] = al;
        }

      // S = Q K^T
      float sacc[2][NT][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) sacc[mt][nt][i] = 0.f;
      const uint32_t* krow[NT];
      bool kvalid[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int key = nt * 8 + gq;
        kvalid[nt] = key < a.Tk;
        krow[nt] = sK + (size_t)((kvalid[nt] ? key : 0) * a.G + (int)g) * P.pitch;
      }
      for (int ss = 0; ss < P.nss; ++ss) {
        const int wb = ss * 16 + 4 * tq;
        uint32_t af[2][2][4], bf[2][NT][2];   // [mma of the super-step]...
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool ok = wb + c < P.hd2;   // words past head_dim: zero in both operands
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            af[c >> 1][mt][(c & 1) * 2 + 0] = ok ? qrow[mt][0][wb + c] : 0u;
            af[c >> 1][mt][(c & 1) * 2 + 1] = ok ? qrow[mt][1][wb + c] : 0u;
          }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) bf[c >> 1][nt][c & 1] = (ok && kvalid[nt]) ? krow[nt][wb + c] : 0u;
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          if (ss * 16 + 2 * h2 < P.hd2) {   // warp-uniform: the second MMA of the last super-step may be all padding
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
              for (int nt = 0; nt < NT; ++nt)
                mma_16816_bf16(sacc[mt][nt], af[h2][mt][0], af[h2][mt][1], af[h2][mt][2], af[h2][mt][3], bf[h2][nt][0],
                               bf[h2][nt][1]);
          }
        }
      }

      // exact softmax; lane holds keys nt*8 + 2*tq + {0,1} of rows gq (c0,c1) and gq+8 (c2,c3); P normalised -> bf16
      uint32_t pa[2][KK][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float m = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int key = nt * 8 + 2 * tq + i;
              float x = sacc[mt][nt][hf * 2 + i] * a.scale_log2;
              x = ((allow[mt][hf] >> key) & 1u) ? x : -INFINITY;
              sacc[mt][nt][hf * 2 + i] = x;
              m = fmaxf(m, x);
            }
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          const float mref = (m == -INFINITY) ? 0.f : m;
          float l = 0.f;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float p = ex2(sacc[mt][nt][hf * 2 + i] - mref);   // 0 for masked keys
              sacc[mt][nt][hf * 2 + i] = p;
              l += p;
            }
          l += __shfl_xor_sync(0xffffffffu, l, 1);
          l += __shfl_xor_sync(0xffffffffu, l, 2);
          const float inv = l > 0.f ? 1.f / l : 0.f;   // rows with no allowed key (or q_valid == 0) give zeros
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            // A fragment of P.V: k-step nt / 2; registers {0,1} = keys 0-7 of the step (rows gq, gq+8), {2,3} = keys 8-15
            pa[mt][nt >> 1][(nt & 1) * 2 + hf] = pack_bf16x2(sacc[mt][nt][hf * 2] * inv, sacc[mt][nt][hf * 2 + 1] * inv);
          }
          if (NT & 1) pa[mt][KK - 1][2 + hf] = 0u;   // KMAX == 8: keys 8-15 of the only k-step do not exist
        }
      }

      // O = P V, one n-tile (8 columns = output words 16*A + 4*tq + c, tq = 0..3) at a time; the bf16 result
      // overwrites this warp's own Q rows
      for (int nt2 = 0; nt2 < 4 * P.nss; ++nt2) {
        if (16 * (nt2 >> 2) + (nt2 & 3) >= P.hd2) continue;   // warp-uniform: the tile is all padding
        float oacc[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) oacc[mt][i] = 0.f;
        const uint32_t* vcol = sVt + (size_t)(nt2 * 8 + gq) * VW;
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) {
          const uint32_t b0 = vcol[kk * 8 + tq];
          const uint32_t b1 = (KMAX >= 16) ? vcol[(kk * 8 + 4 + tq) % VW] : 0u;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
            mma_16816_bf16(oacc[mt], pa[mt][kk][0], pa[mt][kk][1], pa[mt][kk][2], pa[mt][kk][3], b0, b1);
        }
        const int w = 16 * (nt2 >> 2) + 4 * tq + (nt2 & 3);   // word (columns 2w, 2w+1) of the output row
        if (w < P.hd2) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            if (rvalid[mt][0]) qrow[mt][0][w] = pack_bf16x2(oacc[mt][0], oacc[mt][1]);
            if (rvalid[mt][1]) qrow[mt][1][w] = pack_bf16x2(oacc[mt][2], oacc[mt][3]);
          }
        }
      }
    }

    // ---- write O[n] from the stage's Q block
    VATS_SHORT_TRACE(4)
    if (kBulk && P.o_bulk) {
      fence_proxy_async_smem();
      __syncthreads();
      VATS_SHORT_TRACE(5)
      if (io_lane) {
        bulk_store_1d(a.o + n * a.os_n + (long long)t0 * a.os_t, smem_u32(sQ), (uint32_t)q_rows * (uint32_t)P.hd2 * 4u);
        bulk_commit_group();
      }
      VATS_SHORT_TRACE(6)
    } else {
      __syncthreads();
      for (int row = warp; row < q_rows; row += nwarps) {
        unsigned t, h;
        short_fastdiv((unsigned)row, P.div_H, (unsigned)a.H, &t, &h);
        uint32_t* dst = reinterpret_cast<uint32_t*>(a.o + n * a.os_n + (long long)(t0 + (int)t) * a.os_t + (long long)h * a.os_h);
        for (int w = lane; w < P.hd2; w += 32) dst[w] = sQ[row * P.pitch + w];
      }
      __syncthreads();   // the stage is refilled by a later iteration's prefetch
    }
    if (++s == P.stages) {
      s = 0;
      ph ^= 1u;
    }
  }
  if (kBulk && P.o_bulk && io_lane) bulk_wait_group0();
}

}  // namespace vats

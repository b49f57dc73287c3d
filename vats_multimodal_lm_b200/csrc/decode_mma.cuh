// decode_mma.cuh — KV-cache single-query GQA + sliding-window attention, TMA-fed and HBM-bound.
//
// Same contract as decode.cuh (reference src/optimized_attention.py:508-516 + 709-714, cache KVCache :169-287); this
// is the fast path for TMA-addressable caches (even head_dim <= 128, 16-byte aligned base and strides — the drop-in
// KVCache rounds the head stride up to 8 elements, so hd 60 qualifies).
//
// The kernel is a byte pump: decode is 4 flop/byte, so everything is organised around keeping 128 KB of K/V in
// flight per SM, spending almost no issue slots per byte, and never stalling the consumers at an item boundary.
//   * persistent CTAs (one per SM) walk a static list of work items (sequence b, KV group g, head batch, split);
//   * warps NCW..2NCW-1 are producers (one elected lane each, producer p feeds consumer p): per stage of SK keys
//     they issue the TMA boxes of K and V (128B-swizzled) into an mbarrier ring that runs across item boundaries —
//     the ring never drains between items;
//   * warps 0-3 are consumers; stage i belongs to warp i % 4 and the ring depth is a multiple of 4, so a slot is
//     always consumed by the same warp (required: a consumer waits on a slot by phase parity, which is only sound
//     if it saw the slot's previous phase complete itself — TMA completions of different slots are unordered) and
//     the warps never wait for each other inside an item.
//     Per stage a warp does  S[heads x 32] = Q K^T  and  O[heads x hd] += P V  with mma.sync.m16n8k16 (bf16 in, fp32
//     accumulate; the <= 8 query heads of the group are the M rows, padded to 16) — operands come straight from the
//     swizzled tiles with ldmatrix / ldmatrix.trans, conflict-free.  That is ~300 instructions per 16 KB of cache
//     instead of ~2000 on the FP32 pipe; the tensor pipe is <5 % busy, which is the point: the SM only moves bytes.
//   * online softmax per stage in registers, quad shuffles for the row max / sum;
//   * at the end of an item a consumer warp leaves its (m, l, O) in its scratch rows (mbarrier hand-off) and goes on;
//     warp 2*NCW, the flush warp, merges the warps' rows, writes bf16 (one split) or the fp32 partial to the workspace,
//     and the LAST split to reach a (b, g, head batch) unit's counter merges the splits — no second launch, and no
//     CTA barrier, global fence or atomic on the consumers' path.
#pragma once
#include "decode.cuh"
#include "ptx.cuh"

namespace vats {

constexpr int kDmChunkAlign = 32;  // split chunks are multiples of this many keys (covers both stage sizes)
constexpr int kDmMaxHeads = 8;  // query heads per item (M rows of the MMA tile that are live)

struct DecodeMmaParams {
  DecodeParams d;       // pointers, strides, chunk, num_splits, workspace (ws_acc, ws_ml)
  int hpg_tile;         // heads per item (<= 8)
  int num_items;        // B * G * head_batches * num_splits
  int stages;           // ring depth
  int* counters;        // [B * G * head_batches], zero on entry, left zero on exit
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// All consumer threads meet here; the warp is re-converged first (bar.sync is the .aligned form).
template <int NTHREADS>
__device__ __forceinline__ void consumer_barrier() {
  __syncwarp();
  asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
}

struct DmItem {
  int b, g, hb, split;
  int ks, nkeys, nst;  // first key, number of keys, number of 32-key stages
};

template <int SK>
__device__ __forceinline__ DmItem dm_decode_item(const DecodeMmaParams& P, int item) {
  const DecodeParams& d = P.d;
  DmItem it;
  it.split = item % d.num_splits;
  int u = item / d.num_splits;
  it.hb = u % d.head_batches;
  u /= d.head_batches;
  it.g = u % d.G;
  it.b = u / d.G;
  const int L = d.seq_lens[it.b];
  int lo = 0;
  if (d.left >= 0) lo = max(0, L - 1 - d.left);
  it.ks = lo + it.split * d.chunk;
  const int ke = min(L, it.ks + d.chunk);
  it.nkeys = max(0, ke - it.ks);
  it.nst = (it.nkeys + SK - 1) / SK;
  return it;
}

// HD: tile width = head dim rounded up to 16 / 32 / 64 / 128 (d.hd <= HD is the real head dim: TMA zero-fills the
// columns past it, e.g. hd 60 in a cache whose head stride is 64).  Shared memory: [stages][K halves | V halves][32 keys][128 B] + merge scratch.
template <int HD, int NCW, int SK>
__global__ void __launch_bounds__((2 * NCW + 1) * 32, 1)
decode_mma_kernel(const DecodeMmaParams P, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v) {
  using namespace ptx;
  constexpr int kDmConsumerWarps = NCW;
  constexpr int HALVES = (HD + 63) / 64;           // 64-element (128 B) swizzle regions per row
  constexpr int HALF_BYTES = SK * 128;             // one TMA box: SK keys x 128 B
  constexpr int NSUB = SK / 16;                    // 16-key sub-blocks (MMA k / n granularity) per stage
  constexpr int TILE_BYTES = HALVES * HALF_BYTES;  // K (or V) of one stage
  constexpr int STAGE_BYTES = 2 * TILE_BYTES;
  constexpr int KSTEPS = HD / 16;                  // MMA k-steps of Q K^T
  constexpr int NT_O = HD / 8;                     // n-tiles of the output accumulator
  constexpr int SCR_ROW = HD + 4;                  // fp32 words per scratch row (padding vs bank conflicts)

  extern __shared__ unsigned char smem_raw[];
  const DecodeParams& d = P.d;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* sbase = smem_raw + (base - raw);
  const uint32_t ring = base;
  float* scr = reinterpret_cast<float*>(sbase + (size_t)P.stages * STAGE_BYTES);  // [4 warps][8 rows][SCR_ROW]
  float* scr_ml = scr + kDmConsumerWarps * kDmMaxHeads * SCR_ROW;                 // [4][8][2]
  uint64_t* full = reinterpret_cast<uint64_t*>(scr_ml + kDmConsumerWarps * kDmMaxHeads * 2);
  uint64_t* empty = full + P.stages;
  uint64_t* scr_full = empty + P.stages;          // [NCW] consumer w has put its (m, l, O) of an item into its scratch rows
  uint64_t* scr_empty = scr_full + kDmConsumerWarps;   // [NCW] ... and the flush warp has read them

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int w = 0; w < kDmConsumerWarps; ++w) {
      mbar_init(smem_u32(&scr_full[w]), 1);
      mbar_init(smem_u32(&scr_empty[w]), 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == 2 * kDmConsumerWarps) {
    // =================================================================== flush warp: the end of every item, off the
    // consumers' critical path.  It merges the NCW warps' (m, l, O) from the scratch rows, writes the bf16 output
    // (one split) or the item's partial to the workspace, and — splits only — signals the unit's counter; the last
    // split to arrive merges them.  The gpu-scope fences and the atomic round trip (micro-seconds) used to stall
    // all consumer warps behind CTA barriers at every item end: 31 us of a 191 us launch.
    uint32_t fph = 0u;
    for (int item = blockIdx.x; item < P.num_items; item += gridDim.x) {
      const DmItem it = dm_decode_item<SK>(P, item);
      const int h0 = it.g * d.hpg + it.hb * P.hpg_tile;
      const int nh = min(P.hpg_tile, d.hpg - it.hb * P.hpg_tile);
      const int unit = (it.b * d.G + it.g) * d.head_batches + it.hb;
      for (int w = 0; w < kDmConsumerWarps; ++w) mbar_wait(smem_u32(&scr_full[w]), fph);
      fph ^= 1u;
      // lane owns output columns lane, lane + 32, ...; the head loop is unrolled so every load is in flight at once
      constexpr int CPL = (HD + 31) / 32;
#pragma unroll
      for (int h = 0; h < kDmMaxHeads; ++h) {
        if (h < nh) {
          float M = -INFINITY;
#pragma unroll
          for (int w = 0; w < kDmConsumerWarps; ++w) M = fmaxf(M, scr_ml[(w * kDmMaxHeads + h) * 2]);
          const float Mref = (M == -INFINITY) ? 0.f : M;
          float acc[CPL], L = 0.f;
#pragma unroll
          for (int c = 0; c < CPL; ++c) acc[c] = 0.f;
#pragma unroll
          for (int w = 0; w < kDmConsumerWarps; ++w) {
            const float mw = scr_ml[(w * kDmMaxHeads + h) * 2];
            const float f = (mw == -INFINITY) ? 0.f : ex2(mw - Mref);
            L += f * scr_ml[(w * kDmMaxHeads + h) * 2 + 1];
#pragma unroll
            for (int c = 0; c < CPL; ++c)
              if (c * 32 + lane < HD) acc[c] += f * scr[(w * kDmMaxHeads + h) * SCR_ROW + c * 32 + lane];
          }
          const int hq = h0 + h;
          if (d.num_splits == 1) {
            const float invL = L > 0.f ? 1.f / L : 0.f;
#pragma unroll
            for (int c = 0; c < CPL; ++c)
              if (c * 32 + lane < d.hd)
                d.o[it.b * d.os_b + (long long)hq * d.os_h + c * 32 + lane] = __float2bfloat16(acc[c] * invL);
          } else {
            const long long slot = (long long)(it.b * d.H + hq) * d.num_splits + it.split;
#pragma unroll
            for (int c = 0; c < CPL; ++c)
              if (c * 32 + lane < HD) d.ws_acc[slot * HD + c * 32 + lane] = acc[c];
            if (lane == 0) {
              d.ws_ml[slot * 2] = M;
              d.ws_ml[slot * 2 + 1] = L;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0)
        for (int w = 0; w < kDmConsumerWarps; ++w) mbar_arrive(smem_u32(&scr_empty[w]));   // the consumers may refill
      if (d.num_splits > 1) {
        // ---- the last split of this unit to arrive merges all of them.  One lane fences: __syncwarp orders the
        //      other lanes' partial stores before it and the gpu-scope fence is cumulative.
        int last = 0;
        if (lane == 0) {
          __threadfence();
          const int old = atomicAdd(&P.counters[unit], 1);
          __threadfence();
          last = (old == d.num_splits - 1) ? 1 : 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          float accm[kDmMaxHeads][CPL], Mh[kDmMaxHeads], Lh[kDmMaxHeads];
#pragma unroll
          for (int h = 0; h < kDmMaxHeads; ++h) {
            Mh[h] = -INFINITY;
            Lh[h] = 0.f;
#pragma unroll
            for (int c = 0; c < CPL; ++c) accm[h][c] = 0.f;
          }
          // two passes over the splits, every head's loads of a pass issued together (L2 latency paid once per pass)
          for (int sp = 0; sp < d.num_splits; ++sp) {
#pragma unroll
            for (int h = 0; h < kDmMaxHeads; ++h)
              if (h < nh) Mh[h] = fmaxf(Mh[h], __ldcg(d.ws_ml + ((long long)(it.b * d.H + h0 + h) * d.num_splits + sp) * 2));
          }
          for (int sp = 0; sp < d.num_splits; ++sp) {
            float2 ml[kDmMaxHeads];
            float v[kDmMaxHeads][CPL];
#pragma unroll
            for (int h = 0; h < kDmMaxHeads; ++h) {
              ml[h] = make_float2(-INFINITY, 0.f);
              const long long slot = (long long)(it.b * d.H + h0 + h) * d.num_splits + sp;
              if (h < nh) ml[h] = __ldcg(reinterpret_cast<const float2*>(d.ws_ml + slot * 2));
#pragma unroll
              for (int c = 0; c < CPL; ++c) {
                v[h][c] = 0.f;
                if (h < nh && c * 32 + lane < HD) v[h][c] = __ldcg(d.ws_acc + slot * HD + c * 32 + lane);
              }
            }
#pragma unroll
            for (int h = 0; h < kDmMaxHeads; ++h) {
              const float Mref = (Mh[h] == -INFINITY) ? 0.f : Mh[h];
              const float f = (ml[h].x == -INFINITY) ? 0.f : ex2(ml[h].x - Mref);
              Lh[h] += f * ml[h].y;
#pragma unroll
              for (int c = 0; c < CPL; ++c) accm[h][c] += f * v[h][c];
            }
          }
#pragma unroll
          for (int h = 0; h < kDmMaxHeads; ++h) {
            const float invL = Lh[h] > 0.f ? 1.f / Lh[h] : 0.f;
#pragma unroll
            for (int c = 0; c < CPL; ++c)
              if (h < nh && c * 32 + lane < d.hd)
                d.o[it.b * d.os_b + (long long)(h0 + h) * d.os_h + c * 32 + lane] = __float2bfloat16(accm[h][c] * invL);
          }
          if (lane == 0) P.counters[unit] = 0;  // leave the workspace ready for the next call
        }
      }
    }
    return;
  }

  if (warp >= kDmConsumerWarps) {
    // =================================================================== producers: warp NCW + p feeds consumer p.
    // One issuing thread needs ~300 cycles per TMA box (expect_tx, descriptor moves, barrier wait; measured with
    // tools/micro/tma_bw.cu), far less than the TMA unit can take: one producer per consumer keeps the HBM stream
    // full.  Ring slots / phases advance incrementally — an integer division costs this thread ~125 cycles.
    const int pw = warp - kDmConsumerWarps;
    if (lane == 0) {
      prefetch_tmap(&tmap_k);
      prefetch_tmap(&tmap_v);
      int slot = pw;        // slot of this producer's next stage (stages % NCW == 0: slot == pw mod NCW always)
      uint32_t ph = 0u;     // its phase parity
      int gs_mod = 0;       // (global stage index of the item's first stage) mod NCW
      for (int item = blockIdx.x; item < P.num_items; item += gridDim.x) {
        const DmItem it = dm_decode_item<SK>(P, item);
        for (int j = (pw - gs_mod) & (kDmConsumerWarps - 1); j < it.nst; j += kDmConsumerWarps) {
          mbar_wait(smem_u32(&empty[slot]), ph ^ 1u, 0x10000000u | (uint32_t)j);
          const uint32_t bar = smem_u32(&full[slot]);
          mbar_expect_tx(bar, STAGE_BYTES);
          const uint32_t dst = ring + (uint32_t)slot * STAGE_BYTES;
          const int k0 = it.ks + j * SK;
#pragma unroll
          for (int h = 0; h < HALVES; ++h) {
            tma_load_4d(dst + h * HALF_BYTES, &tmap_k, bar, 64 * h, it.g, k0, it.b);
            tma_load_4d(dst + TILE_BYTES + h * HALF_BYTES, &tmap_v, bar, 64 * h, it.g, k0, it.b);
          }
          slot += kDmConsumerWarps;
          if (slot >= P.stages) { slot -= P.stages; ph ^= 1u; }
        }
        gs_mod = (gs_mod + it.nst) & (kDmConsumerWarps - 1);
      }
    }
    return;
  }

  // ===================================================================== consumers
  const int quad = lane >> 2;   // MMA row of c0/c1 == query head within the item
  const int qlane = lane & 3;
  int slot = warp;      // ring slot of this warp's next stage
  uint32_t cph = 0u;    // its phase parity
  int gs_mod = 0;       // (global stage index of the item's first stage) mod NCW
  uint32_t sph = 0u;    // phase parity of this warp's scratch hand-off

  for (int item = blockIdx.x; item < P.num_items; item += gridDim.x) {
    const DmItem it = dm_decode_item<SK>(P, item);
    const int h0 = it.g * d.hpg + it.hb * P.hpg_tile;
    const int nh = min(P.hpg_tile, d.hpg - it.hb * P.hpg_tile);

    // ---- Q fragments (A operand, rows = heads; rows >= nh and rows 8..15 are zero)
    uint32_t qa[KSTEPS][2];
    {
      const __nv_bfloat16* qrow = d.q + it.b * d.qs_b + (long long)(h0 + quad) * d.qs_h;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        qa[ks][0] = 0u;
        qa[ks][1] = 0u;
        if (quad < nh && it.nst > 0) {
          // head dims below HD (60 in a 64-wide tile): the columns past hd are zero-filled by TMA in K / V and must
          // be zero (and unread) in Q
          if (HD == d.hd || ks * 16 + qlane * 2 < d.hd)
            qa[ks][0] = *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + qlane * 2);
          if (HD == d.hd || ks * 16 + 8 + qlane * 2 < d.hd)
            qa[ks][1] = *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8 + qlane * 2);
        }
      }
    }

    float oacc[NT_O][4];
#pragma unroll
    for (int n = 0; n < NT_O; ++n) oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f;
    float m_run = -INFINITY;  // scaled-log2 units, row `quad`
    float l_run = 0.f;        // this thread's share of the row sum (quad-reduced at the end of the item)

    for (int j = (warp - gs_mod) & (kDmConsumerWarps - 1); j < it.nst; j += kDmConsumerWarps) {
      const int s = slot;
      mbar_wait(smem_u32(&full[s]), cph, 0x20000000u | (uint32_t)j);
      slot += kDmConsumerWarps;
      if (slot >= P.stages) { slot -= P.stages; cph ^= 1u; }
      const uint32_t kt = ring + (uint32_t)s * STAGE_BYTES;
      const uint32_t vt = kt + TILE_BYTES;
      const int kcount = min(SK, it.nkeys - j * SK);

      // ---- S[16 x 32] = Q K^T : n-tile nt = keys [8 nt, 8 nt + 8)
      float sacc[2 * NSUB][4];
#pragma unroll
      for (int nt = 0; nt < 2 * NSUB; ++nt) sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
      {
        // ldmatrix.x4: matrices (nt even, k lo) (nt even, k hi) (nt odd, k lo) (nt odd, k hi) of a 16-key sub-block.
        // Loads are issued in batches of KB k-steps ahead of their MMAs so the smem latency is paid once per batch.
        const int mi = lane >> 3, rr = lane & 7;
        constexpr int KB = KSTEPS < 4 ? KSTEPS : 4;
#pragma unroll
        for (int kb = 0; kb < KSTEPS; kb += KB) {
          uint32_t bf[KB][NSUB][4];
#pragma unroll
          for (int kk = 0; kk < KB; ++kk) {
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
              const int ks = kb + kk;
              const int key = sub * 16 + (mi >> 1) * 8 + rr;
              const int chunk = ks * 2 + (mi & 1);  // 16-byte chunk of the key row
              const uint32_t addr = kt + (uint32_t)(chunk >> 3) * HALF_BYTES + (uint32_t)key * 128u +
                                    (uint32_t)(((chunk & 7) ^ (key & 7)) << 4);
              ldmatrix_x4(addr, bf[kk][sub][0], bf[kk][sub][1], bf[kk][sub][2], bf[kk][sub][3]);
            }
          }
#pragma unroll
          for (int kk = 0; kk < KB; ++kk) {
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
              const int ks = kb + kk;
              mma_bf16_16816(sacc[sub * 2 + 0], qa[ks][0], 0u, qa[ks][1], 0u, bf[kk][sub][0], bf[kk][sub][1]);
              mma_bf16_16816(sacc[sub * 2 + 1], qa[ks][0], 0u, qa[ks][1], 0u, bf[kk][sub][2], bf[kk][sub][3]);
            }
          }
        }
      }

      // ---- mask the tail of a partial stage, scale, online softmax for row `quad` (c0, c1 of every n-tile)
      float mt = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 2 * NSUB; ++nt) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int key = nt * 8 + qlane * 2 + i;
          float sv = sacc[nt][i] * d.scale_log2;
          sv = key < kcount ? sv : -INFINITY;
          sacc[nt][i] = sv;
          mt = fmaxf(mt, sv);
        }
      }
      mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 1));
      mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 2));
      const float m_new = fmaxf(m_run, mt);
      const float mref = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = (m_run == -INFINITY) ? 0.f : ex2(m_run - mref);
      m_run = m_new;
      l_run *= corr;
      if (__any_sync(0xffffffffu, corr != 1.f)) {
#pragma unroll
        for (int n = 0; n < NT_O; ++n) {
          oacc[n][0] *= corr;
          oacc[n][1] *= corr;
        }
      }
      uint32_t pa[NSUB][2];  // A fragments of P for the 16-key sub-blocks: [sub][k lo / k hi]
#pragma unroll
      for (int nt = 0; nt < 2 * NSUB; ++nt) {
        const float p0 = ex2(sacc[nt][0] - mref);
        const float p1 = ex2(sacc[nt][1] - mref);
        l_run += p0 + p1;
        pa[nt >> 1][nt & 1] = pack_bf16x2(p0, p1);
      }

      // ---- a partial stage may hold stale cache rows past the sequence end: make them finite (0 * NaN = NaN)
      if (kcount < SK) {
        for (int idx = lane; idx < (SK - kcount) * HALVES * 8; idx += 32) {
          const int row = kcount + idx / (HALVES * 8);
          const int c = idx % (HALVES * 8);
          *reinterpret_cast<uint4*>(sbase + (vt - base) + (size_t)(c >> 3) * HALF_BYTES + (size_t)row * 128 +
                                    (size_t)(((c & 7) ^ (row & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();  // these generic-proxy stores must be ordered before the slot's next TMA fill
        __syncwarp();
      }

      // ---- O[16 x HD] += P V : ldmatrix.x4.trans fetches (keys lo, n-tile 2c) (keys hi, 2c) (keys lo, 2c+1) (keys hi, 2c+1)
      {
        const int mi = lane >> 3, rr = lane & 7;
        constexpr int NPB = (NT_O / 2) < 8 ? (NT_O / 2) : 8;  // n-tile pairs per batch
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
#pragma unroll
          for (int nb = 0; nb < NT_O / 2; nb += NPB) {
            uint32_t bf[NPB][4];
#pragma unroll
            for (int i = 0; i < NPB; ++i) {
              const int np = nb + i;
              const int key = sub * 16 + (mi & 1) * 8 + rr;
              const int chunk = np * 2 + (mi >> 1);
              const uint32_t addr = vt + (uint32_t)(chunk >> 3) * HALF_BYTES + (uint32_t)key * 128u +
                                    (uint32_t)(((chunk & 7) ^ (key & 7)) << 4);
              ldmatrix_x4_trans(addr, bf[i][0], bf[i][1], bf[i][2], bf[i][3]);
            }
#pragma unroll
            for (int i = 0; i < NPB; ++i) {
              const int np = nb + i;
              mma_bf16_16816(oacc[np * 2 + 0], pa[sub][0], 0u, pa[sub][1], 0u, bf[i][0], bf[i][1]);
              mma_bf16_16816(oacc[np * 2 + 1], pa[sub][0], 0u, pa[sub][1], 0u, bf[i][2], bf[i][3]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
    }
    gs_mod = (gs_mod + it.nst) & (kDmConsumerWarps - 1);

    // ---- hand this warp's (m, l, O) of the item to the flush warp through its scratch rows and move on: no CTA-wide
    //      barrier, no global fence on this path
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
    mbar_wait(smem_u32(&scr_empty[warp]), sph ^ 1u);   // the previous item's rows have been read
    if (quad < kDmMaxHeads) {
      float* row = scr + (warp * kDmMaxHeads + quad) * SCR_ROW;
#pragma unroll
      for (int n = 0; n < NT_O; ++n) {
        row[n * 8 + qlane * 2] = oacc[n][0];
        row[n * 8 + qlane * 2 + 1] = oacc[n][1];
      }
      if (qlane == 0) {
        scr_ml[(warp * kDmMaxHeads + quad) * 2] = m_run;
        scr_ml[(warp * kDmMaxHeads + quad) * 2 + 1] = l_run;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&scr_full[warp]));
    sph ^= 1u;
  }
}

template <int HD, int NCW, int SK>
__host__ inline size_t decode_mma_smem_bytes(int stages) {
  constexpr int kDmConsumerWarps = NCW;
  constexpr int HALVES = (HD + 63) / 64;
  constexpr int STAGE_BYTES = 2 * HALVES * SK * 128;
  constexpr int SCR_ROW = HD + 4;
  return 1024 + (size_t)stages * STAGE_BYTES + (size_t)kDmConsumerWarps * kDmMaxHeads * (SCR_ROW + 2) * sizeof(float) +
         (size_t)stages * 16 + (size_t)2 * kDmConsumerWarps * 8 + 64;
}

}  // namespace vats

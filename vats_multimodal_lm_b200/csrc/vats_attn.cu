// vats_attn.cu — C-ABI entry points (include/vats_attn.h): validation, kernel choice, TMA descriptors, launches.
//
// No CPU fallback and no second backend: every path below ends in one of the three sm_100a kernels
// (prefill_tc_kernel, prefill_simt_kernel, decode_split_kernel [+ decode_combine_kernel]) or in an error.
#include "../../include/vats_attn.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>

#include "decode.cuh"
#include "decode_mma.cuh"
#include "mask.cuh"
#include "prefill_simt.cuh"
#include "prefill_tc.cuh"
#include "prefill_short.cuh"
#include "prefill_mid.cuh"
#include "decode_prepare.cuh"
#include "repack.cuh"
#include "backward.cuh"

namespace {

thread_local char g_err[512] = "";
thread_local int g_launches = 0;
thread_local int g_last_kernel = 0;   // VATS_LAUNCHED_* of the last successful launch on this thread
unsigned long long* g_trace = nullptr;  // debug timeline buffer (device memory), see vats_attn_debug_set_trace
int g_trace_cap = 0;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess) return fail(VATS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

// ---- device gate: sm_100 only
constexpr int kMaxDevices = 64;
thread_local int g_dev = 0;   // current device of the calling thread, refreshed by check_device() at every entry point

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: remember what was set per
// device, not per process — one thread may launch on cuda:0 and then on cuda:1.
struct SmemAttrCache {
  size_t set[kMaxDevices] = {};
};
template <typename Kernel>
cudaError_t ensure_dyn_smem(Kernel kernel, size_t smem, SmemAttrCache& cache) {
  const bool tracked = g_dev >= 0 && g_dev < kMaxDevices;
  if (tracked && smem <= cache.set[g_dev]) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && tracked) cache.set[g_dev] = smem;
  return e;
}

int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return fail(VATS_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
  g_dev = dev;
  static thread_local int cached_dev = -1;
  static thread_local int cached_ok = 0;
  if (cached_dev == dev) return cached_ok ? VATS_OK : fail(VATS_ERR_UNSUPPORTED, "device %d is not sm_100", dev);
  int major = 0, minor = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  cached_dev = dev;
  cached_ok = (major == 10);
  if (!cached_ok)
    return fail(VATS_ERR_UNSUPPORTED, "device %d is sm_%d%d; the kernels are built for sm_100a only", dev, major, minor);
  return VATS_OK;
}

// ---- cuTensorMapEncodeTiled through the runtime (libcuda is not linked at build time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// A [N, T, heads, hd] bf16 tensor (strides in elements, hd contiguous) as a 4-D tiled map, dims (hd, heads, T, N),
// box (64 x 1 x 128 x 1), 128-byte swizzle; columns past hd and rows past T are zero-filled by TMA.
// TMA needs a 16-byte aligned base and every stride a multiple of 16 bytes (8 elements) — and the box start address
// (base + coordinates) 16-byte aligned too, which rules out selecting a head through an unaligned inner coordinate.
// Head dims such as 60 or 66 therefore cannot use TMA; they go through the kernel's LDG staging path, which only
// needs 4-byte aligned rows (even strides).
enum class LoadMode { kNone, kTma, kLdg };

LoadMode plan_load(const void* ptr, int hd, const int64_t s[3]) {
  const bool tma = (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0 && s[0] % 8 == 0 && s[1] % 8 == 0 && s[2] % 8 == 0;
  if (tma) return LoadMode::kTma;
  const bool ldg = (reinterpret_cast<uintptr_t>(ptr) & 3u) == 0 && hd % 2 == 0 && s[0] % 2 == 0 && s[1] % 2 == 0 &&
                   s[2] % 2 == 0;
  return ldg ? LoadMode::kLdg : LoadMode::kNone;
}

int encode_map(CUtensorMap* map, const void* ptr, int N, int T, int heads, int hd, const int64_t s[3],
               int box_rows = 128, int box_heads = 1) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(VATS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[4] = {(cuuint64_t)hd, (cuuint64_t)heads, (cuuint64_t)T, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)s[2] * 2, (cuuint64_t)s[1] * 2, (cuuint64_t)s[0] * 2};
  // extent-1 dimensions may carry any stride in PyTorch; give TMA something legal
  if (heads == 1 || strides[0] == 0) strides[0] = (cuuint64_t)((hd + 7) / 8 * 8) * 2;
  if (T == 1 || strides[1] == 0) strides[1] = strides[0] * (cuuint64_t)heads;
  if (N == 1 || strides[2] == 0) strides[2] = strides[1] * (cuuint64_t)T;
  const cuuint32_t box[4] = {64, (cuuint32_t)box_heads, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS)
    return fail(VATS_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (CUresult %d) dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu)",
                (int)rc, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
                (unsigned long long)strides[2]);
  return VATS_OK;
}

// O as rows of 32-bit words for the resident-K/V kernel's untiled stores: dims (H*hd/2, Tq, N), box (pack*hd/2,
// 32/pack, 1), no swizzle — the warp's 32 (token, head) rows are 32/pack contiguous runs of pack*hd elements.
int encode_rows_map(CUtensorMap* map, const void* ptr, int N, int T, int H, int hd, const int64_t s[3], int pack) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(VATS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[3] = {(cuuint64_t)H * hd / 2, (cuuint64_t)T, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)s[1] * 2, (cuuint64_t)s[0] * 2};
  if (T == 1 || strides[0] == 0) strides[0] = (cuuint64_t)H * hd * 2;
  if (N == 1 || strides[1] == 0) strides[1] = strides[0] * (cuuint64_t)T;
  const cuuint32_t box[3] = {(cuuint32_t)(pack * hd / 2), (cuuint32_t)(32 / pack), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS)
    return fail(VATS_ERR_CUDA, "cuTensorMapEncodeTiled (row store map) failed (CUresult %d)", (int)rc);
  return VATS_OK;
}

int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;  // B200; also the answer on a box without a GPU (workspace sizing must not fail there)
  }
  return sms;
}

struct PrefillArgs {
  const void *q, *k, *v;
  void* o;
  const uint8_t *q_valid, *k_valid;
  int N, Tq, Tk, H, G, hd;
  const int64_t *qs, *ks, *vs, *os;
  float scale;
  int causal, left, right;
  void* ws = nullptr;       // caller-owned scratch (vats_attn_prefill_workspace_bytes), may be NULL
  size_t ws_bytes = 0;
  // fused output gather (vats_attn_prefill_gather): `o` is unused, tiles go to o_ranks[0..world)
  void* const* o_ranks = nullptr;
  int world = 0, rank = 0, seq_off = 0, head_off = 0, N_total = 0, H_total = 0;
  float logit_bound = 0.f;   // > 0: the caller guarantees |q.k| <= logit_bound for every (query, key) pair (qk-norm)
};

// bound on the scaled-log2 logits; 2 % head room for the bf16 rounding of unit-norm q / k
float bound_log2_of(const PrefillArgs& A) { return A.logit_bound * 1.02f * A.scale * 1.4426950408889634f; }

int validate_prefill(const PrefillArgs& A) {
  if (A.N < 0 || A.Tq < 0 || A.Tk < 0) return fail(VATS_ERR_INVALID_ARGUMENT, "negative size");
  if (A.H <= 0 || A.G <= 0 || A.hd <= 0) return fail(VATS_ERR_INVALID_ARGUMENT, "H, G, hd must be positive");
  if (A.H % A.G != 0)
    return fail(VATS_ERR_INVALID_ARGUMENT, "num_heads (%d) must be divisible by query_groups (%d)", A.H, A.G);
  if (!(A.scale > 0.f) || !std::isfinite(A.scale))
    return fail(VATS_ERR_INVALID_ARGUMENT, "scale must be a positive finite float (got %g)", (double)A.scale);
  if (!A.qs || !A.ks || !A.vs || !A.os) return fail(VATS_ERR_INVALID_ARGUMENT, "stride arrays must not be NULL");
  if (A.N > 0 && A.Tq > 0 && (!A.q || !A.o)) return fail(VATS_ERR_INVALID_ARGUMENT, "q / o must not be NULL");
  if (A.N > 0 && A.Tk > 0 && (!A.k || !A.v)) return fail(VATS_ERR_INVALID_ARGUMENT, "k / v must not be NULL");
  return VATS_OK;
}

struct TcPlan {
  LoadMode q, k, v;
};

bool tc_legal(const PrefillArgs& A, TcPlan* pl) {
  if (A.hd > 128 || A.Tk <= 0) return false;
  pl->q = plan_load(A.q, A.hd, A.qs);
  pl->k = plan_load(A.k, A.hd, A.ks);
  pl->v = plan_load(A.v, A.hd, A.vs);
  return pl->q != LoadMode::kNone && pl->k != LoadMode::kNone && pl->v != LoadMode::kNone;
}

void fill_common(vats::PrefillParams& p, const PrefillArgs& A) {
  p.q = reinterpret_cast<const __nv_bfloat16*>(A.q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(A.k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(A.v);
  p.o = reinterpret_cast<__nv_bfloat16*>(A.o);
  p.q_valid = A.q_valid;
  p.k_valid = A.k_valid;
  p.N = A.N; p.Tq = A.Tq; p.Tk = A.Tk; p.H = A.H; p.G = A.G; p.hd = A.hd;
  p.hpg = A.H / A.G;
  p.qs_n = A.qs[0]; p.qs_t = A.qs[1]; p.qs_h = A.qs[2];
  p.ks_n = A.ks[0]; p.ks_t = A.ks[1]; p.ks_h = A.ks[2];
  p.vs_n = A.vs[0]; p.vs_t = A.vs[1]; p.vs_h = A.vs[2];
  p.os_n = A.os[0]; p.os_t = A.os[1]; p.os_h = A.os[2];
  p.scale_log2 = A.scale * 1.4426950408889634f;
  p.mask.Tq = A.Tq; p.mask.Tk = A.Tk;
  p.mask.causal = A.causal ? 1 : 0;
  p.mask.left = A.left; p.mask.right = A.right;
}

int launch_simt(const PrefillArgs& A, cudaStream_t st) {
  if (A.hd > vats::kSimtMaxHd) return fail(VATS_ERR_UNSUPPORTED, "head_dim %d > %d is not supported", A.hd, vats::kSimtMaxHd);
  vats::PrefillParams p;
  fill_common(p, A);
  const int total_rows = A.Tq * p.hpg;
  const int rb = (total_rows + vats::kSimtRows - 1) / vats::kSimtRows;
  if (A.G > 65535 || A.N > 65535)
    return fail(VATS_ERR_UNSUPPORTED, "SIMT kernel grid limit: G and N must be <= 65535 (got G=%d N=%d)", A.G, A.N);
  dim3 grid(rb, A.G, A.N);
  const size_t smem = (size_t)(2 * vats::kSimtTileN + vats::kSimtRows) * (A.hd + 1) * sizeof(float);
  const int cpln = (A.hd + 31) / 32;
#define VATS_SIMT_CASE(C)                                                                                         \
  case C: {                                                                                                       \
    CUDA_TRY(cudaFuncSetAttribute(vats::prefill_simt_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                  (int)smem));                                                                    \
    vats::prefill_simt_kernel<C><<<grid, vats::kSimtWarps * 32, smem, st>>>(p);                                   \
  } break;
  switch (cpln) {
    VATS_SIMT_CASE(1)
    VATS_SIMT_CASE(2)
    VATS_SIMT_CASE(3)
    VATS_SIMT_CASE(4)
    VATS_SIMT_CASE(5)
    VATS_SIMT_CASE(6)
    VATS_SIMT_CASE(7)
    VATS_SIMT_CASE(8)
    default:
      return fail(VATS_ERR_UNSUPPORTED, "head_dim %d not supported by the SIMT kernel", A.hd);
  }
#undef VATS_SIMT_CASE
  CUDA_TRY(cudaGetLastError());
  g_launches = 1;
  g_last_kernel = VATS_LAUNCHED_PREFILL_SIMT;
  return VATS_OK;
}

// Short sequences (Tk <= 32) with 4-byte aligned rows: one CTA per sequence, see prefill_short.cuh.
// Returns -1 when the geometry does not qualify (the caller then uses the generic warp kernel).
int launch_short(const PrefillArgs& A, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("VATS_PREFILL_SHORT");  // tuning knob: 0 = always the generic warp kernel
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled || A.Tk > 32 || A.Tk < 1 || A.Tq < 1 || A.N < 1) return -1;
  if (plan_load(A.q, A.hd, A.qs) == LoadMode::kNone || plan_load(A.k, A.hd, A.ks) == LoadMode::kNone ||
      plan_load(A.v, A.hd, A.vs) == LoadMode::kNone || plan_load(A.o, A.hd, A.os) == LoadMode::kNone)
    return -1;
  vats::ShortParams P;
  std::memset(&P, 0, sizeof(P));
  fill_common(P.a, A);
  const int hpg = P.a.hpg;
  const int kmax = A.Tk <= 8 ? 8 : (A.Tk <= 16 ? 16 : 32);
  P.hd2 = A.hd / 2;
  const long long q_rows = (long long)A.Tq * A.H, kv_rows = (long long)A.Tk * A.G;
  // bulk (TMA 1-D) staging: every per-sequence block dense and 16-byte aligned
  auto dense16 = [&](const void* ptr, const int64_t* st3, int heads, long long rows) {
    return st3[2] == A.hd && st3[1] == (int64_t)heads * A.hd && (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0 &&
           (rows * A.hd * 2) % 16 == 0 && (A.N == 1 || (st3[0] * 2) % 16 == 0);
  };
  bool bulk = dense16(A.q, A.qs, A.H, q_rows) && dense16(A.k, A.ks, A.G, kv_rows) && dense16(A.v, A.vs, A.G, kv_rows);
  {
    static int bulk_env = -1;
    if (bulk_env < 0) {
      const char* e = getenv("VATS_PREFILL_SHORT_BULK");  // tuning knob: 0 = cp.async staging even for dense inputs
      bulk_env = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (!bulk_env) bulk = false;
  }
  bool o_bulk = bulk && dense16(A.o, A.os, A.H, q_rows);
  P.pitch = bulk ? P.hd2 : (P.hd2 | 1);
  P.kv_rows = (int)kv_rows;
  P.kv_words = ((int)kv_rows * P.pitch + 3) & ~3;
  P.nss = (P.hd2 + 15) / 16;
  P.vt_words = P.nss * 32 * (kmax / 2);
  // Work item = (sequence, chunk of query tokens): the whole sequence when its Q block fits a stage of a two-CTA-per-SM
  // launch, otherwise the largest token chunk that does (cross-attention: thousands of image tokens against 16 text
  // tokens).  Every item re-stages the sequence's K / V (a few KB).
  {
    const long long per_cta = (227 * 1024) / 2 - 1024 - 64 - 8LL * P.vt_words * 4;        // bytes, 8 compute warps
    const long long q_budget = per_cta / 2 / 4 - 2LL * P.kv_words;                         // words per stage for Q
    const long long row_words = (long long)A.H * P.pitch;
    long long tq = A.Tq;
    if (tq * row_words > q_budget) tq = q_budget / row_words;
    if (tq < 1) return -1;
    P.tq_chunk = (int)tq;
    P.chunks = (A.Tq + P.tq_chunk - 1) / P.tq_chunk;
    if (P.chunks > 1) {   // sub-blocks of tokens must stay 16-byte multiples for the bulk copies
      const bool tok16 = ((long long)A.H * A.hd * 2) % 16 == 0;
      if (!tok16) {
        if (bulk) return -1;   // (pitch was chosen for the bulk image; such shapes go to the generic kernel)
      }
      o_bulk = o_bulk && tok16;
    }
    if ((long long)A.N * P.chunks > 0x7fffffffLL) return -1;
    P.num_items = (long long)A.N * P.chunks;
  }
  P.o_bulk = o_bulk ? 1 : 0;
  P.q_rows = P.tq_chunk * A.H;
  P.q_words = (P.q_rows * P.pitch + 3) & ~3;
  P.m_rows = P.tq_chunk * hpg;
  P.m_blocks = (P.m_rows + 31) / 32;
  // launch shape: threads per CTA, ring depth, prefetch distance, CTAs per SM
  int threads = 288, stages = 2, dist = 1, per_sm = 2;
  {
    static int cfg[4] = {-1, 0, 0, 0};
    if (cfg[0] < 0) {
      const char* e = getenv("VATS_PREFILL_SHORT_CFG");  // tuning knob: "threads,stages,dist,ctas_per_sm"
      if (!e || sscanf(e, "%d,%d,%d,%d", &cfg[0], &cfg[1], &cfg[2], &cfg[3]) != 4) cfg[0] = 0;
    }
    if (cfg[0] > 0) {
      threads = cfg[0]; stages = cfg[1]; dist = cfg[2]; per_sm = cfg[3];
    }
  }
  if (threads != 160 && threads != 288) threads = 288;   // 4 or 8 compute warps + the data-movement warp
  const int cwarps = threads / 32 - 1;
  if (per_sm < 1 || per_sm > 2) per_sm = 2;
  if (stages > vats::kShortMaxStages) stages = vats::kShortMaxStages;
  auto fits = [&](int st_, int per) {
    return vats::short_smem_bytes(P.q_words, P.kv_words, P.vt_words, st_, cwarps) <= (size_t)(227 * 1024) / per - 1024;
  };
  if (!fits(2, per_sm)) per_sm = 1;   // large sequences: one CTA per SM
  while (stages > 2 && !fits(stages, per_sm)) --stages;
  if (stages < 2 || !fits(stages, per_sm)) return -1;
  if (dist >= stages) dist = stages - 1;
  if (dist < 1) dist = 1;
  P.stages = stages;
  P.dist = dist;
  const size_t smem = vats::short_smem_bytes(P.q_words, P.kv_words, P.vt_words, stages, cwarps);
  P.no_mask = (!A.causal && A.left < 0 && A.right < 0 && !A.q_valid && !A.k_valid) ? 1 : 0;
  vats::tc_find_divisor((unsigned)A.H, P.div_H);
  vats::tc_find_divisor((unsigned)A.G, P.div_G);
  vats::tc_find_divisor((unsigned)hpg, P.div_hpg);
  vats::tc_find_divisor((unsigned)P.m_blocks, P.div_mb);
  vats::tc_find_divisor((unsigned)P.hd2, P.div_hd2);
  vats::tc_find_divisor((unsigned)P.chunks, P.div_chunks);
  int grid = per_sm * sm_count();
  if ((long long)grid > P.num_items) grid = (int)P.num_items;
  P.trace = g_trace;
  P.trace_cap = g_trace_cap;
#define VATS_SHORT_LAUNCH(K, B)                                                                                    \
  {                                                                                                                \
    static thread_local SmemAttrCache set;                                                                         \
    CUDA_TRY(ensure_dyn_smem(vats::prefill_short_kernel<K, B>, smem, set));                                        \
    vats::prefill_short_kernel<K, B><<<(unsigned)grid, threads, smem, st>>>(P);                                    \
  }
#define VATS_SHORT_CASE(K)               \
  if (kmax == K) {                       \
    if (bulk) VATS_SHORT_LAUNCH(K, true) \
    else VATS_SHORT_LAUNCH(K, false)     \
  }
  VATS_SHORT_CASE(8)
  VATS_SHORT_CASE(16)
  VATS_SHORT_CASE(32)
#undef VATS_SHORT_CASE
#undef VATS_SHORT_LAUNCH
  CUDA_TRY(cudaGetLastError());
  g_launches = 1;
  g_last_kernel = VATS_LAUNCHED_PREFILL_SHORT;
  return VATS_OK;
}

int launch_tc(const PrefillArgs& A, const TcPlan& pl, cudaStream_t st);

// q / k / v with rows TMA cannot address: one streaming repack into the CALLER's scratch buffer (head stride rounded
// up to 8 elements), then the TMA-fed kernel on the copies.  Returns -1 if no (or too small a) scratch was given —
// the caller then uses the kernel's own cp.async staging variant.  The library allocates nothing.
size_t tc_repack_offsets(const PrefillArgs& A, const TcPlan& pl, size_t off[4]) {
  const int hd_pad = (A.hd + 7) / 8 * 8;
  const int Ts[3] = {A.Tq, A.Tk, A.Tk};
  const int heads[3] = {A.H, A.G, A.G};
  const LoadMode modes[3] = {pl.q, pl.k, pl.v};
  off[0] = 0;
  for (int i = 0; i < 3; ++i) {
    const size_t bytes = modes[i] == LoadMode::kLdg ? (size_t)A.N * Ts[i] * heads[i] * hd_pad * 2 : 0;
    off[i + 1] = off[i] + ((bytes + 255) & ~(size_t)255);
  }
  return off[3];
}

int launch_mid(const PrefillArgs& A, const TcPlan& pl, cudaStream_t st);

int launch_tc_repacked(const PrefillArgs& A, const TcPlan& pl, cudaStream_t st, bool mid = false) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("VATS_PREFILL_REPACK");  // tuning knob: 0 = stage inside the kernel instead
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled) return -1;
  const int hd_pad = (A.hd + 7) / 8 * 8;
  const void* src[3] = {A.q, A.k, A.v};
  const int64_t* str[3] = {A.qs, A.ks, A.vs};
  const int Ts[3] = {A.Tq, A.Tk, A.Tk};
  const int heads[3] = {A.H, A.G, A.G};
  const LoadMode modes[3] = {pl.q, pl.k, pl.v};
  size_t off[4];
  const size_t need = tc_repack_offsets(A, pl, off);
  if (!A.ws || A.ws_bytes < need || (reinterpret_cast<uintptr_t>(A.ws) & 15u) != 0) return -1;
  void* ws = A.ws;
  vats::RepackParams R;
  std::memset(&R, 0, sizeof(R));
  R.hd2 = A.hd / 2;
  R.hd_pad = hd_pad;
  int64_t new_str[3][3];
  PrefillArgs B = A;
  B.ws = nullptr;
  B.ws_bytes = 0;
  long long max_rows = 0;
  int nrep = 0;
  for (int i = 0; i < 3; ++i) {
    if (modes[i] != LoadMode::kLdg) continue;
    vats::RepackTensor& t = R.t[nrep++];
    t.src = reinterpret_cast<const __nv_bfloat16*>(src[i]);
    t.dst = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(ws) + off[i]);
    t.s_n = str[i][0]; t.s_t = str[i][1]; t.s_h = str[i][2];
    t.T = Ts[i]; t.heads = heads[i];
    t.rows = (long long)A.N * Ts[i] * heads[i];
    if (t.rows > max_rows) max_rows = t.rows;
    new_str[i][2] = hd_pad;
    new_str[i][1] = (int64_t)heads[i] * hd_pad;
    new_str[i][0] = (int64_t)Ts[i] * heads[i] * hd_pad;
    if (i == 0) { B.q = t.dst; B.qs = new_str[0]; }
    if (i == 1) { B.k = t.dst; B.ks = new_str[1]; }
    if (i == 2) { B.v = t.dst; B.vs = new_str[2]; }
  }
  static int chunk_env = -1;
  if (chunk_env < 0) {
    const char* e = getenv("VATS_REPACK_CHUNK");  // tuning knob: 0 = the row-per-warp copy kernel
    chunk_env = (e && atoi(e) == 0) ? 0 : 1;
  }
  const long long cpr = hd_pad / 8;
  if (chunk_env && max_rows * cpr < 0x7fffffffLL) {
    // one thread per 16-byte destination chunk (repack.cuh)
    vats::tc_find_divisor((unsigned)cpr, R.div_cpr);
    for (int i = 0; i < nrep; ++i) {
      vats::tc_find_divisor((unsigned)R.t[i].heads, R.t[i].div_heads);
      vats::tc_find_divisor((unsigned)R.t[i].T, R.t[i].div_T);
    }
    const long long per_block = (long long)vats::kRepackChunkThreads * vats::kRepackChunkUnroll;
    long long blocks = (max_rows * cpr + per_block - 1) / per_block;
    const long long cap = (long long)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    vats::repack_chunk_kernel<<<dim3((unsigned)blocks, (unsigned)nrep), vats::kRepackChunkThreads, 0, st>>>(R);
  } else {
    long long blocks = (max_rows + vats::kRepackWarps * vats::kRepackRows - 1) / (vats::kRepackWarps * vats::kRepackRows);
    const long long cap = (long long)sm_count() * 64;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    vats::repack_kernel<<<dim3((unsigned)blocks, (unsigned)nrep), vats::kRepackWarps * 32, 0, st>>>(R);
  }
  int rc = VATS_OK;
  if (cudaGetLastError() != cudaSuccess) {
    rc = fail(VATS_ERR_CUDA, "repack kernel launch failed");
  } else {
    TcPlan pl2{LoadMode::kTma, LoadMode::kTma, LoadMode::kTma};
    rc = mid ? launch_mid(B, pl2, st) : launch_tc(B, pl2, st);
  }
  if (rc == VATS_OK) g_launches += 1;
  return rc;
}

int launch_tc(const PrefillArgs& A, const TcPlan& pl, cudaStream_t st) {
  vats::TcParams P;
  std::memset(&P, 0, sizeof(P));
  fill_common(P.a, A);
  P.hd_pad = (A.hd + 15) / 16 * 16;
  P.regions = (P.hd_pad + 63) / 64;
  P.q_blocks = (A.Tq + vats::kTcBlockM - 1) / vats::kTcBlockM;
  P.pairs = (P.a.hpg + 1) / 2;
  P.no_band = (!A.causal && A.left < 0 && A.right < 0) ? 1 : 0;
  P.bounded = A.logit_bound > 0.f ? 1 : 0;
  P.bound_log2 = bound_log2_of(A);
  // one staging mode per launch: if any of q / k / v cannot be addressed by TMA, all three use the LDG loaders
  const bool any_ldg = pl.q == LoadMode::kLdg || pl.k == LoadMode::kLdg || pl.v == LoadMode::kLdg;
  if (any_ldg) {
    const int rc2 = launch_tc_repacked(A, pl, st);
    if (rc2 >= 0) return rc2;
  }
  {
    auto al8 = [&](const void* ptr, const int64_t* st3) {
      return (reinterpret_cast<uintptr_t>(ptr) & 7u) == 0 && st3[0] % 4 == 0 && st3[1] % 4 == 0 && st3[2] % 4 == 0;
    };
    P.ldg_vec = (A.hd % 4 == 0 && al8(A.q, A.qs) && al8(A.k, A.ks) && al8(A.v, A.vs)) ? 2 : 1;
  }
  P.o_vec16 = ((reinterpret_cast<uintptr_t>(A.world > 0 ? A.o_ranks[0] : A.o) & 15u) == 0 && A.os[0] % 8 == 0 && A.os[1] % 8 == 0 &&
               A.os[2] % 8 == 0 && A.hd % 8 == 0)
                  ? 1
                  : 0;
  {
    // The two softmax warpgroups take turns in the exponential phase when an item visits few KV tiles (they start
    // in lock-step there: cfg3 0.207 vs 0.225 ms); on long items they drift apart by themselves and the hand-over only
    // costs (cfg5 1 016 vs 1 035 TFLOP/s).
    static int order = -1;   // -1: by geometry
    static bool read = false;
    if (!read) {
      const char* e = getenv("VATS_PREFILL_ORDER_SOFTMAX");  // tuning knob: 0 / 1 force
      if (e) order = atoi(e) ? 1 : 0;
      read = true;
    }
    long long tiles = (A.Tk + vats::kTcBlockN - 1) / vats::kTcBlockN;
    if (A.left >= 0 && (A.causal || A.right >= 0)) {
      const long long band = (long long)A.left + (A.causal ? 0 : A.right) + vats::kTcBlockM;   // keys a query block can see
      const long long bt = band / vats::kTcBlockN + 2;
      if (bt < tiles) tiles = bt;
    }
    P.order_softmax = order >= 0 ? order : (tiles <= 8 ? 1 : 0);
  }
  // ring depths: fill the 227 KB of shared memory (also pins one CTA per SM, which owns all 512 TMEM columns)
  const int tile_bytes = P.regions * vats::kTcRegionBytes;
  const int budget = 227 * 1024 - 1024 - (int)sizeof(vats::TcSmemBarriers) - 2 * tile_bytes;
  // O leaves through per-warp staging tiles (32 KB of shared memory) whenever the K/V rings keep two slots each:
  // as TMA tile stores when O is TMA-addressable, else as coalesced 32-bit stores when its rows are 4-byte aligned.
  {
    const LoadMode om = plan_load(A.world > 0 ? A.o_ranks[0] : A.o, A.hd, A.os);
    static int o_env = -1;
    if (o_env < 0) {
      const char* e = getenv("VATS_PREFILL_O_STAGE");  // tuning knob: 0 = per-thread row stores only
      o_env = (e && atoi(e) == 0) ? 0 : 1;
    }
    const bool room = (budget - 8 * vats::kTcOStageBytes) / tile_bytes >= 4;
    P.o_stage = (!o_env || !room || om == LoadMode::kNone) ? 0 : (om == LoadMode::kTma ? 1 : 2);
  }
  int stages = (budget - (P.o_stage ? 8 * vats::kTcOStageBytes : 0)) / tile_bytes;
  int nk = stages / 2, nv = stages - nk;
  if (nk > vats::kTcMaxStages) nk = vats::kTcMaxStages;
  if (nv > vats::kTcMaxStages) nv = vats::kTcMaxStages;
  if (nk < 1 || nv < 1) return fail(VATS_ERR_UNSUPPORTED, "tile does not fit shared memory");
  P.nk = nk;
  P.nv = nv;
  const size_t smem = vats::tc_smem_bytes(P.regions, nk, nv, P.o_stage);

  CUtensorMap mq, mk, mv, mo;
  std::memset(&mq, 0, sizeof(mq));
  std::memset(&mk, 0, sizeof(mk));
  std::memset(&mv, 0, sizeof(mv));
  std::memset(&mo, 0, sizeof(mo));
  int rc;
  if (!any_ldg) {
    if ((rc = encode_map(&mq, A.q, A.N, A.Tq, A.H, A.hd, A.qs)) != VATS_OK) return rc;
    if ((rc = encode_map(&mk, A.k, A.N, A.Tk, A.G, A.hd, A.ks)) != VATS_OK) return rc;
    if ((rc = encode_map(&mv, A.v, A.N, A.Tk, A.G, A.hd, A.vs)) != VATS_OK) return rc;
  }
  vats::TcPeerMaps peer_maps;
  std::memset(&peer_maps, 0, sizeof(peer_maps));
  if (A.world > 0) {
    if (P.o_stage != 1)
      return fail(VATS_ERR_UNSUPPORTED, "fused gather needs a TMA-addressable output (16-byte aligned base and strides, "
                                        "head_dim a multiple of 8) and room for the staging tiles");
    for (int r = 0; r < A.world; ++r) {
      if ((reinterpret_cast<uintptr_t>(A.o_ranks[r]) & 15u) != 0)
        return fail(VATS_ERR_INVALID_ARGUMENT, "o_ranks[%d] is not 16-byte aligned", r);
      if ((rc = encode_map(&peer_maps.m[r], A.o_ranks[r], A.N_total, A.Tq, A.H_total, A.hd, A.os, 32)) != VATS_OK) return rc;
    }
    P.peers = A.world;
    P.peer_first = (A.rank + 1) % A.world;
    P.seq_off = A.seq_off;
    P.head_off = A.head_off;
  } else if (P.o_stage == 1 && (rc = encode_map(&mo, A.o, A.N, A.Tq, A.H, A.hd, A.os, 32)) != VATS_OK) {
    return rc;
  }

  const long long ctas = (long long)A.N * A.G * P.pairs * P.q_blocks;
  if (ctas > 0x7fffffffLL) return fail(VATS_ERR_UNSUPPORTED, "grid too large");
  P.num_work = (int)ctas;
  vats::tc_find_divisor((unsigned)P.q_blocks, P.div_qb);
  vats::tc_find_divisor((unsigned)P.pairs, P.div_pairs);
  vats::tc_find_divisor((unsigned)A.G, P.div_g);
  P.trace = g_trace;
  P.trace_cap = g_trace_cap;
  int grid = sm_count();
  if (grid > P.num_work) grid = P.num_work;
  static thread_local SmemAttrCache smem_set[2];
  if (any_ldg)
    CUDA_TRY(ensure_dyn_smem(vats::prefill_tc_kernel<true>, smem, smem_set[1]));
  else
    CUDA_TRY(ensure_dyn_smem(vats::prefill_tc_kernel<false>, smem, smem_set[0]));
  if (any_ldg)
    vats::prefill_tc_kernel<true><<<(unsigned)grid, vats::kTcThreads, smem, st>>>(P, mq, mk, mv, mo, peer_maps);
  else
    vats::prefill_tc_kernel<false><<<(unsigned)grid, vats::kTcThreads, smem, st>>>(P, mq, mk, mv, mo, peer_maps);
  CUDA_TRY(cudaGetLastError());
  g_launches = 1;
  g_last_kernel = any_ldg ? VATS_LAUNCHED_PREFILL_TC_LDG : VATS_LAUNCHED_PREFILL_TC;
  return VATS_OK;
}

// Sequences of at most 256 keys: K / V of a (sequence, KV group) resident in shared memory for all of the group's heads
// and query blocks, single-pass softmax, two tile slots (prefill_mid.cuh).  Returns -1 when the geometry does not
// qualify (the caller then uses the 128 x 128 tile kernel).
bool mid_enabled() {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("VATS_PREFILL_MID");  // tuning knob: 0 = always the 128 x 128 tile kernel
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  return enabled != 0;
}

bool mid_legal(const PrefillArgs& A, const TcPlan& pl) {
  if (A.Tk < 1 || A.Tk > 256 || A.hd > 128 || A.hd % 2 != 0) return false;
  if (pl.q == LoadMode::kNone || pl.k == LoadMode::kNone || pl.v == LoadMode::kNone) return false;
  if (plan_load(A.o, A.hd, A.os) == LoadMode::kNone) return false;
  if ((long long)A.N * A.G > 0x3fffffffLL) return false;
  return true;
}

int launch_mid(const PrefillArgs& A, const TcPlan& pl, cudaStream_t st) {
  if (!mid_legal(A, pl)) return -1;
  vats::MidParams P;
  std::memset(&P, 0, sizeof(P));
  fill_common(P.a, A);
  const int hpg = P.a.hpg;
  P.hd_pad = (A.hd + 15) / 16 * 16;
  P.regions = (P.hd_pad + 63) / 64;
  P.n_pad = (A.Tk + 15) / 16 * 16;
  if (2 * P.n_pad + P.hd_pad <= 512) {   // one O accumulator behind the two S slots
    P.o_shared = 1;
    P.slot_cols = P.n_pad;
    P.o_off = 2 * P.n_pad;
  } else {                               // O inside each slot's consumed S columns
    P.o_shared = 0;
    P.slot_cols = vats::kMidSlotCols;
    P.o_off = (P.n_pad / 2 + 15) / 16 * 16;
  }
  if (const char* e = getenv("VATS_PREFILL_MID_OSHARED")) {   // tuning knob: 0 = always the in-slot accumulator
    if (atoi(e) == 0) {
      P.o_shared = 0;
      P.slot_cols = vats::kMidSlotCols;
      P.o_off = (P.n_pad / 2 + 15) / 16 * 16;
    }
  }
  // pack the group's heads into the tile rows when their count is a power of two (TMA box {64, pack, 128 / pack})
  int pack = 1;
  if ((hpg & (hpg - 1)) == 0) pack = hpg > 32 ? 32 : hpg;
  P.pack = pack;
  P.pack_shift = 0;
  while ((1 << P.pack_shift) < pack) ++P.pack_shift;
  P.tok_per_tile = 128 >> P.pack_shift;
  P.q_tiles = (A.Tq + P.tok_per_tile - 1) / P.tok_per_tile;
  P.head_sets = hpg / pack;
  const bool any_ldg = pl.q == LoadMode::kLdg || pl.k == LoadMode::kLdg || pl.v == LoadMode::kLdg;
  P.simple_mask = (!A.causal && A.left < 0 && A.right < 0 && !A.q_valid && !A.k_valid) ? 1 : 0;
  const long long tpi = (long long)P.q_tiles * P.head_sets;
  if (tpi > 0x3fffffLL) return -1;
  P.tiles_per_item = (int)tpi;
  P.num_items = A.N * A.G;
  if (any_ldg) {
    // rows TMA cannot address (dense hd 66 / 60): with caller-owned scratch, one streaming repack + the TMA-fed
    // kernel (cfg4a dense: 0.20 + 0.43 ms) beats the in-kernel cp.async staging (2.0 ms: 96 loader threads issuing
    // 4-byte copies with per-copy address arithmetic)
    const int rc2 = launch_tc_repacked(A, pl, st, /*mid=*/true);
    if (rc2 >= 0) return rc2;
  }
  {
    auto al8 = [&](const void* ptr, const int64_t* st3) {
      return (reinterpret_cast<uintptr_t>(ptr) & 7u) == 0 && st3[0] % 4 == 0 && st3[1] % 4 == 0 && st3[2] % 4 == 0;
    };
    P.ldg_vec = (A.hd % 4 == 0 && al8(A.q, A.qs) && al8(A.k, A.ks) && al8(A.v, A.vs)) ? 2 : 1;
  }
  P.o_stage = plan_load(A.o, A.hd, A.os) == LoadMode::kTma ? 1 : 2;   // (3 is decided below, once o_bufs is known)
  P.bounded = A.logit_bound > 0.f ? 1 : 0;
  P.bound_log2 = bound_log2_of(A);
  const long long q_bytes = 2LL * P.regions * vats::kMidQRegionBytes;
  const long long kv_stage = 2LL * P.regions * P.n_pad * 128;
  long long nkv = 0;
  for (int bufs = 2; bufs >= 1 && nkv < 1; --bufs) {   // two staging tiles per epilogue warp when they fit
    const long long budget = 227LL * 1024 - 1024 - (long long)sizeof(vats::MidBarriers) - q_bytes -
                             4LL * bufs * vats::kTcOStageBytes;
    nkv = budget / kv_stage;
    P.o_bufs = bufs;
  }
  if (nkv < 1) return -1;
  if (nkv > vats::kMidMaxKv) nkv = vats::kMidMaxKv;
  {
    static int nkv_env = -1;
    if (nkv_env < 0) {
      const char* e = getenv("VATS_PREFILL_MID_STAGES");  // tuning knob: K / V ring depth
      nkv_env = e ? atoi(e) : 0;
    }
    if (nkv_env >= 1 && nkv_env < nkv) nkv = nkv_env;
  }
  P.nkv = (int)nkv;
  const size_t smem = vats::mid_smem_bytes(P.regions, P.n_pad, P.nkv, P.o_bufs);
  vats::tc_find_divisor((unsigned)A.G, P.div_g);
  vats::tc_find_divisor((unsigned)P.q_tiles, P.div_qt);
  P.trace = g_trace;
  P.trace_cap = g_trace_cap;

  CUtensorMap mq, mk, mv, mo;
  std::memset(&mq, 0, sizeof(mq));
  std::memset(&mk, 0, sizeof(mk));
  std::memset(&mv, 0, sizeof(mv));
  std::memset(&mo, 0, sizeof(mo));
  int rc;
  if (!any_ldg) {
    if ((rc = encode_map(&mq, A.q, A.N, A.Tq, A.H, A.hd, A.qs, P.tok_per_tile, pack)) != VATS_OK) return rc;
    if ((rc = encode_map(&mk, A.k, A.N, A.Tk, A.G, A.hd, A.ks, P.n_pad)) != VATS_OK) return rc;
    if ((rc = encode_map(&mv, A.v, A.N, A.Tk, A.G, A.hd, A.vs, P.n_pad)) != VATS_OK) return rc;
  }
  {
    // O dense within a token (head stride == hd) and pack*hd*2 a multiple of 16 bytes: the warp's rows are 32/pack
    // 16-byte aligned runs — one untiled TMA store per warp instead of one tile store per 64 columns, and it also
    // serves dense head dims TMA tiles cannot address (66, 60)
    static int rows_env = -1;
    if (rows_env < 0) {
      const char* e = getenv("VATS_PREFILL_MID_ROWSTORE");  // tuning knob: 0 = tile stores / coalesced stores
      rows_env = (e && atoi(e) == 0) ? 0 : 1;
    }
    const bool rows_ok = rows_env && A.os[2] == A.hd && (pack * A.hd * 2) % 16 == 0 && pack * A.hd <= 512 &&
                         (A.Tq == 1 || (A.os[1] * 2) % 16 == 0) && (A.N == 1 || (A.os[0] * 2) % 16 == 0) &&
                         (reinterpret_cast<uintptr_t>(A.o) & 15u) == 0 &&
                         (size_t)32 * A.hd * 2 <= (size_t)P.o_bufs * vats::kTcOStageBytes;
    if (rows_ok) P.o_stage = 3;
  }
  if (P.o_stage == 1 && (rc = encode_map(&mo, A.o, A.N, A.Tq, A.H, A.hd, A.os, 32 >> P.pack_shift, pack)) != VATS_OK)
    return rc;
  if (P.o_stage == 3 && (rc = encode_rows_map(&mo, A.o, A.N, A.Tq, A.H, A.hd, A.os, pack)) != VATS_OK) return rc;
  int grid = sm_count();
  if (grid > P.num_items) grid = P.num_items;
  static thread_local SmemAttrCache smem_set[4];
#define VATS_MID_LAUNCH(L, S, IDX)                                                                      \
  {                                                                                                     \
    CUDA_TRY(ensure_dyn_smem(vats::prefill_mid_kernel<L, S>, smem, smem_set[IDX]));                     \
    vats::prefill_mid_kernel<L, S><<<(unsigned)grid, vats::kMidThreads, smem, st>>>(P, mq, mk, mv, mo); \
  }
  if (any_ldg) {
    if (P.simple_mask) VATS_MID_LAUNCH(true, true, 3) else VATS_MID_LAUNCH(true, false, 2)
  } else {
    if (P.simple_mask) VATS_MID_LAUNCH(false, true, 1) else VATS_MID_LAUNCH(false, false, 0)
  }
#undef VATS_MID_LAUNCH
  CUDA_TRY(cudaGetLastError());
  g_launches = 1;
  g_last_kernel = VATS_LAUNCHED_PREFILL_MID;
  return VATS_OK;
}

int choose_kernel(const PrefillArgs& A, TcPlan* pl) {
  const bool legal = tc_legal(A, pl);
  if (!legal) return VATS_KERNEL_SIMT;
  // Below 32 keys a 128 x 128 tile is > 3/4 padding and the problem is bandwidth-bound (ViT-3D temporal pass: 8 tokens
  // per sequence): the CUDA-core class (short-sequence kernel first).  Few QUERY tokens against many keys (chunked
  // prefill tails, speculative-token verification) stay on the tensor-core kernel: it streams K/V once per head pair,
  // which measured 10-60x faster than the generic warp kernel (64 x 4 queries vs 8 192 keys: 0.50 vs 5.1 ms).
  // Many query tokens against a handful of keys (image-generation cross-attention: 10 368 image tokens x 16 text
  // tokens) also run faster on 128-row tiles (0.154 ms) than on the short-sequence kernel's token chunks (0.229 ms;
  // generic warp kernel 0.379 ms): the cut is at two query blocks.
  if (A.Tk < 32 && A.Tq < 256) return VATS_KERNEL_SIMT;
  // At most 256 keys: the whole K / V of a (sequence, KV group) is one tile — the resident-K/V kernel (ViT spatial
  // passes, text encoders, cross-attention contexts, short prompts).
  // It parallelises over (sequence, KV group) items only: with fewer items than half the SMs (a batch-1 prompt: 8
  // items) the tile kernel's finer items (head pairs x query blocks) win — cfg1 T=32: 0.029 vs 0.039 ms.
  if (A.Tk <= 256 && (long long)A.N * A.G >= sm_count() / 2 && mid_enabled() && mid_legal(A, *pl)) return VATS_KERNEL_MID;
  return VATS_KERNEL_TCGEN05;
}

int prefill_impl(const PrefillArgs& A, int kernel, void* stream) {
  g_launches = 0;
  int rc = validate_prefill(A);
  if (rc != VATS_OK) return rc;
  if ((rc = check_device()) != VATS_OK) return rc;
  if (A.N == 0 || A.Tq == 0) return VATS_OK;  // empty output (reference: T == 0 returns an empty tensor, :405-407)
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TcPlan pl{LoadMode::kNone, LoadMode::kNone, LoadMode::kNone};
  int auto_choice = choose_kernel(A, &pl);
  if (kernel == VATS_KERNEL_AUTO) {
    kernel = auto_choice;
    if (kernel == VATS_KERNEL_MID) {
      const int rc2 = launch_mid(A, pl, st);
      if (rc2 >= 0) return rc2;
      kernel = VATS_KERNEL_TCGEN05;   // K / V of one group do not fit shared memory
    }
  }
  if (kernel == VATS_KERNEL_TCGEN05) {
    if (!tc_legal(A, &pl))
      return fail(VATS_ERR_UNSUPPORTED,
                  "geometry not supported by the tcgen05 kernel (hd=%d: need 1 <= Tk, even hd <= 128, even strides and "
                  "4-byte aligned bases)",
                  A.hd);
    return launch_tc(A, pl, st);
  }
  if (kernel == VATS_KERNEL_MID) {
    if (!tc_legal(A, &pl) || !mid_legal(A, pl))
      return fail(VATS_ERR_UNSUPPORTED,
                  "geometry not supported by the resident-K/V kernel (need 1 <= Tk <= 256, even hd <= 128, even strides, "
                  "4-byte aligned bases; got Tk=%d hd=%d)", A.Tk, A.hd);
    const int rc2 = launch_mid(A, pl, st);
    if (rc2 >= 0) return rc2;
    return fail(VATS_ERR_UNSUPPORTED, "resident-K/V kernel: K / V of one KV group do not fit shared memory (Tk=%d hd=%d)",
                A.Tk, A.hd);
  }
  if (kernel == VATS_KERNEL_SIMT) {
    const int rc = launch_short(A, st);
    return rc >= 0 ? rc : launch_simt(A, st);
  }
  return fail(VATS_ERR_INVALID_ARGUMENT, "unknown kernel selector %d", kernel);
}

__global__ void debug_mask_kernel(uint8_t* out, const uint8_t* q_valid, const uint8_t* k_valid, int N, int Tq, int Tk,
                                  vats::MaskParams mp) {
  const long long total = (long long)N * Tq * Tk;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % Tk);
    const int i = (int)((idx / Tk) % Tq);
    const int n = (int)(idx / ((long long)Tk * Tq));
    bool ok = vats::allowed_geom(mp, i, j);
    if (ok && q_valid) ok = q_valid[(long long)n * Tq + i] != 0;
    if (ok && k_valid) ok = k_valid[(long long)n * Tk + j] != 0;
    out[idx] = ok ? 1 : 0;
  }
}

template <int VEC, int LPK, int CPL>
int launch_decode_hpg(const vats::DecodeParams& p, int hpg_tile, dim3 grid, size_t smem, cudaStream_t st) {
  constexpr int U = 8;
  switch (hpg_tile) {
    case 1: vats::decode_split_kernel<VEC, LPK, CPL, 1, U><<<grid, vats::kDecodeThreads, smem, st>>>(p); break;
    case 2: vats::decode_split_kernel<VEC, LPK, CPL, 2, U><<<grid, vats::kDecodeThreads, smem, st>>>(p); break;
    case 3: vats::decode_split_kernel<VEC, LPK, CPL, 3, U><<<grid, vats::kDecodeThreads, smem, st>>>(p); break;
    case 4: vats::decode_split_kernel<VEC, LPK, CPL, 4, U><<<grid, vats::kDecodeThreads, smem, st>>>(p); break;
    default: return fail(VATS_ERR_UNSUPPORTED, "internal: bad head tile %d", hpg_tile);
  }
  return VATS_OK;
}

int decode_chunk_and_splits(int S_max, int left, int* chunk, int* splits) {
  long long window = S_max;
  if (left >= 0 && (long long)left + 1 < window) window = (long long)left + 1;
  if (window < 1) window = 1;
  int ch = 256;
  *chunk = ch;
  *splits = (int)((window + ch - 1) / ch);
  return 0;
}


// Split-K plan of the TMA + mma.sync decode kernel: enough items for ~20 per SM (static round-robin tail <= 5 %),
// chunks are multiples of the 32-key stage.
void decode_mma_plan(int B, int G, int hpg, int S_max, int left, int* hpg_tile, int* head_batches, int* chunk,
                     int* splits) {
  long long window = S_max;
  if (left >= 0 && (long long)left + 1 < window) window = (long long)left + 1;
  if (window < 1) window = 1;
  *hpg_tile = hpg < vats::kDmMaxHeads ? hpg : vats::kDmMaxHeads;
  *head_batches = (hpg + *hpg_tile - 1) / *hpg_tile;
  const long long units = (long long)B * G * (*head_batches);
  // Items are walked round-robin by one persistent CTA per SM: pick the split count whose last wave is fullest,
  // charging ~0.4 % per extra split for its merge / partial traffic.  Splits never go below 64 keys.
  const long long sms = sm_count();
  long long max_ns = (window + 63) / 64;
  if (max_ns > 64) max_ns = 64;
  if (max_ns < 1) max_ns = 1;
  long long want = 1;
  double best = -1.0;
  for (long long ns = 1; ns <= max_ns; ++ns) {
    const long long items = units * ns;
    const long long waves = (items + sms - 1) / sms;
    const double eff = (double)items / (double)(waves * sms) - 0.004 * (double)ns;
    if (eff > best + 1e-9) {
      best = eff;
      want = ns;
    }
  }
  long long ch = (window + want - 1) / want;
  ch = (ch + vats::kDmChunkAlign - 1) / vats::kDmChunkAlign * vats::kDmChunkAlign;
  *chunk = (int)ch;
  *splits = (int)((window + ch - 1) / ch);
}

// tile width of the mma decode kernel for a head dim (template parameter HD)
int decode_mma_tile_hd(int hd) { return hd <= 16 ? 16 : (hd <= 32 ? 32 : (hd <= 64 ? 64 : 128)); }

// bytes of the split counters at the head of the decode workspace (one int per (sequence, KV group, head batch))
size_t decode_counter_bytes(int B, int H, int G) {
  const int hpg = H / G;
  const int hb = (hpg + vats::kDmMaxHeads - 1) / vats::kDmMaxHeads;
  return ((size_t)B * G * hb * sizeof(int) + 255) / 256 * 256;
}

size_t decode_mma_ws_bytes(int B, int H, int G, int hd, int S_max, int left) {
  hd = decode_mma_tile_hd(hd);
  int tile, hb, chunk, ns;
  decode_mma_plan(B, G, H / G, S_max, left, &tile, &hb, &chunk, &ns);
  const size_t counters = decode_counter_bytes(B, H, G);
  return counters + (ns > 1 ? (size_t)B * H * ns * ((size_t)hd + 2) * sizeof(float) : 0);
}

template <int HD, int NCW, int SK>
int launch_decode_mma(vats::DecodeMmaParams& P, const CUtensorMap& mk, const CUtensorMap& mv, cudaStream_t st) {
  // ring depth: two stages per consumer warp (8 x 16 KB = 128 KB in flight per SM).  Measured on cfg2 with the flush
  // warp in place: 4 stages 0.195 ms, 8 stages 0.165 ms, 12 stages (all 227 KB) 0.174 ms — more bytes in flight than
  // the latency-bandwidth product needs only add DRAM page conflicts.
  int max_stages = 48;
  while (vats::decode_mma_smem_bytes<HD, NCW, SK>(max_stages) > 227 * 1024 && max_stages > 2) --max_stages;
  // Tiles of <= 64 columns (hd 60 / 64: 8 KB stages) take three stages per consumer warp: cfg2-medium 8 stages 0.135,
  // 12 stages 0.129, 16 stages 0.138, 24 stages 0.143 ms (tools/run_workload.py cfg2m --time, launch latency included).
  // (64-key stages for these tiles — 16 KB per stage like a 128-column tile's — measured 0.118 against 0.114 ms.)
  const int want_stages = (HD <= 64 ? 3 : 2) * NCW;
  int stages = want_stages < max_stages ? want_stages : max_stages;
  if (const char* e = getenv("VATS_DECODE_STAGES")) {  // tuning / debugging knob
    const int want = atoi(e);
    if (want >= 2 && want <= max_stages) stages = want;
  }
  // The ring depth MUST be a multiple of the consumer-warp count, so that a slot is always consumed by the same
  // warp: a consumer waits on a slot's "full" barrier by phase parity, which is only sound if it has itself seen
  // the slot's previous phase complete (TMA completions of different slots are not ordered).
  stages -= stages % NCW;
  if (stages < NCW) return fail(VATS_ERR_UNSUPPORTED, "decode: shared memory too small for the stage ring");
  const size_t smem = vats::decode_mma_smem_bytes<HD, NCW, SK>(stages);
  P.stages = stages;
  static thread_local SmemAttrCache smem_set;
  CUDA_TRY(ensure_dyn_smem(vats::decode_mma_kernel<HD, NCW, SK>, smem, smem_set));
  int grid = sm_count();
  if (grid > P.num_items) grid = P.num_items;
  vats::decode_mma_kernel<HD, NCW, SK><<<grid, (2 * NCW + 1) * 32, smem, st>>>(P, mk, mv);
  CUDA_TRY(cudaGetLastError());
  return VATS_OK;
}

// Consumer-warp count / stage size of the mma decode kernel: (4 warps, 32-key stages) or (8 warps, 16-key stages).
int decode_mma_ncw() {
  static int ncw = 0;
  if (ncw == 0) {
    const char* e = getenv("VATS_DECODE_CONSUMER_WARPS");  // tuning knob
    ncw = (e && atoi(e) == 8) ? 8 : 4;  // measured on B200: 4 warps x 32-key stages 198 us, 8 x 16-key 279 us (cfg2)
  }
  return ncw;
}
int decode_mma_stage_keys() { return decode_mma_ncw() == 8 ? 16 : 32; }

template <int HD>
int launch_decode_mma_ncw(vats::DecodeMmaParams& P, const CUtensorMap& mk, const CUtensorMap& mv, cudaStream_t st) {
  return decode_mma_ncw() == 8 ? launch_decode_mma<HD, 8, 16>(P, mk, mv, st)
                               : launch_decode_mma<HD, 4, 32>(P, mk, mv, st);
}

}  // namespace

extern "C" {

int vats_attn_version(void) { return VATS_ATTN_VERSION; }
void vats_attn_debug_set_trace(void* dev_buf, int capacity) {
  g_trace = reinterpret_cast<unsigned long long*>(dev_buf);
  g_trace_cap = capacity;
}
const char* vats_attn_last_error(void) { return g_err; }
int vats_attn_last_launch_count(void) { return g_launches; }
int vats_attn_last_kernel(void) { return g_last_kernel; }

int vats_attn_prefill_ex(const void* q, const void* k, const void* v, void* o, const uint8_t* q_valid,
                         const uint8_t* k_valid, int N, int Tq, int Tk, int H, int G, int hd,
                         const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                         const int64_t o_strides[3], float scale, int causal, int left, int right, int kernel,
                         void* stream) {
  PrefillArgs A{q, k, v, o, q_valid, k_valid, N, Tq, Tk, H, G, hd, q_strides, k_strides, v_strides, o_strides,
                scale, causal, left, right};
  return prefill_impl(A, kernel, stream);
}

int vats_attn_prefill_ws(const void* q, const void* k, const void* v, void* o, const uint8_t* q_valid,
                         const uint8_t* k_valid, int N, int Tq, int Tk, int H, int G, int hd,
                         const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                         const int64_t o_strides[3], float scale, int causal, int left, int right, int kernel,
                         float logit_bound, void* workspace, size_t workspace_bytes, void* stream) {
  PrefillArgs A{q, k, v, o, q_valid, k_valid, N, Tq, Tk, H, G, hd, q_strides, k_strides, v_strides, o_strides,
                scale, causal, left, right};
  A.ws = workspace;
  A.ws_bytes = workspace ? workspace_bytes : 0;
  if (logit_bound < 0.f || !std::isfinite(logit_bound))
    return fail(VATS_ERR_INVALID_ARGUMENT, "logit_bound must be >= 0 and finite (0 = unknown)");
  A.logit_bound = logit_bound;
  return prefill_impl(A, kernel, stream);
}

int vats_attn_prefill_gather(const void* q, const void* k, const void* v, void* const* o_ranks, int world, int rank,
                             int seq_offset, int head_offset, int N_total, int H_total, const uint8_t* q_valid,
                             const uint8_t* k_valid, int N, int Tq, int Tk, int H, int G, int hd,
                             const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                             const int64_t o_strides[3], float scale, int causal, int left, int right,
                             float logit_bound, void* workspace, size_t workspace_bytes, void* stream) {
  g_launches = 0;
  if (logit_bound < 0.f || !std::isfinite(logit_bound))
    return fail(VATS_ERR_INVALID_ARGUMENT, "logit_bound must be >= 0 and finite (0 = unknown)");
  if (!o_ranks || world < 1 || world > vats::kTcMaxPeers || rank < 0 || rank >= world)
    return fail(VATS_ERR_INVALID_ARGUMENT, "fused gather: need 1 <= world <= %d output pointers and 0 <= rank < world",
                vats::kTcMaxPeers);
  if (seq_offset < 0 || head_offset < 0 || seq_offset + N > N_total || head_offset + H > H_total)
    return fail(VATS_ERR_INVALID_ARGUMENT, "fused gather: the local [%d, %d] block at (%d, %d) does not fit the gathered "
                "[%d, Tq, %d, hd] tensor", N, H, seq_offset, head_offset, N_total, H_total);
  for (int r = 0; r < world; ++r)
    if (!o_ranks[r]) return fail(VATS_ERR_INVALID_ARGUMENT, "fused gather: o_ranks[%d] is NULL", r);
  PrefillArgs A{q, k, v, o_ranks[rank], q_valid, k_valid, N, Tq, Tk, H, G, hd, q_strides, k_strides, v_strides, o_strides,
                scale, causal, left, right};
  A.ws = workspace;
  A.ws_bytes = workspace ? workspace_bytes : 0;
  A.o_ranks = o_ranks;
  A.world = world;
  A.rank = rank;
  A.seq_off = seq_offset;
  A.head_off = head_offset;
  A.N_total = N_total;
  A.H_total = H_total;
  A.logit_bound = logit_bound;
  int rc = validate_prefill(A);
  if (rc != VATS_OK) return rc;
  if ((rc = check_device()) != VATS_OK) return rc;
  if (N == 0 || Tq == 0) return VATS_OK;
  TcPlan pl{LoadMode::kNone, LoadMode::kNone, LoadMode::kNone};
  if (!tc_legal(A, &pl))
    return fail(VATS_ERR_UNSUPPORTED, "fused gather runs on the tcgen05 tile kernel: geometry not supported (hd=%d)", hd);
  return launch_tc(A, pl, reinterpret_cast<cudaStream_t>(stream));
}

size_t vats_attn_prefill_backward_workspace_bytes(int N, int Tq, int H) {
  if (N <= 0 || Tq <= 0 || H <= 0) return 0;
  return (size_t)2 * N * H * Tq * sizeof(float);   // log-sum-exp and D = rowsum(dO o O), fp32 [N, H, Tq] each
}

int vats_attn_prefill_backward(const void* q, const void* k, const void* v, const void* o, const void* dout, void* dq,
                               void* dk, void* dv, const uint8_t* q_valid, const uint8_t* k_valid, int N, int Tq, int Tk,
                               int H, int G, int hd, const int64_t q_strides[3], const int64_t k_strides[3],
                               const int64_t v_strides[3], const int64_t o_strides[3], const int64_t do_strides[3],
                               float scale, int causal, int left, int right, void* workspace, size_t workspace_bytes,
                               void* stream) {
  g_launches = 0;
  PrefillArgs A{q, k, v, const_cast<void*>(o), q_valid, k_valid, N, Tq, Tk, H, G, hd, q_strides, k_strides, v_strides,
                o_strides, scale, causal, left, right};
  int rc = validate_prefill(A);
  if (rc != VATS_OK) return rc;
  if (!do_strides) return fail(VATS_ERR_INVALID_ARGUMENT, "stride arrays must not be NULL");
  if ((rc = check_device()) != VATS_OK) return rc;
  if (N == 0 || (Tq == 0 && Tk == 0)) return VATS_OK;
  if (!dq || !dk || !dv || (Tq > 0 && !dout)) return fail(VATS_ERR_INVALID_ARGUMENT, "dout / dq / dk / dv must not be NULL");
  if (hd > 128 || hd % 2 != 0)
    return fail(VATS_ERR_UNSUPPORTED, "backward supports even head_dim <= 128 (got %d)", hd);
  {
    auto al4 = [](const void* p, const int64_t* s3) {
      return (reinterpret_cast<uintptr_t>(p) & 3u) == 0 && s3[0] % 2 == 0 && s3[1] % 2 == 0 && s3[2] % 2 == 0;
    };
    if (!al4(q, q_strides) || !al4(k, k_strides) || !al4(v, v_strides) || !al4(o, o_strides) || !al4(dout, do_strides) ||
        (reinterpret_cast<uintptr_t>(dq) & 3u) || (reinterpret_cast<uintptr_t>(dk) & 3u) || (reinterpret_cast<uintptr_t>(dv) & 3u))
      return fail(VATS_ERR_UNSUPPORTED, "backward needs 4-byte aligned rows (even strides, 4-byte aligned bases)");
  }
  const size_t need = vats_attn_prefill_backward_workspace_bytes(N, Tq, H);
  if (Tq > 0 && (!workspace || workspace_bytes < need))
    return fail(VATS_ERR_WORKSPACE, "backward workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  if (N > 65535 || H > 65535) return fail(VATS_ERR_UNSUPPORTED, "backward grid limit: N and H must be <= 65535");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  vats::BwdParams P;
  std::memset(&P, 0, sizeof(P));
  fill_common(P.a, A);
  P.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  P.dos_n = do_strides[0]; P.dos_t = do_strides[1]; P.dos_h = do_strides[2];
  P.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  P.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  P.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  P.lse = reinterpret_cast<float*>(workspace);
  P.dsum = P.lse + (size_t)N * H * Tq;
  P.scale = scale;
  P.hd_pad = (hd + 15) / 16 * 16;
  {
    auto al16 = [&](const void* ptr, const int64_t* st3) {
      return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0 && st3[0] % 8 == 0 && st3[1] % 8 == 0 && st3[2] % 8 == 0;
    };
    P.vec16 = (al16(A.q, A.qs) && al16(A.k, A.ks) && al16(A.v, A.vs) && al16(A.o, A.os) && al16(dout, do_strides)) ? 1 : 0;
  }
  const size_t smem = vats::bwd_smem_bytes(P.hd_pad);
  const size_t smem_dq = vats::bwd_dq_smem_bytes(P.hd_pad);
  const int ks = P.hd_pad / 16;
  const dim3 grid_q((Tq + vats::kBwdBM - 1) / vats::kBwdBM, H, N);
  const dim3 grid_k((Tk + vats::kBwdBN - 1) / vats::kBwdBN, G, N);
  if (Tk == 0) {   // nothing to attend: all gradients are zero
    CUDA_TRY(cudaMemsetAsync(dq, 0, (size_t)N * Tq * H * hd * 2, st));
    return VATS_OK;
  }
#define VATS_BWD_CASE(KS)                                                                                         \
  case KS: {                                                                                                      \
    static thread_local SmemAttrCache c1, c2;                                                                     \
    CUDA_TRY(ensure_dyn_smem(vats::attn_bwd_dq_kernel<KS>, smem_dq, c1));                                            \
    CUDA_TRY(ensure_dyn_smem(vats::attn_bwd_dkv_kernel<KS>, smem, c2));                                           \
    if (Tq > 0) vats::attn_bwd_dq_kernel<KS><<<grid_q, vats::kBwdThreads, smem_dq, st>>>(P);                         \
    vats::attn_bwd_dkv_kernel<KS><<<grid_k, vats::kBwdThreads, smem, st>>>(P);                                    \
  } break;
  switch (ks) {
    VATS_BWD_CASE(1)
    VATS_BWD_CASE(2)
    VATS_BWD_CASE(3)
    VATS_BWD_CASE(4)
    VATS_BWD_CASE(5)
    VATS_BWD_CASE(6)
    VATS_BWD_CASE(7)
    VATS_BWD_CASE(8)
    default:
      return fail(VATS_ERR_UNSUPPORTED, "backward: head_dim %d not supported", hd);
  }
#undef VATS_BWD_CASE
  CUDA_TRY(cudaGetLastError());
  g_launches = Tq > 0 ? 2 : 1;
  g_last_kernel = VATS_LAUNCHED_BACKWARD;
  return VATS_OK;
}

size_t vats_attn_prefill_workspace_bytes(int N, int Tq, int Tk, int H, int G, int hd, const int64_t q_strides[3],
                                         const int64_t k_strides[3], const int64_t v_strides[3], const void* q,
                                         const void* k, const void* v) {
  if (N <= 0 || Tq <= 0 || Tk <= 0 || H <= 0 || G <= 0 || hd <= 0 || H % G != 0 || !q_strides || !k_strides ||
      !v_strides)
    return 0;
  PrefillArgs A{q, k, v, nullptr, nullptr, nullptr, N, Tq, Tk, H, G, hd, q_strides, k_strides, v_strides, q_strides,
                1.f, 0, -1, -1};
  TcPlan pl{LoadMode::kNone, LoadMode::kNone, LoadMode::kNone};
  const int choice = choose_kernel(A, &pl);
  if (choice != VATS_KERNEL_TCGEN05 && choice != VATS_KERNEL_MID) return 0;
  if (pl.q != LoadMode::kLdg && pl.k != LoadMode::kLdg && pl.v != LoadMode::kLdg) return 0;
  size_t off[4];
  return tc_repack_offsets(A, pl, off);
}

int vats_attn_prefill(const void* q, const void* k, const void* v, void* o, const uint8_t* q_valid,
                      const uint8_t* k_valid, int N, int Tq, int Tk, int H, int G, int hd,
                      const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                      const int64_t o_strides[3], float scale, int causal, int left, int right, void* stream) {
  return vats_attn_prefill_ex(q, k, v, o, q_valid, k_valid, N, Tq, Tk, H, G, hd, q_strides, k_strides, v_strides,
                              o_strides, scale, causal, left, right, VATS_KERNEL_AUTO, stream);
}

int vats_attn_prefill_plan(int N, int Tq, int Tk, int H, int G, int hd, const int64_t q_strides[3],
                           const int64_t k_strides[3], const int64_t v_strides[3], const int64_t o_strides[3],
                           const void* q, const void* k, const void* v) {
  PrefillArgs A{q, k, v, nullptr, nullptr, nullptr, N, Tq, Tk, H, G, hd, q_strides, k_strides, v_strides, o_strides,
                1.f, 0, -1, -1};
  if (H <= 0 || G <= 0 || hd <= 0 || H % G != 0 || !q_strides || !k_strides || !v_strides) return -1;
  TcPlan pl;
  return choose_kernel(A, &pl);
}

int vats_attn_prefill_prepare(const void* q_in, const void* k_in, const void* v_in, int in_dtype, void* q_out,
                              void* k_out, void* v_out, const float* cos_table, const float* sin_table, int N, int T,
                              int H, int G, int hd, int pos0, const int64_t qin_strides[3],
                              const int64_t kin_strides[3], const int64_t vin_strides[3],
                              const int64_t qout_strides[3], const int64_t kout_strides[3],
                              const int64_t vout_strides[3], int qk_norm, float eps, void* stream) {
  g_launches = 0;
  if (N < 0 || T < 0 || pos0 < 0) return fail(VATS_ERR_INVALID_ARGUMENT, "negative size");
  if (H <= 0 || G <= 0 || hd <= 0) return fail(VATS_ERR_INVALID_ARGUMENT, "H, G, hd must be positive");
  if (in_dtype != 0 && in_dtype != 1) return fail(VATS_ERR_INVALID_ARGUMENT, "in_dtype must be 0 (bf16) or 1 (fp32)");
  if (!qin_strides || !kin_strides || !vin_strides || !qout_strides || !kout_strides || !vout_strides)
    return fail(VATS_ERR_INVALID_ARGUMENT, "stride arrays must not be NULL");
  if ((cos_table == nullptr) != (sin_table == nullptr))
    return fail(VATS_ERR_INVALID_ARGUMENT, "cos_table and sin_table must both be given or both be NULL");
  if (cos_table != nullptr && hd % 2 != 0) return fail(VATS_ERR_UNSUPPORTED, "RoPE needs an even head_dim (got %d)", hd);
  if (hd > 2 * 32 * vats::kPrepareMaxPairs) return fail(VATS_ERR_UNSUPPORTED, "head_dim %d > 256 is not supported", hd);
  if (!(eps >= 0.f)) return fail(VATS_ERR_INVALID_ARGUMENT, "eps must be >= 0");
  int rc = check_device();
  if (rc != VATS_OK) return rc;
  if (N == 0 || T == 0) return VATS_OK;
  if (!q_in || !k_in || !v_in || !q_out || !k_out || !v_out)
    return fail(VATS_ERR_INVALID_ARGUMENT, "tensor pointers must not be NULL");
  vats::PrefillPrepareParams p;
  std::memset(&p, 0, sizeof(p));
  p.q_in = q_in; p.k_in = k_in; p.v_in = v_in; p.in_fp32 = in_dtype;
  p.q_out = reinterpret_cast<__nv_bfloat16*>(q_out);
  p.k_out = reinterpret_cast<__nv_bfloat16*>(k_out);
  p.v_out = reinterpret_cast<__nv_bfloat16*>(v_out);
  p.cos_table = cos_table; p.sin_table = sin_table;
  p.N = N; p.T = T; p.H = H; p.G = G; p.hd = hd; p.pos0 = pos0;
  p.qi_n = qin_strides[0]; p.qi_t = qin_strides[1]; p.qi_h = qin_strides[2];
  p.ki_n = kin_strides[0]; p.ki_t = kin_strides[1]; p.ki_h = kin_strides[2];
  p.vi_n = vin_strides[0]; p.vi_t = vin_strides[1]; p.vi_h = vin_strides[2];
  p.qo_n = qout_strides[0]; p.qo_t = qout_strides[1]; p.qo_h = qout_strides[2];
  p.ko_n = kout_strides[0]; p.ko_t = kout_strides[1]; p.ko_h = kout_strides[2];
  p.vo_n = vout_strides[0]; p.vo_t = vout_strides[1]; p.vo_h = vout_strides[2];
  p.qk_norm = qk_norm ? 1 : 0;
  p.eps = eps;
  const long long rows = (long long)N * T * (H + 2 * G);
  long long blocks = (rows + vats::kPrepareWarps - 1) / vats::kPrepareWarps;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  vats::prefill_prepare_kernel<<<(unsigned)blocks, vats::kPrepareWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  CUDA_TRY(cudaGetLastError());
  g_launches = 1;
  g_last_kernel = VATS_LAUNCHED_PREFILL_PREPARE;
  return VATS_OK;
}

int vats_attn_prefill_prepare_table(const void* q_in, const void* k_in, const void* v_in, int in_dtype, void* q_out,
                                    void* k_out, void* v_out, const float* cos_table, const float* sin_table,
                                    const int32_t* partner, int No, int Ni, int T, int H, int G, int hd,
                                    const int64_t qin_strides[4], const int64_t kin_strides[4],
                                    const int64_t vin_strides[4], const int64_t qout_strides[3],
                                    const int64_t kout_strides[3], const int64_t vout_strides[3], int qk_norm, float eps,
                                    void* stream) {
  g_launches = 0;
  if (No < 0 || Ni < 0 || T < 0) return fail(VATS_ERR_INVALID_ARGUMENT, "negative size");
  if (H <= 0 || G <= 0 || hd <= 0) return fail(VATS_ERR_INVALID_ARGUMENT, "H, G, hd must be positive");
  if (in_dtype != 0 && in_dtype != 1) return fail(VATS_ERR_INVALID_ARGUMENT, "in_dtype must be 0 (bf16) or 1 (fp32)");
  if (!qin_strides || !kin_strides || !vin_strides || !qout_strides || !kout_strides || !vout_strides)
    return fail(VATS_ERR_INVALID_ARGUMENT, "stride arrays must not be NULL");
  const bool any = cos_table || sin_table || partner;
  if (any && !(cos_table && sin_table && partner))
    return fail(VATS_ERR_INVALID_ARGUMENT, "cos_table, sin_table and partner must all be given or all be NULL");
  if (hd > vats::kPrepareTableMaxHd) return fail(VATS_ERR_UNSUPPORTED, "head_dim %d > 256 is not supported", hd);
  if (!(eps >= 0.f)) return fail(VATS_ERR_INVALID_ARGUMENT, "eps must be >= 0");
  int rc = check_device();
  if (rc != VATS_OK) return rc;
  if (No == 0 || Ni == 0 || T == 0) return VATS_OK;
  if (!q_in || !k_in || !v_in || !q_out || !k_out || !v_out)
    return fail(VATS_ERR_INVALID_ARGUMENT, "tensor pointers must not be NULL");
  vats::PrepareTableParams p;
  std::memset(&p, 0, sizeof(p));
  p.q_in = q_in; p.k_in = k_in; p.v_in = v_in; p.in_fp32 = in_dtype;
  p.q_out = reinterpret_cast<__nv_bfloat16*>(q_out);
  p.k_out = reinterpret_cast<__nv_bfloat16*>(k_out);
  p.v_out = reinterpret_cast<__nv_bfloat16*>(v_out);
  p.cos_table = cos_table; p.sin_table = sin_table; p.partner = partner;
  p.No = No; p.Ni = Ni; p.T = T; p.H = H; p.G = G; p.hd = hd;
  p.q_no = qin_strides[0]; p.q_ni = qin_strides[1]; p.q_t = qin_strides[2]; p.q_h = qin_strides[3];
  p.k_no = kin_strides[0]; p.k_ni = kin_strides[1]; p.k_t = kin_strides[2]; p.k_h = kin_strides[3];
  p.v_no = vin_strides[0]; p.v_ni = vin_strides[1]; p.v_t = vin_strides[2]; p.v_h = vin_strides[3];
  p.qo_n = qout_strides[0]; p.qo_t = qout_strides[1]; p.qo_h = qout_strides[2];
  p.ko_n = kout_strides[0]; p.ko_t = kout_strides[1]; p.ko_h = kout_strides[2];
  p.vo_n = vout_strides[0]; p.vo_t = vout_strides[1]; p.vo_h = vout_strides[2];
  p.qk_norm = qk_norm ? 1 : 0;
  p.eps = eps;
  const long long rows = (long long)No * Ni * T * (H + 2 * G);
  long long blocks = (rows + vats::kPrepareWarps - 1) / vats::kPrepareWarps;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  vats::prefill_prepare_table_kernel<<<(unsigned)blocks, vats::kPrepareWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  CUDA_TRY(cudaGetLastError());
  g_launches = 1;
  g_last_kernel = VATS_LAUNCHED_PREFILL_PREPARE;
  return VATS_OK;
}

int vats_attn_decode_prepare(const void* q_in, const void* k_in, const void* v_in, int in_dtype, void* q_out,
                             void* k_cache, void* v_cache, const int32_t* seq_lens, const float* cos_table,
                             const float* sin_table, int B, int H, int G, int hd, int S_max,
                             const int64_t qin_strides[2], const int64_t kin_strides[2], const int64_t vin_strides[2],
                             const int64_t qout_strides[2], const int64_t k_strides[3], const int64_t v_strides[3],
                             int qk_norm, float eps, void* stream) {
  g_launches = 0;
  if (B < 0 || S_max < 0) return fail(VATS_ERR_INVALID_ARGUMENT, "negative size");
  if (H <= 0 || G <= 0 || hd <= 0) return fail(VATS_ERR_INVALID_ARGUMENT, "H, G, hd must be positive");
  if (in_dtype != 0 && in_dtype != 1) return fail(VATS_ERR_INVALID_ARGUMENT, "in_dtype must be 0 (bf16) or 1 (fp32)");
  if (!qin_strides || !kin_strides || !vin_strides || !qout_strides || !k_strides || !v_strides)
    return fail(VATS_ERR_INVALID_ARGUMENT, "stride arrays must not be NULL");
  if ((cos_table == nullptr) != (sin_table == nullptr))
    return fail(VATS_ERR_INVALID_ARGUMENT, "cos_table and sin_table must both be given or both be NULL");
  if (cos_table != nullptr && hd % 2 != 0) return fail(VATS_ERR_UNSUPPORTED, "RoPE needs an even head_dim (got %d)", hd);
  if (hd > 2 * 32 * vats::kPrepareMaxPairs) return fail(VATS_ERR_UNSUPPORTED, "head_dim %d > 256 is not supported", hd);
  if (!(eps >= 0.f)) return fail(VATS_ERR_INVALID_ARGUMENT, "eps must be >= 0");
  int rc = check_device();
  if (rc != VATS_OK) return rc;
  if (B == 0) return VATS_OK;
  if (!q_in || !k_in || !v_in || !q_out || !k_cache || !v_cache || !seq_lens)
    return fail(VATS_ERR_INVALID_ARGUMENT, "tensor pointers must not be NULL");
  vats::PrepareParams p;
  std::memset(&p, 0, sizeof(p));
  p.q_in = q_in; p.k_in = k_in; p.v_in = v_in; p.in_fp32 = in_dtype;
  p.q_out = reinterpret_cast<__nv_bfloat16*>(q_out);
  p.k_cache = reinterpret_cast<__nv_bfloat16*>(k_cache);
  p.v_cache = reinterpret_cast<__nv_bfloat16*>(v_cache);
  p.seq_lens = seq_lens; p.cos_table = cos_table; p.sin_table = sin_table;
  p.B = B; p.H = H; p.G = G; p.hd = hd; p.S_max = S_max;
  p.qi_b = qin_strides[0]; p.qi_h = qin_strides[1];
  p.ki_b = kin_strides[0]; p.ki_h = kin_strides[1];
  p.vi_b = vin_strides[0]; p.vi_h = vin_strides[1];
  p.qo_b = qout_strides[0]; p.qo_h = qout_strides[1];
  p.ks_b = k_strides[0]; p.ks_t = k_strides[1]; p.ks_h = k_strides[2];
  p.vs_b = v_strides[0]; p.vs_t = v_strides[1]; p.vs_h = v_strides[2];
  p.qk_norm = qk_norm ? 1 : 0;
  p.eps = eps;
  const long long rows = (long long)B * (H + 2 * G);
  const long long blocks = (rows + vats::kPrepareWarps - 1) / vats::kPrepareWarps;
  if (blocks > 0x7fffffffLL) return fail(VATS_ERR_UNSUPPORTED, "grid too large");
  vats::decode_prepare_kernel<<<(unsigned)blocks, vats::kPrepareWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  CUDA_TRY(cudaGetLastError());
  g_launches = 1;
  g_last_kernel = VATS_LAUNCHED_DECODE_PREPARE;
  return VATS_OK;
}

size_t vats_attn_decode_workspace_bytes(int B, int H, int G, int hd, int S_max, int left) {
  if (B <= 0 || H <= 0 || hd <= 0 || S_max <= 0) return 0;
  if (G <= 0 || H % G != 0) return 0;
  // layout: [split counters of the mma kernel | fp32 partials of whichever kernel runs].  The CUDA-core kernel's
  // partials start behind the counter region too, so it can never leave a non-zero counter behind.
  int chunk, splits;
  decode_chunk_and_splits(S_max, left, &chunk, &splits);
  const size_t simt = decode_counter_bytes(B, H, G) +
                      (splits <= 1 ? 16 : (size_t)B * H * splits * ((size_t)hd + 2) * sizeof(float));
  const size_t mma = decode_mma_ws_bytes(B, H, G, hd, S_max, left);
  return simt > mma ? simt : mma;
}

int vats_attn_decode(const void* q, const void* k_cache, const void* v_cache, void* o, const int32_t* seq_lens,
                     int B, int H, int G, int hd, int S_max, const int64_t q_strides[2], const int64_t k_strides[3],
                     const int64_t v_strides[3], const int64_t o_strides[2], float scale, int left, void* workspace,
                     size_t workspace_bytes, void* stream) {
  g_launches = 0;
  if (B < 0 || S_max < 0) return fail(VATS_ERR_INVALID_ARGUMENT, "negative size");
  if (H <= 0 || G <= 0 || hd <= 0) return fail(VATS_ERR_INVALID_ARGUMENT, "H, G, hd must be positive");
  if (H % G != 0)
    return fail(VATS_ERR_INVALID_ARGUMENT, "num_heads (%d) must be divisible by query_groups (%d)", H, G);
  if (!(scale > 0.f) || !std::isfinite(scale))
    return fail(VATS_ERR_INVALID_ARGUMENT, "scale must be a positive finite float (got %g)", (double)scale);
  if (!q_strides || !k_strides || !v_strides || !o_strides)
    return fail(VATS_ERR_INVALID_ARGUMENT, "stride arrays must not be NULL");
  int rc = check_device();
  if (rc != VATS_OK) return rc;
  if (B == 0) return VATS_OK;
  if (!q || !o || !seq_lens) return fail(VATS_ERR_INVALID_ARGUMENT, "q / o / seq_lens must not be NULL");
  if (S_max > 0 && (!k_cache || !v_cache)) return fail(VATS_ERR_INVALID_ARGUMENT, "k_cache / v_cache must not be NULL");
  if (hd > 256) return fail(VATS_ERR_UNSUPPORTED, "decode supports head_dim <= 256 (got %d)", hd);
  if (hd % 2 != 0) return fail(VATS_ERR_UNSUPPORTED, "decode needs an even head_dim (got %d)", hd);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  vats::DecodeParams p;
  std::memset(&p, 0, sizeof(p));
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(k_cache);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v_cache);
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.seq_lens = seq_lens;
  p.B = B; p.H = H; p.G = G; p.hd = hd; p.S_max = S_max;
  p.hpg = H / G;
  p.qs_b = q_strides[0]; p.qs_h = q_strides[1];
  p.ks_b = k_strides[0]; p.ks_t = k_strides[1]; p.ks_h = k_strides[2];
  p.vs_b = v_strides[0]; p.vs_t = v_strides[1]; p.vs_h = v_strides[2];
  p.os_b = o_strides[0]; p.os_h = o_strides[1];
  p.scale_log2 = scale * 1.4426950408889634f;
  p.left = left;
  decode_chunk_and_splits(S_max > 0 ? S_max : 1, left, &p.chunk, &p.num_splits);
  if (p.num_splits > 1) {
    const size_t need = vats_attn_decode_workspace_bytes(B, H, G, hd, S_max, left);
    if (!workspace || workspace_bytes < need)
      return fail(VATS_ERR_WORKSPACE, "decode workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    p.ws_acc = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + decode_counter_bytes(B, H, G));
    p.ws_ml = p.ws_acc + (size_t)B * H * p.num_splits * hd;
  }

  // ---- fast path: TMA-addressable cache (16-byte aligned base / strides, head_dim multiple of 16 up to 128)
  {
    const bool tma_ok = hd <= 128 && S_max > 0 &&
                        (reinterpret_cast<uintptr_t>(k_cache) & 15u) == 0 &&
                        (reinterpret_cast<uintptr_t>(v_cache) & 15u) == 0 && k_strides[0] % 8 == 0 &&
                        k_strides[1] % 8 == 0 && k_strides[2] % 8 == 0 && v_strides[0] % 8 == 0 &&
                        v_strides[1] % 8 == 0 && v_strides[2] % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(q) & 3u) == 0 && q_strides[0] % 2 == 0 && q_strides[1] % 2 == 0 &&
                        (long long)B * G <= 0x3fffffffLL;
    if (tma_ok) {
      vats::DecodeMmaParams P;
      std::memset(&P, 0, sizeof(P));
      P.d = p;
      decode_mma_plan(B, G, p.hpg, S_max, left, &P.hpg_tile, &P.d.head_batches, &P.d.chunk, &P.d.num_splits);
      const size_t need = decode_mma_ws_bytes(B, H, G, hd, S_max, left);
      if (!workspace || workspace_bytes < need)
        return fail(VATS_ERR_WORKSPACE, "decode workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
      const size_t counters = decode_counter_bytes(B, H, G);
      P.counters = reinterpret_cast<int*>(workspace);
      P.d.ws_acc = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + counters);
      P.d.ws_ml = P.d.ws_acc + (size_t)B * H * P.d.num_splits * decode_mma_tile_hd(hd);
      const long long items = (long long)B * G * P.d.head_batches * P.d.num_splits;
      if (items > 0x7fffffffLL) return fail(VATS_ERR_UNSUPPORTED, "decode: too many work items");
      P.num_items = (int)items;
      CUtensorMap mk, mv;
      if ((rc = encode_map(&mk, k_cache, B, S_max, G, hd, k_strides, decode_mma_stage_keys())) != VATS_OK) return rc;
      if ((rc = encode_map(&mv, v_cache, B, S_max, G, hd, v_strides, decode_mma_stage_keys())) != VATS_OK) return rc;
      switch (decode_mma_tile_hd(hd)) {
        case 16: rc = launch_decode_mma_ncw<16>(P, mk, mv, st); break;
        case 32: rc = launch_decode_mma_ncw<32>(P, mk, mv, st); break;
        case 64: rc = launch_decode_mma_ncw<64>(P, mk, mv, st); break;
        default: rc = launch_decode_mma_ncw<128>(P, mk, mv, st); break;
      }
      if (rc != VATS_OK) return rc;
      g_launches = 1;
      g_last_kernel = VATS_LAUNCHED_DECODE_MMA;
      return VATS_OK;
    }
  }

  // vector width: limited by head_dim, base alignment and strides of q, k and v
  auto align_of = [](const void* ptr, std::initializer_list<int64_t> strides, int hd_) {
    int vec = 8;
    while (vec > 2) {
      bool ok = (reinterpret_cast<uintptr_t>(ptr) % (vec * 2) == 0) && (hd_ % vec == 0);
      for (int64_t s : strides) ok = ok && (s % vec == 0);
      if (ok) break;
      vec >>= 1;
    }
    return vec;
  };
  int vec = align_of(q, {q_strides[0], q_strides[1]}, hd);
  vec = std::min(vec, align_of(k_cache, {k_strides[0], k_strides[1], k_strides[2]}, hd));
  vec = std::min(vec, align_of(v_cache, {v_strides[0], v_strides[1], v_strides[2]}, hd));
  {
    bool ok2 = (reinterpret_cast<uintptr_t>(q) % 4 == 0) && (reinterpret_cast<uintptr_t>(k_cache) % 4 == 0) &&
               (reinterpret_cast<uintptr_t>(v_cache) % 4 == 0);
    for (int64_t s : {q_strides[0], q_strides[1], k_strides[0], k_strides[1], k_strides[2], v_strides[0],
                      v_strides[1], v_strides[2]})
      ok2 = ok2 && (s % 2 == 0);
    if (!ok2) return fail(VATS_ERR_UNSUPPORTED, "decode needs 4-byte aligned rows (even strides, 4-byte aligned bases)");
  }

  const int hpg_tile = p.hpg >= 4 ? 4 : p.hpg;
  p.head_batches = (p.hpg + hpg_tile - 1) / hpg_tile;
  if ((long long)B * G > 65535 || p.head_batches > 65535)
    return fail(VATS_ERR_UNSUPPORTED, "decode grid limit: B*G must be <= 65535 (got %lld)", (long long)B * G);
  dim3 grid(p.num_splits, B * G, p.head_batches);
  const size_t smem = (size_t)vats::kDecodeWarps * hpg_tile * hd * sizeof(float);

  // lanes per key row: the smallest power of two covering hd / vec (<= 32, else several chunks per lane)
  const int chunks = (hd + vec - 1) / vec;
  rc = VATS_ERR_UNSUPPORTED;
  if (vec == 8) {
    if (chunks <= 4) rc = launch_decode_hpg<8, 4, 1>(p, hpg_tile, grid, smem, st);
    else if (chunks <= 8) rc = launch_decode_hpg<8, 8, 1>(p, hpg_tile, grid, smem, st);
    else if (chunks <= 16) rc = launch_decode_hpg<8, 16, 1>(p, hpg_tile, grid, smem, st);
    else rc = launch_decode_hpg<8, 32, 1>(p, hpg_tile, grid, smem, st);
  } else if (vec == 4) {
    if (chunks <= 8) rc = launch_decode_hpg<4, 8, 1>(p, hpg_tile, grid, smem, st);
    else if (chunks <= 16) rc = launch_decode_hpg<4, 16, 1>(p, hpg_tile, grid, smem, st);
    else if (chunks <= 32) rc = launch_decode_hpg<4, 32, 1>(p, hpg_tile, grid, smem, st);
    else rc = launch_decode_hpg<4, 32, 2>(p, hpg_tile, grid, smem, st);
  } else {
    if (chunks <= 16) rc = launch_decode_hpg<2, 16, 1>(p, hpg_tile, grid, smem, st);
    else if (chunks <= 32) rc = launch_decode_hpg<2, 32, 1>(p, hpg_tile, grid, smem, st);
    else if (chunks <= 64) rc = launch_decode_hpg<2, 32, 2>(p, hpg_tile, grid, smem, st);
    else rc = launch_decode_hpg<2, 32, 4>(p, hpg_tile, grid, smem, st);
  }
  if (rc != VATS_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  g_launches = 1;
  g_last_kernel = VATS_LAUNCHED_DECODE_SPLIT;
  if (p.num_splits > 1) {
    vats::decode_combine_kernel<<<B * H, 128, 0, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    g_launches = 2;
  }
  return VATS_OK;
}

int vats_attn_debug_mask(uint8_t* out, const uint8_t* q_valid, const uint8_t* k_valid, int N, int Tq, int Tk,
                         int causal, int left, int right, void* stream) {
  if (N < 0 || Tq < 0 || Tk < 0) return fail(VATS_ERR_INVALID_ARGUMENT, "negative size");
  int rc = check_device();
  if (rc != VATS_OK) return rc;
  if ((long long)N * Tq * Tk == 0) return VATS_OK;
  if (!out) return fail(VATS_ERR_INVALID_ARGUMENT, "out must not be NULL");
  vats::MaskParams mp{Tq, Tk, causal ? 1 : 0, left, right};
  debug_mask_kernel<<<592, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, q_valid, k_valid, N, Tq, Tk, mp);
  CUDA_TRY(cudaGetLastError());
  return VATS_OK;
}

int vats_attn_debug_tile_range(int q0, int block_m, int block_n, int Tq, int Tk, int causal, int left, int right,
                               int* first_tile, int* last_tile) {
  if (!first_tile || !last_tile || block_m <= 0 || block_n <= 0)
    return fail(VATS_ERR_INVALID_ARGUMENT, "bad arguments");
  vats::MaskParams mp{Tq, Tk, causal ? 1 : 0, left, right};
  vats::tile_range(mp, q0, block_m, block_n, first_tile, last_tile);
  return VATS_OK;
}

int vats_attn_debug_tile_is_full(int tile, int q0, int block_m, int block_n, int Tq, int Tk, int causal, int left,
                                 int right) {
  vats::MaskParams mp{Tq, Tk, causal ? 1 : 0, left, right};
  return vats::tile_is_full(mp, tile, q0, block_m, block_n) ? 1 : 0;
}

}  // extern "C"

/*
 * vats_attn.h — C-ABI of the B200 (sm_100a) GQA + sliding-window attention core.
 *
 * This is the drop-in boundary for the one hot path of S-VATS31/vats-multimodal-lm that the LLM,
 * the 2D ViT and the 3D ViT share.  The reference has no FFI of its own: its boundary is a single
 * library call, `F.scaled_dot_product_attention`, made from three modules.  Every entry point below
 * names the reference call site it replaces (paths relative to the reference tree):
 *
 *   vats_attn_prefill        src/optimized_attention.py:657-723            (LLM  Attention.forward SDPA branch, incl. the
 *                                                                           mask build at 668-706 and `extend_kv_heads` 486-497)
 *                            src/optimized_attention.py:628-635            (the intended FA2 call: causal + window_size=(left,right))
 *                            src/transformers/vision/vit_2d/optimized_attention.py:348-423   (SpatialAttention._torch_attention)
 *                            src/transformers/vision/vit_3d/optimized_attention.py:185-348   (_grouped_query_attention, spatial and temporal)
 *   vats_attn_decode         src/optimized_attention.py:508-516 + 709-714  (the intended KV-cache single-query step; the cache
 *                                                                           type is KVCache, src/optimized_attention.py:169-287)
 *   vats_attn_decode_prepare src/optimized_attention.py:463-474 + 224-257  (qk-norm, RoPE and cache append of the new token, fused)
 *   vats_attn_debug_mask     the mask predicate of SURVEY.md §8a-0 (src/optimized_attention.py:519-520, 632-634, 673-675;
 *                            vit_3d/optimized_attention.py:276-277) materialised for bit-exact tests
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers to bf16 (uint16 storage) unless stated otherwise.
 *   - Strides are in ELEMENTS, ordered (sequence, token, head); the head_dim axis is contiguous (stride 1).
 *   - K/V carry G (= query_groups) heads and are never expanded: query head h reads K/V head h / (H/G)
 *     (the `repeat_interleave` ordering of utils/attention_utils.py:27).  G == 1 is the MQA case.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - Return value 0 = success; non-zero = error, message available from vats_attn_last_error() (thread-local).
 *   - There is NO CPU fallback and no other backend: a missing GPU, a non-sm_100 device or an unsupported
 *     geometry is an error.
 *   - The caller owns every buffer, including the decode workspace.
 *
 * Mask predicate (normative; `off = Tk - Tq`, bottom-right aligned):
 *   allowed(n,i,j) = (q_valid == NULL || q_valid[n*Tq+i]) && (k_valid == NULL || k_valid[n*Tk+j])
 *                 && (!causal  || j <= i + off)
 *                 && (left  < 0 || j >= i + off - left)
 *                 && (right < 0 || j <= i + off + right)
 *   A query row with no allowed key produces an all-zero output row (never NaN).
 */
#ifndef VATS_ATTN_H_
#define VATS_ATTN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VATS_ATTN_VERSION 100 /* major*100 + minor */

/* error codes */
#define VATS_OK 0
#define VATS_ERR_INVALID_ARGUMENT 1
#define VATS_ERR_UNSUPPORTED 2 /* geometry / layout / device the kernels do not cover */
#define VATS_ERR_CUDA 3        /* a CUDA runtime / driver call failed */
#define VATS_ERR_WORKSPACE 4   /* decode workspace missing or too small */

/* kernel selector for vats_attn_prefill_ex */
#define VATS_KERNEL_AUTO 0    /* shape-based choice (what vats_attn_prefill uses) */
#define VATS_KERNEL_TCGEN05 1 /* TMA + tcgen05/TMEM tile kernel; error if the geometry is not TMA-legal */
#define VATS_KERNEL_SIMT 2    /* CUDA-core warp kernel for tiny / irregular sequences */
#define VATS_KERNEL_MID 3     /* tcgen05 kernel for <= 256 keys: K/V of a KV group resident in shared memory */

/*
 * Prefill / encoder attention:  O[n,i,h,:] = softmax_j( scale * <Q[n,i,h,:], K[n,j,h/(H/G),:]> | allowed ) . V[n,j,h/(H/G),:]
 *
 *   q,o      [N, Tq, H, hd]   k,v  [N, Tk, G, hd]     (any strides; hd contiguous)
 *   q_valid  [N, Tq] uint8 or NULL  — LLM SDPA-path padding semantics (query rows), src/optimized_attention.py:673-675
 *   k_valid  [N, Tk] uint8 or NULL  — ViT-3D key-padding semantics, vit_3d/optimized_attention.py:276-277
 *   scale    softmax scale applied to the logits (reference: self.softmax_scale or 1/sqrt(hd))
 *   causal   non-zero = causal; the caller applies the reference's `if causal: right_window = 0` itself or not —
 *            the predicate above makes both spellings equivalent.
 *   left/right  window; negative = unlimited.
 */
int vats_attn_prefill(const void* q, const void* k, const void* v, void* o,
                      const uint8_t* q_valid, const uint8_t* k_valid,
                      int N, int Tq, int Tk, int H, int G, int hd,
                      const int64_t q_strides[3], const int64_t k_strides[3],
                      const int64_t v_strides[3], const int64_t o_strides[3],
                      float scale, int causal, int left, int right, void* stream);

/* Same, with an explicit kernel choice (VATS_KERNEL_*).  Used by tests to cross-check the two kernels. */
int vats_attn_prefill_ex(const void* q, const void* k, const void* v, void* o,
                         const uint8_t* q_valid, const uint8_t* k_valid,
                         int N, int Tq, int Tk, int H, int G, int hd,
                         const int64_t q_strides[3], const int64_t k_strides[3],
                         const int64_t v_strides[3], const int64_t o_strides[3],
                         float scale, int causal, int left, int right, int kernel, void* stream);

/*
 * Same, with caller-owned scratch.  Tensors whose rows TMA cannot address (head dims such as 60 or 66 in a dense
 * layout: rows only 4- or 8-byte aligned) are streamed once into `workspace` with the head stride rounded up to 8
 * elements and the TMA-fed kernel runs on the copies; vats_attn_prefill_workspace_bytes() says how much that takes
 * (0 = no scratch needed).  Without a workspace (NULL / too small — what vats_attn_prefill and _ex pass) the kernel
 * stages such rows itself with cp.async, which is slower.  The library never allocates device memory.
 *
 * logit_bound > 0 is a promise by the caller that |<q, k>| <= logit_bound for every (query, key) pair — true by
 * construction behind the reference's qk-norm (utils/attention_utils.py:80-102: q and k are unit vectors, RoPE is a
 * rotation; pass 1.0).  Softmax is shift-invariant, so the kernels then use the bound instead of the row maximum:
 * no maximum pass, no cross-warpgroup exchange, no rescaling of the accumulator.  The result is the same softmax
 * (exponent arguments stay within 2 * bound * scale of zero).  0 = unknown: exact row maxima.  A violated promise can
 * overflow the exponentials — it is a contract, not a hint.
 */
int vats_attn_prefill_ws(const void* q, const void* k, const void* v, void* o,
                         const uint8_t* q_valid, const uint8_t* k_valid,
                         int N, int Tq, int Tk, int H, int G, int hd,
                         const int64_t q_strides[3], const int64_t k_strides[3],
                         const int64_t v_strides[3], const int64_t o_strides[3],
                         float scale, int causal, int left, int right, int kernel,
                         float logit_bound, void* workspace, size_t workspace_bytes, void* stream);
size_t vats_attn_prefill_workspace_bytes(int N, int Tq, int Tk, int H, int G, int hd,
                                         const int64_t q_strides[3], const int64_t k_strides[3],
                                         const int64_t v_strides[3],
                                         const void* q, const void* k, const void* v);

/*
 * Prefill with the multi-GPU output gather fused into the kernel's epilogue (units are sharded by
 * batch x KV-head group, the only exchange is the gather of the outputs).  This rank computes its local block
 * q [N, Tq, H, hd] x k/v [N, Tk, G, hd] and every finished O tile is written, while it is still in shared memory, into
 * the gathered tensor [N_total, Tq, H_total, hd] of EVERY rank at (seq_offset + n, :, head_offset + h, :) — TMA tile
 * stores to peer memory over NVLink.  No second pass over O, no collective kernel.
 *   o_ranks[r]  device pointer of rank r's gathered tensor as mapped into THIS process (symmetric memory / CUDA IPC);
 *               all of them share o_strides.  o_ranks[rank] is the local copy.
 * The caller synchronises the ranks around the call: nobody may still read the previous contents when a peer starts
 * writing, and the gathered tensor is complete on a rank once every rank's launch has finished (a barrier on the
 * stream after the call).  Runs on the tcgen05 tile kernel; the output must be TMA-addressable (16-byte aligned base
 * and strides, head_dim % 8 == 0), otherwise VATS_ERR_UNSUPPORTED (use vats_attn_prefill + a collective).
 */
int vats_attn_prefill_gather(const void* q, const void* k, const void* v, void* const* o_ranks, int world, int rank,
                             int seq_offset, int head_offset, int N_total, int H_total,
                             const uint8_t* q_valid, const uint8_t* k_valid,
                             int N, int Tq, int Tk, int H, int G, int hd,
                             const int64_t q_strides[3], const int64_t k_strides[3],
                             const int64_t v_strides[3], const int64_t o_strides[3],
                             float scale, int causal, int left, int right,
                             float logit_bound, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Backward of vats_attn_prefill: the reference trains through the same modules
 * (training/transformers/nlp/loops/training_loop.py:54-65) and torch differentiates its SDPA call
 * (src/optimized_attention.py:709-714).  Given the forward inputs, the forward output o and dL/do:
 *     dq [N, Tq, H, hd],  dk / dv [N, Tk, G, hd]   (DENSE bf16 outputs; dk / dv sum over the H/G heads of a group)
 * Same mask predicate as the forward (rows / keys that are masked get zero gradient).  bf16 operands, fp32
 * accumulation (mma.sync), deterministic.  workspace: vats_attn_prefill_backward_workspace_bytes(N, Tq, H) bytes of
 * device scratch (the recomputed log-sum-exp and rowsum(do * o)); no zero-fill needed.
 */
int vats_attn_prefill_backward(const void* q, const void* k, const void* v, const void* o, const void* dout,
                               void* dq, void* dk, void* dv,
                               const uint8_t* q_valid, const uint8_t* k_valid,
                               int N, int Tq, int Tk, int H, int G, int hd,
                               const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                               const int64_t o_strides[3], const int64_t do_strides[3],
                               float scale, int causal, int left, int right,
                               void* workspace, size_t workspace_bytes, void* stream);
size_t vats_attn_prefill_backward_workspace_bytes(int N, int Tq, int H);

/* Which kernel VATS_KERNEL_AUTO would pick for this geometry (VATS_KERNEL_TCGEN05 / _SIMT / _MID). Host only. */
int vats_attn_prefill_plan(int N, int Tq, int Tk, int H, int G, int hd,
                           const int64_t q_strides[3], const int64_t k_strides[3],
                           const int64_t v_strides[3], const int64_t o_strides[3],
                           const void* q, const void* k, const void* v);

/*
 * KV-cache decode: one query token per sequence against its cache.
 *
 *   q,o       [B, H, hd]              strides (batch, head) in elements, hd contiguous
 *   k_cache   [B, S_max, G, hd]       strides (batch, token, head) in elements
 *   v_cache   same geometry, own strides
 *   seq_lens  [B] int32 — tokens valid in the cache INCLUDING the token being decoded (the caller appends
 *             the new k,v at position seq_lens[b]-1 before the call).  The query sits at position
 *             seq_lens[b]-1 and attends keys  max(0, L-1-left) .. L-1  (left < 0 = all keys 0..L-1).
 *             seq_lens[b] == 0 gives a zero output row.
 *   workspace device scratch of at least vats_attn_decode_workspace_bytes(...) bytes: split counters (the first
 *             bytes; they must be ZERO on entry and are left zero by every completed call, so one zero-filled buffer
 *             serves any number of stream-ordered calls) followed by the fp32 split-K partials.  After a failed /
 *             aborted launch re-zero it.  Head dims TMA can address — even, <= 128, cache base and strides multiples
 *             of 16 bytes (a head stride of 64 for hd 60) — run on decode_mma_kernel; others on the CUDA-core kernel.
 */
int vats_attn_decode(const void* q, const void* k_cache, const void* v_cache, void* o,
                     const int32_t* seq_lens,
                     int B, int H, int G, int hd, int S_max,
                     const int64_t q_strides[2], const int64_t k_strides[3],
                     const int64_t v_strides[3], const int64_t o_strides[2],
                     float scale, int left,
                     void* workspace, size_t workspace_bytes, void* stream);

size_t vats_attn_decode_workspace_bytes(int B, int H, int G, int hd, int S_max, int left);

/*
 * Pre-core producers of a prefill chunk, fused into one launch (prefill half, 1-D RoPE):
 * q, k [N,T,heads,hd] are L2-normalised (utils/attention_utils.py:80-102, if qk_norm) and rotated at position
 * pos0 + t (src/optimized_attention.py:97-143, order :467-474); v is passed through; all three are rounded to bf16
 * once and written with the caller's output strides — typically a head stride rounded up to 8 elements, the layout
 * the tensor-core kernel can fetch with TMA (this replaces the normalise / rotate / cast / pad passes of the PyTorch
 * path).  in_dtype: 0 = bf16, 1 = fp32.  Strides are (sequence, token, head) in elements.
 */
int vats_attn_prefill_prepare(const void* q_in, const void* k_in, const void* v_in, int in_dtype,
                              void* q_out, void* k_out, void* v_out,
                              const float* cos_table, const float* sin_table,
                              int N, int T, int H, int G, int hd, int pos0,
                              const int64_t qin_strides[3], const int64_t kin_strides[3], const int64_t vin_strides[3],
                              const int64_t qout_strides[3], const int64_t kout_strides[3], const int64_t vout_strides[3],
                              int qk_norm, float eps, void* stream);

/*
 * The same for the ViT passes, with the rotation given as tables (SURVEY.md §8f rank 1, 2-D axial and 3-D RoPE; rank 2):
 *     out[c] = xn[c] * cos_table[tok][c] + xn[partner[c]] * sin_table[tok][c]
 * covers every rotary variant of the reference — vit_2d/optimized_attention.py:128-172 (blocks x1, x2, y1, y2),
 * vit_3d/rope_3d.py:97-219 (interleaved pairs of the h / w blocks or of the t block; other columns cos = 1, sin = 0).
 * q_in / k_in / v_in are logical [No, Ni, T, heads, hd] tensors with strides (outer, inner, token, head) — sequence
 * n = no * Ni + ni — so the ViT-3D temporal pass reads its [B, T, S, heads, hd] projections in place (No = B, Ni = S)
 * and writes dense sequences [No * Ni, T, heads, hd(+pad)]: the transposed copy of vit_3d/optimized_attention.py:474-479
 * is never made.  cos_table / sin_table [T, hd] fp32 (sin signed), partner [hd] int32; all three NULL = no rotation.
 */
int vats_attn_prefill_prepare_table(const void* q_in, const void* k_in, const void* v_in, int in_dtype,
                                    void* q_out, void* k_out, void* v_out,
                                    const float* cos_table, const float* sin_table, const int32_t* partner,
                                    int No, int Ni, int T, int H, int G, int hd,
                                    const int64_t qin_strides[4], const int64_t kin_strides[4], const int64_t vin_strides[4],
                                    const int64_t qout_strides[3], const int64_t kout_strides[3], const int64_t vout_strides[3],
                                    int qk_norm, float eps, void* stream);

/*
 * Pre-core step of one cached decode token, fused into one launch (decode part):
 *     q, k = F.normalize(q, eps), F.normalize(k, eps)     utils/attention_utils.py:80-102   (only if qk_norm != 0)
 *     q, k = rope(q), rope(k)  at position seq_lens[b]-1  src/optimized_attention.py:97-143 (interleaved pairs 2i, 2i+1;
 *                                                         cos/sin tables as built by RoPE._update_cache :83-99)
 *     append k, v to the cache at that position           src/optimized_attention.py:224-257 (intended contract)
 * fp32 arithmetic, one rounding to bf16 at the end.  The order (norm, then RoPE) is src/optimized_attention.py:467-474.
 *
 *   q_in [B,H,hd], k_in / v_in [B,G,hd]   bf16 (in_dtype 0) or fp32 (in_dtype 1); strides (batch, head) in elements
 *   q_out [B,H,hd] bf16;   k_cache / v_cache [B,S_max,G,hd] bf16, strides (batch, token, head)
 *   seq_lens [B] int32: cache length INCLUDING the token being written (the same array vats_attn_decode takes);
 *             sequences with seq_lens[b] <= 0 or > S_max are skipped
 *   cos_table / sin_table [>= max position + 1, hd/2] fp32, or both NULL for no rotation
 */
int vats_attn_decode_prepare(const void* q_in, const void* k_in, const void* v_in, int in_dtype,
                             void* q_out, void* k_cache, void* v_cache, const int32_t* seq_lens,
                             const float* cos_table, const float* sin_table,
                             int B, int H, int G, int hd, int S_max,
                             const int64_t qin_strides[2], const int64_t kin_strides[2], const int64_t vin_strides[2],
                             const int64_t qout_strides[2], const int64_t k_strides[3], const int64_t v_strides[3],
                             int qk_norm, float eps, void* stream);

/* Number of kernels the last successful vats_attn_decode / vats_attn_prefill on this thread launched. */
int vats_attn_last_launch_count(void);

/* Which kernel the last successful compute call on this thread ended in (the main kernel, not helpers such as the
 * repack or the split combine).  Lets tests and the benchmark prove which path a call site reaches. */
#define VATS_LAUNCHED_NONE 0
#define VATS_LAUNCHED_PREFILL_TC 1       /* prefill_tc_kernel<false>: TMA + tcgen05 */
#define VATS_LAUNCHED_PREFILL_TC_LDG 2   /* prefill_tc_kernel<true>: cp.async staging + tcgen05 */
#define VATS_LAUNCHED_PREFILL_SHORT 3    /* prefill_short_kernel: <= 32 keys */
#define VATS_LAUNCHED_PREFILL_SIMT 4     /* prefill_simt_kernel: generic CUDA-core fallback */
#define VATS_LAUNCHED_DECODE_MMA 5       /* decode_mma_kernel: TMA + mma.sync split-K */
#define VATS_LAUNCHED_DECODE_SPLIT 6     /* decode_split_kernel (+ decode_combine_kernel) */
#define VATS_LAUNCHED_PREFILL_PREPARE 7
#define VATS_LAUNCHED_DECODE_PREPARE 8
#define VATS_LAUNCHED_PREFILL_MID 9      /* prefill_mid_kernel: 33..256 keys, K/V resident per (sequence, KV group) */
#define VATS_LAUNCHED_BACKWARD 10        /* attention backward kernels */
int vats_attn_last_kernel(void);

/*
 * Materialise the mask predicate with the kernels' own device function: out[n,i,j] = allowed(n,i,j) as 0/1.
 * out is a DEVICE pointer to N*Tq*Tk bytes.
 */
int vats_attn_debug_mask(uint8_t* out, const uint8_t* q_valid, const uint8_t* k_valid,
                         int N, int Tq, int Tk, int causal, int left, int right, void* stream);

/*
 * Host-only: the KV tile range [*first_tile, *last_tile] (inclusive, in units of `block_n` keys) that the
 * tile-skipping logic visits for the query block [q0, q0+block_m), and whether the tiles at the two ends need the
 * per-element predicate.  *first_tile > *last_tile means "no tile".  Pure integer arithmetic shared with the
 * device code; lets CPU tests prove that tile skipping never drops an allowed (i,j) pair.
 */
int vats_attn_debug_tile_range(int q0, int block_m, int block_n, int Tq, int Tk,
                               int causal, int left, int right,
                               int* first_tile, int* last_tile);
/* Host-only: 1 if tile `tile` of `block_n` keys is entirely allowed for every row of the query block (no predicate
 * needed, k_valid aside), 0 otherwise. */
int vats_attn_debug_tile_is_full(int tile, int q0, int block_m, int block_n, int Tq, int Tk,
                                 int causal, int left, int right);

/* Debug only: block 0 of the tensor-core prefill kernel appends (tag, SM clock) records to this device buffer of
 * 4 roles x capacity x {tag, clock} uint64 records (zero it first); NULL switches tracing off. */
void vats_attn_debug_set_trace(void* dev_buf, int capacity);

const char* vats_attn_last_error(void);
int vats_attn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VATS_ATTN_H_ */

// Micro-benchmark: does a running tcgen05.mma stream starve tcgen05.ld (TMEM -> registers) of OTHER columns?
// Warps 0-3 read columns [0, 128) of TMEM in a loop (4 x 32x32b.x32 + wait); warp 4 optionally keeps the tensor core
// busy with M=128 N=256 K=16 bf16 MMAs accumulating into columns [256, 512) from (garbage) shared-memory operands.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_mma_contention tmem_mma_contention.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "../../vats_multimodal_lm_b200/csrc/ptx.cuh"
using namespace vats::ptx;

__global__ void __launch_bounds__(160, 1) probe(int iters, int mma_on, int mma_cols_overlap, long long* out, uint32_t* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t done_bar;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&done_bar), 1);
    fence_mbar_init();
    stop = 0;
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t tmem = tmem_base_s;
  if (warp == 4) {
    if (lane == 0 && mma_on) {
      const uint64_t adesc = make_smem_desc_sw128(base, 16, 1024);
      const uint64_t bdesc = make_smem_desc_sw128(base + 16384, 16, 1024);
      const uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
      const uint32_t d = tmem + (mma_cols_overlap ? 0u : 256u);
      long long n = 0;
      while (!stop) {
        for (int k = 0; k < 16; ++k) mma_ss(d, adesc, bdesc, idesc, 1u);
        tc_commit(smem_u32(&done_bar));
        mbar_wait(smem_u32(&done_bar), (uint32_t)(n & 1));
        ++n;
      }
      out[gridDim.x + blockIdx.x] = n * 16;
    }
  } else {
  const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t r[128];
  uint32_t acc = 0;
  asm volatile("bar.sync 1, 128;");
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    tmem_ld_32x32b_x32(t + 0, r);
    tmem_ld_32x32b_x32(t + 32, r + 32);
    tmem_ld_32x32b_x32(t + 64, r + 64);
    tmem_ld_32x32b_x32(t + 96, r + 96);
    tmem_ld_wait();
    acc += r[0] + r[127];
  }
  const long long t1 = clock64();
  asm volatile("bar.sync 1, 128;");
  if (threadIdx.x == 0) {
    out[blockIdx.x] = t1 - t0;
    stop = 1;
  }
  if (acc == 0xdeadbeef) sink[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 2 * 148 * 8);
  cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(out, 0, 2 * 148 * 8);
    probe<<<148, 160, 66 * 1024>>>(iters, mode > 0, mode == 2, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s %8.1f cycles per 128-column row read (16 KB/warp) ; MMAs issued meanwhile: %lld (%.0f cycles each)  (%s)\n",
           mode == 0 ? "tcgen05.ld alone" : (mode == 1 ? "tcgen05.ld + MMA stream (other columns)" : "tcgen05.ld + MMA stream (same columns)"),
           (double)h[0] / iters, h[148], h[148] ? (double)h[0] / h[148] : 0.0, cudaGetErrorString(e));
  }
  return 0;
}

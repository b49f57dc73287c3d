// Micro-benchmark: TMEM -> register read throughput (tcgen05.ld) per SM, by shape, batch and warps per sub-partition.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_bw tmem_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "../../vats_multimodal_lm_b200/csrc/ptx.cuh"
using namespace vats::ptx;

__device__ __forceinline__ void ld_32x32b_x64(uint32_t taddr, uint32_t* r) {
  tmem_ld_32x32b_x32(taddr, r);
  tmem_ld_32x32b_x32(taddr + 32, r + 32);
}
__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t* r) {   // 16 lanes x 256 bit, x8: 32 regs/thread
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// mode 0: 32x32b.x16 + wait each;  1: 32x32b.x32 + wait each;  2: 4 x (x32) then one wait;  3: 16x256b.x8 + wait;
// 4: x16 issue-only stream, wait every 8;  5: tcgen05.st 32x32b.x32 + wait::st
template <int MODE>
__global__ void probe(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t r[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)((it * 32) & 255);
    if (MODE == 0) {
      tmem_ld_32x32b_x16(tmem + col, r);
      tmem_ld_wait();
      acc += r[0] + r[15];
    } else if (MODE == 1) {
      tmem_ld_32x32b_x32(tmem + col, r);
      tmem_ld_wait();
      acc += r[0] + r[31];
    } else if (MODE == 2) {
      tmem_ld_32x32b_x32(tmem + 0, r);
      tmem_ld_32x32b_x32(tmem + 32, r + 32);
      tmem_ld_32x32b_x32(tmem + 64, r + 64);
      tmem_ld_32x32b_x32(tmem + 96, r + 96);
      tmem_ld_wait();
      acc += r[0] + r[127];
    } else if (MODE == 3) {
      ld_16x256b_x8(tmem + col, r);
      tmem_ld_wait();
      acc += r[0] + r[31];
    } else if (MODE == 4) {
      tmem_ld_32x32b_x16(tmem + col, r + 16 * (it & 7));
      if ((it & 7) == 7) {
        tmem_ld_wait();
        acc += r[0] + r[127];
      }
    } else {
      tmem_st_32x32b_x32(tmem + col, r);
      tmem_st_wait();
    }
  }
  tmem_ld_wait();
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0xdeadbeef) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base_s, 512);
  }
}

template <int MODE>
void run(const char* name, int threads, int bytes_per_iter_per_warp) {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 148 * 8);
  cudaMalloc(&sink, 4);
  const int iters = 4096;
  probe<MODE><<<148, threads>>>(iters, out, sink);
  cudaDeviceSynchronize();
  probe<MODE><<<148, threads>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = (double)h[0];
  const int warps = threads / 32;
  printf("%-34s warps/SM %2d: %8.1f cycles/iter, %7.1f B/clk/SM, %6.1f B/clk per sub-partition  (%s)\n", name, warps,
         cyc / iters, (double)bytes_per_iter_per_warp * warps * iters / cyc,
         (double)bytes_per_iter_per_warp * warps * iters / cyc / 4.0, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(sink);
}

int main() {
  for (int threads : {128, 256, 512}) {
    run<0>("ld 32x32b.x16 + wait", threads, 16 * 128);
    run<1>("ld 32x32b.x32 + wait", threads, 32 * 128);
    run<2>("ld 4 x 32x32b.x32, one wait", threads, 128 * 128);
    run<3>("ld 16x256b.x8 + wait", threads, 32 * 128);
    run<4>("ld 32x32b.x16 stream, wait per 8", threads, 16 * 128);
    run<5>("st 32x32b.x32 + wait", threads, 32 * 128);
  }
  return 0;
}

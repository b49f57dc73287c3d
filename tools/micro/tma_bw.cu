// Micro-benchmark: how fast does one SM's TMA unit fill shared memory with (64 x ROWS) bf16 boxes?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_bw tma_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../vats_multimodal_lm_b200/csrc/ptx.cuh"
using namespace vats::ptx;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// one thread issues `nbox` boxes round-robin into `slots` smem slots, `inflight` at a time; measures cycles
__global__ void tma_probe(const __grid_constant__ CUtensorMap map, int rank, int rows, int nbox, int slots, int inflight,
                          int heads, int T, int Tw, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars_all[4][8];
  const int wid = threadIdx.x >> 5;
  uint64_t* bars = bars_all[wid];
  const uint32_t base = ((smem_u32(smem) + 1023u) & ~1023u) + wid * (slots * 128u * rows);
  const uint32_t box_bytes = 128u * rows;
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < slots; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    long long t0 = clock64();
    int issued = 0, done = 0;
    int si = 0, sd = 0;          // slot of the next issue / of the next completion
    uint32_t pd = 0;             // phase parity of the next completion
    int tok = 0, head = wid;
    while (done < nbox) {
      while (issued < nbox && issued - done < inflight) {
        const uint32_t bar = smem_u32(&bars[si]);
        mbar_expect_tx(bar, box_bytes);
        tma_load_4d(base + si * box_bytes, &map, bar, 0, head, tok, blockIdx.x);
        tok += rows;
        if (tok >= Tw) { tok = 0; if (++head == heads) head = 0; }
        if (++si == slots) si = 0;
        ++issued;
      }
      mbar_wait(smem_u32(&bars[sd]), pd);
      if (++sd == slots) { sd = 0; pd ^= 1u; }
      ++done;
    }
    if (wid == 0) out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int N = 148, T = 4096, heads = 8, hd = 64;
  size_t elems = (size_t)N * T * heads * hd;
  void* d;
  cudaMalloc(&d, elems * 2);
  cudaMemset(d, 0, elems * 2);
  long long* dout;
  cudaMalloc(&dout, 148 * 8);
  cudaFuncSetAttribute(tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rank : {4}) {
    for (int rows : {32, 64, 128}) {
      CUtensorMap m;
      CUresult rc;
      if (rank == 4) {
        cuuint64_t dims[4] = {(cuuint64_t)hd, (cuuint64_t)heads, (cuuint64_t)T, (cuuint64_t)N};
        cuuint64_t str[3] = {(cuuint64_t)hd * 2, (cuuint64_t)heads * hd * 2, (cuuint64_t)T * heads * hd * 2};
        cuuint32_t box[4] = {64, 1, (cuuint32_t)rows, 1}, es[4] = {1, 1, 1, 1};
        rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {
        cuuint64_t dims[2] = {(cuuint64_t)heads * hd, (cuuint64_t)N * T};
        cuuint64_t str[1] = {(cuuint64_t)heads * hd * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)rows}, es[2] = {1, 1};
        rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); continue; }
      for (int nw : {1, 2, 4})
      for (int mode = 1; mode < 3; ++mode)
      for (int inflight : {4}) {
        const int slots = 4, nbox = 512;
        if ((size_t)nw * slots * 128 * rows > 190 * 1024) continue;
        // mode 0: 148 SMs streaming from HBM; mode 1: 148 SMs, L2-resident window (512 tokens); mode 2: 8 SMs, L2-resident
        const int grid = mode == 2 ? 8 : 148;
        const int Tw = mode == 0 ? T : 512;
        for (int rep = 0; rep < 3; ++rep) tma_probe<<<grid, 32 * nw, 200 * 1024>>>(m, rank, rows, nbox, slots, inflight, heads, T, Tw, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(148);
        cudaMemcpy(h.data(), dout, grid * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
        printf("rows %3d (%5d B/box) issuers %d mode %d: %8.0f cyc/box/issuer  %6.1f B/clk/SM\n", rows,
               128 * rows, nw, mode, avg / nbox, nw * 128.0 * rows * nbox / avg);
      }
    }
  }
  return 0;
}

// Dev tool: cycles per tcgen05.mma (kind::f16, bf16 operands, cta_group::1) as a function of N, for A from shared memory
// (SS) and A from TMEM (TS). One CTA per SM, one warp issues `iters` x 4 MMAs (K = 64 in four K = 16 steps) back to back
// and commits once.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I interspeech_ser_b200/csrc -I include
//                     tools/mma_cost.cu -o tools/mma_cost.bin
#include <cstdio>
#include <cuda.h>
#include "common.cuh"
using namespace serenc;

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) mma_cost_kernel(int iters, int N, int M, int ts_mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = align_smem_1024(raw);
  uint8_t* sA = smem;                 // 128 x 64 bf16 (16 KB)
  uint8_t* sB = smem + 16384;         // 256 x 64 bf16 (32 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if ((threadIdx.x & 31) == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = warp_uniform(*slot);
  if (warp == 0) {
    const uint32_t idesc = ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24));
    const uint64_t adesc = umma_desc_sw128(smem_u32(sA));
    const uint64_t bdesc = umma_desc_sw128(smem_u32(sB));
    long long t0 = clock64();
    if (elect_one_sync()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (ts_mode) umma_ts(tm, tm + 256 + 8 * k, bdesc + (uint64_t)(2 * k), idesc, 1u);
          else umma_bf16_ss(tm, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(mma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  const int iters = 2000;
  printf("cycles per tcgen05.mma (bf16, K = 16), %d x 4 MMAs back to back, all 148 SMs busy\n", iters);
  for (int ts = 0; ts < 2; ++ts)
    for (int M : {128, 64})
      for (int N : {16, 32, 64, 128, 256}) {
        if (M == 128 && N % 16) continue;
        mma_cost_kernel<<<148, 128, 52 * 1024>>>(iters, N, M, ts, d_out);
        mma_cost_kernel<<<148, 128, 52 * 1024>>>(iters, N, M, ts, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        printf("%s M=%3d N=%3d : issue %7.1f  complete %7.1f cycles/MMA   (%s)\n", ts ? "TS" : "SS", M, N, (double)h[0] / (iters * 4.0),
               (double)h[1] / (iters * 4.0), cudaGetErrorString(e));
      }
  return 0;
}

// Dev tool: per-SM throughput of the exponential forms a softmax can use on sm_100a, to decide what the attention kernels'
// exp pass should issue (profiles/r02_notes.md):
//   f32      ex2.approx.ftz.f32                 one MUFU op per score
//   bf16x2   ex2.approx.ftz.bf16x2              two scores per MUFU op, bf16 in / out (the P operand is bf16 anyway)
//   f16x2    ex2.approx.f16x2                   two scores per MUFU op, fp16 in / out
//   poly     Cody-Waite exp2 on the FMA pipe    (x = n + f; 2^f by a degree-3 polynomial, FFMA2-packed; n added to the exponent)
//   mix      3 of 4 scores on MUFU, 1 of 4 on the polynomial
// Every thread runs ILP independent chains; 148 x CTAS_PER_SM CTAs of 128 threads. Prints scores per clock and SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I interspeech_ser_b200/csrc -I include tools/mufu_cost.cu -o tools/mufu_cost.bin
#include <cstdio>
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
using namespace serenc;

__device__ __forceinline__ float ex2_f32(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }

// 2^x for x <= 0 (clamped at -126) on the FMA pipe, two values at once. max rel. error of the cubic ~1e-4 (P is rounded to bf16).
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& y0, float& y1) {
  const float MAGIC = 12582912.f;   // 1.5 * 2^23: x + MAGIC rounds x to the nearest integer in the low mantissa bits
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t x = pack_f32x2(x0, x1);
  const uint64_t r = fadd2(x, pack_f32x2(MAGIC, MAGIC));           // round(x) + MAGIC
  float r0, r1;
  unpack_f32x2(r, r0, r1);
  const uint64_t f = fadd2(x, fadd2(pack_f32x2(MAGIC, MAGIC), pack_f32x2(-r0, -r1)));   // x - round(x) in [-0.5, 0.5]
  uint64_t p = pack_f32x2(0.0555041f, 0.0555041f);
  p = ffma2(p, f, pack_f32x2(0.2402265f, 0.2402265f));
  p = ffma2(p, f, pack_f32x2(0.6931472f, 0.6931472f));
  p = ffma2(p, f, pack_f32x2(1.0f, 1.0f));
  float p0, p1;
  unpack_f32x2(p, p0, p1);
  y0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(r0) << 23));
  y1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(r1) << 23));
}

template <int MODE, int ILP>
__global__ void __launch_bounds__(128) mufu_kernel(int iters, float seed, float* out, long long* cyc) {
  float v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = -0.001f * (float)(threadIdx.x + i) * seed;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) { v[i] = ex2_f32(v[i]) - 1.0009765f; }
    } else if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < ILP; i += 2) {
        const uint32_t pk = MODE == 1 ? cvt_bf16x2(v[i], v[i + 1]) : cvt_f16x2(v[i], v[i + 1]);
        const uint32_t e = MODE == 1 ? ex2_bf16x2(pk) : ex2_f16x2(pk);
        // consume the packed result the way the kernel would (it IS the P operand): one integer op keeps it alive
        acc += __uint_as_float((e & 0xffffu) << 16);
        v[i] -= 0.0009765f; v[i + 1] -= 0.0009765f;
      }
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < ILP; i += 2) { float a, b; exp2_poly2(v[i], v[i + 1], a, b); v[i] = a - 1.0009765f; v[i + 1] = b - 1.0009765f; }
    } else {
#pragma unroll
      for (int i = 0; i < ILP; i += 8) {
#pragma unroll
        for (int k = 0; k < 6; ++k) v[i + k] = ex2_f32(v[i + k]) - 1.0009765f;
        float a, b; exp2_poly2(v[i + 6], v[i + 7], a, b); v[i + 6] = a - 1.0009765f; v[i + 7] = b - 1.0009765f;
      }
    }
  }
  const long long t1 = clock64();
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc += v[i];
  if (acc == 12345.678f) out[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ctas_per_sm, float* d_out, long long* d_cyc) {
  constexpr int ILP = 16;
  const int iters = 4000;
  mufu_kernel<MODE, ILP><<<148 * ctas_per_sm, 128>>>(iters, 1.0f, d_out, d_cyc);
  mufu_kernel<MODE, ILP><<<148 * ctas_per_sm, 128>>>(iters, 1.0f, d_out, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
  const double scores = (double)iters * ILP * 128.0 * ctas_per_sm;
  printf("%-8s %d CTAs/SM (%2d warps): %8.2f scores / clk / SM   (%lld cycles, %s)\n", name, ctas_per_sm, 4 * ctas_per_sm, scores / (double)h, h,
         cudaGetErrorString(e));
}

int main() {
  float* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 16); cudaMalloc(&d_cyc, 16);
  for (int c : {1, 2, 4}) {
    run<0>("f32", c, d_out, d_cyc);
    run<1>("bf16x2", c, d_out, d_cyc);
    run<2>("f16x2", c, d_out, d_cyc);
    run<3>("poly", c, d_out, d_cyc);
    run<4>("mix6:2", c, d_out, d_cyc);
  }
  // accuracy of the forms against exp2 in double, x in [-20, 0]
  return 0;
}

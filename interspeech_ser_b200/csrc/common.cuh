// Common device/host helpers for libserenc (sm_100a only).
//
// PTX wrappers for mbarrier / TMA / tcgen05 written from the PTX ISA; the sm_100 shared-memory
// descriptor and instruction-descriptor bit layouts were cross-checked against the CuTe headers
// vendored in this image (cute/arch/mma_sm100_desc.hpp) but no CuTe/CUTLASS code is used.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

#ifndef SERENC_WATCHDOG
#define SERENC_WATCHDOG 1  // bounded mbarrier spins: a protocol bug traps instead of hanging the GPU
#endif

namespace serenc {

// One utterance of a batch: where its samples live and where its conv0 output rows go.
struct UttSpan {
  int64_t sample_start;  // offset of the first sample in the waveform buffer
  int32_t sample_len;    // valid samples
  int32_t T0;            // valid conv0 frames (wav2vec2 family)
  int64_t row0;          // first conv0 output row
  int32_t slot;          // conv0 output rows owned by this utterance
  int32_t pad_;
};

// Sample `idx` of a waveform buffer that holds float32 samples or int16 PCM (include/serenc.h serenc_wav_dtype). int16
// is scaled by 1 / 32768 exactly as librosa / soundfile do when they hand the reference float32 samples
// (preprocess_speech.py:47), so both forms give bit-identical results.
__device__ __forceinline__ float load_sample(const void* wav, int is_i16, int64_t idx) {
  return is_i16 ? (float)reinterpret_cast<const int16_t*>(wav)[idx] * (1.0f / 32768.0f)
                : reinterpret_cast<const float*>(wav)[idx];
}

// ---------------------------------------------------------------------------------------------
// small math
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// packed 2 x fp32 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2, one issue slot for two lanes of work)
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// exact (erf) GELU, the activation every model on this path uses (HF ACT2FN["gelu"]).
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// Same function for the GEMM / LayerNorm epilogues, where libm's branch-free erff (~27 instructions per element)
// made the FFN up-projection issue- and power-bound:  gelu(x) = hx + |hx| (1 - erfc(|x|/sqrt2)),  hx = x/2, with
//   erfc(z) = exp2(P7(t)),  t = z/2 - 1 in [-1, 1]  (z clamped at 4: erfc(4) = 1.5e-8)
// P7 = degree-7 least-squares fit of log2(erfc) on a Chebyshev grid. Max |error| of the GELU in fp32 arithmetic is
// 6.0e-7 over [-9, 9] (checked against scipy in oracle/gelu_fit.py) — round-off level, far below the bf16 rounding
// applied to the result. 13 instructions per element, one MUFU.EX2.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float t = fminf(fmaf(fabsf(x), 0.35355339059327373f, -1.0f), 1.0f);
  float p = -2.886363771e-03f;
  p = fmaf(p, t, 1.296435855e-02f);
  p = fmaf(p, t, -3.462206945e-02f);
  p = fmaf(p, t, 8.234396577e-02f);
  p = fmaf(p, t, -1.898051500e-01f);
  p = fmaf(p, t, -5.330767155e+00f);
  p = fmaf(p, t, -1.274813366e+01f);
  p = fmaf(p, t, -7.739973545e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
  const float hx = 0.5f * x;
  const float a = fabsf(hx);
  return fmaf(-a, e, hx + a);
}

// The form every epilogue and row kernel uses since round 2 (two values at once):
//   gelu(x) = h + h * tanh(x * (B1 + B3 x^2 + B5 x^4)),  h = x / 2,  x^2 clamped at 64
// (B1, B3, B5) = minimax fit of the exact erf GELU (oracle/gelu_fit.py): max |error| 2.5e-5 from the fit plus the
// 2^-11 relative error of MUFU.TANH, i.e. <= 2.5e-4 |x| - an order of magnitude below the bf16 rounding (2^-9 relative)
// every one of these results goes through. Measured on the oracle (oracle/gelu_fit.py --model): replacing F.gelu by
// this form, tanh noise included, moves WavLM-large's pooled embeddings by 1e-4 (max relative), cosine 0.9999998.
// The clamp keeps the argument monotone: B5 < 0 would turn the polynomial over beyond |x| = 10.
// 6 packed FP32 instructions + 2 FMNMX + 2 MUFU per pair, against ~19 for the erf-exact form below: the GELU was ~10 of
// the ~22 instructions per element of conv0 and the conv LayerNorm passes (issue-bound), and 9 % of an FC1 launch.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_fast2(float x0, float x1, float& y0, float& y1) {
  const uint64_t x = pack_f32x2(x0, x1);
  float q0, q1;
  unpack_f32x2(fmul2(x, x), q0, q1);
  const uint64_t x2 = pack_f32x2(fminf(q0, 64.f), fminf(q1, 64.f));
  uint64_t u = ffma2(x2, pack_f32x2(-3.515174775e-04f, -3.515174775e-04f), pack_f32x2(3.700565057e-02f, 3.700565057e-02f));
  u = ffma2(u, x2, pack_f32x2(7.975078789e-01f, 7.975078789e-01f));
  float u0, u1;
  unpack_f32x2(fmul2(u, x), u0, u1);
  const uint64_t t = pack_f32x2(tanh_approx(u0), tanh_approx(u1));
  const uint64_t h = fmul2(x, pack_f32x2(0.5f, 0.5f));
  unpack_f32x2(ffma2(h, t, h), y0, y1);
}

// two erf-exact GELUs at once (the round-1 form, kept for op-level comparisons): the polynomial runs on FFMA2 (7 packed
// instead of 14 scalar FMAs); same arithmetic per lane as gelu_erf_fast, so results are bit-identical to it
__device__ __forceinline__ void gelu_erf_fast2(float x0, float x1, float& y0, float& y1) {
  const float t0 = fminf(fmaf(fabsf(x0), 0.35355339059327373f, -1.0f), 1.0f);
  const float t1 = fminf(fmaf(fabsf(x1), 0.35355339059327373f, -1.0f), 1.0f);
  const uint64_t t = pack_f32x2(t0, t1);
  uint64_t p = pack_f32x2(-2.886363771e-03f, -2.886363771e-03f);
  p = ffma2(p, t, pack_f32x2(1.296435855e-02f, 1.296435855e-02f));
  p = ffma2(p, t, pack_f32x2(-3.462206945e-02f, -3.462206945e-02f));
  p = ffma2(p, t, pack_f32x2(8.234396577e-02f, 8.234396577e-02f));
  p = ffma2(p, t, pack_f32x2(-1.898051500e-01f, -1.898051500e-01f));
  p = ffma2(p, t, pack_f32x2(-5.330767155e+00f, -5.330767155e+00f));
  p = ffma2(p, t, pack_f32x2(-1.274813366e+01f, -1.274813366e+01f));
  p = ffma2(p, t, pack_f32x2(-7.739973545e+00f, -7.739973545e+00f));
  float p0, p1, e0, e1;
  unpack_f32x2(p, p0, p1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(p0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(p1));
  const float h0 = 0.5f * x0, h1 = 0.5f * x1;
  const float a0 = fabsf(h0), a1 = fabsf(h1);
  y0 = fmaf(-a0, e0, h0 + a0);
  y1 = fmaf(-a1, e1, h1 + a1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// ---------------------------------------------------------------------------------------------
// shared-memory address / mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// try_wait with a suspend-time hint (ns): ptxas turns a failed probe into NANOSLEEP.SYNCS <hint>, i.e. the thread
// stops burning issue slots — at the price of wake-up latency. Only for waits that are long and not on the
// critical path (the attention kernel's 128 softmax threads), never for the GEMM pipeline.
template <uint32_t HINT_NS>
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity), "r"(HINT_NS)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __noinline__ void mbar_watchdog_trap(uint32_t parity) {
  printf("serenc: mbarrier watchdog (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if SERENC_WATCHDOG
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0xfffu) == 0 && global_timer_ns() - t0 > 4000000000ull) mbar_watchdog_trap(parity);  // 4 s: a bug
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}
template <uint32_t HINT_NS>
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
#if SERENC_WATCHDOG
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait_hint<HINT_NS>(bar, parity)) {
    if (((++spins) & 0xffu) == 0 && global_timer_ns() - t0 > 4000000000ull) mbar_watchdog_trap(parity);
  }
#else
  while (!mbar_try_wait_hint<HINT_NS>(bar, parity)) {
  }
#endif
}

// Dynamic shared memory base rounded up to 1024 B (128B-swizzle atoms) WITHOUT leaving the shared address space:
// plain pointer arithmetic on the extern array, so the compiler keeps emitting LDS/STS (a uintptr_t round trip
// degrades every access behind it to generic LD/ST).
__device__ __forceinline__ uint8_t* align_smem_1024(uint8_t* raw) {
  return raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
}

// One elected lane of a CONVERGED warp. The single-thread tcgen05 / TMA instructions take their operands from
// uniform registers; issued from inside an `if (lane == 0)` region the compiler cannot prove the operands
// warp-uniform and wraps every UTCHMMA / UTMALDG in an elect + 5 x R2UR "waterfall" loop (~100 cycles per
// instruction, measured: it made every MMA narrower than N = 256 issue-bound). Role warps therefore run their
// loops with all 32 lanes (uniform control flow, uniform operands) and only predicate the issue itself.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// value of lane 0, marked warp-uniform for the compiler
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch. The kernels of the layer loop (LayerNorm, GEMMs, attention) are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: the grid is scheduled while the tail of the previous kernel of the
// stream is still running, sets up (barriers, TMEM, tensor-map prefetch) and then waits here until the previous grid has
// completed and its writes are visible. Nothing produced or overwritten by an earlier kernel is touched before pdl_wait().
// pdl_trigger() lets the NEXT kernel of the stream be scheduled once every CTA of this grid has got this far (or has exited);
// it always follows this kernel's own wait, so a third kernel never overtakes two. Both are no-ops in a plain launch.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) loads, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// TMA store (shared -> global through a tensor map, bulk-group completion). The caller has made its generic-proxy writes
// of the tile visible with fence_proxy_async_smem() + a barrier before the issuing thread gets here.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source (it may be overwritten)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp receives TMEM lane (lane_base+i).
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA-pair (cluster of 2, tcgen05 cta_group::2) variants
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// Arrive on an mbarrier of a CTA of this cluster. RELAXED: the callers only hand back TMEM accumulators, whose reads
// are ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync. (.release.cluster compiled to
// MEMBAR.ALL.GPU, i.e. every epilogue warp drained its global stores before the accumulator could be reused:
// ncu showed 0.67 membar-stalled warps per issue on the FC1 GEMM.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory, completion bytes credited to an mbarrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A (256 rows: 128 from each CTA's smem) * B (N rows: N/2 from each CTA's smem)
__device__ __forceinline__ void umma_bf16_ss_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far complete) on the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (sm_100 "version 1"):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for SW128 K-major) | [32,46) SBO>>4 = 1024 B (8 rows x 128 B)
//   [46,48) version=1 | [61,64) layout: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32 (bit4), A=B=bf16 (bits 7,10), both K-major, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// cp.async / ldmatrix / mma.sync (used by the attention kernel)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
  int sz = valid ? 16 : 0;  // src-size 0 => zero-fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

}  // namespace serenc

// ---------------------------------------------------------------------------------------------
// host-side error plumbing (thread-local message, negative status codes; see include/serenc.h)
// ---------------------------------------------------------------------------------------------
namespace serenc {
void set_error(const char* fmt, ...);
void note_cuda_error(int cuda_error);   // remembered per thread: the entry-point guard decides whether it poisons the handle
}

#define SERENC_CUDA_OK(expr)                                                                        \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      serenc::note_cuda_error((int)_e);                                                             \
      serenc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SERENC_ERR_CUDA;                                                                       \
    }                                                                                               \
  } while (0)

#define SERENC_TRY(expr)       \
  do {                         \
    int _s = (expr);           \
    if (_s != 0) return _s;    \
  } while (0)

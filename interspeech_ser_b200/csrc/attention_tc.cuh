// tcgen05 flash attention for head_dim 64 (WavLM-large, wav2vec2/HuBERT-large, Whisper): packed variable-length,
// non-causal, optional WavLM gated relative-position bias.
//
// One CTA = 128 query rows of one (utterance, head). Both GEMMs of the attention run on the 5th-gen tensor cores
// with accumulators in TMEM; the softmax runs one query row per thread straight out of TMEM, so there are no
// cross-lane reductions and none of the ldmatrix / mma.sync fragment traffic of the previous kernel:
//
//   control thread (warp 4):  TMA Q, K_j, V_j  ->  S = Q K_j^T (tcgen05.mma, N = 128)  ->  ...  ->  O += P_j V_j
//   softmax threads (warps 0-3, thread r <-> query row r <-> TMEM lane r):
//        tcgen05.ld S row -> scale (+ gate * bias[key - query]) -> running max / sum (exp2 domain)
//        -> rescale O in TMEM when the max moved (tcgen05.ld / tcgen05.st) -> P (bf16) into shared memory in the
//        128B-swizzled K-major layout the second MMA reads -> final O / l -> bf16 rows of the output.
//
// Keys are processed 64 at a time. TMEM: S = columns [0, 64), O = columns [64, 128) of a 128-column allocation;
// shared memory Q 16 KB + K 8 KB + V 8 KB + P 16 KB, single-buffered (K_{j+1} is fetched as soon as S_j is
// complete, V_{j+1} as soon as O += P_j V_j is). S_{j+1} = Q K_{j+1}^T is issued the moment the softmax threads
// have finished reading S_j, AHEAD of O += P_j V_j, so that MMA and its commit latency sit under the next row-max
// pass. WavLM: gate[row, head] comes precomputed (wavlm_gate_kernel), the bias window of the whole query tile is
// loaded once under the Q/K/V loads, and the biased scores are written back over S so the exp pass does not redo
// the bias. The running max only moves when it would grow by more than 2^8 (exact), so O is rarely rescaled. FOUR CTAs per SM (<= 80 registers: the softmax makes two passes over S in TMEM instead of holding the
// row) overlap one CTA's softmax with the others' MMAs and loads.
// V is consumed as an MN-major (head-dim contiguous) B operand directly from its row-major [key, d] tile.
#pragma once
#include "attention_params.cuh"
#include "common.cuh"
#include <type_traits>

namespace serenc {

constexpr int FA_BM = 128;   // query rows per CTA (= TMEM lanes)
constexpr int FA_BN = 64;    // keys per block
constexpr int FA_HD = 64;
constexpr int FA_THREADS = 160;
constexpr int FA_Q_BYTES = FA_BM * FA_HD * 2;   // 16 KB
constexpr int FA_KV_BYTES = FA_BN * FA_HD * 2;  // 8 KB
constexpr int FA_P_BYTES = FA_BM * FA_BN * 2;   // 16 KB
constexpr int FA_BAR_BYTES = 64;
constexpr int FA_SMEM_FIXED = FA_Q_BYTES + 2 * FA_KV_BYTES + FA_P_BYTES + FA_BAR_BYTES + 1024;
constexpr int FA_SMEM_LIMIT = 227 * 1024;    // opt-in dynamic shared memory of one CTA (bias window of very long utterances)
// bias-window entries a query tile of an utterance with tmax frames can see (one per (key - query) offset)
inline int fa_window_entries(int tmax) { return FA_BM + FA_BN * ((tmax + FA_BN - 1) / FA_BN); }
inline size_t fa_smem_bytes(bool wavlm, int tmax) {
  return (size_t)FA_SMEM_FIXED + (wavlm ? (size_t)fa_window_entries(tmax) * 4 : 0);
}
constexpr int FA_TMEM_COLS = 128;
constexpr uint32_t FA_WAIT_HINT_NS = 2000;   // softmax threads sleep (NANOSLEEP.SYNCS) instead of spinning on S / O barriers
constexpr int FA_TMEM_S = 0, FA_TMEM_O = 64;
constexpr float FA_RESCALE_LOG2 = 8.0f;      // the running max is only raised when it would grow by more than this

// MN-major (N contiguous), 128B-swizzled B operand: 8-row (K) groups are 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x <= 0 on the FMA pipe, two values at once: x = n + f with n = round(x) (magic-number add), 2^f by a cubic
// (|f| <= 0.5, max relative error ~1e-4: below the bf16 rounding of P), n added into the exponent field by one LEA.
// Used for a share of a thread's exponentials when the MUFU pipe (one ex2 per 8 cycles, warp and scheduler) is the
// bound of the softmax.
__device__ __forceinline__ void exp2_fma2(float x0, float x1, float& y0, float& y1) {
  const float MAGIC = 12582912.f;   // 1.5 * 2^23
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t x = pack_f32x2(x0, x1);
  const uint64_t r = fadd2(x, pack_f32x2(MAGIC, MAGIC));                 // low mantissa bits = round(x)
  const uint64_t f = fadd2(x, fadd2(pack_f32x2(MAGIC, MAGIC), fmul2(r, pack_f32x2(-1.f, -1.f))));   // x - round(x)
  uint64_t p = pack_f32x2(0.0555041f, 0.0555041f);
  p = ffma2(p, f, pack_f32x2(0.2402265f, 0.2402265f));
  p = ffma2(p, f, pack_f32x2(0.6931472f, 0.6931472f));
  p = ffma2(p, f, pack_f32x2(1.0f, 1.0f));
  float p0, p1, r0, r1;
  unpack_f32x2(p, p0, p1);
  unpack_f32x2(r, r0, r1);
  y0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(r0) << 23));
  y1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(r1) << 23));
}
__device__ __forceinline__ void softmax_group_sync() {  // the 128 softmax threads only (named barrier 1)
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

constexpr int FA_TRACE_SLOTS = 48;   // [0,24): softmax thread 0, [24,48): control thread
constexpr int FA_TRACE_CTAS = 64;    // CTAs [0,32) and [2048,2080) of the linearised grid
__device__ __forceinline__ void fa_stamp(long long* tr, int slot) {
  if (tr) tr[slot] = clock64();
}

// gate[row, head] of WavLM's gated relative position bias (HF modeling_wavlm.py:167-176) from the layer input.
// Only the SUMS of outputs 0-3 and 4-7 of gru_rel_pos_linear are used (view(.., 2, 4).sum(-1)), so the [8, 64]
// weight collapses to two 64-vectors. One thread per (row, head): 128 contiguous bytes each, fully coalesced.
__global__ void __launch_bounds__(256)
wavlm_gate_kernel(const bf16* __restrict__ hln, int64_t rows, int d, int heads, const float* __restrict__ gru_w,
                  const float* __restrict__ gru_b, const float* __restrict__ gru_const, float* __restrict__ gate) {
  __shared__ float4 s_w[FA_HD / 2];   // (wa[k], wb[k], wa[k+1], wb[k+1])
  __shared__ float s_b[2];
  if (threadIdx.x < 2 * FA_HD) {
    const int k = threadIdx.x >> 1, g = threadIdx.x & 1;
    const float* w = gru_w + g * 4 * FA_HD + k;
    reinterpret_cast<float*>(s_w)[threadIdx.x] = (__ldg(w) + __ldg(w + FA_HD)) + (__ldg(w + 2 * FA_HD) + __ldg(w + 3 * FA_HD));
  }
  if (threadIdx.x < 2) {
    const float* bb = gru_b + threadIdx.x * 4;
    s_b[threadIdx.x] = (__ldg(bb) + __ldg(bb + 1)) + (__ldg(bb + 2) + __ldg(bb + 3));
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * heads) return;
  const int64_t row = i / heads;
  const int h = (int)(i - row * heads);
  const uint4* x4 = reinterpret_cast<const uint4*>(hln + row * d + h * FA_HD);
  uint4 xr[FA_HD / 8];
#pragma unroll
  for (int c = 0; c < FA_HD / 8; ++c) xr[c] = __ldg(x4 + c);
  float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
#pragma unroll
  for (int c = 0; c < FA_HD / 8; ++c) {
    const uint32_t uu[4] = {xr[c].x, xr[c].y, xr[c].z, xr[c].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 xv = unpack_bf16x2(uu[e]);
      const float4 w = s_w[c * 4 + e];
      a0 = fmaf(xv.x, w.x, a0); b0 = fmaf(xv.x, w.y, b0);
      a1 = fmaf(xv.y, w.z, a1); b1 = fmaf(xv.y, w.w, b1);
    }
  }
  const float ga = 1.f / (1.f + __expf(-((a0 + a1) + s_b[0])));
  const float gb = 1.f / (1.f + __expf(-((b0 + b1) + s_b[1])));
  gate[i] = ga * (gb * __ldg(gru_const + h) - 1.f) + 2.f;
}

// POLY > 0: every POLY-th pair of a thread's exponentials runs on the FMA pipe (exp2_fma2) instead of MUFU.
// CTLHINT: the control warp's long waits (P_j ready) suspend with a time hint instead of spinning: the four control
// warps of an SM all sit on scheduler 0 (warp id 4), next to the first softmax warp of every CTA.
template <bool WAVLM, int POLY = 0, bool CTLHINT = false>
__global__ void __launch_bounds__(FA_THREADS, 4)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
  pdl_wait();      // launched with programmatic stream serialization (common.cuh): q|k|v come from the kernel before
  pdl_trigger();
  extern __shared__ uint8_t fa_smem_raw[];
  uint8_t* smem = align_smem_1024(fa_smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + FA_Q_BYTES;
  uint8_t* sV = sK + FA_KV_BYTES;
  uint8_t* sP = sV + FA_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + FA_P_BYTES);
  float* s_win = reinterpret_cast<float*>(sP + FA_P_BYTES + FA_BAR_BYTES);  // WAVLM: [FA_BM + FA_BN * nkv]
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_p = bars + 4;
  uint64_t* bar_o = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int b = blockIdx.z, h = blockIdx.y;
  const int r0 = p.frame_off[b];
  const int T = p.frame_off[b + 1] - r0;
  const int i0 = blockIdx.x * FA_BM;
  if (i0 >= T) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Tk = p.key_len ? min(T, max(1, p.key_len[b])) : T;   // keys of the utterance that take part (right-padded text batches)
  const int nkv = (Tk + FA_BN - 1) / FA_BN;
  long long* tr = nullptr;
  if (p.trace && (tid == 0 || tid == 128)) {
    const int lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const int idx = lin < 32 ? lin : (lin >= 2048 && lin < 2080 ? lin - 2048 + 32 : -1);
    if (idx >= 0) tr = p.trace + (int64_t)idx * FA_TRACE_SLOTS + (tid == 128 ? 24 : 0);
  }
  fa_stamp(tr, 0);

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      mbar_init(bar_q, 1);
      mbar_init(bar_k, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      fence_mbar_init();
      // the first Q / K / V tiles are requested before the TMEM allocation and the setup barrier: their ~2 000-cycle
      // round trip is the longest item of a CTA's prologue
      mbar_arrive_expect_tx(bar_q, FA_Q_BYTES);
      tma_load_2d(sQ, &tmQ, bar_q, h * FA_HD, r0 + i0);
      mbar_arrive_expect_tx(bar_k, FA_KV_BYTES);
      tma_load_2d(sK, &tmKV, bar_k, p.d + h * FA_HD, r0);
      mbar_arrive_expect_tx(bar_v, FA_KV_BYTES);
      tma_load_2d(sV, &tmKV, bar_v, 2 * p.d + h * FA_HD, r0);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, FA_TMEM_COLS);
    tmem_relinquish();
  } else if (WAVLM) {
    // bias window of this query tile: wbuf[x] = bias_h[x - 127 - i0]; row r later reads win[key] = wbuf[key + 127 - r].
    // Filled here, four independent L2 loads in flight per thread, so that it sits under the TMEM allocation and the
    // Q / K loads (as a dependent loop after the setup barrier it cost ~2 300 cycles of every CTA); the setup barrier
    // below publishes it.
    const float* btab_h = p.btab + (int64_t)h * (2 * WAVLM_MAXD - 1) + (WAVLM_MAXD - 1);
    const int nwin = FA_BM - 1 + FA_BN * nkv;
    for (int x0 = tid; x0 < nwin; x0 += 4 * 128) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int dlt = x0 + u * 128 - (FA_BM - 1) - i0;
        dlt = max(-(WAVLM_MAXD - 1), min(WAVLM_MAXD - 1, dlt));   // buckets saturate at |delta| >= 778
        v[u] = __ldg(btab_h + dlt);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (x0 + u * 128 < nwin) s_win[x0 + u * 128] = v[u];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  fa_stamp(tr, 1);

  if (warp == 4) {
    // ------------------------------ control: TMA + MMA issue ------------------------------
    // The whole warp runs this loop with warp-uniform operands and ONE elected lane issues (see elect_one_sync in
    // common.cuh: from inside `if (lane == 0)` every UTCHMMA / UTMALDG costs an elect + R2UR waterfall, ~100 cycles,
    // which with N = 64 MMAs was most of this thread's time).
    {
      const uint32_t tmem_u = warp_uniform(tmem_base);
      const int r0u = (int)warp_uniform((uint32_t)r0);
      const int nkv = (int)warp_uniform((uint32_t)((Tk + FA_BN - 1) / FA_BN));   // shadows the CTA-wide value: uniform for the compiler
      const int colq = h * FA_HD, colk = p.d + h * FA_HD, colv = 2 * p.d + h * FA_HD;
      constexpr uint32_t idesc_s = umma_idesc_bf16(FA_BM, FA_BN);                 // Q K^T: both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(FA_BM, FA_HD) | (1u << 16);    // P V: B (= V) MN-major
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      const uint64_t kdesc = umma_desc_sw128(smem_u32(sK));
      const uint64_t pdesc = umma_desc_sw128(smem_u32(sP));
      const uint64_t vdesc = umma_desc_sw128_mn(smem_u32(sV));
      mbar_wait(bar_q, 0);
      fa_stamp(tr, 2);
      mbar_wait(bar_k, 0);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < FA_HD / 16; ++k)
          umma_bf16_ss(tmem_u + FA_TMEM_S, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, (uint32_t)(k != 0));
        umma_commit(bar_s);
      }
      __syncwarp();
      for (int j = 0; j < nkv; ++j) {
        const uint32_t ph = (uint32_t)(j & 1);
        const bool more = j + 1 < nkv;
        mbar_wait(bar_s, ph);  // S_j complete => K tile free
        if (j < 4) fa_stamp(tr, 4 + 4 * j);
        if (more && elect_one_sync()) {
          mbar_arrive_expect_tx(bar_k, FA_KV_BYTES);
          tma_load_2d(sK, &tmKV, bar_k, colk, r0u + (j + 1) * FA_BN);
        }
        __syncwarp();
        if (CTLHINT) mbar_wait_relaxed<500>(bar_p, ph); else
        mbar_wait(bar_p, ph);  // P_j in shared memory, every softmax thread is done reading S_j, O rescaled
        if (j < 4) fa_stamp(tr, 5 + 4 * j);
        tc_fence_after();
        if (more) {
          // S_{j+1} goes to the tensor core AHEAD of O += P_j V_j: the softmax of block j+1 (its row-max pass only
          // needs S) starts as soon as possible, and P_j V_j completes underneath it. (Releasing the S columns even
          // earlier - once S_j sits in registers - did not pay: the single K tile is then the late one.)
          mbar_wait(bar_k, ph ^ 1u);
          tc_fence_after();
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < FA_HD / 16; ++k)
              umma_bf16_ss(tmem_u + FA_TMEM_S, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, (uint32_t)(k != 0));
            umma_commit(bar_s);
          }
          __syncwarp();
        }
        mbar_wait(bar_v, ph);
        if (j < 4) fa_stamp(tr, 6 + 4 * j);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < FA_BN / 16; ++k) {
            // A = P: +32 B per 16 keys inside the swizzle row; B = V (MN-major): 16 keys = 16 rows of 128 B
            umma_bf16_ss(tmem_u + FA_TMEM_O, pdesc + (uint64_t)(2 * k), vdesc + (uint64_t)(k * (16 * 128 >> 4)), idesc_o, (uint32_t)((j | k) != 0));
          }
          umma_commit(bar_o);
        }
        __syncwarp();
        if (more) {
          mbar_wait(bar_o, ph);  // O += P_j V_j complete => V tile free
          if (j < 4) fa_stamp(tr, 7 + 4 * j);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(bar_v, FA_KV_BYTES);
            tma_load_2d(sV, &tmKV, bar_v, colv, r0u + (j + 1) * FA_BN);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------ softmax: one query row per thread ------------------------------
    const int row = tid;                       // 0..127 == TMEM lane
    const int qi = i0 + row;                   // query index inside the utterance
    const bool row_valid = qi < T;
    const bool warp_valid = (i0 + warp * 32) < T;  // warp-uniform
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    const float sc2 = p.scale * LOG2E;

    float gate = 0.f;
    const float* win = nullptr;
    if (WAVLM) {
      // gate[row, head] comes precomputed (LayerNorm epilogue or wavlm_gate_kernel); the bias window was filled before
      // the setup barrier
      if (row_valid) gate = __ldg(p.gate + (int64_t)(r0 + qi) * p.heads + h);
      win = s_win + (FA_BM - 1 - row);
      gate *= LOG2E;
    }
    fa_stamp(tr, 2);
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const uint32_t ph = (uint32_t)(j & 1);
      const int j0 = j * FA_BN;
      const int ncols = min(FA_BN, Tk - j0);
      mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_s, ph);
      if (j < 4) fa_stamp(tr, 4 + 4 * j);
      tc_fence_after();

      // pass 1: row maximum of  x = s * scale * log2e (+ gate * bias). Without bias the maximum is taken on the
      // raw scores (scale > 0). Ragged last block: columns >= ncols are excluded; full blocks carry no compares.
      const bool full = (ncols == FA_BN);   // CTA-uniform
      float mx = -INFINITY;
      if (warp_valid) {
#pragma unroll
        for (int c = 0; c < FA_BN / 32; ++c) {
          if (c * 32 < ncols) {  // warp-uniform
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_lane + FA_TMEM_S + c * 32, r);
            tmem_ld_wait();
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if (WAVLM) {   // x = s * scale * log2e + gate * bias, two keys per FMUL2 / FFMA2
              const uint64_t sc22 = pack_f32x2(sc2, sc2), gate2 = pack_f32x2(gate, gate);
              const float* wk = win + j0 + c * 32;
#pragma unroll
              for (int k = 0; k < 32; k += 2) {
                const uint64_t x2 = ffma2(gate2, pack_f32x2(wk[k], wk[k + 1]),
                                          fmul2(pack_f32x2(__uint_as_float(r[k]), __uint_as_float(r[k + 1])), sc22));
                float x0, x1;
                unpack_f32x2(x2, x0, x1);
                r[k] = __float_as_uint(x0);
                r[k + 1] = __float_as_uint(x1);
              }
            }
            if (full) {
#pragma unroll
              for (int k = 0; k < 32; ++k) m4[k & 3] = fmaxf(m4[k & 3], __uint_as_float(r[k]));
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (c * 32 + k < ncols) m4[k & 3] = fmaxf(m4[k & 3], __uint_as_float(r[k]));
            }
            if (WAVLM) tmem_st_32x32b_x32(t_lane + FA_TMEM_S + c * 32, r);   // pass 2 reads x back instead of redoing the bias
            mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
          }
        }
        if (!WAVLM) mx *= sc2;
      }
      if (j < 4) fa_stamp(tr, 5 + 4 * j);
      // lazy running max: keep the old one unless the new maximum exceeds it by more than 2^8 (p stays <= 256,
      // harmless in bf16 / fp32; sum and O are accumulated against the same m, so the result is exact)
      const bool raise = (mx > m_run + FA_RESCALE_LOG2);   // also true for the first block (m_run = -inf)
      const float m_new = raise ? mx : m_run;
      const float alpha = (j == 0) ? 0.f : fast_exp2(m_run - m_new);   // 1 for lanes that keep their max

      if (j > 0) {
        mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_o, ph ^ 1u);  // O += P_{j-1} V_{j-1} complete: O may be rescaled, P overwritten
        tc_fence_after();
        if (warp_valid && __any_sync(0xffffffffu, raise)) {
#pragma unroll
          for (int c = 0; c < FA_HD / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_lane + FA_TMEM_O + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * alpha);
            tmem_st_32x32b_x32(t_lane + FA_TMEM_O + c * 32, r);
          }
        }
      }
      if (WAVLM || j > 0) tmem_st_wait();

      if (j < 4) fa_stamp(tr, 6 + 4 * j);
      // pass 2: p = exp2(x - m), row sum, P (bf16) -> shared memory (K-major SW128: chunk = key / 8, XOR row % 8)
      float rs = 0.f;
      const bool two = FA_BN / 2 < ncols;   // CTA-uniform: the second 32-column chunk holds valid keys
      const float neg_m = -m_new;
      // FULLC (compile-time) = every key of the block is valid: no per-element masking in the hot path
      const uint64_t negm2 = pack_f32x2(neg_m, neg_m), sc22 = pack_f32x2(sc2, sc2);
      auto chunk = [&](uint32_t (&r)[32], const int c, auto fullc) {
        constexpr bool FULLC = decltype(fullc)::value;
        uint64_t acc0 = 0ull, acc1 = 0ull;   // packed fp32 partial sums (bit pattern 0 = +0.0f, +0.0f)
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          const uint64_t s2 = pack_f32x2(__uint_as_float(r[k]), __uint_as_float(r[k + 1]));
          const uint64_t x2 = WAVLM ? fadd2(s2, negm2) : ffma2(s2, sc22, negm2);
          float x0, x1;
          unpack_f32x2(x2, x0, x1);
          float e0, e1;
          constexpr int PM = POLY > 0 ? POLY : 1;
          if (POLY > 0 && ((k >> 1) % PM) == PM - 1) {
            exp2_fma2(x0, x1, e0, e1);
          } else {
            e0 = fast_exp2(x0);
            e1 = fast_exp2(x1);
          }
          if (!FULLC) {
            if (c * 32 + k >= ncols) e0 = 0.f;
            if (c * 32 + k + 1 >= ncols) e1 = 0.f;
          }
          const uint64_t e2 = pack_f32x2(e0, e1);
          if (k & 2) acc1 = fadd2(acc1, e2); else acc0 = fadd2(acc0, e2);
          pk[k >> 1] = pack_bf16x2(e0, e1);
        }
        float a0, a1, a2, a3;
        unpack_f32x2(acc0, a0, a1);
        unpack_f32x2(acc1, a2, a3);
        rs += (a0 + a1) + (a2 + a3);
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          const int ch = c * 4 + k8;
          *reinterpret_cast<uint4*>(sP + row * 128 + ((ch ^ (row & 7)) << 4)) =
              make_uint4(pk[k8 * 4 + 0], pk[k8 * 4 + 1], pk[k8 * 4 + 2], pk[k8 * 4 + 3]);
        }
      };
      if (warp_valid) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_lane + FA_TMEM_S, r);
        tmem_ld_wait();
        if (full) {
          chunk(r, 0, std::true_type{});
          tmem_ld_32x32b_x32(t_lane + FA_TMEM_S + 32, r);
          tmem_ld_wait();
          chunk(r, 1, std::true_type{});
        } else {
          chunk(r, 0, std::false_type{});
          if (two) {
            tmem_ld_32x32b_x32(t_lane + FA_TMEM_S + 32, r);
            tmem_ld_wait();
            chunk(r, 1, std::false_type{});
          } else {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
              const int ch = 4 + k8;
              *reinterpret_cast<uint4*>(sP + row * 128 + ((ch ^ (row & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      fence_proxy_async_smem();   // generic-proxy writes of P -> visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(bar_p);
      if (j < 4) fa_stamp(tr, 7 + 4 * j);
    }

    // ------------------------------ epilogue: O / l -> bf16 ------------------------------
    mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_o, (uint32_t)((nkv - 1) & 1));
    fa_stamp(tr, 20);
    tc_fence_after();
    if (warp_valid) {
      const float inv = 1.f / l_run;
      bf16* orow = p.out + (int64_t)(r0 + qi) * p.d + h * FA_HD;
#pragma unroll
      for (int c = 0; c < FA_HD / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_lane + FA_TMEM_O + c * 32, r);
        tmem_ld_wait();
        if (row_valid) {
#pragma unroll
          for (int k = 0; k < 32; k += 8) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(r[k + 0]) * inv, __uint_as_float(r[k + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(r[k + 2]) * inv, __uint_as_float(r[k + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(r[k + 4]) * inv, __uint_as_float(r[k + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(r[k + 6]) * inv, __uint_as_float(r[k + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c * 32 + k) = u;
          }
        }
      }
    }
  }

  fa_stamp(tr, 21);
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FA_TMEM_COLS);
  }
  fa_stamp(tr, 22);
}

}  // namespace serenc

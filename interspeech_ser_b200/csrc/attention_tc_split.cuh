// tcgen05 flash attention for head_dim 64, second generation (WavLM-large, wav2vec2 / HuBERT-large, Whisper, RoBERTa):
// packed variable-length, non-causal, optional WavLM gated relative-position bias, optional key-length mask.
//
// What the first-generation kernel (attention_tc.cuh: 4 CTAs/SM, one softmax thread per query row, Q / P in shared
// memory) ran into, measured in round 1 (profiles/r01_notes.md, r01c_ncu_attention.csv) and re-derived in round 2
// (profiles/r02_notes.md):
//   * an SS-form tcgen05.mma with N = 64 occupies the operand path for ~75 cycles (43 + N/2: its 4 KB A tile is re-read
//     from shared memory by every K-step), so the 8 MMAs of a 128 x 64 block cost ~600 cycles of a server that all CTAs
//     of an SM share: 2 400 cycles per round of four CTAs, above the 2 048 cycles the round's exponentials need on MUFU;
//   * each CTA's block is one serial chain S -> row max -> exp -> P -> PV of ~2 500 cycles in which a softmax warp owns 64
//     scores per thread; four such chains per SM leave both servers ~55 % busy.
// This kernel removes both:
//   * Q lives in TENSOR MEMORY (128 lanes x 32 columns of bf16 pairs, written once per CTA by the softmax threads straight
//     from global memory), P is written over the consumed S columns as bf16 pairs, so BOTH MMAs are TS-form (A from TMEM:
//     10 + N/2 = 42 cycles at N = 64, tools/mma_cost.cu) and neither Q nor P ever touches shared memory;
//   * TWO softmax threads per query row: warps w and w + 4 share TMEM lane quarter w and split the 64 keys of a block in
//     halves, so a thread holds its 32 scores in registers for the whole block (ONE tcgen05.ld per block instead of two
//     passes over TMEM), the pair exchanges its half-row maxima through shared memory behind a 64-thread named barrier,
//     and row sums stay per thread until the epilogue. A block's chain per warp drops to ~650 cycles; with two CTAs per
//     SM there are again 16 softmax warps per SM, now bound by the MUFU rate (16 ex2 / clk / SM);
//   * S is double-buffered in TMEM and K / V are two-slot rings prefetched two blocks ahead (the protocol of
//     attention_tc_wide.cuh), so MMA and TMA latencies sit under the other buffer's softmax;
//   * the last, ragged key block runs its MMAs at N = K = ceil16(valid keys) instead of 64.
// TMEM: 256 columns per CTA (S0 | S1 | O | Q), two CTAs per SM. Shared memory: K ring 16 KB + V ring 16 KB + exchange
// buffers (+ the WavLM bias window).
#pragma once
#include "attention_tc_wide.cuh"

namespace serenc {

constexpr int FS_SOFTMAX_THREADS = 256;
constexpr int FS_THREADS = FS_SOFTMAX_THREADS + 32;
constexpr int FS_KV_BYTES = FA_BN * FA_HD * 2;       // 8 KB: one ring slot
constexpr int FS_TMEM_COLS = 256;
constexpr int FS_TMEM_S0 = 0, FS_TMEM_S1 = 64, FS_TMEM_O = 128, FS_TMEM_Q = 192;
constexpr int FS_BAR_BYTES = 128;
constexpr int FS_XCHG_BYTES = (2 * 2 * FA_BM + 2 * FA_BM) * 4;   // half-row maxima [parity][half][row] + half-row sums [half][row]
constexpr int FS_SMEM_FIXED = 4 * FS_KV_BYTES + FS_BAR_BYTES + FS_XCHG_BYTES + 1024;
inline size_t fs_smem_bytes(bool wavlm, int tmax) {
  return (size_t)FS_SMEM_FIXED + (wavlm ? (size_t)fa_window_entries(tmax) * 4 : 0);
}

__device__ __forceinline__ void pair_barrier(int quarter) {   // warps `quarter` and `quarter + 4`: named barriers 1..4, 64 threads
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}

template <bool WAVLM>
__global__ void __launch_bounds__(FS_THREADS, 2)
attention_tc_split_kernel(const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
  extern __shared__ uint8_t fs_smem_raw[];
  uint8_t* smem = align_smem_1024(fs_smem_raw);
  uint8_t* sK = smem;                       // [2 slots] 64 keys x 128 B, 128B-swizzled
  uint8_t* sV = sK + 2 * FS_KV_BYTES;       // [2 slots]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * FS_KV_BYTES);
  uint64_t* bar_k = bars + 0;   // [2] K slot full
  uint64_t* bar_v = bars + 2;   // [2] V slot full
  uint64_t* bar_s = bars + 4;   // [2] S buffer complete
  uint64_t* bar_p = bars + 6;   // P_j in TMEM, O rescaled (256 arrivals)
  uint64_t* bar_o = bars + 7;   // O += P_j V_j complete
  uint64_t* bar_q = bars + 8;   // Q in TMEM (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* s_mx = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + FS_BAR_BYTES);   // [2 parities][2 halves][128 rows]
  float* s_l = s_mx + 2 * 2 * FA_BM;                                                          // [2 halves][128 rows]
  float* s_win = s_l + 2 * FA_BM;                                                             // WAVLM: [FA_BM + FA_BN * nkv]

  const int b = blockIdx.z, h = blockIdx.y;
  const int r0 = p.frame_off[b];
  const int T = p.frame_off[b + 1] - r0;
  const int i0 = blockIdx.x * FA_BM;
  if (i0 >= T) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Tk = p.key_len ? min(T, max(1, p.key_len[b])) : T;   // keys that take part (see AttnParams::key_len)
  const int nkv = (Tk + FA_BN - 1) / FA_BN;
  long long* tr = nullptr;   // debug clock stamps (tools/trace_attn.py): [0, 24) softmax thread 0, [24, 48) control thread
  if (p.trace && (tid == 0 || tid == FS_SOFTMAX_THREADS)) {
    const int lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const int idx = lin < 32 ? lin : (lin >= 2048 && lin < 2080 ? lin - 2048 + 32 : -1);
    if (idx >= 0) tr = p.trace + (int64_t)idx * FA_TRACE_SLOTS + (tid == FS_SOFTMAX_THREADS ? 24 : 0);
  }
  fa_stamp(tr, 0);

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmKV);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_k + i, 1);
        mbar_init(bar_v + i, 1);
        mbar_init(bar_s + i, 1);
      }
      mbar_init(bar_p, FS_SOFTMAX_THREADS);
      mbar_init(bar_o, 1);
      mbar_init(bar_q, FS_SOFTMAX_THREADS);
      fence_mbar_init();
      // the first K / V tiles are requested before the TMEM allocation and the setup barrier
      const int colk = p.d + h * FA_HD, colv = 2 * p.d + h * FA_HD;
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        if (s2 < nkv) {
          mbar_arrive_expect_tx(bar_k + s2, FS_KV_BYTES);
          tma_load_2d(sK + s2 * FS_KV_BYTES, &tmKV, bar_k + s2, colk, r0 + s2 * FA_BN);
        }
      }
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        if (s2 < nkv) {
          mbar_arrive_expect_tx(bar_v + s2, FS_KV_BYTES);
          tma_load_2d(sV + s2 * FS_KV_BYTES, &tmKV, bar_v + s2, colv, r0 + s2 * FA_BN);
        }
      }
    }
    __syncwarp();
    tmem_alloc(tmem_slot, FS_TMEM_COLS);
    tmem_relinquish();
  } else if (WAVLM) {
    // bias window of this query tile: wbuf[x] = bias_h[x - 127 - i0]; row r later reads win[key] = wbuf[key + 127 - r]
    const float* btab_h = p.btab + (int64_t)h * (2 * WAVLM_MAXD - 1) + (WAVLM_MAXD - 1);
    const int nwin = FA_BM - 1 + FA_BN * nkv;
    for (int x0 = tid; x0 < nwin; x0 += 4 * FS_SOFTMAX_THREADS) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int dlt = x0 + u * FS_SOFTMAX_THREADS - (FA_BM - 1) - i0;
        dlt = max(-(WAVLM_MAXD - 1), min(WAVLM_MAXD - 1, dlt));   // buckets saturate at |delta| >= 778
        v[u] = __ldg(btab_h + dlt);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (x0 + u * FS_SOFTMAX_THREADS < nwin) s_win[x0 + u * FS_SOFTMAX_THREADS] = v[u];
    }
  }
  // this thread's half of its query row, straight from global memory (64 B), requested before the setup barrier
  const int quarter = warp & 3, half = (warp >> 2) & 1;
  const int row = quarter * 32 + lane;       // query row of the tile == TMEM lane
  const int qi = i0 + row;
  const bool row_valid = (warp < 8) && qi < T;
  uint4 qv[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) qv[c] = make_uint4(0u, 0u, 0u, 0u);
  if (row_valid) {
    const uint4* qsrc = reinterpret_cast<const uint4*>(p.qkv + (int64_t)(r0 + qi) * p.ld_qkv + h * FA_HD + half * 32);
#pragma unroll
    for (int c = 0; c < 4; ++c) qv[c] = __ldg(qsrc + c);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  fa_stamp(tr, 1);

  if (warp == 8) {
    // ------------------------------ control: TMA + MMA issue (whole warp, one elected lane issues) ------------------------------
    const uint32_t tmem_u = warp_uniform(tmem_base);
    const int r0u = (int)warp_uniform((uint32_t)r0);
    const int nk = (int)warp_uniform((uint32_t)nkv);
    const int n_last = (int)warp_uniform((uint32_t)(((Tk - (nkv - 1) * FA_BN) + 15) & ~15));   // keys of the last block, rounded up to the MMA granule
    const int colk = p.d + h * FA_HD, colv = 2 * p.d + h * FA_HD;
    constexpr uint32_t idesc_s_full = umma_idesc_bf16(FA_BM, FA_BN);                 // Q K^T: A (TMEM) and B both K-major
    const uint32_t idesc_s_last = umma_idesc_bf16(FA_BM, n_last);
    constexpr uint32_t idesc_o = umma_idesc_bf16(FA_BM, FA_HD) | (1u << 16);         // P V: A from TMEM, B (= V) MN-major
    const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK));
    const uint64_t vdesc0 = umma_desc_sw128_mn(smem_u32(sV));
    auto issue_s = [&](int i) {   // S_i = Q K_i^T into buffer i & 1; K-step k reads Q columns [8k, 8k + 8) and K bytes [32k, 32k + 32) of every row
      const uint64_t kdesc = kdesc0 + (uint64_t)((i & 1) * (FS_KV_BYTES >> 4));
      const uint32_t idesc = (i == nk - 1) ? idesc_s_last : idesc_s_full;
#pragma unroll
      for (int k = 0; k < FA_HD / 16; ++k)
        umma_bf16_ts(tmem_u + ((i & 1) ? FS_TMEM_S1 : FS_TMEM_S0), tmem_u + FS_TMEM_Q + 8 * k, kdesc + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
      umma_commit(bar_s + (i & 1));
    };
    mbar_wait(bar_q, 0);      // every softmax thread has written its half row of Q to TMEM
    fa_stamp(tr, 2);
    mbar_wait(bar_k, 0);
    tc_fence_after();
    if (elect_one_sync()) issue_s(0);
    __syncwarp();
    if (nk > 1) {
      mbar_wait(bar_k + 1, 0);
      tc_fence_after();
      if (elect_one_sync()) issue_s(1);
      __syncwarp();
    }
    mbar_wait(bar_s, 0);   // S_0 complete => K slot 0 free
    if (nk > 2 && elect_one_sync()) {
      mbar_arrive_expect_tx(bar_k, FS_KV_BYTES);
      tma_load_2d(sK, &tmKV, bar_k, colk, r0u + 2 * FA_BN);
    }
    __syncwarp();
    for (int j = 0; j < nk; ++j) {
      const int sl = j & 1;
      const uint32_t ph2 = (uint32_t)((j >> 1) & 1);
      if (j + 1 < nk) {
        mbar_wait(bar_s + (sl ^ 1), (uint32_t)(((j + 1) >> 1) & 1));   // S_{j+1} complete => its K slot is free
        if (j + 3 < nk && elect_one_sync()) {
          mbar_arrive_expect_tx(bar_k + (sl ^ 1), FS_KV_BYTES);
          tma_load_2d(sK + (sl ^ 1) * FS_KV_BYTES, &tmKV, bar_k + (sl ^ 1), colk, r0u + (j + 3) * FA_BN);
        }
        __syncwarp();
      }
      if (j < 4) fa_stamp(tr, 4 + 4 * j);
      mbar_wait(bar_p, (uint32_t)(j & 1));   // P_j in TMEM (over S_j), O rescaled
      if (j < 4) fa_stamp(tr, 5 + 4 * j);
      mbar_wait(bar_v + sl, ph2);
      if (j < 4) fa_stamp(tr, 6 + 4 * j);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t p_tmem = tmem_u + (sl ? FS_TMEM_S1 : FS_TMEM_S0);
        const uint64_t vdesc = vdesc0 + (uint64_t)(sl * (FS_KV_BYTES >> 4));
        const int ksteps = (j == nk - 1) ? (n_last >> 4) : (FA_BN / 16);
        for (int k = 0; k < ksteps; ++k) {
          // A = P: keys [16k, 16k + 16) sit in 8 columns at +8k for the first half row, at 32 + 8(k - 2) for the second
          // (each softmax thread writes over its OWN consumed S columns); B = V (MN-major): 16 keys = 16 rows of 128 B
          const uint32_t pcol = (uint32_t)(k < 2 ? 8 * k : 32 + 8 * (k - 2));
          umma_bf16_ts(tmem_u + FS_TMEM_O, p_tmem + pcol, vdesc + (uint64_t)(k * (16 * 128 >> 4)), idesc_o, (uint32_t)((j | k) != 0));
        }
        umma_commit(bar_o);
      }
      __syncwarp();
      if (j + 2 < nk) {
        mbar_wait(bar_o, (uint32_t)(j & 1));   // O += P_j V_j complete => V slot free, P_j (= S buffer j & 1) consumed
        if (j < 4) fa_stamp(tr, 7 + 4 * j);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar_v + sl, FS_KV_BYTES);
          tma_load_2d(sV + sl * FS_KV_BYTES, &tmKV, bar_v + sl, colv, r0u + (j + 2) * FA_BN);
        }
        __syncwarp();
        mbar_wait(bar_k + sl, ph2 ^ 1u);       // K_{j+2} (requested one block ago)
        tc_fence_after();
        if (elect_one_sync()) issue_s(j + 2);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ softmax: two threads per query row ------------------------------
    const bool warp_valid = (i0 + quarter * 32) < T;   // warp-uniform, identical for both warps of a pair
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int my_lo = half * 32;                       // this thread's keys of a block: [my_lo, my_lo + 32)
    constexpr float LOG2E = 1.4426950408889634f;
    const float sc2 = p.scale * LOG2E;

    // Q: 16 columns of bf16 pairs per thread (dims [32 half, 32 half + 32) of row `row`), zero rows past the utterance
    {
      uint32_t qr[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        qr[4 * c + 0] = qv[c].x; qr[4 * c + 1] = qv[c].y; qr[4 * c + 2] = qv[c].z; qr[4 * c + 3] = qv[c].w;
      }
      tmem_st_32x32b_x16(t_lane + FS_TMEM_Q + half * 16, qr);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_q);
    }

    float gate = 0.f;
    const float* win = nullptr;
    if (WAVLM) {
      if (row_valid) gate = __ldg(p.gate + (int64_t)(r0 + qi) * p.heads + h) * LOG2E;
      win = s_win + (FA_BM - 1 - row) + my_lo;
    }
    fa_stamp(tr, 2);
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int j0 = j * FA_BN;
      const int ncols = min(FA_BN, Tk - j0);
      const int nmine = max(0, min(32, ncols - my_lo));   // valid keys in this thread's half (warp-uniform)
      const uint32_t t_s = t_lane + ((j & 1) ? FS_TMEM_S1 : FS_TMEM_S0);
      mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_s + (j & 1), (uint32_t)((j >> 1) & 1));
      if (j < 4) fa_stamp(tr, 4 + 4 * j);
      tc_fence_after();

      // scores of this half row -> registers (kept for the whole block), x = s * scale * log2e (+ gate * bias)
      uint32_t r[32];
      float mx = -INFINITY;
      const bool active = warp_valid && nmine > 0;
      if (active) {
        tmem_ld_32x32b_x32(t_s + my_lo, r);
        tmem_ld_wait();
        if (WAVLM) {
          const uint64_t sc22 = pack_f32x2(sc2, sc2), gate2 = pack_f32x2(gate, gate);
          const float* wk = win + j0;
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
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (nmine == 32) {
#pragma unroll
          for (int k = 0; k < 32; ++k) m4[k & 3] = fmaxf(m4[k & 3], __uint_as_float(r[k]));
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (k < nmine) m4[k & 3] = fmaxf(m4[k & 3], __uint_as_float(r[k]));
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        if (!WAVLM) mx *= sc2;
      }
      // the pair agrees on the row maximum of the block (double-buffered by block parity: a slot is rewritten two
      // blocks later, behind the next block's barrier)
      float* xm = s_mx + (j & 1) * (2 * FA_BM);
      xm[half * FA_BM + row] = mx;
      pair_barrier(quarter);
      mx = fmaxf(mx, xm[(half ^ 1) * FA_BM + row]);
      if (j < 4) fa_stamp(tr, 5 + 4 * j);
      // lazy running max (see attention_tc.cuh): raised only when it would grow by more than 2^8
      const bool raise = (mx > m_run + FA_RESCALE_LOG2);
      const float m_new = raise ? mx : m_run;
      const float alpha = (j == 0) ? 0.f : fast_exp2(m_run - m_new);

      if (j > 0 && warp_valid && __any_sync(0xffffffffu, raise)) {
        mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_o, (uint32_t)((j - 1) & 1));  // O += P_{j-1} V_{j-1} complete: O may be rescaled
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {   // this thread's 32 of the 64 output columns
          uint32_t o[16];
          tmem_ld_32x32b_x16(t_lane + FS_TMEM_O + my_lo + c * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
          tmem_st_32x32b_x16(t_lane + FS_TMEM_O + my_lo + c * 16, o);
        }
      }
      if (j < 4) fa_stamp(tr, 6 + 4 * j);

      // p = exp2(x - m), row sum, P as bf16 pairs over this thread's own S columns [my_lo, my_lo + 16)
      float rs = 0.f;
      if (active) {
        const float neg_m = -m_new;
        const uint64_t negm2 = pack_f32x2(neg_m, neg_m), sc22 = pack_f32x2(sc2, sc2);
        uint64_t acc0 = 0ull, acc1 = 0ull;   // packed fp32 partial sums (bit pattern 0 = +0.0f, +0.0f)
        uint32_t pk[16];
        auto expo = [&](auto fullc) {
          constexpr bool FULLC = decltype(fullc)::value;
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const uint64_t s2 = pack_f32x2(__uint_as_float(r[k]), __uint_as_float(r[k + 1]));
            const uint64_t x2 = WAVLM ? fadd2(s2, negm2) : ffma2(s2, sc22, negm2);
            float x0, x1;
            unpack_f32x2(x2, x0, x1);
            float e0 = fast_exp2(x0), e1 = fast_exp2(x1);
            if (!FULLC) {
              if (k >= nmine) e0 = 0.f;
              if (k + 1 >= nmine) e1 = 0.f;
            }
            const uint64_t e2 = pack_f32x2(e0, e1);
            if (k & 2) acc1 = fadd2(acc1, e2); else acc0 = fadd2(acc0, e2);
            pk[k >> 1] = pack_bf16x2(e0, e1);
          }
        };
        if (nmine == 32) expo(std::true_type{}); else expo(std::false_type{});
        float a0, a1, a2, a3;
        unpack_f32x2(acc0, a0, a1);
        unpack_f32x2(acc1, a2, a3);
        rs = (a0 + a1) + (a2 + a3);
        tmem_st_32x32b_x16(t_s + my_lo, pk);
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      // every block waits for the previous P V (normally long complete) so that bar_o is never more than one phase ahead
      if (j > 0) mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_o, (uint32_t)((j - 1) & 1));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p);
      if (j < 4) fa_stamp(tr, 7 + 4 * j);
    }

    // ------------------------------ epilogue: O / l -> bf16 ------------------------------
    s_l[half * FA_BM + row] = l_run;
    mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_o, (uint32_t)((nkv - 1) & 1));
    fa_stamp(tr, 20);
    tc_fence_after();
    pair_barrier(quarter);
    if (warp_valid) {
      const float inv = 1.f / (l_run + s_l[(half ^ 1) * FA_BM + row]);
      bf16* orow = p.out + (int64_t)(r0 + qi) * p.d + h * FA_HD + my_lo;
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_lane + FS_TMEM_O + my_lo, o);
      tmem_ld_wait();
      if (row_valid) {
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[k + 0]) * inv, __uint_as_float(o[k + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[k + 2]) * inv, __uint_as_float(o[k + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[k + 4]) * inv, __uint_as_float(o[k + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[k + 6]) * inv, __uint_as_float(o[k + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + k) = u;
        }
      }
    }
  }

  fa_stamp(tr, 21);
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FS_TMEM_COLS);
  }
  fa_stamp(tr, 22);
}

}  // namespace serenc

// tcgen05 flash attention, third generation: ONE kernel for head_dim 64 (WavLM-large with its gated relative-position
// bias, wav2vec2 / HuBERT-large, Whisper, RoBERTa), 80 (HuBERT-xlarge) and 120 (wav2vec2-xls-r-2b); packed variable-length,
// non-causal, optional key-length mask (HF modeling_wavlm.py:147-271, modeling_wav2vec2.py:438-549,
// modeling_whisper.py:215-357, modeling_roberta.py:190-254).
//
// Skeleton of attention_tc_wide.cuh (one CTA = 128 query rows of one (utterance, head); one softmax thread per query row
// = TMEM lane; 64 keys per block; one control warp issuing TMA and MMAs; K / V two-slot rings prefetched two blocks
// ahead; S double-buffered in TMEM; P written as bf16 pairs over the consumed S columns and read by a TS-form MMA; two
// CTAs per SM). What round 2 measured and changed (profiles/r02_notes.md):
//   * the exp pass is bound per WARP (one MUFU.EX2 per 8 cycles and warp), so what decides the kernel is how much of a
//     block's chain S -> max -> exp -> P is NOT exp: with two chains per scheduler the duty has to be > 50 %. The old
//     chain made two passes over TMEM with four serialised tcgen05.ld round trips (1 750 cycles per block, 29 % duty).
//     Now a thread loads its 64 scores ONCE, both tcgen05.ld in flight together, keeps them in registers for the maximum
//     and for the exponentials (~900 cycles, 57 % duty);
//   * Q lives in TENSOR MEMORY for head_dim 64 / 80 (written once per CTA by the softmax threads straight from global
//     memory), so S = Q K^T is a TS-form MMA too (10 + N/2 instead of 43 + N/2 cycles, tools/mma_cost.cu) and Q needs
//     neither a TMA round trip nor 16-32 KB of shared memory (head_dim 120 does not fit: 60 more TMEM columns);
//   * WavLM's gated relative-position bias (bias window of the query tile in shared memory, gate per (row, head) from the
//     LayerNorm) is a template flag here, so WavLM gets the deep pipeline as well;
//   * the last, ragged key block runs its MMAs at N = K = ceil16(valid keys) instead of 64;
//   * a tried-and-dropped variant (attention_tc_split.cuh, A/B arm): two softmax threads per row. Halving a thread's
//     scores doubles the per-block fixed cost (waits, exchange, fences) per score and the two warps of a row share a
//     scheduler: 678 us against 561 us per Whisper launch.
// TMEM: 256 columns per CTA: S0 [0, 64) | S1 [64, 128) | O [128, 128 + ON) | Q [128 + ON, ...).
#pragma once
#include "attention_tc_wide.cuh"

namespace serenc {

template <int HD>
struct Fa3Cfg {
  static constexpr int NCH = (HD + 63) / 64;           // 64-column chunks of a head in shared memory
  static constexpr int KSTEPS = (HD + 15) / 16;        // K-steps of S = Q K^T
  static constexpr int ON = KSTEPS * 16;               // N of the P V MMA (64 / 80 / 128)
  static constexpr bool Q_TMEM = HD <= 80;             // Q as the TMEM A operand of S = Q K^T
  static constexpr int QCOLS = HD / 2;                 // TMEM columns of Q (bf16 pairs): 32 / 40
  static constexpr int Q_CHUNK = FA_BM * 128;          // one 64-column chunk of Q in shared memory: 16 KB (head_dim 120 only)
  static constexpr int KV_CHUNK = FA_BN * 128;         // one 64-column chunk of K / V: 8 KB
  static constexpr int Q_BYTES = Q_TMEM ? 0 : NCH * Q_CHUNK;
  static constexpr int KV_BYTES = NCH * KV_CHUNK;      // one ring slot
  static constexpr int BAR_BYTES = 128;
  static constexpr int SMEM_FIXED = Q_BYTES + 4 * KV_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 256;
  static constexpr int TMEM_S0 = 0, TMEM_S1 = 64, TMEM_O = 128, TMEM_Q = 128 + ON;
  static_assert(!Q_TMEM || TMEM_Q + QCOLS <= TMEM_COLS, "Q does not fit into the CTA's tensor memory");
};
template <int HD>
inline size_t fa3_smem_bytes(bool wavlm, int tmax) {
  return (size_t)Fa3Cfg<HD>::SMEM_FIXED + (wavlm ? (size_t)fa_window_entries(tmax) * 4 : 0);
}

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// 2^x for x <= 0 on the FMA pipe, two values at once (x = n + f, n = round(x); 2^f by a cubic, n added to the exponent
// field). Max relative error ~1e-4, below the bf16 rounding of P. POLY > 0 moves every POLY-th pair of a thread's
// exponentials here (A/B arm: on B200 the FMA-pipe form costs ~6 issue slots per pair against 2 for MUFU).
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& y0, float& y1) {
  const float MAGIC = 12582912.f;   // 1.5 * 2^23
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t x = pack_f32x2(x0, x1);
  const uint64_t r = fadd2(x, pack_f32x2(MAGIC, MAGIC));
  float r0, r1;
  unpack_f32x2(r, r0, r1);
  const uint64_t f = fadd2(x, fadd2(pack_f32x2(MAGIC, MAGIC), pack_f32x2(-r0, -r1)));
  uint64_t p = pack_f32x2(0.0555041f, 0.0555041f);
  p = ffma2(p, f, pack_f32x2(0.2402265f, 0.2402265f));
  p = ffma2(p, f, pack_f32x2(0.6931472f, 0.6931472f));
  p = ffma2(p, f, pack_f32x2(1.0f, 1.0f));
  float p0, p1;
  unpack_f32x2(p, p0, p1);
  y0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(r0) << 23));
  y1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(r1) << 23));
}

template <int HD, bool WAVLM, int POLY = 0>
__global__ void __launch_bounds__(FA_THREADS, 2)
attention_tc_v3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
  using C = Fa3Cfg<HD>;
  static_assert(!WAVLM || HD == FA_HD, "the gated relative position bias exists for head_dim 64 only");
  extern __shared__ uint8_t fa_smem_raw[];
  uint8_t* smem = align_smem_1024(fa_smem_raw);
  uint8_t* sQ = smem;                       // head_dim 120 only
  uint8_t* sK = sQ + C::Q_BYTES;            // [2 slots]
  uint8_t* sV = sK + 2 * C::KV_BYTES;       // [2 slots]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * C::KV_BYTES);
  uint64_t* bar_q = bars + 0;   // Q in shared memory (TMA) or in TMEM (128 arrivals)
  uint64_t* bar_k = bars + 1;   // [2]: K slot full
  uint64_t* bar_v = bars + 3;   // [2]: V slot full
  uint64_t* bar_s = bars + 5;   // [2]: S buffer complete
  uint64_t* bar_p = bars + 7;   // P_j in TMEM (128 arrivals)
  uint64_t* bar_o = bars + 8;   // O += P_j V_j complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* s_win = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + C::BAR_BYTES);   // WAVLM: [FA_BM + FA_BN * nkv]

  const int b = blockIdx.z, h = blockIdx.y;
  const int r0 = p.frame_off[b];
  const int T = p.frame_off[b + 1] - r0;
  const int i0 = blockIdx.x * FA_BM;
  if (i0 >= T) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Tk = p.key_len ? min(T, max(1, p.key_len[b])) : T;   // keys that take part (see AttnParams::key_len)
  const int nkv = (Tk + FA_BN - 1) / FA_BN;
  long long* tr = nullptr;   // debug clock stamps, same slots as attention_tc_kernel (tools/trace_attn.py)
  if (p.trace && (tid == 0 || tid == 128)) {
    const int lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const int idx = lin < 32 ? lin : (lin >= 2048 && lin < 2080 ? lin - 2048 + 32 : -1);
    if (idx >= 0) tr = p.trace + (int64_t)idx * FA_TRACE_SLOTS + (tid == 128 ? 24 : 0);
  }
  fa_stamp(tr, 0);

  const int qi = i0 + tid;                   // tid 0..127 == TMEM lane == query row of the tile
  const bool row_valid = tid < FA_BM && qi < T;
  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      mbar_init(bar_q, C::Q_TMEM ? FA_BM : 1);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_k + i, 1);
        mbar_init(bar_v + i, 1);
        mbar_init(bar_s + i, 1);
      }
      mbar_init(bar_p, FA_BM);
      mbar_init(bar_o, 1);
      fence_mbar_init();
      // first tiles requested before the TMEM allocation and the setup barrier (the longest item of the prologue)
      const int slot_q = h, slot_k = p.heads + h, slot_v = 2 * p.heads + h;
      if (!C::Q_TMEM) {
        mbar_arrive_expect_tx(bar_q, C::Q_BYTES);
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) tma_load_3d(sQ + c * C::Q_CHUNK, &tmQ, bar_q, c * 64, slot_q, r0 + i0);
      }
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        if (s2 < nkv) {
          mbar_arrive_expect_tx(bar_k + s2, C::KV_BYTES);
#pragma unroll
          for (int c = 0; c < C::NCH; ++c) tma_load_3d(sK + s2 * C::KV_BYTES + c * C::KV_CHUNK, &tmKV, bar_k + s2, c * 64, slot_k, r0 + s2 * FA_BN);
        }
      }
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        if (s2 < nkv) {
          mbar_arrive_expect_tx(bar_v + s2, C::KV_BYTES);
#pragma unroll
          for (int c = 0; c < C::NCH; ++c) tma_load_3d(sV + s2 * C::KV_BYTES + c * C::KV_CHUNK, &tmKV, bar_v + s2, c * 64, slot_v, r0 + s2 * FA_BN);
        }
      }
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  } else if (WAVLM) {
    // bias window of this query tile: wbuf[x] = bias_h[x - 127 - i0]; row r later reads win[key] = wbuf[key + 127 - r].
    // Filled here, four independent L2 loads in flight per thread, under the TMEM allocation and the first K / V loads.
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
  // this thread's query row straight from global memory (HD bf16 = 128 / 160 B), requested before the setup barrier
  uint4 qv[C::Q_TMEM ? HD / 8 : 1];
  if constexpr (C::Q_TMEM) {
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) qv[c] = make_uint4(0u, 0u, 0u, 0u);
    if (row_valid) {
      const uint4* qsrc = reinterpret_cast<const uint4*>(p.qkv + (int64_t)(r0 + qi) * p.ld_qkv + h * HD);
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) qv[c] = __ldg(qsrc + c);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  fa_stamp(tr, 1);

  if (warp == 4) {
    // ------------------------------ control: TMA + MMA issue (whole warp, one elected lane issues) ------------------------------
    const uint32_t tmem_u = warp_uniform(tmem_base);
    const int r0u = (int)warp_uniform((uint32_t)r0);
    const int nk = (int)warp_uniform((uint32_t)nkv);
    const int n_last = (int)warp_uniform((uint32_t)(((Tk - (nkv - 1) * FA_BN) + 15) & ~15));   // keys of the last block, rounded up to the MMA granule
    const int slot_k = p.heads + h, slot_v = 2 * p.heads + h;   // head slots of the rank-3 map
    constexpr uint32_t idesc_s_full = umma_idesc_bf16(FA_BM, FA_BN);           // Q K^T: both K-major
    const uint32_t idesc_s_last = umma_idesc_bf16(FA_BM, n_last);
    constexpr uint32_t idesc_o = umma_idesc_bf16(FA_BM, C::ON) | (1u << 16);   // P V: A from TMEM, B (= V) MN-major
    const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
    const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK));
    const uint64_t vdesc0 = umma_desc_sw128_mn_wide(smem_u32(sV), C::KV_CHUNK);
    auto load_kv = [&](uint8_t* dst, uint64_t* bar, int slot, int key0) {   // one ring slot: NCH chunks of 64 keys
      mbar_arrive_expect_tx(bar, C::KV_BYTES);
#pragma unroll
      for (int c = 0; c < C::NCH; ++c) tma_load_3d(dst + c * C::KV_CHUNK, &tmKV, bar, c * 64, slot, r0u + key0);
    };
    auto issue_s = [&](int i) {   // S_i = Q K_i^T into buffer i & 1: K-step k reads chunk k / 4 at +32 B * (k % 4)
      const uint64_t kdesc = kdesc0 + (uint64_t)((i & 1) * (C::KV_BYTES >> 4));
      const uint32_t idesc = (i == nk - 1) ? idesc_s_last : idesc_s_full;
      const uint32_t d_tmem = tmem_u + ((i & 1) ? C::TMEM_S1 : C::TMEM_S0);
#pragma unroll
      for (int k = 0; k < C::KSTEPS; ++k) {
        const uint64_t ko = (uint64_t)((k >> 2) * (C::KV_CHUNK >> 4) + 2 * (k & 3));
        if (C::Q_TMEM) {
          umma_bf16_ts(d_tmem, tmem_u + C::TMEM_Q + 8 * k, kdesc + ko, idesc, (uint32_t)(k != 0));
        } else {
          const uint64_t qo = (uint64_t)((k >> 2) * (C::Q_CHUNK >> 4) + 2 * (k & 3));
          umma_bf16_ss(d_tmem, qdesc + qo, kdesc + ko, idesc, (uint32_t)(k != 0));
        }
      }
      umma_commit(bar_s + (i & 1));
    };
    mbar_wait(bar_q, 0);
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
    if (nk > 2 && elect_one_sync()) load_kv(sK, bar_k, slot_k, 2 * FA_BN);
    __syncwarp();
    for (int j = 0; j < nk; ++j) {
      const int sl = j & 1;
      const uint32_t ph2 = (uint32_t)((j >> 1) & 1);
      if (j + 1 < nk) {
        mbar_wait(bar_s + (sl ^ 1), (uint32_t)(((j + 1) >> 1) & 1));   // S_{j+1} complete => its K slot is free
        if (j + 3 < nk && elect_one_sync()) load_kv(sK + (sl ^ 1) * C::KV_BYTES, bar_k + (sl ^ 1), slot_k, (j + 3) * FA_BN);
        __syncwarp();
      }
      if (j < 4) fa_stamp(tr, 4 + 4 * j);
      mbar_wait(bar_p, (uint32_t)(j & 1));   // P_j in TMEM (over S_j), O rescaled
      if (j < 4) fa_stamp(tr, 5 + 4 * j);
      mbar_wait(bar_v + sl, ph2);
      if (j < 4) fa_stamp(tr, 6 + 4 * j);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t p_tmem = tmem_u + (sl ? C::TMEM_S1 : C::TMEM_S0);
        const uint64_t vdesc = vdesc0 + (uint64_t)(sl * (C::KV_BYTES >> 4));
        const int ksteps = (j == nk - 1) ? (n_last >> 4) : (FA_BN / 16);
        for (int k = 0; k < ksteps; ++k) {
          // A = P: 8 TMEM columns (bf16 pairs) per 16 keys; B = V (MN-major): 16 keys = 16 rows of 128 B in every chunk
          umma_bf16_ts(tmem_u + C::TMEM_O, p_tmem + 8 * k, vdesc + (uint64_t)(k * (16 * 128 >> 4)), idesc_o, (uint32_t)((j | k) != 0));
        }
        umma_commit(bar_o);
      }
      __syncwarp();
      if (j + 2 < nk) {
        mbar_wait(bar_o, (uint32_t)(j & 1));   // O += P_j V_j complete => V slot free, P_j (= S buffer j & 1) consumed
        if (j < 4) fa_stamp(tr, 7 + 4 * j);
        if (elect_one_sync()) load_kv(sV + sl * C::KV_BYTES, bar_v + sl, slot_v, (j + 2) * FA_BN);
        __syncwarp();
        // S_{j+2} into the buffer P_j sat in (MMAs on different accumulators are not documented as ordered, hence
        // only after the completion above); K_{j+2} was requested one block ago
        mbar_wait(bar_k + sl, ph2 ^ 1u);
        tc_fence_after();
        if (elect_one_sync()) issue_s(j + 2);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ softmax: one query row per thread ------------------------------
    const bool warp_valid = (i0 + warp * 32) < T;  // warp-uniform
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    const float sc2 = p.scale * LOG2E;

    if constexpr (C::Q_TMEM) {   // Q row -> TMEM as bf16 pairs (column c = dims 2c, 2c + 1); rows past the utterance are zeros
      uint32_t qr[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        qr[4 * c + 0] = qv[c].x; qr[4 * c + 1] = qv[c].y; qr[4 * c + 2] = qv[c].z; qr[4 * c + 3] = qv[c].w;
      }
      tmem_st_32x32b_x32(t_lane + C::TMEM_Q, qr);
      if constexpr (C::QCOLS > 32) {
        uint32_t q8[8];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint4 u = qv[(HD / 8 > 8) ? 8 + c : 0];
          q8[4 * c + 0] = u.x; q8[4 * c + 1] = u.y; q8[4 * c + 2] = u.z; q8[4 * c + 3] = u.w;
        }
        tmem_st_32x32b_x8(t_lane + C::TMEM_Q + 32, q8);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_q);
    }

    float gate = 0.f;
    const float* win = nullptr;
    if (WAVLM) {
      // gate[row, head] comes precomputed (LayerNorm epilogue or wavlm_gate_kernel); the bias window was filled before
      // the setup barrier
      if (row_valid) gate = __ldg(p.gate + (int64_t)(r0 + qi) * p.heads + h) * LOG2E;
      win = s_win + (FA_BM - 1 - tid);
    }
    fa_stamp(tr, 2);
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int j0 = j * FA_BN;
      const int ncols = min(FA_BN, Tk - j0);
      const bool full = (ncols == FA_BN);   // CTA-uniform
      const bool two = FA_BN / 2 < ncols;   // CTA-uniform: the second 32-column chunk holds valid keys
      const uint32_t t_s = t_lane + ((j & 1) ? C::TMEM_S1 : C::TMEM_S0);
      mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_s + (j & 1), (uint32_t)((j >> 1) & 1));
      if (j < 4) fa_stamp(tr, 4 + 4 * j);
      tc_fence_after();

      // the row's 64 scores -> registers, both loads in flight together; they stay there for the maximum AND the
      // exponentials (the previous generation re-read S from TMEM for its second pass: four serialised round trips)
      uint32_t ra[32], rb[32];
      float mx = -INFINITY;
      if (warp_valid) {
        tmem_ld_32x32b_x32(t_s, ra);
        if (two) tmem_ld_32x32b_x32(t_s + 32, rb);
        tmem_ld_wait();
        auto prep = [&](uint32_t (&r)[32], const int c) {   // x = s * scale * log2e + gate * bias (WavLM) ; running maximum
          if (WAVLM) {
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
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          if (full) {
#pragma unroll
            for (int k = 0; k < 32; ++k) m4[k & 3] = fmaxf(m4[k & 3], __uint_as_float(r[k]));
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (c * 32 + k < ncols) m4[k & 3] = fmaxf(m4[k & 3], __uint_as_float(r[k]));
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
        };
        prep(ra, 0);
        if (two) prep(rb, 1);
        if (!WAVLM) mx *= sc2;   // without bias the maximum is taken on the raw scores (scale > 0)
      }
      if (j < 4) fa_stamp(tr, 5 + 4 * j);
      // lazy running max (see attention_tc.cuh): raised only when it would grow by more than 2^8
      const bool raise = (mx > m_run + FA_RESCALE_LOG2);
      const float m_new = raise ? mx : m_run;
      const float alpha = (j == 0) ? 0.f : fast_exp2(m_run - m_new);

      if (j > 0 && warp_valid && __any_sync(0xffffffffu, raise)) {
        mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_o, (uint32_t)((j - 1) & 1));  // O += P_{j-1} V_{j-1} complete: O may be rescaled
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < C::ON / 16; ++c) {   // 16 columns at a time: never past the O region (Q sits right behind it)
          uint32_t o[16];
          tmem_ld_32x32b_x16(t_lane + C::TMEM_O + c * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
          tmem_st_32x32b_x16(t_lane + C::TMEM_O + c * 16, o);
        }
      }
      if (j < 4) fa_stamp(tr, 6 + 4 * j);

      // p = exp2(x - m), row sum, P as bf16 pairs over columns [0, 32) of this S buffer
      float rs = 0.f;
      const float neg_m = -m_new;
      const uint64_t negm2 = pack_f32x2(neg_m, neg_m), sc22 = pack_f32x2(sc2, sc2);
      auto chunk = [&](const uint32_t (&r)[32], const int c, auto fullc) {
        constexpr bool FULLC = decltype(fullc)::value;   // every key of the block is valid: no per-element masking
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
            exp2_poly2(x0, x1, e0, e1);
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
        tmem_st_32x32b_x16(t_s + c * 16, pk);   // chunk 1 lands on S columns [16, 32): both chunks sit in registers already
      };
      if (warp_valid) {
        if (full) {
          chunk(ra, 0, std::true_type{});
          chunk(rb, 1, std::true_type{});
        } else {
          chunk(ra, 0, std::false_type{});
          if (two) chunk(rb, 1, std::false_type{});   // otherwise the P V MMA stops after ceil(ncols / 16) K-steps
        }
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
    mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_o, (uint32_t)((nkv - 1) & 1));
    fa_stamp(tr, 20);
    tc_fence_after();
    if (warp_valid) {
      const float inv = 1.f / l_run;
      bf16* orow = p.out + (int64_t)(r0 + qi) * p.d + h * HD;
      constexpr int OCH = (HD + 31) / 32;   // 32-column chunks, all loads in flight together (reads past O touch Q / spare columns only)
      uint32_t o[OCH][32];
#pragma unroll
      for (int c = 0; c < OCH; ++c) tmem_ld_32x32b_x32(t_lane + C::TMEM_O + c * 32, o[c]);
      tmem_ld_wait();
      if (row_valid) {
#pragma unroll
        for (int c = 0; c < OCH; ++c) {
#pragma unroll
          for (int k = 0; k < 32; k += 8) {
            if (c * 32 + k < HD) {   // compile-time: HD is a multiple of 8
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(o[c][k + 0]) * inv, __uint_as_float(o[c][k + 1]) * inv);
              u.y = pack_bf16x2(__uint_as_float(o[c][k + 2]) * inv, __uint_as_float(o[c][k + 3]) * inv);
              u.z = pack_bf16x2(__uint_as_float(o[c][k + 4]) * inv, __uint_as_float(o[c][k + 5]) * inv);
              u.w = pack_bf16x2(__uint_as_float(o[c][k + 6]) * inv, __uint_as_float(o[c][k + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + c * 32 + k) = u;
            }
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
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
  fa_stamp(tr, 22);
}

}  // namespace serenc

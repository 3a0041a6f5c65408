// tcgen05 flash attention, deep-pipelined variant: head_dim 80 (HuBERT-xlarge), 120 (wav2vec2-xls-r-2b) and 64
// without bias (Whisper, wav2vec2/HuBERT-large): packed variable-length, non-causal (HF modeling_wav2vec2.py:438-549,
// modular_hubert.py, modeling_whisper.py:215-357 = plain SDPA with key padding).
//
// Same skeleton as attention_tc.cuh (one CTA = 128 query rows of one (utterance, head), softmax one query row per
// thread straight out of TMEM, 64 keys per block, one control warp issuing TMA and MMAs). What differs:
//
//  * Wide heads: Q, K and V live in shared memory as 64-column chunks ([rows x 128 B], 128B-swizzled), loaded through a
//    RANK-3 tensor map {head_dim, 3 * heads, rows} over the packed [sum_T, 3d] projection buffer: the box of the second
//    chunk runs past the end of the head and TMA zero-fills it, so neither the neighbouring head's columns nor a
//    zeroing pass ever reach shared memory. S = Q K^T runs ceil(head_dim / 16) K-steps (5 / 8); O += P V is ONE MMA per
//    16 keys with N = 80 / 128, V consumed MN-major across both chunks (leading-dimension byte offset = chunk pitch).
//  * The per-CTA clock trace of the first version (single K / V tile, P through shared memory) showed every block
//    paced by a TMA round trip (~1 700 cycles from HBM) + the S MMA in series: K_{j+1} could only be fetched once S_j
//    was done. Now K and V are two-slot rings (K_{j+3} is requested when S_{j+1} completes, V_{j+2} when P_j V_j does)
//    and S is double-buffered in TMEM (S_{j+2} is issued right behind P_j V_j), so the chain of a block is the softmax.
//  * P never goes through shared memory: the softmax threads write it as packed bf16 pairs over the first 32 columns
//    of the S buffer they just consumed (tcgen05.st) and O += P V reads its A operand from TMEM (tcgen05.mma TS form:
//    10 + N/2 instead of 43 + N/2 cycles per MMA, tools/mma_cost.cu; no STS, no proxy fence, 16 KB less shared memory).
//  * TMEM: 256 columns per CTA (S0 | S1 | O up to 128), two CTAs per SM.
#pragma once
#include "attention_tc.cuh"

namespace serenc {

template <int HD>
struct FawCfg {
  static constexpr int NCH = (HD + 63) / 64;           // 64-column chunks of a head in shared memory
  static constexpr int KSTEPS = (HD + 15) / 16;        // K-steps of S = Q K^T
  static constexpr int ON = KSTEPS * 16;               // N of the P V MMA (64 / 80 / 128)
  static constexpr int OCH = (ON + 31) / 32;           // 32-column TMEM chunks of O the softmax threads touch
  static constexpr int Q_CHUNK = FA_BM * 128;          // one 64-column chunk of Q: 16 KB
  static constexpr int KV_CHUNK = FA_BN * 128;         // one 64-column chunk of K / V: 8 KB
  static constexpr int Q_BYTES = NCH * Q_CHUNK;
  static constexpr int KV_BYTES = NCH * KV_CHUNK;      // one ring slot
  static constexpr int SMEM_BYTES = Q_BYTES + 4 * KV_BYTES + 128 + 1024;
  static constexpr int TMEM_COLS = 256;
  static constexpr int TMEM_S0 = 0, TMEM_S1 = 64, TMEM_O = 128;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// MN-major 128B-swizzled B operand spanning several 64-column chunks: 8-key groups 1024 B apart, chunks `chunk_pitch` apart
__device__ __forceinline__ uint64_t umma_desc_sw128_mn_wide(uint32_t smem_addr, uint32_t chunk_pitch) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((chunk_pitch >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// D[tmem] (+)= A[tmem: 128 lanes x 8 columns of bf16 pairs] * B[smem descriptor]   (K = 16)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <int HD, bool CTLHINT = false>
__global__ void __launch_bounds__(FA_THREADS, 2)
attention_tc_wide_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
  pdl_wait();      // launched with programmatic stream serialization (common.cuh): q|k|v come from the kernel before
  pdl_trigger();
  using C = FawCfg<HD>;
  extern __shared__ uint8_t fa_smem_raw[];
  uint8_t* smem = align_smem_1024(fa_smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::Q_BYTES;            // [2 slots]
  uint8_t* sV = sK + 2 * C::KV_BYTES;       // [2 slots]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * C::KV_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;   // [2]: K slot full
  uint64_t* bar_v = bars + 3;   // [2]: V slot full
  uint64_t* bar_s = bars + 5;   // [2]: S buffer complete
  uint64_t* bar_p = bars + 7;   // P_j in TMEM (128 arrivals)
  uint64_t* bar_o = bars + 8;   // O += P_j V_j complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

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

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      mbar_init(bar_q, 1);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_k + i, 1);
        mbar_init(bar_v + i, 1);
        mbar_init(bar_s + i, 1);
      }
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      fence_mbar_init();
      // first tiles requested before the TMEM allocation and the setup barrier (the longest item of the prologue)
      const int slot_q = h, slot_k = p.heads + h, slot_v = 2 * p.heads + h;
      mbar_arrive_expect_tx(bar_q, C::Q_BYTES);
#pragma unroll
      for (int c = 0; c < C::NCH; ++c) tma_load_3d(sQ + c * C::Q_CHUNK, &tmQ, bar_q, c * 64, slot_q, r0 + i0);
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
    const int slot_q = h, slot_k = p.heads + h, slot_v = 2 * p.heads + h;   // head slots of the rank-3 map
    constexpr uint32_t idesc_s = umma_idesc_bf16(FA_BM, FA_BN);                // Q K^T: both K-major
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
#pragma unroll
      for (int k = 0; k < C::KSTEPS; ++k) {
        const uint64_t qo = (uint64_t)((k >> 2) * (C::Q_CHUNK >> 4) + 2 * (k & 3));
        const uint64_t ko = (uint64_t)((k >> 2) * (C::KV_CHUNK >> 4) + 2 * (k & 3));
        umma_bf16_ss(tmem_u + ((i & 1) ? C::TMEM_S1 : C::TMEM_S0), qdesc + qo, kdesc + ko, idesc_s, (uint32_t)(k != 0));
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
      if (CTLHINT) mbar_wait_relaxed<500>(bar_p, (uint32_t)(j & 1)); else
      mbar_wait(bar_p, (uint32_t)(j & 1));   // P_j in TMEM (over S_j), O rescaled
      if (j < 4) fa_stamp(tr, 5 + 4 * j);
      mbar_wait(bar_v + sl, ph2);
      if (j < 4) fa_stamp(tr, 6 + 4 * j);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t p_tmem = tmem_u + (sl ? C::TMEM_S1 : C::TMEM_S0);
        const uint64_t vdesc = vdesc0 + (uint64_t)(sl * (C::KV_BYTES >> 4));
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k) {
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
    const int qi = i0 + tid;                   // tid 0..127 == TMEM lane == query row of the tile
    const bool row_valid = qi < T;
    const bool warp_valid = (i0 + warp * 32) < T;  // warp-uniform
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    const float sc2 = p.scale * LOG2E;

    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int j0 = j * FA_BN;
      const int ncols = min(FA_BN, Tk - j0);
      const uint32_t t_s = t_lane + ((j & 1) ? C::TMEM_S1 : C::TMEM_S0);
      mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_s + (j & 1), (uint32_t)((j >> 1) & 1));
      if (j < 4) fa_stamp(tr, 4 + 4 * j);
      tc_fence_after();

      // pass 1: row maximum of the raw scores (scale > 0); columns >= ncols of a ragged last block are excluded
      const bool full = (ncols == FA_BN);   // CTA-uniform
      float mx = -INFINITY;
      if (warp_valid) {
#pragma unroll
        for (int c = 0; c < FA_BN / 32; ++c) {
          if (c * 32 < ncols) {  // warp-uniform
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_s + c * 32, r);
            tmem_ld_wait();
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
          }
        }
        mx *= sc2;
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
        for (int c = 0; c < C::OCH; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_lane + C::TMEM_O + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * alpha);
          tmem_st_32x32b_x32(t_lane + C::TMEM_O + c * 32, r);
        }
      }
      if (j < 4) fa_stamp(tr, 6 + 4 * j);

      // pass 2: p = exp2(s * scale * log2e - m), row sum, P as bf16 pairs over columns [0, 32) of this S buffer
      float rs = 0.f;
      const bool two = FA_BN / 2 < ncols;   // CTA-uniform
      const float neg_m = -m_new;
      // FULLC (compile-time) = every key of the block is valid: no per-element masking in the hot path
      const uint64_t negm2 = pack_f32x2(neg_m, neg_m), sc22 = pack_f32x2(sc2, sc2);
      auto chunk = [&](const int c, auto fullc) {
        constexpr bool FULLC = decltype(fullc)::value;
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_s + c * 32, r);
        tmem_ld_wait();
        uint64_t acc0 = 0ull, acc1 = 0ull;   // packed fp32 partial sums (bit pattern 0 = +0.0f, +0.0f)
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          const uint64_t x2 = ffma2(pack_f32x2(__uint_as_float(r[k]), __uint_as_float(r[k + 1])), sc22, negm2);
          float x0, x1;
          unpack_f32x2(x2, x0, x1);
          float e0 = fast_exp2(x0), e1 = fast_exp2(x1);
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
        tmem_st_32x32b_x16(t_s + c * 16, pk);   // chunk 1 lands on S columns [16, 32): consumed by chunk 0 already
      };
      if (warp_valid) {
        if (full) {
          chunk(0, std::true_type{});
          chunk(1, std::true_type{});
        } else {
          chunk(0, std::false_type{});
          if (two) {
            chunk(1, std::false_type{});
          } else {
            uint32_t z[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) z[k] = 0u;
            tmem_st_32x32b_x16(t_s + 16, z);
          }
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
#pragma unroll
      for (int c = 0; c < C::OCH; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_lane + C::TMEM_O + c * 32, r);
        tmem_ld_wait();
        if (row_valid) {
#pragma unroll
          for (int k = 0; k < 32; k += 8) {
            if (c * 32 + k < HD) {   // compile-time: HD is a multiple of 8
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

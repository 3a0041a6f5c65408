// tcgen05 flash attention for head_dim 80 (HuBERT-xlarge) and 120 (wav2vec2-xls-r-2b): packed variable-length,
// non-causal, no bias (HF modeling_wav2vec2.py:438-549, modular_hubert.py attention = plain SDPA with key padding).
//
// Same structure as attention_tc.cuh (one CTA = 128 query rows of one (utterance, head), softmax one query row per
// thread straight out of TMEM, P through shared memory, 64 keys per block); what changes with the head width:
//
//  * A head is wider than one 128-byte swizzle row, so Q, K and V live in shared memory as TWO 64-column chunks
//    ([rows x 128 B], 128B-swizzled). They are loaded through a RANK-3 tensor map {head_dim, 3 * heads, rows} over the
//    packed [sum_T, 3d] projection buffer: the box of the second chunk runs past the end of the head and TMA zero-fills
//    it, so neither the neighbouring head's columns nor a separate zeroing pass ever reach shared memory.
//  * S = Q K^T runs ceil(head_dim / 16) K-steps (5 / 8: the last step of head_dim 120 is half zeros on both sides);
//    O += P V is ONE MMA per 16 keys with N = 80 / 128, V consumed MN-major across both chunks (leading-dimension
//    byte offset = chunk pitch).
//  * TMEM: 256 columns per CTA (O needs up to 128), two CTAs per SM. The 128 columns left of O hold S DOUBLE-BUFFERED:
//    S_{j+1} = Q K_{j+1}^T is issued as soon as K_{j+1} has landed, without waiting for the softmax of block j.
#pragma once
#include "attention_tc.cuh"

namespace serenc {

template <int HD>
struct FawCfg {
  static constexpr int KSTEPS = (HD + 15) / 16;        // K-steps of S = Q K^T
  static constexpr int ON = KSTEPS * 16;               // N of the P V MMA (80 / 128)
  static constexpr int OCH = (ON + 31) / 32;           // 32-column TMEM chunks of O the softmax threads touch
  static constexpr int Q_CHUNK = FA_BM * 128;          // one 64-column chunk of Q: 16 KB
  static constexpr int KV_CHUNK = FA_BN * 128;         // one 64-column chunk of K / V: 8 KB
  static constexpr int Q_BYTES = 2 * Q_CHUNK;
  static constexpr int KV_BYTES = 2 * KV_CHUNK;
  static constexpr int SMEM_BYTES = Q_BYTES + 2 * KV_BYTES + FA_P_BYTES + FA_BAR_BYTES + 1024;
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

template <int HD>
__global__ void __launch_bounds__(FA_THREADS, 2)
attention_tc_wide_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
  using C = FawCfg<HD>;
  extern __shared__ uint8_t fa_smem_raw[];
  uint8_t* smem = align_smem_1024(fa_smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::Q_BYTES;
  uint8_t* sV = sK + C::KV_BYTES;
  uint8_t* sP = sV + C::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + FA_P_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;   // [2]: one per S buffer
  uint64_t* bar_p = bars + 5;
  uint64_t* bar_o = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int b = blockIdx.z, h = blockIdx.y;
  const int r0 = p.frame_off[b];
  const int T = p.frame_off[b + 1] - r0;
  const int i0 = blockIdx.x * FA_BM;
  if (i0 >= T) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkv = (T + FA_BN - 1) / FA_BN;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      mbar_init(bar_q, 1);
      mbar_init(bar_k, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s + 0, 1);
      mbar_init(bar_s + 1, 1);
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------ control: TMA + MMA issue (whole warp, one elected lane issues) ------------------------------
    const uint32_t tmem_u = warp_uniform(tmem_base);
    const int r0u = (int)warp_uniform((uint32_t)r0);
    const int nkvu = (int)warp_uniform((uint32_t)nkv);
    const int slot_q = h, slot_k = p.heads + h, slot_v = 2 * p.heads + h;   // head slots of the rank-3 map
    constexpr uint32_t idesc_s = umma_idesc_bf16(FA_BM, FA_BN);                // Q K^T: both K-major
    constexpr uint32_t idesc_o = umma_idesc_bf16(FA_BM, C::ON) | (1u << 16);   // P V: B (= V) MN-major
    const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
    const uint64_t kdesc = umma_desc_sw128(smem_u32(sK));
    const uint64_t pdesc = umma_desc_sw128(smem_u32(sP));
    const uint64_t vdesc = umma_desc_sw128_mn_wide(smem_u32(sV), C::KV_CHUNK);
    auto issue_s = [&](int buf) {   // S[buf] = Q K^T: K-step k reads chunk k / 4 at +32 B * (k % 4)
#pragma unroll
      for (int k = 0; k < C::KSTEPS; ++k) {
        const uint64_t qo = (uint64_t)((k >> 2) * (C::Q_CHUNK >> 4) + 2 * (k & 3));
        const uint64_t ko = (uint64_t)((k >> 2) * (C::KV_CHUNK >> 4) + 2 * (k & 3));
        umma_bf16_ss(tmem_u + (buf ? C::TMEM_S1 : C::TMEM_S0), qdesc + qo, kdesc + ko, idesc_s, (uint32_t)(k != 0));
      }
      umma_commit(bar_s + buf);
    };
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(bar_q, C::Q_BYTES);
      tma_load_3d(sQ, &tmQ, bar_q, 0, slot_q, r0u + i0);
      tma_load_3d(sQ + C::Q_CHUNK, &tmQ, bar_q, 64, slot_q, r0u + i0);
      mbar_arrive_expect_tx(bar_k, C::KV_BYTES);
      tma_load_3d(sK, &tmKV, bar_k, 0, slot_k, r0u);
      tma_load_3d(sK + C::KV_CHUNK, &tmKV, bar_k, 64, slot_k, r0u);
      mbar_arrive_expect_tx(bar_v, C::KV_BYTES);
      tma_load_3d(sV, &tmKV, bar_v, 0, slot_v, r0u);
      tma_load_3d(sV + C::KV_CHUNK, &tmKV, bar_v, 64, slot_v, r0u);
    }
    __syncwarp();
    mbar_wait(bar_q, 0);
    mbar_wait(bar_k, 0);
    tc_fence_after();
    if (elect_one_sync()) issue_s(0);
    __syncwarp();
    for (int j = 0; j < nkvu; ++j) {
      const uint32_t ph = (uint32_t)(j & 1);
      const bool more = j + 1 < nkvu;
      mbar_wait(bar_s + (j & 1), (uint32_t)((j >> 1) & 1));   // S_j complete => K tile free
      if (more) {
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar_k, C::KV_BYTES);
          tma_load_3d(sK, &tmKV, bar_k, 0, slot_k, r0u + (j + 1) * FA_BN);
          tma_load_3d(sK + C::KV_CHUNK, &tmKV, bar_k, 64, slot_k, r0u + (j + 1) * FA_BN);
        }
        __syncwarp();
        // S_{j+1} into the other buffer: its previous tenant S_{j-1} was released by bar_p(j-1), waited on in iteration j-1
        mbar_wait(bar_k, ph ^ 1u);
        tc_fence_after();
        if (elect_one_sync()) issue_s((j + 1) & 1);
        __syncwarp();
      }
      mbar_wait(bar_p, ph);   // P_j in shared memory, O rescaled
      mbar_wait(bar_v, ph);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k) {
          // A = P: +32 B per 16 keys inside the swizzle row; B = V (MN-major): 16 keys = 16 rows of 128 B in every chunk
          umma_bf16_ss(tmem_u + C::TMEM_O, pdesc + (uint64_t)(2 * k), vdesc + (uint64_t)(k * (16 * 128 >> 4)), idesc_o, (uint32_t)((j | k) != 0));
        }
        umma_commit(bar_o);
      }
      __syncwarp();
      if (more) {
        mbar_wait(bar_o, ph);   // O += P_j V_j complete => V tile free
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar_v, C::KV_BYTES);
          tma_load_3d(sV, &tmKV, bar_v, 0, slot_v, r0u + (j + 1) * FA_BN);
          tma_load_3d(sV + C::KV_CHUNK, &tmKV, bar_v, 64, slot_v, r0u + (j + 1) * FA_BN);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ softmax: one query row per thread ------------------------------
    const int row = tid;                       // 0..127 == TMEM lane
    const int qi = i0 + row;
    const bool row_valid = qi < T;
    const bool warp_valid = (i0 + warp * 32) < T;  // warp-uniform
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    const float sc2 = p.scale * LOG2E;

    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const uint32_t ph = (uint32_t)(j & 1);
      const int j0 = j * FA_BN;
      const int ncols = min(FA_BN, T - j0);
      const uint32_t t_s = t_lane + ((j & 1) ? C::TMEM_S1 : C::TMEM_S0);
      mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_s + (j & 1), (uint32_t)((j >> 1) & 1));
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
      // lazy running max (see attention_tc.cuh): raised only when it would grow by more than 2^8
      const bool raise = (mx > m_run + FA_RESCALE_LOG2);
      const float m_new = raise ? mx : m_run;
      const float alpha = (j == 0) ? 0.f : fast_exp2(m_run - m_new);

      if (j > 0) {
        mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_o, ph ^ 1u);  // O += P_{j-1} V_{j-1} complete: O may be rescaled, P overwritten
        tc_fence_after();
        if (warp_valid && __any_sync(0xffffffffu, raise)) {
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
        tmem_st_wait();
      }

      // pass 2: p = exp2(s * scale * log2e - m), row sum, P (bf16) -> shared memory (K-major SW128)
      float rs = 0.f;
      const bool two = FA_BN / 2 < ncols;   // CTA-uniform
      const float neg_m = -m_new;
      auto chunk = [&](uint32_t (&r)[32], const int c) {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          float e = fast_exp2(fmaf(__uint_as_float(r[k]), sc2, neg_m));
          if (!full && c * 32 + k >= ncols) e = 0.f;
          s4[k & 3] += e;
          r[k] = __float_as_uint(e);
        }
        rs += (s4[0] + s4[1]) + (s4[2] + s4[3]);
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r[k8 * 8 + 0]), __uint_as_float(r[k8 * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(r[k8 * 8 + 2]), __uint_as_float(r[k8 * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(r[k8 * 8 + 4]), __uint_as_float(r[k8 * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(r[k8 * 8 + 6]), __uint_as_float(r[k8 * 8 + 7]));
          const int ch = c * 4 + k8;
          *reinterpret_cast<uint4*>(sP + row * 128 + ((ch ^ (row & 7)) << 4)) = u;
        }
      };
      if (warp_valid) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_s, r);
        tmem_ld_wait();
        chunk(r, 0);
        if (two) {
          tmem_ld_32x32b_x32(t_s + 32, r);
          tmem_ld_wait();
          chunk(r, 1);
        } else {
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {
            const int ch = 4 + k8;
            *reinterpret_cast<uint4*>(sP + row * 128 + ((ch ^ (row & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      fence_proxy_async_smem();   // generic-proxy writes of P -> visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(bar_p);
    }

    // ------------------------------ epilogue: O / l -> bf16 ------------------------------
    mbar_wait_relaxed<FA_WAIT_HINT_NS>(bar_o, (uint32_t)((nkv - 1) & 1));
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

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace serenc

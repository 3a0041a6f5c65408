// conv0 of the wav2vec2 / HuBERT / WavLM feature encoder on the tensor cores:
//   Conv1d(1 -> 512, k = 10, s = 5) (+ bias) -> LayerNorm(512) -> GELU, channels-last bf16 output
// (WavLMLayerNormConvLayer layer 0, HF modeling_wavlm.py:703-727; feat_extract_norm = "layer" only - the GroupNorm
// variant of the base-size checkpoints keeps conv0_kernel<1|2> in frontend_norm.cuh).
//
// Round 1's conv0_kernel kept the 160 filter taps of a lane's 16 channels in registers and was ALU-bound: 22 issued
// instructions per output element (10 tap FMAs as 5 FFMA2, LayerNorm, GELU), 12 % occupancy at 224 registers,
// 1 145 GB/s on a kernel whose work is 3.3 MB of output per audio-second. Here the 10-tap dot products are ONE
// tcgen05.mma chain per 128 frames:
//   * C_in = 1, so a frame's "im2col" row is its 10 input samples. bf16 operands alone would lose the fp32 waveform's
//     precision, so sample and weight are split x = xh + xl, w = wh + wl (bf16 each) and the K dimension holds the three
//     significant cross terms: K = 32 = [xh (10) | xl (10) | xh (10) | 0 0] against [wh | wh | wl | 0 0]; the dropped
//     xl * wl term is 2^-16 relative, far below the bf16 rounding of the output;
//   * A (128 frames x 32) is written by the producer warp straight into the 128B-swizzled K-major layout the MMA reads
//     (double-buffered), B (512 channels x 32) is built once per CTA; D = 128 x 512 fp32 fills the CTA's 512 TMEM
//     columns (two N = 256 MMAs per K-step);
//   * the epilogue has the WHOLE 512-channel row of a frame in one TMEM lane: LayerNorm statistics need no shuffles. Eight
//     warps, two per lane quarter (columns [0, 256) and [256, 512)), exchange their half-row (pivot, sum, sum of squares)
//     through shared memory, then normalise + affine + GELU + bf16-pack their half and store it through the same
//     XOR-swizzled staging tile as the GEMM's bf16 epilogue (16-byte stores, 64 contiguous bytes per row and block).
// Persistent: one CTA per SM walks (utterance, 128-frame tile) items, so the 64 KB weight tile is built once per SM.
#pragma once
#include "frontend_norm.cuh"

namespace serenc {

constexpr int C0T_BM = 128;                       // frames per tile (= TMEM lanes)
constexpr int C0T_THREADS = 288;                  // 8 epilogue warps + 1 producer / MMA warp
constexpr int C0T_A_BYTES = C0T_BM * 128;         // 16 KB: 128 rows of one 128-byte swizzle row (64 B used)
constexpr int C0T_B_BYTES = CONV0_C * 128;        // 64 KB
constexpr int C0T_XS = C0T_BM * CONV0_S + CONV0_K;   // samples a tile reads (+ tail)
constexpr int C0T_STAGING = 8 * 32 * 64;          // per epilogue warp: 32 rows x 64 B, XOR-swizzled
constexpr int C0T_SMEM = 2 * C0T_A_BYTES + C0T_B_BYTES + C0T_STAGING + 3 * CONV0_C * 4 + 2 * C0T_XS * 4 + 2 * 3 * C0T_BM * 4 + 128 + 1024;

__device__ __forceinline__ void c0t_pair_barrier(int quarter) {   // the two epilogue warps of a lane quarter
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}
// 32 K-values of one A / B row (8 bf16 per 16-byte chunk) -> shared memory, K-major 128B swizzle (chunk ^ (row & 7))
__device__ __forceinline__ void c0t_store_row(uint8_t* tile, int row, const uint32_t (&v)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(tile + row * 128 + ((c ^ (row & 7)) << 4)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
// [hi (10) | lo (10) | hi (10) | 0 0] (samples) or [hi | hi | lo | 0 0] (weights, REPEAT_HI) of ten fp32 values, as bf16 pairs
template <bool REPEAT_HI>
__device__ __forceinline__ void c0t_split_row(const float (&x)[CONV0_K], uint32_t (&v)[16]) {
  float hi[CONV0_K], lo[CONV0_K];
#pragma unroll
  for (int j = 0; j < CONV0_K; ++j) {
    hi[j] = __bfloat162float(__float2bfloat16(x[j]));
    lo[j] = x[j] - hi[j];
  }
  float k[32];
#pragma unroll
  for (int j = 0; j < CONV0_K; ++j) {
    k[j] = hi[j];
    k[10 + j] = REPEAT_HI ? hi[j] : lo[j];
    k[20 + j] = REPEAT_HI ? lo[j] : hi[j];
  }
  k[30] = k[31] = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = pack_bf16x2(k[2 * i], k[2 * i + 1]);
}

__global__ void __launch_bounds__(C0T_THREADS, 1)
conv0_tc_kernel(const void* __restrict__ wav, int wav_i16, const Conv0Utt* __restrict__ utts,
                const int32_t* __restrict__ tile_off /*[B+1]: prefix sum of ceil(slot / 128)*/, int batch,
                const float2* __restrict__ stats,   // nullptr: input already normalised
                const float* __restrict__ w,        // [512][10]
                const float* __restrict__ bias,     // [512] or nullptr
                const float* __restrict__ gamma, const float* __restrict__ beta, bf16* __restrict__ out) {
  extern __shared__ uint8_t c0t_smem_raw[];
  uint8_t* smem = align_smem_1024(c0t_smem_raw);
  uint8_t* sA = smem;                                   // [2] A tiles
  uint8_t* sB = sA + 2 * C0T_A_BYTES;                   // weights
  uint8_t* sStage = sB + C0T_B_BYTES;                   // [8 warps][32 rows][64 B]
  float* s_par = reinterpret_cast<float*>(sStage + C0T_STAGING);   // [3][512]: bias, gamma, beta
  float* s_xs = s_par + 3 * CONV0_C;                    // [2][C0T_XS] normalised samples of a tile
  float* s_st = s_xs + 2 * C0T_XS;                      // [2 halves][3][128 rows]: pivot, sum, sum of squares about the pivot
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_st + 2 * 3 * C0T_BM);
  uint64_t* bar_full = bars + 0;    // accumulator complete (tcgen05.commit)
  uint64_t* bar_empty = bars + 1;   // accumulator drained (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total_tiles = tile_off[batch];

  // ---- once per CTA: barriers, TMEM, weights (split hi / lo), affine parameters ----
  if (warp == 8) {
    if (lane == 0) {
      mbar_init(bar_full, 1);
      mbar_init(bar_empty, 256);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int n = tid; n < CONV0_C; n += C0T_THREADS) {
    float wr[CONV0_K];
#pragma unroll
    for (int j = 0; j < CONV0_K; ++j) wr[j] = __ldg(w + n * CONV0_K + j);
    uint32_t v[16];
    c0t_split_row<true>(wr, v);
    c0t_store_row(sB, n, v);
    s_par[n] = bias ? __ldg(bias + n) : 0.f;
    s_par[CONV0_C + n] = __ldg(gamma + n);
    s_par[2 * CONV0_C + n] = __ldg(beta + n);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto find_utt = [&](int tile) {   // largest b with tile_off[b] <= tile
    int lo = 0, hi = batch - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile_off[mid] <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
  };

  if (warp == 8) {
    // ------------------------------ producer + MMA issuer ------------------------------
    const uint32_t tmem_u = warp_uniform(tmem_base);
    constexpr uint32_t idesc = umma_idesc_bf16(C0T_BM, 256);
    const uint64_t bdesc = umma_desc_sw128(smem_u32(sB));
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int b = find_utt(tile);
      const Conv0Utt u = utts[b];
      const int t0 = (tile - tile_off[b]) * C0T_BM;
      float mean = 0.f, rstd = 1.f;
      if (stats) {
        const float2 st = stats[b];
        mean = st.x;
        rstd = st.y;
      }
      float* xs = s_xs + buf * C0T_XS;
      {
        // all of a lane's loads in flight before the first use (a rolled loop would pay one HBM round trip per 32 samples)
        constexpr int NLD = (C0T_XS + 31) / 32;
        float xv[NLD];
#pragma unroll
        for (int k = 0; k < NLD; ++k) {
          const int i = lane + 32 * k;
          const int64_t s = (int64_t)t0 * CONV0_S + i;
          xv[k] = (i < C0T_XS && s < u.sample_len) ? load_sample(wav, wav_i16, u.sample_start + s) : mean;
        }
#pragma unroll
        for (int k = 0; k < NLD; ++k) {
          const int i = lane + 32 * k;
          const int64_t s = (int64_t)t0 * CONV0_S + i;
          if (i < C0T_XS) xs[i] = s < u.sample_len ? (xv[k] - mean) * rstd : 0.f;
        }
      }
      __syncwarp();
      uint8_t* a_tile = sA + buf * C0T_A_BYTES;
#pragma unroll 1
      for (int r = lane; r < C0T_BM; r += 32) {
        float xr[CONV0_K];
#pragma unroll
        for (int j = 0; j < CONV0_K; ++j) xr[j] = xs[r * CONV0_S + j];
        uint32_t v[16];
        c0t_split_row<false>(xr, v);
        c0t_store_row(a_tile, r, v);
      }
      fence_proxy_async_smem();
      __syncwarp();
      mbar_wait(bar_empty, (uint32_t)((it & 1) ^ 1));   // previous tile's accumulator drained (passes at once for the first tile)
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t adesc = umma_desc_sw128(smem_u32(a_tile));
#pragma unroll
        for (int nh = 0; nh < 2; ++nh) {
#pragma unroll
          for (int k = 0; k < 2; ++k)   // K = 32: two K-steps of 16 (+32 B inside the swizzle row)
            umma_bf16_ss(tmem_u + (uint32_t)(nh * 256), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(nh * (256 * 128 >> 4) + 2 * k), idesc,
                         (uint32_t)(k != 0));
        }
        umma_commit(bar_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------ epilogue: LayerNorm + GELU over the 512 channels of every frame ------------------------------
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;                       // frame inside the tile == TMEM lane
    const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 256);
    uint4* st4 = reinterpret_cast<uint4*>(sStage + warp * (32 * 64));   // [32 rows][4 chunks of 16 B], chunk index XOR-swizzled by row
    const float* pb = s_par + half * 256;
    const float* pg = s_par + CONV0_C + half * 256;
    const float* pe = s_par + 2 * CONV0_C + half * 256;
    const int sr = lane >> 2, sc = lane & 3;                   // coalesced phase: row 8 i + lane / 4, 16-byte segment lane % 4
    const bool has_bias = bias != nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int b = find_utt(tile);
      const Conv0Utt u = utts[b];
      const int t0 = (tile - tile_off[b]) * C0T_BM;
      mbar_wait(bar_full, (uint32_t)(it & 1));
      tc_fence_after();

      // pass 1: half-row statistics about a pivot (this half's first value + bias): stable one-pass variance.
      // Both passes keep the NEXT 32-column block's tcgen05.ld in flight while the current one is processed.
      uint32_t ra[32], rb[32];
      tmem_ld_32x32b_x32(t_addr, ra);
      tmem_ld_wait();
      tmem_ld_32x32b_x32(t_addr + 32u, rb);
      const float pivot = __uint_as_float(ra[0]) + pb[0];
      const uint64_t npv2 = pack_f32x2(-pivot, -pivot);
      uint64_t s2 = 0ull, q2 = 0ull;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t (&r)[32] = (c & 1) ? rb : ra;
        if (c > 0) {
          tmem_ld_wait();
          if (c + 1 < 8) tmem_ld_32x32b_x32(t_addr + (uint32_t)((c + 1) * 32), (c & 1) ? ra : rb);
        }
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          uint64_t d = fadd2(pack_f32x2(__uint_as_float(r[k]), __uint_as_float(r[k + 1])), npv2);
          if (has_bias) {   // CTA-uniform
            const float2 bb = *reinterpret_cast<const float2*>(pb + c * 32 + k);
            d = fadd2(d, pack_f32x2(bb.x, bb.y));
          }
          s2 = fadd2(s2, d);
          q2 = ffma2(d, d, q2);
        }
      }
      float s_lo, s_hi, q_lo, q_hi;
      unpack_f32x2(s2, s_lo, s_hi);
      unpack_f32x2(q2, q_lo, q_hi);
      const float sv = s_lo + s_hi, qv = q_lo + q_hi;
      float* xst = s_st + half * (3 * C0T_BM);
      xst[row] = pivot; xst[C0T_BM + row] = sv; xst[2 * C0T_BM + row] = qv;
      c0t_pair_barrier(quarter);
      const float* ost = s_st + (half ^ 1) * (3 * C0T_BM);
      const float p1 = ost[row], s1 = ost[C0T_BM + row], q1 = ost[2 * C0T_BM + row];
      // combine the two halves (256 values each): mean, then both sums of squares re-centred on it
      const float mu = ((sv + 256.f * pivot) + (s1 + 256.f * p1)) * (1.f / CONV0_C);
      const float d0 = mu - pivot, d1 = mu - p1;
      const float var = ((qv - 2.f * d0 * sv + 256.f * d0 * d0) + (q1 - 2.f * d1 * s1 + 256.f * d1 * d1)) * (1.f / CONV0_C);
      const float rs = rsqrtf(fmaxf(var, 0.f) + 1e-5f);
      tmem_ld_32x32b_x32(t_addr, ra);   // first block of pass 2, in flight across the barrier
      c0t_pair_barrier(quarter);   // the exchange buffer is free again for the next tile

      // pass 2: normalise + affine + GELU -> bf16 -> staging -> 16-byte stores (64 contiguous bytes per row and block)
      int64_t orow[4];
      bool ozero[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = t0 + quarter * 32 + 8 * i + sr;
        orow[i] = t < u.slot ? u.row0 + t : -1;
        ozero[i] = t >= u.T0;             // slot padding rows: zeros, as every later layer expects finite values there
      }
      const float nmurs = -mu * rs;
      const uint64_t nmurs2 = pack_f32x2(nmurs, nmurs), rs2 = pack_f32x2(rs, rs);   // (x - mu) rs = x rs + (-mu rs): one FFMA2
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t (&r)[32] = (c & 1) ? rb : ra;
        tmem_ld_wait();
        if (c + 1 < 8) tmem_ld_32x32b_x32(t_addr + (uint32_t)((c + 1) * 32), (c & 1) ? ra : rb);
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const int col = c * 32 + k8 * 8 + e;
            const float2 gg = *reinterpret_cast<const float2*>(pg + col);
            const float2 be = *reinterpret_cast<const float2*>(pe + col);
            uint64_t v = pack_f32x2(__uint_as_float(r[k8 * 8 + e]), __uint_as_float(r[k8 * 8 + e + 1]));
            if (has_bias) {   // CTA-uniform
              const float2 bb = *reinterpret_cast<const float2*>(pb + col);
              v = fadd2(v, pack_f32x2(bb.x, bb.y));
            }
            v = ffma2(ffma2(v, rs2, nmurs2), pack_f32x2(gg.x, gg.y), pack_f32x2(be.x, be.y));
            float v0, v1;
            unpack_f32x2(v, v0, v1);
            gelu_fast2(v0, v1, v0, v1);
            pk[e >> 1] = pack_bf16x2(v0, v1);
          }
          st4[lane * 4 + (k8 ^ ((lane >> 1) & 3))] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        __syncwarp();
        const int colo = half * 256 + c * 32 + sc * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rl = 8 * i + sr;
          uint4 o = st4[rl * 4 + (sc ^ ((rl >> 1) & 3))];
          if (ozero[i]) o = make_uint4(0u, 0u, 0u, 0u);
          if (orow[i] >= 0) *reinterpret_cast<uint4*>(out + orow[i] * CONV0_C + colo) = o;
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(bar_empty);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace serenc

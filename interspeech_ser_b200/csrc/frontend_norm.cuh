// HBM-bound kernels of the wav2vec2/WavLM/HuBERT path: waveform statistics, conv0 (+LayerNorm+GELU),
// row LayerNorm (+GELU), layout scatter for the positional conv, hidden-state accumulation, masked mean pooling.
#pragma once
#include "common.cuh"

namespace serenc {

// ---------------------------------------------------------------------------------------------
// Per-utterance waveform statistics: mean and 1/sqrt(var + 1e-7) over the valid samples
// (Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm, HF feature_extraction_wav2vec2.py:77-97;
//  population variance, two-pass like numpy's). One block per utterance.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) wav_stats_kernel(const void* __restrict__ wav, int wav_i16,
                                                          const UttSpan* __restrict__ utts,
                                                          float2* __restrict__ stats /*[B] (mean, rstd)*/) {
  const int b = blockIdx.x;
  const int64_t x0 = utts[b].sample_start;
  const int n = utts[b].sample_len;
  __shared__ double red[32];
  __shared__ float s_mean;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)load_sample(wav, wav_i16, x0 + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    double v = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_mean = (float)(v / (double)max(n, 1));
  }
  __syncthreads();
  const float mean = s_mean;
  acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = load_sample(wav, wav_i16, x0 + i) - mean;
    acc += (double)(d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __syncthreads();
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    double v = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
      const float var = (float)(v / (double)max(n, 1));
      stats[b] = make_float2(mean, rsqrtf(var + 1e-7f));
    }
  }
}

// Apply the normalisation (used by the FeatureExtractor surface; the fused encode path normalises inside conv0).
// Padding samples (i >= len) are written as 0, as HF does (feature_extraction_wav2vec2.py:90-93).
__global__ void wav_normalize_kernel(const float* __restrict__ wav, const UttSpan* __restrict__ utts,
                                     const float2* __restrict__ stats, float* __restrict__ out, int64_t out_stride,
                                     int out_len) {
  const int b = blockIdx.y;
  const float2 st = stats[b];
  const float* x = wav + utts[b].sample_start;
  const int n = utts[b].sample_len;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < out_len; i += gridDim.x * blockDim.x)
    out[b * out_stride + i] = i < n ? (x[i] - st.x) * st.y : 0.f;
}

// ---------------------------------------------------------------------------------------------
// conv0: Conv1d(1 -> 512, k=10, s=5) + LayerNorm(512) + exact GELU, channels-last bf16 output
// (WavLMLayerNormConvLayer layer 0, HF modeling_wavlm.py:703-727). C_in = 1, so this is a bandwidth
// kernel: 16 MFLOP but 3.3 MB of output per audio-second.
//
// Output "slot" layout: utterance b owns rows [row0[b], row0[b] + slot[b]) with row0 = 64*R6[b]; rows
// t >= T0[b] inside the slot are written as zeros so every later layer sees finite, deterministic values.
// Lane l of a warp owns channels {64*i + 2*l, 64*i + 2*l + 1 : i < 8} -> coalesced bf16x2 stores.
// ---------------------------------------------------------------------------------------------
constexpr int CONV0_C = 512;
constexpr int CONV0_K = 10;
constexpr int CONV0_S = 5;
constexpr int CONV0_TILE = 256;  // frames per block

typedef UttSpan Conv0Utt;

// MODE 0: LayerNorm over the 512 channels of every frame (feat_extract_norm = "layer").
// MODE 1: GroupNorm statistics pass (feat_extract_norm = "group", HF WavLMGroupNormConvLayer: per-channel mean /
//         variance over the utterance's valid frames): writes per-block partial (sum, sum of squares) per channel.
// MODE 2: GroupNorm apply pass: conv recomputed (10 MACs per output beat a 1 KB round trip), y = gelu(a*scale + shift)
//         with the per-(utterance, channel) scale/shift produced by conv0_gn_finalize_kernel.
template <int MODE>
__global__ void __launch_bounds__(256, 1) conv0_kernel(const void* __restrict__ wav, int wav_i16,
                                                       const Conv0Utt* __restrict__ utts,
                                                       const float2* __restrict__ stats,  // nullptr: already normalised
                                                       const float* __restrict__ w,       // [512][10]
                                                       const float* __restrict__ bias,    // [512] or nullptr
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       bf16* __restrict__ out,
                                                       double2* __restrict__ gn_partial,      // MODE 1: [B][tiles][512]
                                                       const float2* __restrict__ gn_affine,  // MODE 2: [B][512] (scale, shift)
                                                       int tiles_per_utt) {
  // The 160 filter taps of a lane's 16 channels live in REGISTERS for the whole tile, so the inner loop is
  // 160 FFMA + 10 broadcast shared-memory reads per frame (a shared-memory-resident filter made the kernel
  // LDS-bound at ~1 TB/s of output).
  __shared__ float ws[8 * CONV0_C * 2];     // filter staging (coalesced global read, per-lane gather); MODE 1 reuses it for the block reduction
  __shared__ float2 s_b[CONV0_C / 2], s_g[CONV0_C / 2], s_be[CONV0_C / 2];
  __shared__ float xs[CONV0_TILE * CONV0_S + CONV0_K];

  const Conv0Utt u = utts[blockIdx.y];
  const int t0 = blockIdx.x * CONV0_TILE;
  if (t0 >= u.slot) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int i = threadIdx.x; i < CONV0_K * CONV0_C; i += blockDim.x) ws[i] = w[i];
  for (int c = threadIdx.x; c < CONV0_C; c += blockDim.x) {
    reinterpret_cast<float*>(s_b)[c] = bias ? bias[c] : 0.f;
    if (MODE == 0) {
      reinterpret_cast<float*>(s_g)[c] = gamma[c];
      reinterpret_cast<float*>(s_be)[c] = beta[c];
    } else if (MODE == 2) {
      const float2 a = gn_affine[(int64_t)blockIdx.y * CONV0_C + c];
      reinterpret_cast<float*>(s_g)[c] = a.x;
      reinterpret_cast<float*>(s_be)[c] = a.y;
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (stats) {
    const float2 st = stats[blockIdx.y];
    mean = st.x;
    rstd = st.y;
  }
  for (int i = threadIdx.x; i < CONV0_TILE * CONV0_S + CONV0_K; i += blockDim.x) {
    const int64_t s = (int64_t)t0 * CONV0_S + i;
    xs[i] = s < u.sample_len ? (load_sample(wav, wav_i16, u.sample_start + s) - mean) * rstd : 0.f;
  }
  __syncthreads();

  // Channel pairs (64 i + 2 lane, + 1) are carried as packed fp32x2 values: the 160 FMAs of a frame issue as 80
  // FFMA2, and the LayerNorm / GELU arithmetic as FADD2 / FMUL2 / FFMA2 (per-lane IEEE results, half the issue slots).
  uint64_t wr[8][CONV0_K];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j < CONV0_K; ++j)
      wr[i][j] = pack_f32x2(ws[(64 * i + 2 * lane) * CONV0_K + j], ws[(64 * i + 2 * lane + 1) * CONV0_K + j]);
  }

  const int t_end = min(CONV0_TILE, u.slot - t0);
  float gsum[16], gsq[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) gsum[i] = gsq[i] = 0.f;
  for (int f = warp; f < t_end; f += 8) {
    const int t = t0 + f;
    uint32_t* orow = reinterpret_cast<uint32_t*>(out + (u.row0 + t) * CONV0_C);
    if (t >= u.T0) {  // slot padding rows: zeros (warp-uniform)
      if (MODE != 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) orow[32 * i + lane] = 0u;
      }
      continue;
    }
    uint64_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 b = s_b[32 * i + lane];
      a[i] = pack_f32x2(b.x, b.y);
    }
#pragma unroll
    for (int j = 0; j < CONV0_K; ++j) {
      const float xv = xs[f * CONV0_S + j];
      const uint64_t xv2 = pack_f32x2(xv, xv);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = ffma2(wr[i][j], xv2, a[i]);
    }
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a0, a1;
        unpack_f32x2(a[i], a0, a1);
        gsum[2 * i] += a0;
        gsum[2 * i + 1] += a1;
        gsq[2 * i] = fmaf(a0, a0, gsq[2 * i]);
        gsq[2 * i + 1] = fmaf(a1, a1, gsq[2 * i + 1]);
      }
      continue;
    }
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 g = s_g[32 * i + lane], be = s_be[32 * i + lane];
        float v0, v1, y0, y1;
        unpack_f32x2(ffma2(a[i], pack_f32x2(g.x, g.y), pack_f32x2(be.x, be.y)), v0, v1);
        gelu_fast2(v0, v1, y0, y1);
        orow[32 * i + lane] = pack_bf16x2(y0, y1);
      }
      continue;
    }
    // LayerNorm statistics in ONE butterfly: sum and sum of squares about a per-frame pivot (this lane-0 channel's
    // value, broadcast) travel through the five shuffle rounds together, so a frame pays one reduction latency
    // instead of two dependent ones (8 warps per SM cannot hide them). Shifting by the pivot keeps the one-pass
    // variance free of the |mean| >> std cancellation.
    float a00, a01;
    unpack_f32x2(a[0], a00, a01);
    const float pivot = __shfl_sync(0xffffffffu, a00, 0);
    const uint64_t npv2 = pack_f32x2(-pivot, -pivot);
    uint64_t s2 = 0ull, q2 = 0ull;   // (+0.0f, +0.0f)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a[i] = fadd2(a[i], npv2);          // values about the pivot, reused below
      s2 = fadd2(s2, a[i]);
      q2 = ffma2(a[i], a[i], q2);
    }
    float s_lo, s_hi, q_lo, q_hi;
    unpack_f32x2(s2, s_lo, s_hi);
    unpack_f32x2(q2, q_lo, q_hi);
    float sv = s_lo + s_hi, qv = q_lo + q_hi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sv += __shfl_xor_sync(0xffffffffu, sv, o);
      qv += __shfl_xor_sync(0xffffffffu, qv, o);
    }
    const float mu = sv * (1.f / CONV0_C);                       // mean about the pivot
    const float rs = rsqrtf(fmaxf(qv * (1.f / CONV0_C) - mu * mu, 0.f) + 1e-5f);
    const uint64_t nmu2 = pack_f32x2(-mu, -mu);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fadd2(a[i], nmu2);        // centred values
    const uint64_t rs2 = pack_f32x2(rs, rs);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 g = s_g[32 * i + lane], be = s_be[32 * i + lane];
      float v0, v1, y0, y1;
      unpack_f32x2(ffma2(fmul2(a[i], rs2), pack_f32x2(g.x, g.y), pack_f32x2(be.x, be.y)), v0, v1);
      gelu_fast2(v0, v1, y0, y1);
      orow[32 * i + lane] = pack_bf16x2(y0, y1);
    }
  }
  if (MODE == 1) {
    // fixed-order block reduction over the 8 warps (deterministic), in double
    __syncthreads();                       // xs / ws are dead: reuse ws as the exchange buffer
    float* ex = ws;                        // [8 warps][512][2]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = 64 * i + 2 * lane;
      ex[(warp * CONV0_C + c) * 2 + 0] = gsum[2 * i];
      ex[(warp * CONV0_C + c) * 2 + 1] = gsq[2 * i];
      ex[(warp * CONV0_C + c + 1) * 2 + 0] = gsum[2 * i + 1];
      ex[(warp * CONV0_C + c + 1) * 2 + 1] = gsq[2 * i + 1];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < CONV0_C; c += blockDim.x) {
      double sm = 0.0, sq = 0.0;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) {
        sm += (double)ex[(wv * CONV0_C + c) * 2 + 0];
        sq += (double)ex[(wv * CONV0_C + c) * 2 + 1];
      }
      gn_partial[((int64_t)blockIdx.y * tiles_per_utt + blockIdx.x) * CONV0_C + c] = make_double2(sm, sq);
    }
  }
}

// per (utterance, channel): mean / biased variance over the valid frames -> scale = gamma * rstd, shift = beta - mean*scale
__global__ void conv0_gn_finalize_kernel(const double2* __restrict__ partial, const Conv0Utt* __restrict__ utts,
                                         int tiles_per_utt, const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float2* __restrict__ affine) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= CONV0_C) return;
  const Conv0Utt u = utts[b];
  const int ntile = (u.T0 + CONV0_TILE - 1) / CONV0_TILE;
  double sm = 0.0, sq = 0.0;
  for (int t = 0; t < ntile; ++t) {
    const double2 p = partial[((int64_t)b * tiles_per_utt + t) * CONV0_C + c];
    sm += p.x;
    sq += p.y;
  }
  const double n = (double)max(u.T0, 1);
  const double mean = sm / n;
  const double var = fmax(sq / n - mean * mean, 0.0);
  const float scale = gamma[c] * (float)(1.0 / sqrt(var + 1e-5));
  affine[(int64_t)b * CONV0_C + c] = make_float2(scale, beta[c] - (float)mean * scale);
}

// ---------------------------------------------------------------------------------------------
// Row LayerNorm over C = 128*NV channels, one warp per row, fp32 statistics (eps 1e-5), optional GELU,
// optional input gather (in_rowmap) and output scatter (out_rowmap). Serves the conv-layer LN+GELU
// (bf16 -> bf16 in place), the feature-projection LN (bf16 gather -> bf16), the two LNs of every
// transformer layer (fp32 residual stream -> bf16 GEMM operand) and the final LN (fp32 -> fp32).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load4<bf16>(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// WavLM gate fused into the LayerNorm that produces the attention input (HF modeling_wavlm.py:167-176):
// gate[row, head] = a * (b * const_head - 1) + 2,  (a, b) = sigmoid(sums of outputs 0-3 / 4-7 of gru_rel_pos_linear).
struct LnGate {
  const float* w2 = nullptr;      // [2, 64] the 8 x 64 weight summed over each group of four outputs
  const float* b2 = nullptr;      // [2]
  const float* gconst = nullptr;  // [heads]
  float* out = nullptr;           // [rows, heads], heads = row width / 64
};

// Rows per warp: narrow rows give a warp too little to do per trip to HBM, so it keeps several rows in flight.
// (Round 2 re-measured the choice with the RPW_ template parameter below, profiles/r02k_ln_rpw.log: width 512 at 1 / 2 / 4 / 8
// rows per warp = 457 / 421 / 424 / 637 us for 908 k rows (6.6 TB/s at 2-4, the HBM copy peak), width 1 024 at 1-2 / 4 = 36.9 /
// 45.1 us for 28 258 rows: the built-in choices stand.)
template <int NV>
struct LnRows {
  static constexpr int RPW = NV <= 4 ? 4 : (NV <= 8 ? 2 : 1);
};

template <int NV, typename TIn, typename TOut, bool GELU, int RPW_ = LnRows<NV>::RPW>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const TIn* __restrict__ in, int64_t ld_in,
                                                              TOut* __restrict__ out, int64_t ld_out,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int64_t rows,
                                                              const int32_t* __restrict__ in_rowmap,
                                                              const int32_t* __restrict__ out_rowmap, float eps,
                                                              bf16* __restrict__ out2 = nullptr, int64_t ld_out2 = 0,
                                                              const LnGate gate = LnGate()) {
  // out2: optional second, bf16 copy of the result (post-LN layers: the fp32 residual stream is normalised in place
  // and the same values feed the next GEMM)
  // gate: optional WavLM gate of the NEXT attention (head_dim 64: the 2 heads of a 128-column chunk are its two
  // half-warps), computed from the fp32 normalised row that is in registers anyway
  constexpr int RPW = RPW_;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  pdl_wait();
  pdl_trigger();
  if (row0 >= rows) return;
  // gate constants first: their (L2) latency must overlap the row loads, not follow the statistics
  float4 gwa = make_float4(0.f, 0.f, 0.f, 0.f), gwb = gwa;
  float gb0 = 0.f, gb1 = 0.f, gcst = 0.f;
  if (NV <= 8 && gate.out) {   // this lane's 4 columns of every chunk sit at offset 4 * (lane % 16) inside their head
    gwa = __ldg(reinterpret_cast<const float4*>(gate.w2 + 4 * (lane & 15)));
    gwb = __ldg(reinterpret_cast<const float4*>(gate.w2 + 64 + 4 * (lane & 15)));
    gb0 = __ldg(gate.b2);
    gb1 = __ldg(gate.b2 + 1);
    if ((lane & 15) < NV) gcst = __ldg(gate.gconst + 2 * (lane & 15) + (lane >> 4));
  }
  int64_t irow[RPW], orow[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int64_t row = row0 + r;
    irow[r] = orow[r] = -1;
    if (row < rows) {
      irow[r] = in_rowmap ? (int64_t)in_rowmap[row] : row;
      orow[r] = out_rowmap ? (int64_t)out_rowmap[row] : row;
      if (irow[r] < 0 || orow[r] < 0) irow[r] = orow[r] = -1;
    }
  }
  float4 v[RPW][NV];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      v[r][i] = irow[r] >= 0 ? load4<TIn>(in + irow[r] * ld_in + (lane + 32 * i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  constexpr float invC = 1.f / (128.f * NV);
  float mu[RPW], rs[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    // packed fp32x2 arithmetic (FADD2 / FFMA2): same two-pass statistics, half the issue slots
    uint64_t s2 = 0ull;   // (+0.0f, +0.0f)
#pragma unroll
    for (int i = 0; i < NV; ++i) s2 = fadd2(s2, fadd2(pack_f32x2(v[r][i].x, v[r][i].y), pack_f32x2(v[r][i].z, v[r][i].w)));
    float s_lo, s_hi;
    unpack_f32x2(s2, s_lo, s_hi);
    mu[r] = warp_sum(s_lo + s_hi) * invC;
    const uint64_t nmu2 = pack_f32x2(-mu[r], -mu[r]);
    uint64_t q2 = 0ull;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint64_t ab = fadd2(pack_f32x2(v[r][i].x, v[r][i].y), nmu2), cd = fadd2(pack_f32x2(v[r][i].z, v[r][i].w), nmu2);
      q2 = ffma2(ab, ab, q2);
      q2 = ffma2(cd, cd, q2);
    }
    float q_lo, q_hi;
    unpack_f32x2(q2, q_lo, q_hi);
    rs[r] = rsqrtf(warp_sum(q_lo + q_hi) * invC + eps);
  }
  // gate partials: after the first exchange (xor 8) lanes with bit 3 clear carry a_i, lanes with bit 3 set b_i
  float gu[RPW][NV <= 8 ? NV : 1];
  const bool ghi = (lane & 8) != 0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (orow[r] >= 0) {
      const uint64_t nmu2 = pack_f32x2(-mu[r], -mu[r]), rs2 = pack_f32x2(rs[r], rs[r]);
      unpack_f32x2(ffma2(fmul2(fadd2(pack_f32x2(v[r][i].x, v[r][i].y), nmu2), rs2), pack_f32x2(g.x, g.y), pack_f32x2(be.x, be.y)), o.x, o.y);
      unpack_f32x2(ffma2(fmul2(fadd2(pack_f32x2(v[r][i].z, v[r][i].w), nmu2), rs2), pack_f32x2(g.z, g.w), pack_f32x2(be.z, be.w)), o.z, o.w);
      if (GELU) {
        gelu_fast2(o.x, o.y, o.x, o.y);
        gelu_fast2(o.z, o.w, o.z, o.w);
      }
      store4<TOut>(out + orow[r] * ld_out + c, o);
      if (out2) store4<bf16>(out2 + orow[r] * ld_out2 + c, o);
      }
      if (NV <= 8 && gate.out) {
        const float a = (o.x * gwa.x + o.y * gwa.y) + (o.z * gwa.z + o.w * gwa.w);
        const float b = (o.x * gwb.x + o.y * gwb.y) + (o.z * gwb.z + o.w * gwb.w);
        gu[r][NV <= 8 ? i : 0] = (ghi ? b : a) + __shfl_xor_sync(0xffffffffu, ghi ? a : b, 8);
      }
    }
  }
  if (NV <= 8 && gate.out) {   // WavLM has head_dim 64 and at most 16 heads
    // Head 2 i + (lane / 16) needs its partial sums added over the 16 lanes of its half-warp. The remaining values
    // are reduced TRANSPOSED: at every step a lane hands half of its values to its partner and keeps the other half,
    // so 4 + 2 + 1 more shuffles leave a_i (b_i) fully summed in lane i (8 + i) of the half-warp.
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = i < NV ? gu[r][i < NV && NV <= 8 ? i : 0] : 0.f;
#pragma unroll
      for (int w = 4; w > 0; w >>= 1) {
        const bool hi = (lane & w) != 0;
#pragma unroll
        for (int j = 0; j < w; ++j) {
          const float send = hi ? t[j] : t[j + w];
          const float keep = hi ? t[j + w] : t[j];
          t[j] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
      }
      const float sb = __shfl_xor_sync(0xffffffffu, t[0], 8);   // b_i from lane 8 + i
      const int i = lane & 15;
      if (orow[r] >= 0 && i < NV) {
        const int head = 2 * i + (lane >> 4);
        const float s0 = 1.f / (1.f + __expf(-(t[0] + gb0)));
        const float s1 = 1.f / (1.f + __expf(-(sb + gb1)));
        gate.out[orow[r] * (2 * NV) + head] = s0 * (s1 * gcst - 1.f) + 2.f;
      }
    }
  }
}

// valid rows of the conv6 output (slot layout) -> packed rows, no normalisation (HuBERT-base: feat_proj_layer_norm=False)
__global__ void gather_rows_bf16_kernel(const bf16* __restrict__ in, int64_t ld_in, bf16* __restrict__ out, int64_t ld_out,
                                        int64_t rows, int cols8, const int32_t* __restrict__ in_rowmap) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols8) return;
  const int64_t r = idx / cols8;
  const int c = (int)(idx - r * cols8) * 8;
  *reinterpret_cast<uint4*>(out + r * ld_out + c) = *reinterpret_cast<const uint4*>(in + (int64_t)in_rowmap[r] * ld_in + c);
}

// ---------------------------------------------------------------------------------------------
// fp32 packed residual stream -> bf16 rows of the positional-conv input buffer (zero gaps between
// utterances = the conv's own zero padding; channels regrouped to g*cg_pad + c). Buffer is pre-zeroed.
// ---------------------------------------------------------------------------------------------
__global__ void scatter_posconv_in_kernel(const float* __restrict__ x, int64_t rows, int d, int cg, int cg_pad,
                                          const int32_t* __restrict__ gap_row /*[rows]*/, bf16* __restrict__ out,
                                          int64_t ld_out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 channels
  const int d4 = d >> 2;
  if (idx >= rows * d4) return;
  const int64_t r = idx / d4;
  const int c = (int)(idx - r * d4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(x + r * d + c);
  const int g = c / cg, cc = c - g * cg;  // cg % 4 == 0
  store4<bf16>(out + (int64_t)gap_row[r] * ld_out + g * cg_pad + cc, v);
}

// acc = (first ? 0 : acc) + scale * x      (mean over selected hidden states, preprocess_speech.py:56-63)
__global__ void accum_scaled_kernel(float* __restrict__ acc, const float* __restrict__ x, int64_t n4, float scale,
                                    int first) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  float4 a = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(acc)[i];
  a.x = fmaf(scale, v.x, a.x); a.y = fmaf(scale, v.y, a.y);
  a.z = fmaf(scale, v.z, a.z); a.w = fmaf(scale, v.w, a.w);
  reinterpret_cast<float4*>(acc)[i] = a;
}

// ---------------------------------------------------------------------------------------------
// Masked mean pooling over the valid frames of each utterance (lora_wavlm/model.py:189-195 of the
// reference): out[b, :] = sum_t x[off[b] + t, :] / n_b, t < n_b. Packed layout => the mask is the
// frame range. Block = (128 columns, utterance b); 8 warps stride over frames, fixed-order smem
// reduction => bitwise deterministic regardless of how utterances are sharded over GPUs.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) masked_mean_pool_kernel(const float* __restrict__ x, int d,
                                                                const int32_t* __restrict__ frame_off /*[B+1]*/,
                                                                const int32_t* __restrict__ n_keep /*[B] or null*/,
                                                                float* __restrict__ out /*[B, d]*/, float scale,
                                                                int accumulate) {
  // out = (accumulate ? out : 0) + scale * mean: the mean (or weighted sum) over selected hidden states commutes with
  // the mean over frames, so a pooled-only caller never materialises the [sum_T, d] average (emit_hidden).
  __shared__ float4 red[8][32];
  const int b = blockIdx.y;
  const int c = blockIdx.x * 128 + (threadIdx.x & 31) * 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = frame_off[b];
  int n = frame_off[b + 1] - r0;
  if (n_keep) n = min(n, n_keep[b]);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < d) {
    for (int t = warp; t < n; t += 8) {
      const float4 v = *reinterpret_cast<const float4*>(x + (int64_t)(r0 + t) * d + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && c < d) {
    float4 a = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      a.x += red[w][lane].x; a.y += red[w][lane].y; a.z += red[w][lane].z; a.w += red[w][lane].w;
    }
    const float inv = scale / (float)max(n, 1);
    float4* o = reinterpret_cast<float4*>(out + (int64_t)b * d + c);
    float4 r = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
    if (accumulate) {
      const float4 prev = *o;
      r.x += prev.x; r.y += prev.y; r.z += prev.z; r.w += prev.w;
    }
    *o = r;
  }
}

// packed [sum T, d] -> padded [B, Tmax, d] (HF-shaped hidden states; pad frames are zero)
__global__ void unpack_frames_kernel(const float* __restrict__ x, int d, const int32_t* __restrict__ frame_off,
                                     int Tmax, float* __restrict__ out) {
  const int b = blockIdx.z, t = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= d) return;
  const int r0 = frame_off[b], n = frame_off[b + 1] - r0;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < n) v = *reinterpret_cast<const float4*>(x + (int64_t)(r0 + t) * d + c);
  *reinterpret_cast<float4*>(out + ((int64_t)b * Tmax + t) * d + c) = v;
}

}  // namespace serenc

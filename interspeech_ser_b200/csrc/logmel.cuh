// Whisper log-mel frontend in fp32 (WhisperFeatureExtractor._torch_extract_fbank_features,
// HF feature_extraction_whisper.py:135-164): pad/truncate to 480 000 samples, periodic Hann(400),
// STFT(n_fft=400, hop=160, center=True, reflect) -> 3001 frames, last dropped; |.|^2; 201 -> n_mels
// slaney mel filterbank; log10(max(., 1e-10)); max(x, max_over_utterance - 8); (x + 4) / 4.
//
// n_fft = 400 is not a power of two. The real DFT is folded twice before the multiply-accumulate
// (n <-> 400-n symmetry of cos/sin, then n <-> 200-n symmetry split by bin parity), which leaves 99 cos +
// 99 sin MACs per bin instead of 400 + 400; twiddles come from one 400-entry cos/sin table in shared
// memory. One CTA computes LM_FR consecutive frames of one utterance: samples staged once in shared
// memory (coalesced), power spectrum kept in shared memory, mel projection through a CSR copy of the
// (sparse, triangular) filterbank, output written [B, n_mels, 3000] with frames contiguous.
// Pass 2 applies the per-utterance dynamic-range clamp and affine map in place.
#pragma once
#include "common.cuh"

namespace serenc {

constexpr int LM_NFFT = 400;
constexpr int LM_HOP = 160;
constexpr int LM_BINS = 201;
constexpr int LM_NSAMP = 480000;
constexpr int LM_FRAMES = 3000;
constexpr int LM_FR = 8;                                    // frames per CTA
constexpr int LM_SPAN = (LM_FR - 1) * LM_HOP + LM_NFFT;     // samples touched by one CTA
constexpr int LM_FOLD = 400;                                // per-frame folded layout, see below
constexpr int LM_THREADS = 256;

// order-preserving float <-> uint map so atomicMax works on signed floats
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct LogmelTables {
  const float* hann;      // [400]
  const float* costab;    // [400] cos(2 pi m / 400)
  const float* sintab;    // [400]
  const int32_t* mel_ptr; // [n_mels + 1] CSR row pointers
  const int32_t* mel_bin; // [nnz]
  const float* mel_w;     // [nnz]
  int n_mels;
};

// per-frame folded buffer (floats):
//   [0,100)   ce_even[n] = e[n] + e[200-n]   (n = 1..99 used), [100,200) ce_odd[n] = e[n] - e[200-n]
//   [200,300) so_even[n] = o[n] - o[200-n],                    [300,400) so_odd[n]  = o[n] + o[200-n]
//   slots n = 0 of each quarter hold the specials: xw[0], xw[200], e[100], o[100]
__global__ void __launch_bounds__(LM_THREADS) logmel_power_kernel(const float* __restrict__ wav,
                                                                   const UttSpan* __restrict__ utts,
                                                                   const LogmelTables tb, float* __restrict__ out,
                                                                   uint32_t* __restrict__ umax /*[B], ordered*/) {
  __shared__ float s_x[LM_SPAN];
  __shared__ float s_cos[LM_NFFT], s_sin[LM_NFFT];
  __shared__ float s_fold[LM_FR][LM_FOLD];
  __shared__ float s_pw[LM_FR][LM_BINS + 3];
  __shared__ float s_max[LM_THREADS / 32];

  const int b = blockIdx.y;
  const int f0 = blockIdx.x * LM_FR;
  const int tid = threadIdx.x;
  const float* x = wav + utts[b].sample_start;
  const int n_valid = min(utts[b].sample_len, LM_NSAMP);  // truncation; beyond n_valid the padded signal is 0

  for (int i = tid; i < LM_NFFT; i += LM_THREADS) {
    s_cos[i] = tb.costab[i];
    s_sin[i] = tb.sintab[i];
  }
  // padded signal index of the first sample of frame f is 160 f - 200 (center=True), reflect at both ends
  const int base = f0 * LM_HOP - LM_NFFT / 2;
  for (int i = tid; i < LM_SPAN; i += LM_THREADS) {
    int s = base + i;
    if (s < 0) s = -s;
    if (s >= LM_NSAMP) s = 2 * (LM_NSAMP - 1) - s;
    s_x[i] = (s < n_valid) ? x[s] : 0.f;
  }
  __syncthreads();

  // window + double fold
  for (int i = tid; i < LM_FR * 100; i += LM_THREADS) {
    const int f = i / 100, n = i - f * 100;
    const float* xf = s_x + f * LM_HOP;
    float* fo = s_fold[f];
    if (n == 0) {
      fo[0] = xf[0] * tb.hann[0];
      fo[100] = xf[200] * tb.hann[200];
      const float a = xf[100] * tb.hann[100], c = xf[300] * tb.hann[300];
      fo[200] = a + c;  // e[100]
      fo[300] = a - c;  // o[100]
    } else {
      const float x1 = xf[n] * tb.hann[n], x2 = xf[400 - n] * tb.hann[400 - n];
      const float x3 = xf[200 - n] * tb.hann[200 - n], x4 = xf[200 + n] * tb.hann[200 + n];
      const float e1 = x1 + x2, o1 = x1 - x2;  // e[n], o[n]
      const float e2 = x3 + x4, o2 = x3 - x4;  // e[200-n], o[200-n]
      fo[n] = e1 + e2;
      fo[100 + n] = e1 - e2;
      fo[200 + n] = o1 - o2;
      fo[300 + n] = o1 + o2;
    }
  }
  __syncthreads();

  // DFT bins: item = (frame, k), k fastest so a warp mostly shares the frame (broadcast operand reads)
  for (int i = tid; i < LM_FR * LM_BINS; i += LM_THREADS) {
    const int f = i / LM_BINS, k = i - f * LM_BINS;
    const float* fo = s_fold[f];
    const int odd = k & 1;
    const float* ce = fo + (odd ? 100 : 0);
    const float* so = fo + (odd ? 300 : 200);
    float re0 = 0.f, re1 = 0.f, im0 = 0.f, im1 = 0.f;
    int m = 0;
#pragma unroll 3
    for (int n = 1; n < 100; n += 2) {
      m += k; if (m >= LM_NFFT) m -= LM_NFFT;
      re0 = fmaf(ce[n], s_cos[m], re0);
      im0 = fmaf(so[n], s_sin[m], im0);
      if (n + 1 < 100) {
        m += k; if (m >= LM_NFFT) m -= LM_NFFT;
        re1 = fmaf(ce[n + 1], s_cos[m], re1);
        im1 = fmaf(so[n + 1], s_sin[m], im1);
      }
    }
    // specials: n = 0, n = 200 (sign (-1)^k), n = 100 (cos(pi k/2), sin(pi k/2))
    const float sgn = odd ? -1.f : 1.f;
    const int q = k & 3;
    const float c100 = (q == 0) ? 1.f : (q == 2 ? -1.f : 0.f);
    const float s100 = (q == 1) ? 1.f : (q == 3 ? -1.f : 0.f);
    const float re = (re0 + re1) + fo[0] + sgn * fo[100] + c100 * fo[200];
    const float im = (im0 + im1) + s100 * fo[300];
    s_pw[f][k] = re * re + im * im;
  }
  __syncthreads();

  // mel projection + log10; item = (mel, frame) with frame fastest (contiguous in the output)
  float lmax = -INFINITY;
  for (int i = tid; i < tb.n_mels * LM_FR; i += LM_THREADS) {
    const int mel = i / LM_FR, f = i - mel * LM_FR;
    const int frame = f0 + f;
    if (frame >= LM_FRAMES) continue;
    float acc = 0.f;
    const int p1 = tb.mel_ptr[mel + 1];
    for (int q = tb.mel_ptr[mel]; q < p1; ++q) acc = fmaf(tb.mel_w[q], s_pw[f][tb.mel_bin[q]], acc);
    const float lv = log10f(fmaxf(acc, 1e-10f));
    out[((int64_t)b * tb.n_mels + mel) * LM_FRAMES + frame] = lv;
    lmax = fmaxf(lmax, lv);
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) s_max[tid >> 5] = lmax;
  __syncthreads();
  if (tid == 0) {
    float m = s_max[0];
#pragma unroll
    for (int w = 1; w < LM_THREADS / 32; ++w) m = fmaxf(m, s_max[w]);
    atomicMax(umax + b, float_to_ordered(m));
  }
}

__global__ void logmel_finalize_kernel(float* __restrict__ out, const uint32_t* __restrict__ umax, int64_t per_utt) {
  const int b = blockIdx.y;
  const float floor_v = ordered_to_float(umax[b]) - 8.0f;
  float4* o = reinterpret_cast<float4*>(out + (int64_t)b * per_utt);
  const int64_t n4 = per_utt >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = o[i];
    v.x = (fmaxf(v.x, floor_v) + 4.0f) * 0.25f;
    v.y = (fmaxf(v.y, floor_v) + 4.0f) * 0.25f;
    v.z = (fmaxf(v.z, floor_v) + 4.0f) * 0.25f;
    v.w = (fmaxf(v.w, floor_v) + 4.0f) * 0.25f;
    o[i] = v;
  }
}

// [B, n_mels, 3000] fp32 -> channels-last bf16 rows with one zero row either side of each utterance
// (conv1's padding=1): out[b*(3002) + 1 + t, c]. Tile transpose through shared memory.
__global__ void __launch_bounds__(256) mel_to_rows_kernel(const float* __restrict__ mel, int n_mels,
                                                           bf16* __restrict__ out, int64_t ld_out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, t = t0 + tx;
    tile[j][tx] = (c < n_mels && t < LM_FRAMES) ? mel[((int64_t)b * n_mels + c) * LM_FRAMES + t] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int t = t0 + j, c = c0 + tx;
    if (t < LM_FRAMES && c < n_mels)
      out[((int64_t)b * (LM_FRAMES + 2) + 1 + t) * ld_out + c] = __float2bfloat16(tile[tx][j]);
  }
}

// x[b, t, :] = pos[t, :]   (residual stream pre-initialised with the sinusoidal positions; the conv2 GEMM
// epilogue then adds gelu(conv2(.)) in place; HF modeling_whisper.py:619-625)
__global__ void broadcast_rows_kernel(const float* __restrict__ pos, int64_t per_utt4, float* __restrict__ x) {
  const int b = blockIdx.y;
  const float4* p = reinterpret_cast<const float4*>(pos);
  float4* o = reinterpret_cast<float4*>(x) + (int64_t)b * per_utt4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_utt4; i += (int64_t)gridDim.x * blockDim.x)
    o[i] = p[i];
}

}  // namespace serenc

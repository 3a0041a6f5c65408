// Whisper log-mel frontend in fp32 (WhisperFeatureExtractor._torch_extract_fbank_features,
// HF feature_extraction_whisper.py:135-164): pad/truncate to 480 000 samples, periodic Hann(400),
// STFT(n_fft=400, hop=160, center=True, reflect) -> 3001 frames, last dropped; |.|^2; 201 -> n_mels
// slaney mel filterbank; log10(max(., 1e-10)); max(x, max_over_utterance - 8); (x + 4) / 4.
//
// n_fft = 400 is not a power of two. The real DFT is folded twice before the multiply-accumulate
// (n <-> 400-n symmetry of cos/sin, then n <-> 200-n symmetry split by bin parity), which leaves 99 cos +
// 99 sin MACs per bin instead of 400 + 400.
//
// One CTA computes LM_FR = 16 consecutive frames of one utterance; one THREAD owns LM_BPT = 2 frequency bins and keeps
// their real / imaginary sums of all 16 frames in registers (as fp32x2 pairs of frames, FFMA2). Per folded sample n it
// needs its own twiddle (cos, sin)(2 pi n k / 400) - one coalesced 8-byte load from a precomputed [99][256] matrix
// (L2-resident, 203 KB) - and the folded samples of the 16 frames, stored frame-contiguous in shared memory so that
// they arrive as broadcast LDS.128. The lower half of the CTA owns the even bins, the upper half the odd ones (the fold
// differs by bin parity, so a warp reads one address). That is 10 loads per 32 FMAs; the previous form (one (frame, bin) per thread,
// twiddles gathered from a 400-entry shared-memory table with bank conflicts) issued 4 loads per 2 FMAs and ran at
// 1.5 TFMA/s. Power spectrum -> shared memory -> mel projection through a CSR copy of the (sparse, triangular)
// filterbank, output written [B, n_mels, 3000] with frames contiguous. Pass 2 applies the per-utterance
// dynamic-range clamp and affine map in place.
#pragma once
#include "common.cuh"

namespace serenc {

constexpr int LM_NFFT = 400;
constexpr int LM_HOP = 160;
constexpr int LM_BINS = 201;
constexpr int LM_NSAMP = 480000;
constexpr int LM_FRAMES = 3000;
constexpr int LM_FR = 16;                                   // frames per CTA
constexpr int LM_SPAN = (LM_FR - 1) * LM_HOP + LM_NFFT;     // samples touched by one CTA
constexpr int LM_FOLD = 400;                                // per-frame folded layout, see below
constexpr int LM_BPT = 2;                                    // bins per thread (registers: LM_BPT x LM_FR sums, re + im); measured 1 / 2 / 4: 0.74 / 0.64 / 0.89 ms
constexpr int LM_THREADS = 256 / LM_BPT;

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
  const float2* twid;     // [99][256]: (cos, sin)(2 pi n k / 400) for n = 1..99, k = 2 * (t % 128) + t / 128 of thread t
  const int32_t* mel_ptr; // [n_mels + 1] CSR row pointers
  const int32_t* mel_bin; // [nnz]
  const float* mel_w;     // [nnz]
  int n_mels;
};

// folded samples in shared memory: s_fold[q][n][f], f = frame inside the CTA (contiguous), q =
//   0: ce_even[n] = e[n] + e[200-n]   1: ce_odd[n] = e[n] - e[200-n]      (e[n] = xw[n] + xw[400-n])
//   2: so_even[n] = o[n] - o[200-n]   3: so_odd[n]  = o[n] + o[200-n]     (o[n] = xw[n] - xw[400-n])
// for n = 1..99; the n = 0 rows hold the specials xw[0], xw[200], e[100], o[100]
__global__ void __launch_bounds__(LM_THREADS) logmel_power_kernel(const void* __restrict__ wav, int wav_i16,
                                                                   const UttSpan* __restrict__ utts,
                                                                   const LogmelTables tb, float* __restrict__ out,
                                                                   uint32_t* __restrict__ umax /*[B], ordered*/) {
  constexpr int PW_LD = LM_BINS + 3;
  constexpr int XP = (LM_SPAN > LM_FR * PW_LD) ? LM_SPAN : LM_FR * PW_LD;
  __shared__ __align__(16) float s_fold[4][100][LM_FR];
  __shared__ float s_xp[XP];           // samples of the CTA's span, later the power spectra [LM_FR][PW_LD]
  __shared__ float s_hann[LM_NFFT];
  __shared__ float s_max[LM_THREADS / 32];

  const int b = blockIdx.y;
  const int f0 = blockIdx.x * LM_FR;
  const int tid = threadIdx.x;
  const int64_t x0 = utts[b].sample_start;
  const int n_valid = min(utts[b].sample_len, LM_NSAMP);  // truncation; beyond n_valid the padded signal is 0

  for (int i = tid; i < LM_NFFT; i += LM_THREADS) s_hann[i] = tb.hann[i];
  // padded signal index of the first sample of frame f is 160 f - 200 (center=True), reflect at both ends
  const int base = f0 * LM_HOP - LM_NFFT / 2;
  for (int i = tid; i < LM_SPAN; i += LM_THREADS) {
    int s = base + i;
    if (s < 0) s = -s;
    if (s >= LM_NSAMP) s = 2 * (LM_NSAMP - 1) - s;
    s_xp[i] = (s >= 0 && s < n_valid) ? load_sample(wav, wav_i16, x0 + s) : 0.f;
  }
  __syncthreads();

  // window + double fold; item = (n, frame) with the frame fastest (contiguous stores)
  for (int i = tid; i < LM_FR * 100; i += LM_THREADS) {
    const int n = i / LM_FR, f = i - n * LM_FR;
    const float* xf = s_xp + f * LM_HOP;
    if (n == 0) {
      s_fold[0][0][f] = xf[0] * s_hann[0];
      s_fold[1][0][f] = xf[200] * s_hann[200];
      const float a = xf[100] * s_hann[100], c = xf[300] * s_hann[300];
      s_fold[2][0][f] = a + c;  // e[100]
      s_fold[3][0][f] = a - c;  // o[100]
    } else {
      const float x1 = xf[n] * s_hann[n], x2 = xf[400 - n] * s_hann[400 - n];
      const float x3 = xf[200 - n] * s_hann[200 - n], x4 = xf[200 + n] * s_hann[200 + n];
      const float e1 = x1 + x2, o1 = x1 - x2;  // e[n], o[n]
      const float e2 = x3 + x4, o2 = x3 - x4;  // e[200-n], o[200-n]
      s_fold[0][n][f] = e1 + e2;
      s_fold[1][n][f] = e1 - e2;
      s_fold[2][n][f] = o1 - o2;
      s_fold[3][n][f] = o1 + o2;
    }
  }
  __syncthreads();   // s_xp (samples) is dead from here on

  // DFT: this thread's LM_BPT bins for all LM_FR frames. Thread t: parity = upper half of the CTA, bins
  // k_j = 2 * (t % H + j * H) + parity with H = LM_THREADS / 2 lanes per parity (twiddle column = (k_j / 2) + 128 * parity).
  // Every folded sample loaded from shared memory feeds LM_BPT FMAs: broadcast LDS.128 costs 4 shared-memory
  // wavefronts whatever the number of lanes that need it, so bins per thread is what buys the FMA rate.
  constexpr int H = LM_THREADS / 2;
  const int odd = tid / H;
  const int idx0 = tid - odd * H;
  uint64_t re2[LM_BPT][LM_FR / 2], im2[LM_BPT][LM_FR / 2];  // frame pairs (f, f + 1)
#pragma unroll
  for (int j = 0; j < LM_BPT; ++j)
#pragma unroll
    for (int q = 0; q < LM_FR / 2; ++q) re2[j][q] = im2[j][q] = 0ull;
  {
    const float2* tw = tb.twid + idx0 + 128 * odd;     // + j * H per bin; columns of bins > 200 hold zeros
    const uint4* ce = reinterpret_cast<const uint4*>(&s_fold[odd ? 1 : 0][0][0]);   // [n][LM_FR / 4]
    const uint4* so = reinterpret_cast<const uint4*>(&s_fold[odd ? 3 : 2][0][0]);
    float2 t_nxt[LM_BPT];
#pragma unroll
    for (int j = 0; j < LM_BPT; ++j) t_nxt[j] = __ldg(tw + j * H);
#pragma unroll 1
    for (int n = 1; n < 100; ++n) {
      uint64_t c2[LM_BPT], s2[LM_BPT];
#pragma unroll
      for (int j = 0; j < LM_BPT; ++j) {
        c2[j] = pack_f32x2(t_nxt[j].x, t_nxt[j].x);
        s2[j] = pack_f32x2(t_nxt[j].y, t_nxt[j].y);
        if (n < 99) t_nxt[j] = __ldg(tw + n * 256 + j * H);
      }
#pragma unroll
      for (int q = 0; q < LM_FR / 4; ++q) {
        const uint4 cv = ce[n * (LM_FR / 4) + q], sv = so[n * (LM_FR / 4) + q];
        const uint64_t ca = pack_f32x2(__uint_as_float(cv.x), __uint_as_float(cv.y)), cb = pack_f32x2(__uint_as_float(cv.z), __uint_as_float(cv.w));
        const uint64_t sa = pack_f32x2(__uint_as_float(sv.x), __uint_as_float(sv.y)), sb = pack_f32x2(__uint_as_float(sv.z), __uint_as_float(sv.w));
#pragma unroll
        for (int j = 0; j < LM_BPT; ++j) {
          re2[j][2 * q] = ffma2(ca, c2[j], re2[j][2 * q]);
          re2[j][2 * q + 1] = ffma2(cb, c2[j], re2[j][2 * q + 1]);
          im2[j][2 * q] = ffma2(sa, s2[j], im2[j][2 * q]);
          im2[j][2 * q + 1] = ffma2(sb, s2[j], im2[j][2 * q + 1]);
        }
      }
    }
    // specials: n = 0, n = 200 (sign (-1)^k), n = 100 (cos(pi k/2), sin(pi k/2))
    const float sgn = odd ? -1.f : 1.f;
#pragma unroll
    for (int j = 0; j < LM_BPT; ++j) {
      const int k = 2 * (idx0 + j * H) + odd;
      if (k < LM_BINS) {
        const int q4 = k & 3;
        const float c100 = (q4 == 0) ? 1.f : (q4 == 2 ? -1.f : 0.f);
        const float s100 = (q4 == 1) ? 1.f : (q4 == 3 ? -1.f : 0.f);
#pragma unroll
        for (int q = 0; q < LM_FR / 2; ++q) {
          float r0, r1, i0, i1;
          unpack_f32x2(re2[j][q], r0, r1);
          unpack_f32x2(im2[j][q], i0, i1);
          const int fa = 2 * q, fb = 2 * q + 1;
          const float rea = r0 + s_fold[0][0][fa] + sgn * s_fold[1][0][fa] + c100 * s_fold[2][0][fa];
          const float reb = r1 + s_fold[0][0][fb] + sgn * s_fold[1][0][fb] + c100 * s_fold[2][0][fb];
          const float ima = i0 + s100 * s_fold[3][0][fa];
          const float imb = i1 + s100 * s_fold[3][0][fb];
          s_xp[fa * PW_LD + k] = rea * rea + ima * ima;
          s_xp[fb * PW_LD + k] = reb * reb + imb * imb;
        }
      }
    }
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
    for (int q = tb.mel_ptr[mel]; q < p1; ++q) acc = fmaf(tb.mel_w[q], s_xp[f * PW_LD + tb.mel_bin[q]], acc);
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

// Flash-style (streaming softmax) non-causal self-attention over packed, variable-length utterances.
//
//   scores[i, j] = scale * q_i . k_j  (+ gate[i] * bias_h[j - i]  for WavLM)   ; keys j >= T are excluded
//   out[i, :]    = softmax_j(scores[i, :]) @ V
//
// * wav2vec2 / HuBERT / Whisper: plain SDPA (HF modeling_wav2vec2.py:438-549, modeling_whisper.py:215-357).
// * WavLM: gated relative position bias (HF modeling_wavlm.py:147-271). The [H, T, T] bias tensor HF
//   materialises is never built: bias_h[delta] is a per-head Toeplitz vector (bucket table expanded once at
//   create time, saturating for |delta| >= 778 so a 2*1024-1 entry table covers every length), and the
//   per-row gate  g = a * (b * const_h - 1) + 2,  (a, b) = sigmoid(sum4(gru_rel_pos_linear(x_i,head)))
//   is computed in the kernel prologue from the layer input.
//
// THIS FILE: the shared parameter block and the first-generation kernel, kept as the A/B arm (SERENC_ATTN_MMA_SYNC=1)
// and as the fallback for combinations the tcgen05 kernels do not cover (gated bias with head_dim 80 / 120: no such
// checkpoint exists). The production kernels are attention_tc.cuh (head_dim 64, WavLM bias) and attention_tc_wide.cuh
// (head_dim 80 / 120).
// Tensor-core path here: mma.sync m16n8k16 bf16 (fp32 accumulate), ldmatrix from XOR-swizzled shared memory,
// cp.async double-buffered K/V tiles. One CTA = 64 query rows of one (utterance, head); 4 warps x 16 rows.
// head_dim 64 / 80 / 120 (80 and 120 are zero-padded to 128 columns in shared memory only).
#pragma once
#include "attention_params.cuh"
#include "common.cuh"

namespace serenc {

constexpr int ATT_BM = 64;
constexpr int ATT_BN = 64;
constexpr int ATT_THREADS = 128;
template <int HD>
struct AttnCfg {
  static constexpr int DP = (HD == 64) ? 64 : 128;          // padded row width in smem (elements)
  static constexpr int CHUNKS = DP / 8;                     // 16-byte chunks per row
  static constexpr int VCHUNKS = HD / 8;                    // chunks that hold data
  static constexpr int KSTEPS = (HD + 15) / 16;
  static constexpr int NT_O = (HD + 7) / 8;                 // output n-tiles
  static constexpr int NP_O = (NT_O + 1) / 2;               // pairs
  static constexpr int TILE_BYTES = ATT_BN * DP * 2;
  static constexpr int SMEM_BYTES = 5 * TILE_BYTES;         // Q + 2K + 2V
};

template <int HD>
__device__ __forceinline__ void att_load_tile(uint8_t* smem_tile, const bf16* gbase, int64_t ld, int row0, int nrows,
                                              int tid) {
  using C = AttnCfg<HD>;
  // 64 rows x CHUNKS chunks, 128 threads
  for (int i = tid; i < ATT_BN * C::CHUNKS; i += ATT_THREADS) {
    const int r = i / C::CHUNKS, c = i - r * C::CHUNKS;
    const bool valid = (r < nrows) && (c < C::VCHUNKS);
    const bf16* src = valid ? gbase + (int64_t)(row0 + r) * ld + c * 8 : gbase;
    cp_async_16(smem_tile + r * (C::DP * 2) + ((c ^ (r & 7)) << 4), src, valid);
  }
}

// One 16-query x 64-key tile step of one warp: S = Q K^T, scale + gated bias + mask, online softmax, O += P V.
// FULL = all 64 keys valid (no mask, no group skipping; keeps the instruction stream branch-free so the MMAs and
// exp2s of different 8-key groups interleave); !FULL = the utterance's last, ragged key tile.
template <int HD, bool WAVLM, bool FULL>
__device__ __forceinline__ void att_tile_step(const uint8_t* kt_s, const uint8_t* vt_s, const uint32_t (&qf)[AttnCfg<HD>::KSTEPS][4],
                                              float (&o_acc)[2 * AttnCfg<HD>::NP_O][4], float (&m_run)[2], float (&l_run)[2],
                                              const float (&gate_r)[2], const float* bwin, float sc2, int nvalid, int warp,
                                              int lane) {
  using C = AttnCfg<HD>;
  const int g = lane >> 2, t4 = lane & 3;
  const int npairs = FULL ? 4 : ((nvalid + 15) >> 4);
  float s[8][4];
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    s[2 * np][0] = s[2 * np][1] = s[2 * np][2] = s[2 * np][3] = 0.f;
    s[2 * np + 1][0] = s[2 * np + 1][1] = s[2 * np + 1][2] = s[2 * np + 1][3] = 0.f;
    if (FULL || np < npairs) {
#pragma unroll
      for (int ks = 0; ks < C::KSTEPS; ++ks) {
        const int mi = lane >> 3;
        const int key = np * 16 + (mi >> 1) * 8 + (lane & 7);
        const int chunk = 2 * ks + (mi & 1);
        uint32_t kb[4];
        ldmatrix_x4(kb, smem_u32(kt_s + key * (C::DP * 2) + ((chunk ^ (key & 7)) << 4)));
        mma_bf16_16816(s[2 * np], qf[ks], kb[0], kb[1]);
        mma_bf16_16816(s[2 * np + 1], qf[ks], kb[2], kb[3]);
      }
    }
  }
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (FULL || (nt >> 1) < npairs) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + 2 * t4 + (e & 1);
        float v = s[nt][e] * sc2;
        if (WAVLM) v = fmaf(gate_r[e >> 1], bwin[col - (warp * 16 + g + (e >> 1) * 8) + 63], v);
        if (!FULL && col >= nvalid) v = -INFINITY;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
  }
  float alpha[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    const float m_new = fmaxf(m_run[r], mx[r]);  // finite: every key tile holds at least one valid key
    alpha[r] = exp2f(m_run[r] - m_new);
    m_run[r] = m_new;
    l_run[r] *= alpha[r];
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (FULL || (nt >> 1) < npairs) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(s[nt][e] - m_run[e >> 1]);
        s[nt][e] = pv;
        l_run[e >> 1] += pv;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2 * C::NP_O; ++i) {
    o_acc[i][0] *= alpha[0]; o_acc[i][1] *= alpha[0];
    o_acc[i][2] *= alpha[1]; o_acc[i][3] *= alpha[1];
  }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    if (FULL || kk < npairs) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < C::NP_O; ++dp) {
        const int mi = lane >> 3;
        const int key = kk * 16 + (mi & 1) * 8 + (lane & 7);
        const int chunk = 2 * dp + (mi >> 1);
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, smem_u32(vt_s + key * (C::DP * 2) + ((chunk ^ (key & 7)) << 4)));
        mma_bf16_16816(o_acc[2 * dp], pa, vb[0], vb[1]);
        mma_bf16_16816(o_acc[2 * dp + 1], pa, vb[2], vb[3]);
      }
    }
  }
}

template <int HD, bool WAVLM>
__global__ void __launch_bounds__(ATT_THREADS, (HD == 64) ? 4 : 2) attention_fwd_kernel(const AttnParams p) {
  using C = AttnCfg<HD>;
  extern __shared__ __align__(128) uint8_t att_smem[];
  __shared__ float s_bwin[2][128];
  __shared__ float s_gate[ATT_BM];
  __shared__ float4 s_gw[WAVLM ? HD * 2 : 1];  // gru_rel_pos_linear weight, [k][8 outputs] as two float4

  const int b = blockIdx.z, h = blockIdx.y;
  const int r0 = p.frame_off[b];
  const int T = p.frame_off[b + 1] - r0;
  const int i0 = blockIdx.x * ATT_BM;
  if (i0 >= T) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;

  uint8_t* sQ = att_smem;
  uint8_t* sK = att_smem + C::TILE_BYTES;
  uint8_t* sV = att_smem + 3 * C::TILE_BYTES;

  const bf16* qbase = p.qkv + (int64_t)r0 * p.ld_qkv + h * HD;
  const bf16* kbase = qbase + p.d;
  const bf16* vbase = qbase + 2 * p.d;

  att_load_tile<HD>(sQ, qbase, p.ld_qkv, i0, T - i0, tid);
  att_load_tile<HD>(sK, kbase, p.ld_qkv, 0, T, tid);
  att_load_tile<HD>(sV, vbase, p.ld_qkv, 0, T, tid);
  cp_async_commit();

  if (WAVLM) {
    // gate for the 64 query rows (HF modeling_wavlm.py:167-176): two threads per row, each half of the head dims;
    // the 8 x HD weight is staged transposed in shared memory so every k costs two broadcast float4 reads.
    float* gw = reinterpret_cast<float*>(s_gw);
    for (int i = tid; i < 8 * HD; i += ATT_THREADS) {
      const int o = i / HD, k = i - o * HD;
      gw[k * 8 + o] = __ldg(p.gru_w + i);
    }
    __syncthreads();
    const int row = tid >> 1, half = tid & 1;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
    if (i0 + row < T) {
      const bf16* x = p.hln + (int64_t)(r0 + i0 + row) * p.d + h * HD + half * (HD / 2);
#pragma unroll 4
      for (int k = 0; k < HD / 2; k += 2) {
        const float2 xv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + k));
        const int kk = half * (HD / 2) + k;
        const float4 w0 = s_gw[kk * 2], w1 = s_gw[kk * 2 + 1], w2 = s_gw[kk * 2 + 2], w3 = s_gw[kk * 2 + 3];
        acc[0] = fmaf(xv.x, w0.x, acc[0]); acc[1] = fmaf(xv.x, w0.y, acc[1]);
        acc[2] = fmaf(xv.x, w0.z, acc[2]); acc[3] = fmaf(xv.x, w0.w, acc[3]);
        acc[4] = fmaf(xv.x, w1.x, acc[4]); acc[5] = fmaf(xv.x, w1.y, acc[5]);
        acc[6] = fmaf(xv.x, w1.z, acc[6]); acc[7] = fmaf(xv.x, w1.w, acc[7]);
        acc[0] = fmaf(xv.y, w2.x, acc[0]); acc[1] = fmaf(xv.y, w2.y, acc[1]);
        acc[2] = fmaf(xv.y, w2.z, acc[2]); acc[3] = fmaf(xv.y, w2.w, acc[3]);
        acc[4] = fmaf(xv.y, w3.x, acc[4]); acc[5] = fmaf(xv.y, w3.y, acc[5]);
        acc[6] = fmaf(xv.y, w3.z, acc[6]); acc[7] = fmaf(xv.y, w3.w, acc[7]);
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 1);
    if (half == 0) {
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        sa += acc[o] + __ldg(p.gru_b + o);
        sb += acc[o + 4] + __ldg(p.gru_b + o + 4);
      }
      const float ga = 1.f / (1.f + __expf(-sa));
      const float gb = 1.f / (1.f + __expf(-sb));
      s_gate[row] = ga * (gb * __ldg(p.gru_const + h) - 1.f) + 2.f;
    }
  }

  const int nkt = (T + ATT_BN - 1) / ATT_BN;
  constexpr float LOG2E = 1.4426950408889634f;
  const float sc2 = p.scale * LOG2E;

  uint32_t qf[C::KSTEPS][4];
  float o_acc[2 * C::NP_O][4];
#pragma unroll
  for (int i = 0; i < 2 * C::NP_O; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  float gate_r[2] = {0.f, 0.f};
  // Tile quantisation: a warp whose 16 query rows are all past the utterance end does no math (it still helps
  // with the cooperative loads), and 16-key groups past the last valid key are skipped in the ragged last tile:
  // T = 199 pays for 208 x 208, not 256 x 256.
  const bool warp_active = (i0 + warp * 16) < T;

  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    const int j0 = kt * ATT_BN;
    if (kt + 1 < nkt) {
      att_load_tile<HD>(sK + (buf ^ 1) * C::TILE_BYTES, kbase, p.ld_qkv, j0 + ATT_BN, T - j0 - ATT_BN, tid);
      att_load_tile<HD>(sV + (buf ^ 1) * C::TILE_BYTES, vbase, p.ld_qkv, j0 + ATT_BN, T - j0 - ATT_BN, tid);
      cp_async_commit();
    }
    if (WAVLM) {
      // bias window: x in [0,127): delta = j0 - i0 - 63 + x  (element (i,j) uses x = (j-j0) - (i-i0) + 63)
      if (tid < 127) {
        int idx = j0 - i0 - 63 + tid + (WAVLM_MAXD - 1);
        idx = max(0, min(2 * WAVLM_MAXD - 2, idx));
        s_bwin[buf][tid] = __ldg(p.btab + (int64_t)h * (2 * WAVLM_MAXD - 1) + idx) * LOG2E;
      }
    }
    if (kt + 1 < nkt) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();

    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < C::KSTEPS; ++ks) {
        const int row = warp * 16 + (lane & 15);
        const int chunk = 2 * ks + (lane >> 4);
        ldmatrix_x4(qf[ks], smem_u32(sQ + row * (C::DP * 2) + ((chunk ^ (row & 7)) << 4)));
      }
      if (WAVLM) {
        gate_r[0] = s_gate[warp * 16 + g];
        gate_r[1] = s_gate[warp * 16 + g + 8];
      }
    }
    if (warp_active) {
      const uint8_t* kt_s = sK + buf * C::TILE_BYTES;
      const uint8_t* vt_s = sV + buf * C::TILE_BYTES;
      if (j0 + ATT_BN <= T)
        att_tile_step<HD, WAVLM, true>(kt_s, vt_s, qf, o_acc, m_run, l_run, gate_r, s_bwin[buf], sc2, ATT_BN, warp, lane);
      else
        att_tile_step<HD, WAVLM, false>(kt_s, vt_s, qf, o_acc, m_run, l_run, gate_r, s_bwin[buf], sc2, T - j0, warp, lane);
    }
    __syncthreads();
  }

  // ---- finalize ----
  if (!warp_active) return;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int row_a = i0 + warp * 16 + g, row_b = row_a + 8;
  bf16* oa = p.out + (int64_t)(r0 + row_a) * p.d + h * HD;
  bf16* ob = p.out + (int64_t)(r0 + row_b) * p.d + h * HD;
#pragma unroll
  for (int nt = 0; nt < 2 * C::NP_O; ++nt) {
    const int col = nt * 8 + 2 * t4;
    if (col < HD) {
      if (row_a < T) *reinterpret_cast<uint32_t*>(oa + col) = pack_bf16x2(o_acc[nt][0] * inv0, o_acc[nt][1] * inv0);
      if (row_b < T) *reinterpret_cast<uint32_t*>(ob + col) = pack_bf16x2(o_acc[nt][2] * inv1, o_acc[nt][3] * inv1);
    }
  }
}

}  // namespace serenc

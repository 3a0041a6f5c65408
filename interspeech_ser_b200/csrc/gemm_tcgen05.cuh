// Persistent, warp-specialised bf16 GEMM for sm_100a:  D[M, N] = epilogue(A[M, K] * W[N, K]^T)
//
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> 4..8-stage shared-memory ring -> tcgen05.mma (one issuing
//   thread, fp32 accumulators in TMEM, double-buffered so the epilogue of tile i overlaps the mainloop of
//   tile i+1) -> tcgen05.ld epilogue with fused bias / exact GELU / fp32 residual add / bf16 down-cast and an
//   optional output row map.
//
// Every Linear / Conv1d on the embedding-extraction path is an implicit GEMM over channels-last activations;
// no im2col buffer exists. The K loop is decomposed into (tap, 64-channel block):
//     K-block kb -> tap = kb / a_kpt, cc = kb % a_kpt
//     A tile     =  map[tap % a_stride] at (column g*a_group_stride + cc*64, row m0 + tap / a_stride)
//   * nn.Linear (q/k/v/out/fc1/fc2/feature projection): one tap, a_stride = 1.
//   * strided Conv1d (conv1..6 of the wav2vec2 feature encoder, HF modeling_wavlm.py:703-727; Whisper
//     conv1/conv2, modeling_whisper.py:567-571): out row m reads input rows s*m + tap. For s = 2 the input is
//     seen through two tensor maps with row stride 2*C (even rows / odd rows), so a tap is just a row offset.
//   * grouped positional Conv1d (k=128, groups=16; modeling_wavlm.py:48-90): s = 1, 128 taps, the group picks
//     the column block. W is packed [N, tap*C_pad + c] to match.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..7 = epilogue (warp w reads TMEM lanes 32*(w%4) .. +31, one accumulator row per thread).
#pragma once
#include "common.cuh"

namespace serenc {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 256;

struct GemmParams {
  int64_t M;         // rows of A (tensor-map extent)
  int n_per_group;   // valid output columns per group (multiple of 8)
  int groups;        // 1 for dense layers, 16 for the positional conv
  int num_kb;        // K / 64 (K tail, if any, is zero-filled by TMA on both operands)
  int tiles_m;
  int tiles_n;       // per group
  int a_kpt;         // K-blocks (of 64 channels) per tap; == num_kb for a plain Linear
  int a_stride;      // temporal stride of the conv (1 or 2) = number of A tensor maps in use
  int a_group_stride;  // column offset per group in A
  const float* bias;     // [groups * n_per_group] or nullptr
  const float* resid;    // fp32, same indexing as out_f32, or nullptr
  float* out_f32;        // may be nullptr
  int64_t ld_f32;
  bf16* out_bf16;        // may be nullptr
  int64_t ld_bf16;
  const int32_t* rowmap; // [M] -> output row, <0 = skip; nullptr = identity
  int act;               // 0 = none, 1 = exact GELU (applied before the residual add)
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // 128 / 256 / 512: power of two
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                         const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);  // every epilogue thread arrives
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_m = p.tiles_n * p.groups;
  const int num_tiles = p.tiles_m * tiles_per_m;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_t = tile / tiles_per_m;
        const int rem = tile - m_t * tiles_per_m;
        const int g = rem / p.tiles_n;
        const int n_t = rem - g * p.tiles_n;
        const int m0 = m_t * GEMM_BM;
        const int wrow0 = g * p.n_per_group + n_t * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int tap = kb / p.a_kpt;
          const int cc = kb - tap * p.a_kpt;
          const int acol = g * p.a_group_stride + cc * GEMM_BK;
          if (p.a_stride == 2) {
            tma_load_2d(sA + stage * Cfg::A_BYTES, (tap & 1) ? &tmA1 : &tmA0, &full_bar[stage], acol, m0 + (tap >> 1));
          } else {
            tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA0, &full_bar[stage], acol, m0 + tap);
          }
          tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * GEMM_BK, wrow0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // +32 B per UMMA_K step inside the 128 B swizzle row: start-address field advances by 2
            umma_bf16_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                         (uint32_t)((kb | k) != 0));
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    const int ew = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_t = tile / tiles_per_m;
      const int rem = tile - m_t * tiles_per_m;
      const int g = rem / p.tiles_n;
      const int n_t = rem - g * p.tiles_n;
      const int64_t row = (int64_t)m_t * GEMM_BM + ew * 32 + lane;
      const int n0 = n_t * BN;

      int64_t orow = -1;
      if (row < p.M) orow = p.rowmap ? (int64_t)p.rowmap[row] : row;
      const bool row_ok = orow >= 0;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);

#pragma unroll 1
      for (int c = 0; c < BN / 16; ++c) {
        const int col0 = n0 + c * 16;
        if (col0 >= p.n_per_group) break;  // warp-uniform
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_row + (uint32_t)(c * 16), r);
        tmem_ld_wait();
        if (row_ok) {
          const int gcol = g * p.n_per_group + col0;
          const int nvalid = min(16, p.n_per_group - col0);  // 8 or 16
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              if (j < nvalid) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + gcol + j));
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
          }
          if (p.resid) {
            const float* rp = p.resid + orow * p.ld_f32 + gcol;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              if (j < nvalid) {
                const float4 q = *reinterpret_cast<const float4*>(rp + j);
                v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
              }
            }
          }
          if (p.out_f32) {
            float* op = p.out_f32 + orow * p.ld_f32 + gcol;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              if (j < nvalid) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          }
          if (p.out_bf16) {
            bf16* op = p.out_bf16 + orow * p.ld_bf16 + gcol;
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              if (j < nvalid) {
                uint4 u;
                u.x = pack_bf16x2(v[j], v[j + 1]);
                u.y = pack_bf16x2(v[j + 2], v[j + 3]);
                u.z = pack_bf16x2(v[j + 4], v[j + 5]);
                u.w = pack_bf16x2(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(op + j) = u;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace serenc

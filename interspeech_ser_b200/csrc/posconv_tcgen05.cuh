// Positional convolution embedding (grouped Conv1d, k = 128, stride 1) as a slab-reuse implicit GEMM on tcgen05.
// HF *PositionalConvEmbedding (modeling_wavlm.py:37-90, modeling_wav2vec2.py:326-379, modular_hubert.py:40-87).
//
// The generic implicit GEMM (gemm_tcgen05.cuh) treats every tap as its own K-block and reloads a 128-row A tile per
// tap - but for a stride-1 convolution the tile of tap t+1 is the tile of tap t shifted by ONE row. At 128 taps and
// 64-channel groups that is 128x redundant L2 -> shared-memory traffic, and the layer ran at the L2's bandwidth
// (~10 TB/s, 320 TFLOP/s). Here one CTA owns 256 output rows of one group:
//
//   * the activation SLAB  [256 + taps - 1 rows] x [cg_pad channels]  is loaded ONCE (TMA, 128B swizzle, 64-channel
//     panels); tap t, row half hf reads it through a UMMA descriptor whose start address is simply shifted by
//     (128 hf + t) rows = (128 hf + t) * 128 B. Measured (tools/diag_posconv.py): the 128B swizzle is a function of
//     the absolute shared-memory address on both the TMA and the UMMA side, so a 128B-aligned (not 1024B-aligned)
//     start address needs NO base-offset field - setting (address >> 7) & 7 there gives wrong data;
//   * only the weights stream: one [BN x 64] K-block per (tap, channel block) through a TMA ring, each used for BOTH
//     128-row halves (two accumulators in TMEM), so W traffic per output row is halved as well;
//   * accumulators are double-buffered in TMEM (2 x 2 x BN columns); the fused epilogue (bias + exact GELU + fp32
//     residual add, gap-layout -> packed row map) is the generic one;
//   * groups narrower than their padding (HuBERT-xlarge: 80 channels in a 128-channel panel pair, 80 outputs in a 128-column
//     tile) issue only the K = 16 steps that hold real channels (5 of 8 per tap) and N = 80 MMAs, and the weight box of a
//     K-block holds 80 rows instead of 128: at these shapes the kernel is bound by MMA issue and by the weight stream out
//     of L2 (4 MB per 256-row tile with full boxes).
#pragma once
#include "gemm_tcgen05.cuh"

namespace serenc {

constexpr int PC_BM = 256;          // output rows per tile (two 128-row accumulators)

struct PosConvCfg {
  int taps, kpt;        // taps; 64-channel panels per tap (cg_pad / 64)
  int slab_rows;        // 256 + taps - 1, rounded up to a multiple of 128 (TMA boxes of 128 rows)
  int slab_bufs;        // 1 or 2
  int cg;               // input channels per group that are not padding (80 of 128 for HuBERT-xlarge): the K = 16 steps of a
                        // panel that hold only zero padding are not issued
  int n_mma;            // output columns per MMA: n_per_group rounded up to 16 (80 instead of the 128-column tile)
};

template <int BN>
struct PosConvSmem {
  static constexpr int W_BYTES = BN * GEMM_BK * 2;
  static constexpr int W_STAGES = BN == 64 ? 8 : 4;
  static constexpr int TMEM_COLS = 4 * BN;   // 2 buffers x 2 row halves x BN columns (256 / 512)
  static size_t bytes(const PosConvCfg& c) {
    return (size_t)c.slab_bufs * c.kpt * c.slab_rows * 128 + (size_t)W_STAGES * W_BYTES + GEMM_STAGING_BYTES + 256 + 1024;
  }
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
posconv_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p,
                       const PosConvCfg cfg) {
  using S = PosConvSmem<BN>;
  constexpr int PC_W_STAGES = S::W_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const int panel_bytes = cfg.slab_rows * 128;
  const int slab_bytes = cfg.kpt * panel_bytes;
  uint8_t* sSlab = smem;                                         // [slab_bufs][kpt][slab_rows][128 B]
  uint8_t* sW = sSlab + cfg.slab_bufs * slab_bytes;              // [PC_W_STAGES][BN][128 B]
  float* staging = reinterpret_cast<float*>(sW + PC_W_STAGES * S::W_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(staging) + GEMM_STAGING_BYTES);
  uint64_t* w_full = bars;                        // [PC_W_STAGES]
  uint64_t* w_empty = w_full + PC_W_STAGES;       // [PC_W_STAGES]
  uint64_t* slab_full = w_empty + PC_W_STAGES;    // [2]
  uint64_t* slab_empty = slab_full + 2;           // [2]  commit after the tile's last MMA
  uint64_t* tfull_bar = slab_empty + 2;           // [2]  accumulators (both halves) ready
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]  accumulators drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == GEMM_WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == GEMM_WARP_MMA && lane == 0) {
    for (int i = 0; i < PC_W_STAGES; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&slab_full[i], 1);
      mbar_init(&slab_empty[i], 1);
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], GEMM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == GEMM_WARP_ALLOC) {
    tmem_alloc(tmem_slot, S::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.tiles_m * p.groups;   // tile -> (m_t = tile / groups, g = tile % groups)
  const int num_kb = cfg.taps * cfg.kpt;

  if (warp == GEMM_WARP_TMA) {
    // ------------------------------ TMA producer (whole warp, one elected lane issues) ------------------------------
    {
      int stage = 0;
      uint32_t phase = 0, n = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
        const int m_t = tile / p.groups, g = tile - m_t * p.groups;
        const int m0 = m_t * PC_BM;
        // slab: buffer sb is free once the MMAs of the tile that used it last have completed
        const uint32_t sb = cfg.slab_bufs == 2 ? (n & 1u) : 0u;
        const uint32_t use = cfg.slab_bufs == 2 ? (n >> 1) : n;
        mbar_wait(&slab_empty[sb], (use & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&slab_full[sb], (uint32_t)slab_bytes);
          for (int cc = 0; cc < cfg.kpt; ++cc)
            for (int r = 0; r < cfg.slab_rows; r += 128)
              tma_load_2d(sSlab + sb * slab_bytes + cc * panel_bytes + r * 128, &tmA, &slab_full[sb],
                          g * p.a_group_stride + cc * GEMM_BK, m0 + r);
        }
        __syncwarp();
        const int wrow0 = g * p.n_per_group;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&w_empty[stage], phase ^ 1u);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&w_full[stage], (uint32_t)cfg.n_mma * 128u);   // the W box holds n_mma rows, not BN
            tma_load_2d(sW + stage * S::W_BYTES, &tmW, &w_full[stage], kb * GEMM_BK, wrow0);
          }
          __syncwarp();
          if (++stage == PC_W_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == GEMM_WARP_MMA) {
    // ------------------------------ MMA issuer (whole warp, one elected lane issues) ------------------------------
    {
      const uint32_t idesc = umma_idesc_bf16(GEMM_BM, cfg.n_mma);
      const uint32_t tmem_u = warp_uniform(tmem_base);
      int stage = 0;
      uint32_t phase = 0, n = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
        const uint32_t sb = cfg.slab_bufs == 2 ? (n & 1u) : 0u;
        const uint32_t use = cfg.slab_bufs == 2 ? (n >> 1) : n;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        mbar_wait(&slab_full[sb], use & 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)(acc * 2 * BN);
        const uint32_t slab_addr = smem_u32(sSlab + sb * slab_bytes);
        int kb = 0;
        for (int t = 0; t < cfg.taps; ++t) {
          for (int cc = 0; cc < cfg.kpt; ++cc, ++kb) {
            mbar_wait(&w_full[stage], phase);
            tc_fence_after();
            const uint64_t bdesc = umma_desc_sw128(smem_u32(sW + stage * S::W_BYTES));
            // rows [128 hf + t, 128 hf + t + 128) of the slab panel: a row-shifted view
            const uint64_t adesc0 = umma_desc_sw128(slab_addr + (uint32_t)(cc * panel_bytes + t * 128));
            const uint64_t adesc1 = umma_desc_sw128(slab_addr + (uint32_t)(cc * panel_bytes + (GEMM_BM + t) * 128));
            const int ks = min(GEMM_BK / 16, (cfg.cg - cc * GEMM_BK + 15) / 16);   // K steps of this panel with real channels (>= 1)
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < GEMM_BK / 16; ++k)
                if (k < ks) umma_bf16_ss(d_tmem, adesc0 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
#pragma unroll
              for (int k = 0; k < GEMM_BK / 16; ++k)
                if (k < ks) umma_bf16_ss(d_tmem + (uint32_t)BN, adesc1 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
              umma_commit(&w_empty[stage]);
            }
            __syncwarp();
            if (++stage == PC_W_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        if (elect_one_sync()) {
          umma_commit(&slab_empty[sb]);
          umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < GEMM_EPI_WARPS) {
    // ------------------------------ epilogue: warp -> (row half, TMEM lane quarter) ------------------------------
    const int ew = warp & 3;
    const int hf = warp >> 2;
    float* st = staging + warp * (32 * GEMM_ST_LD);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_t = tile / p.groups, g = tile - m_t * p.groups;
      const uint32_t t_addr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * 2 * BN + hf * BN);
      gemm_epilogue_warp<BN, false>(p, st, t_addr, (int64_t)m_t * PC_BM + hf * GEMM_BM + ew * 32, g, 0, lane, &tfull_bar[acc], acc_phase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == GEMM_WARP_ALLOC) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

}  // namespace serenc

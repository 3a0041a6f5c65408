// Persistent, warp-specialised bf16 GEMMs for sm_100a:  D[M, N] = epilogue(A[M, K] * W[N, K]^T)
//
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> multi-stage shared-memory ring -> tcgen05.mma (one issuing
//   thread, fp32 accumulators in TMEM, double-buffered so the epilogue of tile i overlaps the mainloop of
//   tile i+1) -> tcgen05.ld epilogue with fused bias / exact GELU / fp32 residual add / bf16 down-cast and an
//   optional output row map.
//
// Two kernels share the operand addressing and the epilogue:
//   * gemm_bf16_tcgen05_2cta_kernel — the workhorse. A CTA pair (cluster of 2 = one TPC) computes a 256 x 256
//     tile with tcgen05.mma.cta_group::2: each CTA stages only its 128 rows of A and its 128 rows of W per
//     K-block (32 KB instead of the 48 KB a 128 x 256 single-CTA tile needs), which is what lifts the kernel off
//     the L2->SM bandwidth ceiling measured with the single-CTA version (profiles/r01_notes.md).
//   * gemm_bf16_tcgen05_kernel<BN> (BN = 64 | 128) — single-CTA 128 x BN tiles for narrow outputs (the grouped
//     positional conv has 64..120 columns per group) and for problems too small to fill 74 CTA pairs.
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
// Warp roles (384 threads): warps 0..7 = epilogue, warp 8 = TMA producer, warp 9 = MMA issuer, warp 10 = TMEM
// allocator. (The single-thread producer / issuer roles sit on the HIGHEST warp ids: the SM's issue arbiter
// favours higher warp ids, and a GELU-heavy epilogue warp must never delay a tcgen05.mma or TMA issue.)
// Epilogue warp w reads TMEM lanes 32*(w%4) .. +31 (one accumulator row per thread, the only
// shape tcgen05.ld offers) for its half of the tile's columns, transposes each 32x32 block through a padded
// shared-memory tile and then touches global memory with lanes running along a row: bias / GELU / residual /
// stores are all 128-byte-coalesced (the row-per-thread layout would cost 32 sectors per request).
// The CTA-pair kernel's fp32 output (the residual stream of out-projection / FC2) takes a second route: residual blocks
// in and result blocks out by TMA through swizzled shared memory, no transpose (gemm2_epilogue_tma).
// Both kernels are launched with programmatic stream serialization: set-up first, pdl_wait() before the first operand.
#pragma once
#include "common.cuh"

namespace serenc {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_WARP_TMA = 8, GEMM_WARP_MMA = 9, GEMM_WARP_ALLOC = 10;
constexpr int GEMM_ST_LD = 36;  // staging row stride in words: 16-byte aligned rows, conflict-free 128-bit accesses
constexpr int GEMM_STAGING_BYTES = GEMM_EPI_WARPS * 32 * GEMM_ST_LD * 4;  // per-warp padded 32x32 fp32 transpose tile
constexpr int GEMM2_BN = 256;     // 2-CTA kernel: tile is 256 (M, 128 per CTA) x 256 (N, W rows split 128 per CTA)
constexpr int GEMM2_STAGES = 5;

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
  long long* trace;      // debug: per-tile clock64 stamps of CTA 0 ([tile][8]); nullptr in production
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 128) ? 5 : 6;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // 128 / 256: power of two
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_STAGING_BYTES + BAR_BYTES + 1024;  // +1024: alignment slack
};

// TMA_NB = 0: epilogue through registers + a transpose staging tile. TMA_NB = 2: fp32 epilogue through TMA with two
// 32 x 32 fp32 blocks (128-byte rows, 128B swizzle) per epilogue warp in place of the staging. (Three blocks cost an operand
// stage; measured, r02o: 4 stages + 3 blocks = 779 TFLOP/s on the out-projection shape against 855 for 5 + 2, FC2 shape
// 1 166 against 1 329 - the operand ring needs its depth.)
constexpr int GEMM2_RING_BLOCK_BYTES = 32 * 32 * 4;
template <int TMA_NB>
struct Gemm2Cfg {
  static constexpr int STAGES = TMA_NB == 3 ? GEMM2_STAGES - 1 : GEMM2_STAGES;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;        // this CTA's 128 rows of A
  static constexpr int B_BYTES = (GEMM2_BN / 2) * GEMM_BK * 2; // this CTA's 128 rows of W
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;        // 32 KB
  static constexpr int TMEM_COLS = 512;                        // 2 x 256 accumulator columns
  static constexpr int EPI_BYTES = TMA_NB ? GEMM_EPI_WARPS * TMA_NB * GEMM2_RING_BLOCK_BYTES : GEMM_STAGING_BYTES;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16 + GEMM_EPI_WARPS * TMA_NB * 8;   // + one mbarrier per ring block
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;      // +1024: alignment slack
  static_assert(SMEM_BYTES <= 227 * 1024, "CTA-pair GEMM exceeds the shared memory of an SM");
};

// ------------------------------------------------------------------------------------------------
// shared pieces
// ------------------------------------------------------------------------------------------------
struct TileCoord {
  int m_t, g, n_t;
};
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile) {
  const int tiles_per_m = p.tiles_n * p.groups;
  TileCoord t;
  t.m_t = tile / tiles_per_m;
  const int rem = tile - t.m_t * tiles_per_m;
  t.g = rem / p.tiles_n;
  t.n_t = rem - t.g * p.tiles_n;
  return t;
}

// Epilogue of one warp for its 32 rows x NCOLS columns of an accumulator.
//   t_addr : TMEM address of (lane quarter, first column)   row0 : first of the warp's 32 rows
//   n0     : first column (within the group) of the warp's column range
//   ready  : mbarrier (and parity) signalled when the accumulator is complete.
//
// Measured with per-tile clock stamps (tools/trace_gemm.py): the epilogue is bound by on-chip traffic, not by
// instructions — TMEM reads (~64 B/clk/SM) plus the shared-memory transpose. Two paths keep it below the
// K = 1024 mainloop (~10 k cycles per 128 x 256 accumulator):
//   * bf16-only outputs (QKV, FC1, conv layers): bias + GELU are applied in the row-per-thread layout tcgen05.ld
//     delivers, the result is packed to bf16 and only 64 B per row go through an XOR-swizzled staging tile
//     (conflict-free 128-bit writes by row and reads by 4-lane row segments); stores are 16 B per lane.
//   * fp32 outputs (residual stream): fp32 staging with 36-word rows; the residual values of block c+1 are
//     fetched while block c is processed, the first block before the accumulator is even ready.
// In both paths the TMEM load of the next 32-column block is issued before the current block is processed.
template <int NCOLS, bool EPI_BF16>
__device__ __forceinline__ void gemm_epilogue_warp(const GemmParams& p, float* st, uint32_t t_addr, int64_t row0, int g,
                                                   int n0, int lane, uint64_t* ready, uint32_t ready_parity,
                                                   long long* trace_after_wait = nullptr) {
  constexpr int NCH = NCOLS / 32;

  if constexpr (EPI_BF16) {
    // coalesced phase: lane -> (row 8*i + lane/4, 16-byte segment lane%4 = 8 bf16 columns)
    const int sr = lane >> 2, sc = lane & 3;
    int64_t orow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = row0 + 8 * i + sr;
      orow[i] = -1;
      if (r < p.M) orow[i] = p.rowmap ? (int64_t)__ldg(p.rowmap + r) : r;
    }
    uint4* st4 = reinterpret_cast<uint4*>(st);   // [32 rows][4 chunks of 16 B], chunk index XOR-swizzled by row
    mbar_wait(ready, ready_parity);
    tc_fence_after();
    if (trace_after_wait) *trace_after_wait = clock64();
    uint32_t ra[32], rb[32];
    if (n0 < p.n_per_group) tmem_ld_32x32b_x32(t_addr, ra);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col0 = n0 + c * 32;
      if (col0 < p.n_per_group) {  // warp-uniform
        uint32_t (&r)[32] = (c & 1) ? rb : ra;
        tmem_ld_wait();
        if (c + 1 < NCH && col0 + 32 < p.n_per_group) tmem_ld_32x32b_x32(t_addr + (uint32_t)((c + 1) * 32), (c & 1) ? ra : rb);
        const int gcol0 = g * p.n_per_group + col0;
        const bool tail = (col0 + 32 > p.n_per_group);   // only multiples of 8 columns are valid (checked on the host)
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[k8 * 8 + e]);
          if (p.bias && !(tail && col0 + k8 * 8 >= p.n_per_group)) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol0 + k8 * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol0 + k8 * 8 + 4));
            unpack_f32x2(fadd2(pack_f32x2(v[0], v[1]), pack_f32x2(b0.x, b0.y)), v[0], v[1]);
            unpack_f32x2(fadd2(pack_f32x2(v[2], v[3]), pack_f32x2(b0.z, b0.w)), v[2], v[3]);
            unpack_f32x2(fadd2(pack_f32x2(v[4], v[5]), pack_f32x2(b1.x, b1.y)), v[4], v[5]);
            unpack_f32x2(fadd2(pack_f32x2(v[6], v[7]), pack_f32x2(b1.z, b1.w)), v[6], v[7]);
          }
          if (p.act == 1) {
#pragma unroll
            for (int e = 0; e < 8; e += 2) gelu_fast2(v[e], v[e + 1], v[e], v[e + 1]);
          }
          uint4 u;
          u.x = pack_bf16x2(v[0], v[1]);
          u.y = pack_bf16x2(v[2], v[3]);
          u.z = pack_bf16x2(v[4], v[5]);
          u.w = pack_bf16x2(v[6], v[7]);
          st4[lane * 4 + (k8 ^ ((lane >> 1) & 3))] = u;
        }
        __syncwarp();
        const int col = col0 + sc * 8;
        if (col < p.n_per_group) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = 8 * i + sr;
            const uint4 u = st4[rl * 4 + (sc ^ ((rl >> 1) & 3))];
            if (orow[i] >= 0) *reinterpret_cast<uint4*>(p.out_bf16 + orow[i] * p.ld_bf16 + g * p.n_per_group + col) = u;
          }
        }
        __syncwarp();
      }
    }
    return;
  } else {
  // ---------------- fp32 / residual path ----------------
  // (Round 2 measured the alternative of reading the accumulator in the 16x256b fragment shape and touching global memory
  //  straight from it, 32-byte sectors, no shared-memory transpose: parity-green but 7-12 % SLOWER on the residual
  //  GEMMs - 636 against 686 TFLOP/s at M = 25 472, N = K = 1 024 - because every LDG / STG then spreads over 8 lines.)
  const int sr = lane >> 3;       // coalesced phase: sub-row 0..3
  const int c4 = (lane & 7) * 4;  // coalesced phase: 4 consecutive columns
  int32_t orow[8];                // output rows this lane touches in the coalesced phase (< 2^31 rows)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = row0 + 4 * i + sr;
    orow[i] = -1;
    if (r < p.M) orow[i] = p.rowmap ? __ldg(p.rowmap + r) : (int32_t)r;
  }
  float4 q[8], qn[8];
  auto fetch_resid = [&](float4 (&dst)[8], int col0) {
    const int col = col0 + c4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < p.n_per_group && orow[i] >= 0)
        dst[i] = *reinterpret_cast<const float4*>(p.resid + (int64_t)orow[i] * p.ld_f32 + g * p.n_per_group + col);
    }
  };
  if (p.resid) fetch_resid(q, n0);

  mbar_wait(ready, ready_parity);
  tc_fence_after();
  if (trace_after_wait) *trace_after_wait = clock64();

#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    const int col0 = n0 + c * 32;
    if (col0 < p.n_per_group) {  // warp-uniform
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_addr + (uint32_t)(c * 32), r);
      if (p.resid && c + 1 < NCH) fetch_resid(qn, col0 + 32);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4*>(st + lane * GEMM_ST_LD + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      __syncwarp();
      const int col = col0 + c4;
      if (col < p.n_per_group) {
        const int gcol = g * p.n_per_group + col;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
        const uint64_t bv01 = pack_f32x2(bv.x, bv.y), bv23 = pack_f32x2(bv.z, bv.w);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (orow[i] < 0) continue;
          float4 v = *reinterpret_cast<const float4*>(st + (4 * i + sr) * GEMM_ST_LD + c4);
          // packed fp32x2 adds (FADD2): the whole step runs at the power cap, every issue slot saved is clock
          unpack_f32x2(fadd2(pack_f32x2(v.x, v.y), bv01), v.x, v.y);
          unpack_f32x2(fadd2(pack_f32x2(v.z, v.w), bv23), v.z, v.w);
          if (p.act == 1) {
            gelu_fast2(v.x, v.y, v.x, v.y);
            gelu_fast2(v.z, v.w, v.z, v.w);
          }
          if (p.resid) {
            unpack_f32x2(fadd2(pack_f32x2(v.x, v.y), pack_f32x2(q[i].x, q[i].y)), v.x, v.y);
            unpack_f32x2(fadd2(pack_f32x2(v.z, v.w), pack_f32x2(q[i].z, q[i].w)), v.z, v.w);
          }
          if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + (int64_t)orow[i] * p.ld_f32 + gcol) = v;
          if (p.out_bf16) {
            uint2 u;
            u.x = pack_bf16x2(v.x, v.y);
            u.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(p.out_bf16 + (int64_t)orow[i] * p.ld_bf16 + gcol) = u;
          }
        }
      }
      if (p.resid) {
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = qn[i];
      }
      __syncwarp();
    }
  }
  }
}

// ------------------------------------------------------------------------------------------------
// single-CTA kernel: 128 x BN tiles
// ------------------------------------------------------------------------------------------------
template <int BN, bool EPI_BF16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                         const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  float* staging = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + GEMM_STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == GEMM_WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
  }
  if (warp == GEMM_WARP_MMA && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], GEMM_EPI_WARPS);  // lane 0 of every epilogue warp arrives
    }
    fence_mbar_init();
  }
  if (warp == GEMM_WARP_ALLOC) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above overlapped the previous kernel's tail; A, the residual and the output are touched below
  pdl_trigger();

  const int num_tiles = p.tiles_m * p.tiles_n * p.groups;

  if (warp == GEMM_WARP_TMA) {
    // ------------------------------ TMA producer (whole warp, one elected lane issues) ------------------------------
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int m0 = tc.m_t * GEMM_BM;
        const int wrow0 = tc.g * p.n_per_group + tc.n_t * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const int tap = kb / p.a_kpt;
          const int cc = kb - tap * p.a_kpt;
          const int acol = tc.g * p.a_group_stride + cc * GEMM_BK;
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            if (p.a_stride == 2) {
              tma_load_2d(sA + stage * Cfg::A_BYTES, (tap & 1) ? &tmA1 : &tmA0, &full_bar[stage], acol, m0 + (tap >> 1));
            } else {
              tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA0, &full_bar[stage], acol, m0 + tap);
            }
            tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * GEMM_BK, wrow0);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == GEMM_WARP_MMA) {
    // ------------------------------ MMA issuer (whole warp, one elected lane issues) ------------------------------
    {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      const uint32_t tmem_u = warp_uniform(tmem_base);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              // +32 B per UMMA_K step inside the 128 B swizzle row: start-address field advances by 2
              umma_bf16_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                           (uint32_t)((kb | k) != 0));
            }
            umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one_sync()) umma_commit(&tfull_bar[acc]);  // accumulator complete
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < GEMM_EPI_WARPS) {
    // ------------------------------ epilogue ------------------------------
    const int ew = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = warp >> 2;        // which half of the tile's columns
    float* st = staging + warp * (32 * GEMM_ST_LD);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile);
      const uint32_t t_addr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
      gemm_epilogue_warp<BN / 2, EPI_BF16>(p, st, t_addr, (int64_t)tc.m_t * GEMM_BM + ew * 32, tc.g,
                                           tc.n_t * BN + half * (BN / 2), lane, &tfull_bar[acc], acc_phase);
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
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 epilogue of the CTA-pair kernel through TMA (plain Linear into the fp32 residual stream: out-projection, FC2)
// ------------------------------------------------------------------------------------------------
// The register-staged path above fetches the residual one 32-column block ahead and its loads sit in front of every block
// (three buffers would spill: 145 registers already). Here a warp owns NB 4 KB shared-memory blocks: the residual block is
// bulk-loaded by TMA NB blocks ahead - across tile boundaries, so the first two blocks of a tile arrive under its mainloop -
// each thread adds its accumulator row to its own 128-byte row of the block in place (row-per-thread is conflict-free under
// the 128B swizzle: lane l touches 16-byte chunk j ^ (l & 7)), and a TMA store writes the block back: no transpose, no
// LDG / STG, no per-row address arithmetic; rows past M and columns past N are clipped by the tensor map.
// A block is (tile, c): rows row0 .. +31 of the warp's lane quarter, columns col0 = n0 + 32 c, c = 0..3; blocks with
// col0 >= N (N tail of the last tile column) do not exist for either the prefetch or the consumer.
template <int BN, int NB>
__device__ __forceinline__ void gemm2_epilogue_tma(const GemmParams& p, const CUtensorMap* tmR, const CUtensorMap* tmO,
                                                   uint8_t* ring, uint64_t* rfull, uint32_t tmem_base, int ew, int half,
                                                   int rank, int pair, int num_pairs, int num_tiles, int lane,
                                                   uint64_t* tfull_bar, uint64_t* tempty_bar) {
  constexpr int RB = GEMM2_RING_BLOCK_BYTES;
  const bool has_res = p.resid != nullptr;
  auto coords = [&](int tile, int c, int& row0, int& col0) {
    const TileCoord tc = decode_tile(p, tile);
    row0 = tc.m_t * (2 * GEMM_BM) + rank * GEMM_BM + ew * 32;
    col0 = tc.n_t * BN + half * (BN / 2) + c * 32;
  };
  // prefetch cursor over the existing blocks of this warp
  int pt = pair, pc = -1;
  auto advance = [&]() {
    for (;;) {
      if (++pc == 4) {
        pc = 0;
        pt += num_pairs;
      }
      if (pt >= num_tiles) return;
      int r0, c0;
      coords(pt, pc, r0, c0);
      if (c0 < p.n_per_group) return;
    }
  };
  int ibuf = 0;   // ring slot of the next block to load
  auto issue_load = [&]() {   // whole warp
    if (pt < num_tiles) {
      if (has_res) {
        int r0, c0;
        coords(pt, pc, r0, c0);
        uint64_t* bar = &rfull[ibuf];
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar, RB);
          tma_load_2d(ring + ibuf * RB, tmR, bar, c0, r0);
        }
        __syncwarp();
      }
      if (++ibuf == NB) ibuf = 0;
      advance();
    }
  };
  if (pair < num_tiles) advance();
#pragma unroll
  for (int i = 0; i < NB; ++i) issue_load();

  int dbuf = 0;          // ring slot of the block being consumed
  uint32_t dphase = 0;   // its mbarrier parity
  int acc = 0;
  uint32_t acc_phase = 0;
  int it = 0;
  const uint32_t sw = (uint32_t)(lane & 7);
  for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
    const bool tracing = p.trace && blockIdx.x == 0 && threadIdx.x == 0;
    if (tracing) p.trace[it * 8 + 4] = clock64();
    const uint32_t t_addr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
    mbar_wait(&tfull_bar[acc], acc_phase);
    tc_fence_after();
    if (tracing) p.trace[it * 8 + 5] = clock64();
    bool released = false;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      int row0, col0;
      coords(tile, c, row0, col0);
      if (col0 >= p.n_per_group) break;   // warp-uniform
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_addr + (uint32_t)(c * 32), r);
      if (has_res) mbar_wait(&rfull[dbuf], dphase);
      uint8_t* blk = ring + dbuf * RB;
      uint8_t* row = blk + lane * 128;
      tmem_ld_wait();
      if (c == 3 || col0 + 32 >= p.n_per_group) {   // the tile's last block is in registers: hand the accumulator back now
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));  // leader's barrier
        released = true;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4* slot = reinterpret_cast<float4*>(row + (((uint32_t)j ^ sw) << 4));
        float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                               __uint_as_float(r[4 * j + 3]));
        if (p.bias) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));
          unpack_f32x2(fadd2(pack_f32x2(v.x, v.y), pack_f32x2(bv.x, bv.y)), v.x, v.y);
          unpack_f32x2(fadd2(pack_f32x2(v.z, v.w), pack_f32x2(bv.z, bv.w)), v.z, v.w);
        }
        if (p.act == 1) {
          gelu_fast2(v.x, v.y, v.x, v.y);
          gelu_fast2(v.z, v.w, v.z, v.w);
        }
        if (has_res) {
          const float4 q = *slot;
          unpack_f32x2(fadd2(pack_f32x2(v.x, v.y), pack_f32x2(q.x, q.y)), v.x, v.y);
          unpack_f32x2(fadd2(pack_f32x2(v.z, v.w), pack_f32x2(q.z, q.w)), v.z, v.w);
        }
        *slot = v;
      }
      fence_proxy_async_smem();   // generic-proxy writes of the block -> visible to the TMA store
      __syncwarp();
      if (elect_one_sync()) {
        tma_store_2d(tmO, blk, col0, row0);
        bulk_commit_group();
        bulk_wait_read_all();     // the slot is refilled right away (ibuf == dbuf here: NB loads are always outstanding)
      }
      __syncwarp();
      if (++dbuf == NB) {
        dbuf = 0;
        dphase ^= 1u;
      }
      issue_load();
    }
    if (!released) {   // a warp whose column half lies past N in this tile column
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
    }
    if (tracing) p.trace[it * 8 + 6] = clock64();
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1u;
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair kernel: 256 x 256 tiles, tcgen05.mma.cta_group::2
// ------------------------------------------------------------------------------------------------
// EPI_BF16: bf16-only outputs. TMA_NB > 0 (fp32 outputs only): the epilogue moves the residual / output blocks with TMA
// (gemm2_epilogue_tma; tmR / tmO are fp32 maps with 32 x 32 boxes, tmR unused without a residual).
template <bool EPI_BF16, int TMA_NB = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                              const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmR,
                              const __grid_constant__ CUtensorMap tmO, const GemmParams p) {
  static_assert(!(EPI_BF16 && TMA_NB), "the TMA epilogue is the fp32 one");
  constexpr bool EPI_TMA = TMA_NB > 0;
  using Cfg = Gemm2Cfg<TMA_NB>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BN = GEMM2_BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  float* staging = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);   // EPI_TMA: the block ring (1 024-aligned)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  [[maybe_unused]] uint64_t* ring_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);   // EPI_TMA: [8 warps][TMA_NB blocks]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = warp_uniform(cluster_ctarank());   // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == GEMM_WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    if constexpr (EPI_TMA) {
      tma_prefetch_desc(&tmR);
      tma_prefetch_desc(&tmO);
    }
  }
  if (warp == GEMM_WARP_MMA && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's producer arrives (expect_tx covers both CTAs' TMA bytes)
      mbar_init(&empty_bar[i], 1);  // multicast tcgen05.commit from the leader
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);                    // multicast tcgen05.commit from the leader
      mbar_init(&tempty_bar[i], 2 * GEMM_EPI_WARPS);  // every epilogue warp of both CTAs (used in the leader only)
    }
    if constexpr (EPI_TMA) {
      for (int i = 0; i < TMA_NB * GEMM_EPI_WARPS; ++i) mbar_init(&ring_bar[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == GEMM_WARP_ALLOC) {
    tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();   // peer barriers initialised + both allocations done before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above overlapped the previous kernel's tail; A, the residual and the output are touched below
  pdl_trigger();

  const int num_tiles = p.tiles_m * p.tiles_n * p.groups;   // tiles_m counts 256-row tiles here

  if (warp == GEMM_WARP_TMA) {
    // ------------------------------ TMA producer (both CTAs; whole warp, one elected lane issues) ------------------------------
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t leader_bar0 = warp_uniform(mapa_shared(smem_u32(&full_bar[0]), 0));   // leader's full_bar[0]
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const TileCoord tc = decode_tile(p, tile);
        const int m0 = tc.m_t * (2 * GEMM_BM) + (int)rank * GEMM_BM;
        const int wrow0 = tc.g * p.n_per_group + tc.n_t * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const uint32_t leader_bar = leader_bar0 + (uint32_t)(stage * 8);
          const int tap = kb / p.a_kpt;
          const int cc = kb - tap * p.a_kpt;
          const int acol = tc.g * p.a_group_stride + cc * GEMM_BK;
          if (elect_one_sync()) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            if (p.a_stride == 2) {
              tma_load_2d_2cta(sA + stage * Cfg::A_BYTES, (tap & 1) ? &tmA1 : &tmA0, leader_bar, acol, m0 + (tap >> 1));
            } else {
              tma_load_2d_2cta(sA + stage * Cfg::A_BYTES, &tmA0, leader_bar, acol, m0 + tap);
            }
            tma_load_2d_2cta(sB + stage * Cfg::B_BYTES, &tmB, leader_bar, kb * GEMM_BK, wrow0);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == GEMM_WARP_MMA) {
    // ------------------------------ MMA issuer (leader CTA only; whole warp, one elected lane issues) ------------------------------
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, BN);
      const uint32_t tmem_u = warp_uniform(tmem_base);
      const bool tracing = p.trace && blockIdx.x == 0 && lane == 0;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        if (tracing) p.trace[it * 8 + 0] = clock64();
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        if (tracing) p.trace[it * 8 + 1] = clock64();
        const uint32_t d_tmem = tmem_u + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == 0 && tracing) p.trace[it * 8 + 2] = clock64();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              umma_bf16_ss_2cta(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                (uint32_t)((kb | k) != 0));
            }
            umma_commit_2cta(&empty_bar[stage]);  // frees this smem slot in BOTH CTAs
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one_sync()) umma_commit_2cta(&tfull_bar[acc]);  // accumulator complete, signalled to both CTAs' epilogues
        __syncwarp();
        if (tracing) p.trace[it * 8 + 3] = clock64();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < GEMM_EPI_WARPS) {
    // ------------------------------ epilogue (both CTAs, own 128 rows) ------------------------------
    const int ew = warp & 3;
    const int half = warp >> 2;
    if constexpr (EPI_TMA) {
      gemm2_epilogue_tma<BN, TMA_NB>(p, &tmR, &tmO, reinterpret_cast<uint8_t*>(staging) + warp * (TMA_NB * GEMM2_RING_BLOCK_BYTES),
                             ring_bar + TMA_NB * warp, tmem_base, ew, half, (int)rank, pair, num_pairs, num_tiles, lane, tfull_bar,
                             tempty_bar);
    } else {
    float* st = staging + warp * (32 * GEMM_ST_LD);
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const TileCoord tc = decode_tile(p, tile);
      const uint32_t t_addr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
      const int64_t row0 = (int64_t)tc.m_t * (2 * GEMM_BM) + (int64_t)rank * GEMM_BM + ew * 32;
      if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[it * 8 + 4] = clock64();
      gemm_epilogue_warp<BN / 2, EPI_BF16>(p, st, t_addr, row0, tc.g, tc.n_t * BN + half * (BN / 2), lane, &tfull_bar[acc],
                                           acc_phase, (p.trace && blockIdx.x == 0 && threadIdx.x == 0) ? p.trace + it * 8 + 5 : nullptr);
      tc_fence_before();
      __syncwarp();
      if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[it * 8 + 6] = clock64();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));  // leader's barrier
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();   // no CTA of the pair may exit while the other can still signal it
  if (warp == GEMM_WARP_ALLOC) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace serenc

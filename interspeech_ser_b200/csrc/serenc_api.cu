// libserenc: C ABI + host-side runtime of the B200-native speech-SSL encoder forward.
// See include/serenc.h for the contract each entry point replaces in the reference.
#include "../../include/serenc.h"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <array>
#include <atomic>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#ifdef SERENC_AB_ARMS   // development build only: the first-generation mma.sync attention and the getenv kernel switches
#include "attention.cuh"
#endif
#include "attention_params.cuh"
#include "attention_tc.cuh"
#include "attention_tc_wide.cuh"
#ifdef SERENC_AB_ARMS
#include "attention_tc_v3.cuh"
#include "attention_tc_split.cuh"
#endif
#include "common.cuh"
#include "frontend_norm.cuh"
#include "conv0_tc.cuh"
#include "gemm_tcgen05.cuh"
#include "posconv_tcgen05.cuh"
#include "logmel.cuh"

using namespace serenc;

// =================================================================================================
// errors
// =================================================================================================
namespace {
thread_local char g_err[1024] = "";
thread_local int g_cuda_err = 0;   // last cudaError_t seen by SERENC_CUDA_OK on this thread
}
namespace serenc {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void note_cuda_error(int e) { g_cuda_err = e; }
}  // namespace serenc

#define SERENC_FAIL(code, ...)   \
  do {                           \
    serenc::set_error(__VA_ARGS__); \
    return (code);               \
  } while (0)

// =================================================================================================
// handle
// =================================================================================================
struct LayerW {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  bf16* w_qkv;  // [3d, d]
  float* b_qkv; // [3d]
  bf16* w_o;    // [d, d]
  float* b_o;
  bf16* w_fc1;  // [ffn, d]
  float* b_fc1;
  bf16* w_fc2;  // [d, ffn]
  float* b_fc2;
  float *gru_w, *gru_b, *gru_const;  // WavLM: [8, hd], [8], [H]
  float *gru_w2, *gru_b2;            // WavLM: the same, summed over the two groups of four outputs: [2, hd], [2]
};

static const int W2V_K[7] = {10, 3, 3, 3, 3, 2, 2};
static const int W2V_S[7] = {5, 2, 2, 2, 2, 2, 2};

struct serenc_handle {
  serenc_config cfg;
  int device = 0;
  int num_sms = 148;
  int head_dim = 64;
  bool finalized = false;
  std::vector<void*> allocs;
  std::set<std::string> loaded;
  std::vector<LayerW> L;

  // wav2vec2-family front end
  float* conv0_w = nullptr;       // [512, 10]
  bf16* conv_w[7] = {nullptr};    // 1..6: [512, k*512] tap-major
  float* conv_b[7] = {nullptr};   // [512] (zeros when conv_bias = 0)
  float* conv_g[7] = {nullptr};
  float* conv_be[7] = {nullptr};
  float *fp_g = nullptr, *fp_be = nullptr, *fp_b = nullptr;
  bf16* fp_w = nullptr;           // [d, 512]
  bf16* pos_w = nullptr;          // [d, taps * cg_pad]
  float* pos_b = nullptr;
  int pos_cg = 0, pos_cg_pad = 0;
  float *fin_g = nullptr, *fin_b = nullptr;
  std::vector<float> rel_embed_host;  // [num_buckets, H]
  float* btab = nullptr;              // [H, 2*WAVLM_MAXD-1]

  // whisper front end
  bf16 *wc1 = nullptr, *wc2 = nullptr;  // [d, 3*n_mels_pad], [d, 3*d]
  float *bc1 = nullptr, *bc2 = nullptr;
  float* pos_emb = nullptr;             // [1500, d]
  int mel_pad = 128;                    // n_mels rounded up to 64
  std::vector<float> mel_filters_host;  // [201, n_mels]
  float *hann = nullptr, *mel_w = nullptr;
  float2* twid = nullptr;   // log-mel twiddle matrix [99][256] (logmel.cuh)
  int32_t *mel_ptr = nullptr, *mel_bin = nullptr;

  std::mutex mu;
  std::map<std::array<uint64_t, 8>, CUtensorMap> tmaps;

  // sticky-fault state (include/serenc.h "kernel faults are sticky")
  std::atomic<int> poisoned{0};
  char poison_msg[512] = "";

  // text encoder (SERENC_ARCH_TEXT): embedding tables, fp32
  float *emb_word = nullptr, *emb_pos = nullptr, *emb_type = nullptr;

  // launch accounting / optional per-class device timing (bench.py's roofline numbers)
  std::atomic<long long> launches{0};
  bool prof = false;
  long long* gemm_trace = nullptr;  // debug: device buffer for per-tile clock stamps of the CTA-pair GEMM (serenc_debug_gemm_trace)
  // A/B switches: constant false in the production build; read from the environment only with -DSERENC_AB_ARMS
  bool force_1cta = false;        // bypass the CTA-pair GEMM
  bool pdl = true;                // programmatic dependent launch for the layer-loop kernels (launch_k)
  bool gemm_no_tma_epilogue = false;  // fp32 epilogue of the CTA-pair GEMM through registers (the round-1 path)
  bool no_posconv_slab = false;   // positional conv through the generic implicit GEMM
  bool force_mma_sync_attn = false;  // attention on the mma.sync kernel
  int attn_deep64 = 0;               // bias-free head_dim-64 attention on the deep-pipelined kernel
  bool attn_split = false;           // head_dim-64 attention on the two-threads-per-row kernel (attention_tc_split.cuh)
  bool attn_v3 = false;              // attention on the one-pass / Q-in-TMEM kernel (attention_tc_v3.cuh)
  int attn_variant = 0;              // bit 0: control-warp waits suspend instead of spinning; bit 1: every 4th pair of exponentials on the FMA pipe; 4: every 2nd + suspend
  int max_smem = 227 * 1024;      // opt-in dynamic shared memory per CTA
  struct ProfRec { int cls; cudaEvent_t a, b; double flops, bytes; int n; };
  std::vector<ProfRec> recs;
};

namespace {
// Counts kernel launches; when profiling is on, brackets them with CUDA events on the launching stream.
struct ProfScope {
  serenc_handle* h;
  cudaStream_t st;
  int idx = -1;
  ProfScope(serenc_handle* h_, int cls, int n_launches, double flops, double bytes, cudaStream_t st_) : h(h_), st(st_) {
    h->launches += n_launches;
    if (h->prof) {
      serenc_handle::ProfRec r;
      r.cls = cls; r.flops = flops; r.bytes = bytes; r.n = n_launches;
      if (cudaEventCreate(&r.a) == cudaSuccess && cudaEventCreate(&r.b) == cudaSuccess) {
        cudaEventRecord(r.a, st);
        std::lock_guard<std::mutex> lk(h->mu);
        idx = (int)h->recs.size();
        h->recs.push_back(r);
      }
    }
  }
  ~ProfScope() {
    if (idx >= 0) {
      cudaEvent_t b;
      { std::lock_guard<std::mutex> lk(h->mu); b = h->recs[idx].b; }
      cudaEventRecord(b, st);
    }
  }
};
}  // namespace

namespace {

// CUDA errors that invalidate the context: every later CUDA call returns them again, so the handle is useless.
bool cuda_error_is_sticky(int e) {
  switch ((cudaError_t)e) {
    case cudaErrorIllegalAddress: case cudaErrorLaunchFailure: case cudaErrorIllegalInstruction:
    case cudaErrorMisalignedAddress: case cudaErrorInvalidAddressSpace: case cudaErrorInvalidPc:
    case cudaErrorHardwareStackError: case cudaErrorAssert: case cudaErrorLaunchTimeout:
    case cudaErrorECCUncorrectable: case cudaErrorUnknown: case cudaErrorDevicesUnavailable:
      return true;
    default:
      return false;
  }
}

// Every entry point that takes a handle runs through this: a poisoned handle answers immediately, and a sticky CUDA
// error seen by the body poisons it.
template <typename F>
int guarded(serenc_handle* h, F&& body) {
  if (h && h->poisoned.load()) {
    serenc::set_error("handle is poisoned by an earlier CUDA fault: %s", h->poison_msg);
    return SERENC_ERR_CUDA;
  }
  g_cuda_err = 0;
  const int rc = body();
  if (rc == SERENC_ERR_CUDA && h && cuda_error_is_sticky(g_cuda_err)) {
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->poisoned.load()) {
      snprintf(h->poison_msg, sizeof(h->poison_msg), "%s", g_err);
      h->poisoned.store(1);
    }
  }
  return rc;
}

template <typename T>
int dev_alloc(serenc_handle* h, T** out, size_t count, bool zero = true) {
  void* p = nullptr;
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = 16;
  SERENC_CUDA_OK(cudaMalloc(&p, bytes));
  if (zero) SERENC_CUDA_OK(cudaMemset(p, 0, bytes));
  h->allocs.push_back(p);
  *out = reinterpret_cast<T*>(p);
  return 0;
}

int upload_f32(float* dst, const float* src, size_t n) {
  SERENC_CUDA_OK(cudaMemcpy(dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}
int upload_bf16(bf16* dst, const std::vector<bf16>& src) {
  SERENC_CUDA_OK(cudaMemcpy(dst, src.data(), src.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  return 0;
}

int64_t shape_numel(const int64_t* shape, int ndim) {
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  return n;
}

// ---------------------------------------------------------------------------------------------
// tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// rank-2 bf16 map: dims {cols, rows}, row stride in bytes, box {64, box_rows}, 128B swizzle, zero OOB fill
int get_tmap(serenc_handle* h, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
             uint32_t box_rows, CUtensorMap* out) {
  std::array<uint64_t, 8> key = {reinterpret_cast<uint64_t>(base), cols, rows, row_stride_bytes, box_rows, 0, 0, 0};
  {
    std::lock_guard<std::mutex> lk(h->mu);
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) SERENC_FAIL(SERENC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (rows == 0) rows = 1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    SERENC_FAIL(SERENC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): base=%p cols=%llu rows=%llu stride=%llu box=%u",
                (int)r, base, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_stride_bytes,
                box_rows);
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->tmaps.size() > 4096) h->tmaps.clear();
  h->tmaps[key] = *out;
  return 0;
}

// rank-2 fp32 map for the TMA epilogue of the CTA-pair GEMM: dims {cols, rows}, box {32, 32} (128-byte rows), 128B swizzle
int get_tmap_f32(serenc_handle* h, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes, CUtensorMap* out) {
  std::array<uint64_t, 8> key = {reinterpret_cast<uint64_t>(base), cols, rows, row_stride_bytes, 32, 0, 0, 4};
  {
    std::lock_guard<std::mutex> lk(h->mu);
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) SERENC_FAIL(SERENC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (rows == 0) rows = 1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    SERENC_FAIL(SERENC_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed (%d): base=%p cols=%llu rows=%llu stride=%llu", (int)r, base,
                (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_stride_bytes);
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->tmaps.size() > 4096) h->tmaps.clear();
  h->tmaps[key] = *out;
  return 0;
}

// rank-3 bf16 map over a packed [rows, 3 * heads * hd] q|k|v buffer seen as {hd, 3 * heads, rows}: box {64, 1, box_rows},
// 128B swizzle. A box that starts at column 64 of a head runs past the head's extent and is zero-filled there.
int get_tmap_heads(serenc_handle* h, const void* base, uint64_t hd, uint64_t slots, uint64_t rows, uint64_t row_stride_bytes,
                   uint32_t box_rows, CUtensorMap* out) {
  std::array<uint64_t, 8> key = {reinterpret_cast<uint64_t>(base), hd, rows, row_stride_bytes, box_rows, slots, 3, 0};
  {
    std::lock_guard<std::mutex> lk(h->mu);
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) SERENC_FAIL(SERENC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (rows == 0) rows = 1;
  cuuint64_t dims[3] = {hd, slots, rows};
  cuuint64_t strides[2] = {hd * 2, row_stride_bytes};
  cuuint32_t box[3] = {64, 1, box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    SERENC_FAIL(SERENC_ERR_CUDA, "cuTensorMapEncodeTiled (rank 3) failed (%d): base=%p hd=%llu slots=%llu rows=%llu stride=%llu box=%u",
                (int)r, base, (unsigned long long)hd, (unsigned long long)slots, (unsigned long long)rows,
                (unsigned long long)row_stride_bytes, box_rows);
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->tmaps.size() > 4096) h->tmaps.clear();
  h->tmaps[key] = *out;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Launch with programmatic stream serialization (common.cuh, pdl_wait): ONLY for kernels that execute pdl_wait() before
// they touch anything an earlier kernel of the stream wrote or still reads - LayerNorm, the GEMMs, the attention kernels.
// Every other kernel is a plain launch, which waits for all of them to complete.
// ---------------------------------------------------------------------------------------------
template <typename... KArgs, typename... Args>
int launch_k(serenc_handle* h, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = h->pdl ? 1 : 0;
  SERENC_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// GEMM launch
// ---------------------------------------------------------------------------------------------
struct GemmCall {
  const bf16* A = nullptr;
  int64_t a_cols = 0;         // channels per input row (tensor-map inner extent)
  int64_t a_rows = 0;         // input rows available
  int64_t a_ld = 0;           // elements between consecutive input rows
  int a_stride = 1;           // temporal stride (1 | 2)
  int a_kpt = 0;              // K-blocks per tap
  int taps = 1;
  int a_group_stride = 0;
  int a_cg_valid = 0;         // grouped conv: input channels per group that are not padding (0: all a_kpt * 64)
  int64_t M = 0;              // output rows
  const bf16* W = nullptr;    // [w_rows, w_k], K index = tap * (a_kpt*64) + channel
  int64_t w_rows = 0;
  int64_t w_k = 0;            // elements per weight row (0: taps * a_kpt * 64)
  int n_per_group = 0;
  int groups = 1;
  const float* bias = nullptr;
  const float* resid = nullptr;
  float* out_f32 = nullptr;
  int64_t ld_f32 = 0;
  bf16* out_bf16 = nullptr;
  int64_t ld_bf16 = 0;
  const int32_t* rowmap = nullptr;
  int act = 0;
  int prof_cls = SERENC_PROF_GEMM_LINEAR;
  double alg_flops = -1.0;    // algorithmic FLOPs (valid rows, unpadded K); < 0: 2*M*N*K of the launch
};

template <int BN>
int launch_gemm_bn(serenc_handle* h, const GemmCall& c, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  GemmParams p;
  p.M = c.M;
  p.n_per_group = c.n_per_group;
  p.groups = c.groups;
  p.num_kb = c.taps * c.a_kpt;
  p.tiles_m = (int)ceil_div64(c.M, GEMM_BM);
  p.tiles_n = ceil_div(c.n_per_group, BN);
  p.a_kpt = c.a_kpt;
  p.a_stride = c.a_stride;
  p.a_group_stride = c.a_group_stride;
  p.bias = c.bias;
  p.resid = c.resid;
  p.out_f32 = c.out_f32;
  p.ld_f32 = c.ld_f32;
  p.out_bf16 = c.out_bf16;
  p.ld_bf16 = c.ld_bf16;
  p.rowmap = c.rowmap;
  p.act = c.act;
  p.trace = nullptr;

  CUtensorMap tA0, tA1, tB;
  const uint64_t s = (uint64_t)c.a_stride;
  // parity-p view of the input: rows p, p+s, p+2s, ...
  SERENC_TRY(get_tmap(h, c.A, (uint64_t)c.a_cols, (uint64_t)ceil_div64(c.a_rows, (int64_t)s), (uint64_t)c.a_ld * s * 2,
                      GEMM_BM, &tA0));
  if (c.a_stride == 2) {
    SERENC_TRY(get_tmap(h, c.A + c.a_ld, (uint64_t)c.a_cols, (uint64_t)(c.a_rows / 2 > 0 ? c.a_rows / 2 : 1),
                        (uint64_t)c.a_ld * 4, GEMM_BM, &tA1));
  } else {
    tA1 = tA0;
  }
  const uint64_t wk = c.w_k > 0 ? (uint64_t)c.w_k : (uint64_t)p.num_kb * GEMM_BK;
  SERENC_TRY(get_tmap(h, c.W, wk, (uint64_t)c.w_rows, wk * 2, BN, &tB));

  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n * p.groups;
  if (tiles <= 0) return 0;
  const int grid = (int)(tiles < h->num_sms ? tiles : h->num_sms);
  const double flops = c.alg_flops >= 0 ? c.alg_flops : 2.0 * (double)c.M * c.n_per_group * c.groups * p.num_kb * GEMM_BK;
  const double bytes = 2.0 * ((double)c.M * c.a_cols / (c.groups > 1 ? 1 : 1) + (double)c.w_rows * p.num_kb * GEMM_BK) +
                       (double)c.M * c.n_per_group * c.groups * ((c.out_f32 ? 4 : 0) + (c.out_bf16 ? 2 : 0) + (c.resid ? 4 : 0));
  ProfScope ps(h, c.prof_cls, 1, flops, bytes, st);
  const bool epi_bf16 = c.out_bf16 && !c.out_f32 && !c.resid;
  if (epi_bf16)
    SERENC_TRY(launch_k(h, gemm_bf16_tcgen05_kernel<BN, true>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, st, tA0, tA1, tB, p));
  else
    SERENC_TRY(launch_k(h, gemm_bf16_tcgen05_kernel<BN, false>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, st, tA0, tA1, tB, p));
  return 0;
}

// CTA-pair kernel: 256 x 256 tiles, launched as clusters of 2
int launch_gemm_2cta(serenc_handle* h, const GemmCall& c, cudaStream_t st) {
  using Cfg = Gemm2Cfg<0>;
  GemmParams p;
  p.M = c.M;
  p.n_per_group = c.n_per_group;
  p.groups = c.groups;
  p.num_kb = c.taps * c.a_kpt;
  p.tiles_m = (int)ceil_div64(c.M, 2 * GEMM_BM);
  p.tiles_n = ceil_div(c.n_per_group, GEMM2_BN);
  p.a_kpt = c.a_kpt;
  p.a_stride = c.a_stride;
  p.a_group_stride = c.a_group_stride;
  p.bias = c.bias;
  p.resid = c.resid;
  p.out_f32 = c.out_f32;
  p.ld_f32 = c.ld_f32;
  p.out_bf16 = c.out_bf16;
  p.ld_bf16 = c.ld_bf16;
  p.rowmap = c.rowmap;
  p.act = c.act;
  p.trace = h->gemm_trace;

  CUtensorMap tA0, tA1, tB;
  const uint64_t s = (uint64_t)c.a_stride;
  SERENC_TRY(get_tmap(h, c.A, (uint64_t)c.a_cols, (uint64_t)ceil_div64(c.a_rows, (int64_t)s), (uint64_t)c.a_ld * s * 2,
                      GEMM_BM, &tA0));
  if (c.a_stride == 2) {
    SERENC_TRY(get_tmap(h, c.A + c.a_ld, (uint64_t)c.a_cols, (uint64_t)(c.a_rows / 2 > 0 ? c.a_rows / 2 : 1),
                        (uint64_t)c.a_ld * 4, GEMM_BM, &tA1));
  } else {
    tA1 = tA0;
  }
  const uint64_t wk = c.w_k > 0 ? (uint64_t)c.w_k : (uint64_t)p.num_kb * GEMM_BK;
  SERENC_TRY(get_tmap(h, c.W, wk, (uint64_t)c.w_rows, wk * 2, GEMM2_BN / 2, &tB));

  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n * p.groups;
  if (tiles <= 0) return 0;
  const int max_pairs = h->num_sms / 2;
  const int pairs = (int)(tiles < max_pairs ? tiles : max_pairs);
  const double flops = c.alg_flops >= 0 ? c.alg_flops : 2.0 * (double)c.M * c.n_per_group * c.groups * p.num_kb * GEMM_BK;
  const double bytes = 2.0 * ((double)c.M * c.a_cols + (double)c.w_rows * p.num_kb * GEMM_BK) +
                       (double)c.M * c.n_per_group * c.groups * ((c.out_f32 ? 4 : 0) + (c.out_bf16 ? 2 : 0) + (c.resid ? 4 : 0));
  ProfScope ps(h, c.prof_cls, 1, flops, bytes, st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see launch_k
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = h->pdl ? 2 : 1;
  const bool epi_bf16 = c.out_bf16 && !c.out_f32 && !c.resid;
  // plain Linear into the fp32 stream (out-projection, FC2): residual / output blocks move by TMA
  const bool epi_tma = !epi_bf16 && c.out_f32 && !c.out_bf16 && !c.rowmap && c.groups == 1 && c.n_per_group % 32 == 0 &&
                       c.ld_f32 % 4 == 0 && (reinterpret_cast<uintptr_t>(c.out_f32) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(c.resid) & 15) == 0 && !h->gemm_no_tma_epilogue;
  if (epi_bf16) {
    SERENC_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_2cta_kernel<true, 0>, tA0, tA1, tB, tB, tB, p));
  } else if (epi_tma) {
    CUtensorMap tR, tO;
    SERENC_TRY(get_tmap_f32(h, c.out_f32, (uint64_t)c.n_per_group, (uint64_t)c.M, (uint64_t)c.ld_f32 * 4, &tO));
    if (c.resid) SERENC_TRY(get_tmap_f32(h, c.resid, (uint64_t)c.n_per_group, (uint64_t)c.M, (uint64_t)c.ld_f32 * 4, &tR));
    else tR = tO;
    cfg.dynamicSmemBytes = Gemm2Cfg<2>::SMEM_BYTES;
    SERENC_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_2cta_kernel<false, 2>, tA0, tA1, tB, tR, tO, p));
  } else {
    SERENC_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_2cta_kernel<false, 0>, tA0, tA1, tB, tB, tB, p));
  }
  return 0;
}

// stride-1 grouped convolution with many taps (the positional conv embedding): slab-reuse kernel
template <int BN>
int launch_posconv_bn(serenc_handle* h, const GemmCall& c, cudaStream_t st) {
  using S = PosConvSmem<BN>;
  PosConvCfg cfg;
  cfg.taps = c.taps;
  cfg.kpt = c.a_kpt;
  cfg.slab_rows = ((PC_BM + c.taps - 1 + 127) / 128) * 128;
  cfg.slab_bufs = 2;
  cfg.cg = (c.a_cg_valid > 0 && c.a_cg_valid <= c.a_kpt * GEMM_BK) ? c.a_cg_valid : c.a_kpt * GEMM_BK;
  cfg.n_mma = (c.n_per_group + 15) / 16 * 16;
  if (cfg.n_mma > BN) cfg.n_mma = BN;
  if (S::bytes(cfg) > (size_t)h->max_smem) cfg.slab_bufs = 1;
  if (S::bytes(cfg) > (size_t)h->max_smem) return -1000;   // caller falls back to the generic implicit GEMM
  GemmParams p;
  p.M = c.M;
  p.n_per_group = c.n_per_group;
  p.groups = c.groups;
  p.num_kb = c.taps * c.a_kpt;
  p.tiles_m = (int)ceil_div64(c.M, PC_BM);
  p.tiles_n = 1;
  p.a_kpt = c.a_kpt;
  p.a_stride = 1;
  p.a_group_stride = c.a_group_stride;
  p.bias = c.bias; p.resid = c.resid; p.out_f32 = c.out_f32; p.ld_f32 = c.ld_f32;
  p.out_bf16 = c.out_bf16; p.ld_bf16 = c.ld_bf16; p.rowmap = c.rowmap; p.act = c.act; p.trace = nullptr;
  CUtensorMap tA, tW;
  SERENC_TRY(get_tmap(h, c.A, (uint64_t)c.a_cols, (uint64_t)c.a_rows, (uint64_t)c.a_ld * 2, GEMM_BM, &tA));
  const uint64_t wk = c.w_k > 0 ? (uint64_t)c.w_k : (uint64_t)p.num_kb * GEMM_BK;
  SERENC_TRY(get_tmap(h, c.W, wk, (uint64_t)c.w_rows, wk * 2, (uint32_t)cfg.n_mma, &tW));
  const int64_t tiles = (int64_t)p.tiles_m * p.groups;
  if (tiles <= 0) return 0;
  const int grid = (int)(tiles < h->num_sms ? tiles : h->num_sms);
  const double flops = c.alg_flops >= 0 ? c.alg_flops : 2.0 * (double)c.M * c.n_per_group * c.groups * p.num_kb * GEMM_BK;
  const double bytes = 2.0 * ((double)c.M * c.a_cols + (double)c.w_rows * p.num_kb * GEMM_BK) +
                       (double)c.M * c.n_per_group * c.groups * ((c.out_f32 ? 4 : 0) + (c.out_bf16 ? 2 : 0) + (c.resid ? 4 : 0));
  // (the opt-in to > 48 KB of dynamic shared memory is per device: serenc_create sets it for both instantiations)
  ProfScope ps(h, c.prof_cls, 1, flops, bytes, st);
  posconv_tcgen05_kernel<BN><<<grid, GEMM_THREADS, S::bytes(cfg), st>>>(tA, tW, p, cfg);
  SERENC_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_gemm(serenc_handle* h, const GemmCall& c, cudaStream_t st) {
  if (c.n_per_group % 8 != 0) SERENC_FAIL(SERENC_ERR_INVALID, "gemm: output width %d not a multiple of 8", c.n_per_group);
  if ((c.a_ld % 8) != 0 || (c.a_group_stride % 8) != 0 || (c.w_k % 8) != 0)
    SERENC_FAIL(SERENC_ERR_INVALID, "gemm: A leading dimension must be a multiple of 8 elements");
  if (c.groups > 1 && c.taps >= 8 && c.a_stride == 1 && c.n_per_group <= 128 && c.out_f32 && !h->no_posconv_slab) {
    const int rc = c.n_per_group <= 64 ? launch_posconv_bn<64>(h, c, st) : launch_posconv_bn<128>(h, c, st);
    if (rc != -1000) return rc;
  }
  if (c.n_per_group <= 64) return launch_gemm_bn<64>(h, c, st);
  // wide outputs with enough work for the CTA pairs: 256 x 256 tiles on tcgen05.mma.cta_group::2
  const int64_t tiles2 = ceil_div64(c.M, 2 * GEMM_BM) * ceil_div(c.n_per_group, GEMM2_BN) * c.groups;
  if (c.n_per_group >= 192 && tiles2 >= h->num_sms / 4 && !h->force_1cta) return launch_gemm_2cta(h, c, st);
  return launch_gemm_bn<128>(h, c, st);
}

// plain Linear: out = A[M, K] W[N, K]^T
GemmCall linear_call(const bf16* A, int64_t M, int K, const bf16* W, int N) {
  GemmCall c;
  c.A = A;
  c.a_cols = K;
  c.a_rows = M;
  c.a_ld = K;
  c.a_stride = 1;
  c.a_kpt = ceil_div(K, GEMM_BK);
  c.taps = 1;
  c.M = M;
  c.W = W;
  c.w_rows = N;
  c.w_k = K;
  c.n_per_group = N;
  c.groups = 1;
  return c;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm launch
// ---------------------------------------------------------------------------------------------
template <typename TIn, typename TOut, bool GELU>
int launch_ln_t(serenc_handle* h, const TIn* in, int64_t ld_in, TOut* out, int64_t ld_out, const float* g, const float* b, int64_t rows,
                int cols, const int32_t* in_map, const int32_t* out_map, float eps, cudaStream_t st, bf16* out2 = nullptr,
                const LnGate& gate = LnGate()) {
  if (rows <= 0) return 0;
  const int nv = cols / 128;
  const dim3 block(256);
  ProfScope ps(h, SERENC_PROF_LAYERNORM, 1, 0.0, (double)rows * cols * (sizeof(TIn) + sizeof(TOut)), st);
#define SERENC_LN_CASE(NV)                                                                                         \
  case NV:                                                                                                         \
    SERENC_TRY(launch_k(h, layernorm_rows_kernel<NV, TIn, TOut, GELU>, dim3((unsigned)ceil_div64(rows, 8 * LnRows<NV>::RPW)), \
                        block, 0, st, in, ld_in, out, ld_out, g, b, rows, in_map, out_map, eps, out2, (int64_t)cols, gate)); \
    break;
  switch (nv) {
    SERENC_LN_CASE(1)
    SERENC_LN_CASE(2)
    SERENC_LN_CASE(3)
    SERENC_LN_CASE(4)
    SERENC_LN_CASE(5)
    SERENC_LN_CASE(6)
    SERENC_LN_CASE(8)
    SERENC_LN_CASE(10)
    SERENC_LN_CASE(12)
    SERENC_LN_CASE(15)
    SERENC_LN_CASE(16)
    default:
      SERENC_FAIL(SERENC_ERR_INVALID, "layernorm: unsupported width %d", cols);
  }
#undef SERENC_LN_CASE
  SERENC_CUDA_OK(cudaGetLastError());
  return 0;
}

bool ln_width_ok(int cols) {
  if (cols % 128) return false;
  const int nv = cols / 128;
  return nv == 1 || nv == 2 || nv == 3 || nv == 4 || nv == 5 || nv == 6 || nv == 8 || nv == 10 || nv == 12 || nv == 15 ||
         nv == 16;
}

// ---------------------------------------------------------------------------------------------
// attention launch
// ---------------------------------------------------------------------------------------------
#ifdef SERENC_AB_ARMS
template <int HD>
int launch_attn_hd(const AttnParams& p, bool wavlm, int tmax, int heads, int batch, cudaStream_t st) {
  const dim3 grid(ceil_div(tmax, ATT_BM), heads, batch), block(ATT_THREADS);
  if (wavlm)
    attention_fwd_kernel<HD, true><<<grid, block, AttnCfg<HD>::SMEM_BYTES, st>>>(p);
  else
    attention_fwd_kernel<HD, false><<<grid, block, AttnCfg<HD>::SMEM_BYTES, st>>>(p);
  SERENC_CUDA_OK(cudaGetLastError());
  return 0;
}
#endif
int launch_attn(serenc_handle* h, const AttnParams& p, bool wavlm, int tmax, int batch, int64_t sum_rows, double alg_flops,
                cudaStream_t st) {
  if (batch <= 0 || tmax <= 0) return 0;
  ProfScope ps(h, SERENC_PROF_ATTENTION, 1, alg_flops, 0.0, st);
  AttnParams pt = p;
  pt.trace = h->gemm_trace;
  pt.heads = h->cfg.heads; pt.batch = batch;
  const dim3 grid(ceil_div(tmax, FA_BM), h->cfg.heads, batch), block(FA_THREADS);
  auto ensure_gate = [&]() -> int {   // stand-alone entry (serenc_op_attention); the encoder stacks fuse the gate into the LayerNorm
    if (!p.gate) SERENC_FAIL(SERENC_ERR_STATE, "attention: no gate buffer");
    if (!p.gate_ready) {
      const int64_t nthr = sum_rows * h->cfg.heads;
      wavlm_gate_kernel<<<(unsigned)ceil_div64(nthr, 256), 256, 0, st>>>(p.hln, sum_rows, p.d, h->cfg.heads, p.gru_w, p.gru_b, p.gru_const, p.gate);
      SERENC_CUDA_OK(cudaGetLastError());
      h->launches += 1;
    }
    return 0;
  };
#ifdef SERENC_AB_ARMS   // measured alternatives, development build only (profiles/r02_notes.md)
  if (h->force_mma_sync_attn) {
    switch (h->head_dim) {
      case 64: return launch_attn_hd<64>(p, wavlm, tmax, h->cfg.heads, batch, st);
      case 80: return launch_attn_hd<80>(p, wavlm, tmax, h->cfg.heads, batch, st);
      case 120: return launch_attn_hd<120>(p, wavlm, tmax, h->cfg.heads, batch, st);
    }
  }
  if (h->head_dim == 64 && h->attn_split) {   // two softmax threads per query row, Q and P in TMEM (attention_tc_split.cuh)
    CUtensorMap tmkv;
    SERENC_TRY(get_tmap(h, p.qkv, (uint64_t)p.ld_qkv, (uint64_t)sum_rows, (uint64_t)p.ld_qkv * 2, FA_BN, &tmkv));
    const size_t smem = fs_smem_bytes(wavlm, tmax);
    if (wavlm) {
      SERENC_TRY(ensure_gate());
      attention_tc_split_kernel<true><<<grid, dim3(FS_THREADS), smem, st>>>(tmkv, pt);
    } else {
      attention_tc_split_kernel<false><<<grid, dim3(FS_THREADS), smem, st>>>(tmkv, pt);
    }
    SERENC_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (h->attn_v3 && (h->head_dim == 64 || !wavlm)) {   // one-pass softmax, Q in TMEM, 2 CTAs/SM for every head_dim (attention_tc_v3.cuh)
    CUtensorMap tmq, tmkv;
    SERENC_TRY(get_tmap_heads(h, p.qkv, (uint64_t)h->head_dim, (uint64_t)3 * h->cfg.heads, (uint64_t)sum_rows, (uint64_t)p.ld_qkv * 2, FA_BM, &tmq));
    SERENC_TRY(get_tmap_heads(h, p.qkv, (uint64_t)h->head_dim, (uint64_t)3 * h->cfg.heads, (uint64_t)sum_rows, (uint64_t)p.ld_qkv * 2, FA_BN, &tmkv));
    if (h->head_dim == 64) {
      const size_t smem = fa3_smem_bytes<64>(wavlm, tmax);
      if (wavlm) {
        SERENC_TRY(ensure_gate());
        attention_tc_v3_kernel<64, true><<<grid, block, smem, st>>>(tmq, tmkv, pt);
      } else if (h->attn_variant & 2) {
        attention_tc_v3_kernel<64, false, 4><<<grid, block, smem, st>>>(tmq, tmkv, pt);
      } else {
        attention_tc_v3_kernel<64, false><<<grid, block, smem, st>>>(tmq, tmkv, pt);
      }
    } else if (h->head_dim == 80) {
      attention_tc_v3_kernel<80, false><<<grid, block, Fa3Cfg<80>::SMEM_FIXED, st>>>(tmq, tmkv, pt);
    } else {
      attention_tc_v3_kernel<120, false><<<grid, block, Fa3Cfg<120>::SMEM_FIXED, st>>>(tmq, tmkv, pt);
    }
    SERENC_CUDA_OK(cudaGetLastError());
    return 0;
  }
#endif
  if (h->head_dim == 64 && (wavlm || !h->attn_deep64)) {
    // head_dim 64 (attention_tc.cuh): Q/K/V tiles through one tensor map over the packed [sum_T, 3d] projection buffer
    CUtensorMap tmq, tmkv;
    SERENC_TRY(get_tmap(h, p.qkv, (uint64_t)p.ld_qkv, (uint64_t)sum_rows, (uint64_t)p.ld_qkv * 2, FA_BM, &tmq));
    SERENC_TRY(get_tmap(h, p.qkv, (uint64_t)p.ld_qkv, (uint64_t)sum_rows, (uint64_t)p.ld_qkv * 2, FA_BN, &tmkv));
    const size_t smem = fa_smem_bytes(wavlm, tmax);
    if (smem > FA_SMEM_LIMIT) SERENC_FAIL(SERENC_ERR_INVALID, "attention: utterance of %d frames exceeds the bias-window capacity", tmax);
    if (wavlm) SERENC_TRY(ensure_gate());
    switch (h->attn_variant) {
#ifdef SERENC_AB_ARMS
      case 1:
        if (wavlm) attention_tc_kernel<true, 0, true><<<grid, block, smem, st>>>(tmq, tmkv, pt);
        else attention_tc_kernel<false, 0, true><<<grid, block, smem, st>>>(tmq, tmkv, pt);
        break;
      case 2:
        if (wavlm) attention_tc_kernel<true, 4, false><<<grid, block, smem, st>>>(tmq, tmkv, pt);
        else attention_tc_kernel<false, 4, false><<<grid, block, smem, st>>>(tmq, tmkv, pt);
        break;
      case 3:
        if (wavlm) attention_tc_kernel<true, 4, true><<<grid, block, smem, st>>>(tmq, tmkv, pt);
        else attention_tc_kernel<false, 4, true><<<grid, block, smem, st>>>(tmq, tmkv, pt);
        break;
      case 4:
        if (wavlm) attention_tc_kernel<true, 2, true><<<grid, block, smem, st>>>(tmq, tmkv, pt);
        else attention_tc_kernel<false, 2, true><<<grid, block, smem, st>>>(tmq, tmkv, pt);
        break;
#endif
      default:
        if (wavlm) SERENC_TRY(launch_k(h, attention_tc_kernel<true>, grid, block, smem, st, tmq, tmkv, pt));
        else SERENC_TRY(launch_k(h, attention_tc_kernel<false>, grid, block, smem, st, tmq, tmkv, pt));
    }
    SERENC_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if ((h->head_dim == 80 || h->head_dim == 120 || h->head_dim == 64) && !wavlm) {
    // wide heads (attention_tc_wide.cuh): rank-3 {head_dim, 3 * heads, rows} maps (zero fill past the head's last column)
    CUtensorMap tmq, tmkv;
    SERENC_TRY(get_tmap_heads(h, p.qkv, (uint64_t)h->head_dim, (uint64_t)3 * h->cfg.heads, (uint64_t)sum_rows, (uint64_t)p.ld_qkv * 2, FA_BM, &tmq));
    SERENC_TRY(get_tmap_heads(h, p.qkv, (uint64_t)h->head_dim, (uint64_t)3 * h->cfg.heads, (uint64_t)sum_rows, (uint64_t)p.ld_qkv * 2, FA_BN, &tmkv));
#ifdef SERENC_AB_ARMS
    if (h->attn_variant & 1) {
      if (h->head_dim == 80) attention_tc_wide_kernel<80, true><<<grid, block, FawCfg<80>::SMEM_BYTES, st>>>(tmq, tmkv, pt);
      else if (h->head_dim == 64) attention_tc_wide_kernel<64, true><<<grid, block, FawCfg<64>::SMEM_BYTES, st>>>(tmq, tmkv, pt);
      else attention_tc_wide_kernel<120, true><<<grid, block, FawCfg<120>::SMEM_BYTES, st>>>(tmq, tmkv, pt);
      SERENC_CUDA_OK(cudaGetLastError());
      return 0;
    }
#endif
    if (h->head_dim == 80)
      SERENC_TRY(launch_k(h, attention_tc_wide_kernel<80>, grid, block, FawCfg<80>::SMEM_BYTES, st, tmq, tmkv, pt));
    else if (h->head_dim == 64)
      SERENC_TRY(launch_k(h, attention_tc_wide_kernel<64>, grid, block, FawCfg<64>::SMEM_BYTES, st, tmq, tmkv, pt));
    else
      SERENC_TRY(launch_k(h, attention_tc_wide_kernel<120>, grid, block, FawCfg<120>::SMEM_BYTES, st, tmq, tmkv, pt));
    SERENC_CUDA_OK(cudaGetLastError());
    return 0;
  }
  SERENC_FAIL(SERENC_ERR_INVALID, "attention: head_dim %d %s the gated relative position bias is not supported", h->head_dim,
              wavlm ? "with" : "without");
}

#ifdef SERENC_AB_ARMS
template <int HD>
int set_attn_attr() {
  SERENC_CUDA_OK(cudaFuncSetAttribute(attention_fwd_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      AttnCfg<HD>::SMEM_BYTES));
  SERENC_CUDA_OK(cudaFuncSetAttribute(attention_fwd_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      AttnCfg<HD>::SMEM_BYTES));
  return 0;
}
#endif

// ---------------------------------------------------------------------------------------------
// workspace carving
// ---------------------------------------------------------------------------------------------
struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* p) : base(reinterpret_cast<uint8_t*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  size_t used() const { return (off + 255) & ~size_t(255); }
};

// row maps built on the device from the per-utterance tables (keeps the per-call H2D copy tiny)
__global__ void w2v_plan_kernel(const int32_t* __restrict__ frame_off, const int32_t* __restrict__ r6, int batch,
                                int pad, int64_t sumT, int64_t mpos, int32_t* __restrict__ fp_gather,
                                int32_t* __restrict__ gap_row, int32_t* __restrict__ pos_rowmap) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < sumT) {
    int lo = 0, hi = batch - 1;  // largest b with frame_off[b] <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (frame_off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    fp_gather[i] = r6[lo] + (int32_t)(i - frame_off[lo]);
    gap_row[i] = (int32_t)i + pad * (lo + 1);
  }
  if (i < mpos) {
    const int64_t q = i + pad;  // gapped row this GEMM row produces
    int lo = 0, hi = batch - 1;  // largest b with gapped start <= q
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if ((int64_t)frame_off[mid] + (int64_t)pad * (mid + 1) <= q) lo = mid; else hi = mid - 1;
    }
    const int64_t t = q - ((int64_t)frame_off[lo] + (int64_t)pad * (lo + 1));
    const int64_t n = frame_off[lo + 1] - frame_off[lo];
    pos_rowmap[i] = (t >= 0 && t < n) ? (int32_t)(frame_off[lo] + t) : -1;
  }
}

__global__ void whisper_plan_kernel(int batch, int32_t* __restrict__ map1 /*[B*3002]*/, int32_t* __restrict__ map2 /*[B*1501]*/) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)batch * 3002) {
    const int t = (int)(i % 3002);
    map1[i] = t < 3000 ? (int32_t)(i + 1) : -1;
  }
  if (i < (int64_t)batch * 1501) {
    const int b = (int)(i / 1501), t = (int)(i % 1501);
    map2[i] = t < 1500 ? b * 1500 + t : -1;
  }
}

int check_ready(serenc_handle* h, int arch) {
  if (!h) SERENC_FAIL(SERENC_ERR_INVALID, "null handle");
  if (!h->finalized) SERENC_FAIL(SERENC_ERR_STATE, "handle not finalized (call serenc_finalize after loading tensors)");
  if (h->poisoned.load()) SERENC_FAIL(SERENC_ERR_CUDA, "handle is poisoned by an earlier CUDA fault: %s", h->poison_msg);
  if (arch >= 0 && h->cfg.arch != arch) SERENC_FAIL(SERENC_ERR_INVALID, "entry point does not match the handle's architecture");
  SERENC_CUDA_OK(cudaSetDevice(h->device));
  return 0;
}

}  // namespace

// =================================================================================================
// pure host helpers (exported; testable without a GPU)
// =================================================================================================
extern "C" int64_t serenc_w2v_num_frames(int64_t n) {
  for (int i = 0; i < 7; ++i) {
    if (n < W2V_K[i]) return 0;
    n = (n - W2V_K[i]) / W2V_S[i] + 1;
  }
  return n;
}

// WavLMAttention._relative_positions_bucket (HF modeling_wavlm.py:253-271), delta = key - query
extern "C" int serenc_wavlm_bucket(int delta, int num_buckets, int max_distance) {
  const int nb = num_buckets / 2;
  int bucket = delta > 0 ? nb : 0;
  const int a = delta < 0 ? -delta : delta;
  const int max_exact = nb / 2;
  if (a < max_exact) return bucket + a;
  // HF evaluates this in fp32
  float v = logf((float)a / (float)max_exact);
  v = v / (float)log((double)max_distance / (double)max_exact);
  v = v * (float)(nb - max_exact);
  int large = max_exact + (int)v;
  if (large > nb - 1) large = nb - 1;
  return bucket + large;
}

extern "C" const char* serenc_last_error(void) { return g_err; }
extern "C" const char* serenc_version(void) { return "serenc 0.1 (sm_100a; tcgen05 GEMM + attention)"; }

extern "C" int serenc_debug_gemm_trace(serenc_handle* h, void* dev_buf) {
  if (!h) SERENC_FAIL(SERENC_ERR_INVALID, "null handle");
  h->gemm_trace = reinterpret_cast<long long*>(dev_buf);
  return 0;
}

extern "C" int serenc_is_poisoned(const serenc_handle* h) { return h && h->poisoned.load() ? 1 : 0; }

extern "C" int serenc_sync(serenc_handle* h, void* stream) {
  if (!h) SERENC_FAIL(SERENC_ERR_INVALID, "null handle");
  return guarded(h, [&]() -> int {
    SERENC_CUDA_OK(cudaSetDevice(h->device));
    SERENC_CUDA_OK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
    SERENC_CUDA_OK(cudaGetLastError());
    return 0;
  });
}

extern "C" int64_t serenc_launch_count(const serenc_handle* h) { return h ? (int64_t)h->launches.load() : 0; }

extern "C" int serenc_set_profiling(serenc_handle* h, int enable) {
  if (!h) SERENC_FAIL(SERENC_ERR_INVALID, "null handle");
  SERENC_CUDA_OK(cudaSetDevice(h->device));
  SERENC_CUDA_OK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(h->mu);
  for (auto& r : h->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  h->recs.clear();
  h->prof = enable != 0;
  return 0;
}

extern "C" int serenc_get_profile(serenc_handle* h, int n_classes, double* ms, double* flops, double* bytes, int64_t* launches) {
  if (!h || !ms || !flops || !bytes || !launches) SERENC_FAIL(SERENC_ERR_INVALID, "null argument");
  SERENC_CUDA_OK(cudaSetDevice(h->device));
  SERENC_CUDA_OK(cudaDeviceSynchronize());
  for (int i = 0; i < n_classes; ++i) { ms[i] = flops[i] = bytes[i] = 0.0; launches[i] = 0; }
  std::lock_guard<std::mutex> lk(h->mu);
  for (auto& r : h->recs) {
    if (r.cls < 0 || r.cls >= n_classes) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
    ms[r.cls] += t; flops[r.cls] += r.flops; bytes[r.cls] += r.bytes; launches[r.cls] += r.n;
  }
  return 0;
}

// =================================================================================================
// lifecycle
// =================================================================================================
extern "C" int serenc_create(const serenc_config* cfg, int device, serenc_handle** out) {
  if (!cfg || !out) SERENC_FAIL(SERENC_ERR_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    SERENC_FAIL(SERENC_ERR_NO_DEVICE, "no CUDA device visible: libserenc has no CPU fallback");
  if (device < 0 || device >= ndev) SERENC_FAIL(SERENC_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
  cudaDeviceProp prop;
  SERENC_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    SERENC_FAIL(SERENC_ERR_NO_DEVICE, "device %d is sm_%d%d; libserenc is built for sm_100a only", device, prop.major, prop.minor);
  SERENC_CUDA_OK(cudaSetDevice(device));

  const int d = cfg->hidden;
  if (cfg->heads <= 0 || d % cfg->heads) SERENC_FAIL(SERENC_ERR_INVALID, "hidden %d not divisible by heads %d", d, cfg->heads);
  const int hd = d / cfg->heads;
  if (hd != 64 && hd != 80 && hd != 120) SERENC_FAIL(SERENC_ERR_INVALID, "head_dim %d unsupported (64, 80, 120)", hd);
  if (!ln_width_ok(d)) SERENC_FAIL(SERENC_ERR_INVALID, "hidden size %d unsupported", d);
  if (cfg->ffn % 64 || cfg->layers < 1 || cfg->layers > 63) SERENC_FAIL(SERENC_ERR_INVALID, "unsupported ffn/layers");

  serenc_handle* h = new serenc_handle();
  h->cfg = *cfg;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->head_dim = hd;
#ifdef SERENC_AB_ARMS
  { const char* e = getenv("SERENC_FORCE_1CTA"); h->force_1cta = e && e[0] == '1'; }
  { const char* e = getenv("SERENC_NO_PDL"); h->pdl = !(e && e[0] == '1'); }
  { const char* e = getenv("SERENC_GEMM_NO_TMA_EPI"); h->gemm_no_tma_epilogue = e && e[0] == '1'; }
  { const char* e = getenv("SERENC_NO_POSCONV_SLAB"); h->no_posconv_slab = e && e[0] == '1'; }
  { const char* e = getenv("SERENC_ATTN_MMA_SYNC"); h->force_mma_sync_attn = e && e[0] == '1'; }
  { const char* e = getenv("SERENC_ATTN_DEEP64"); if (e) h->attn_deep64 = e[0] == '1'; }
  { const char* e = getenv("SERENC_ATTN_SPLIT"); h->attn_split = e && e[0] == '1'; }
  { const char* e = getenv("SERENC_ATTN_V3"); h->attn_v3 = e && e[0] == '1'; }
  { const char* e = getenv("SERENC_ATTN_VARIANT"); h->attn_variant = e ? atoi(e) : 0; }
#endif
  *out = h;

  int st = 0;
  auto A = [&](auto** p, size_t n) { if (!st) st = dev_alloc(h, p, n); };
  h->L.resize(cfg->layers);
  for (auto& l : h->L) {
    A(&l.ln1_g, d); A(&l.ln1_b, d); A(&l.ln2_g, d); A(&l.ln2_b, d);
    A(&l.w_qkv, (size_t)3 * d * d); A(&l.b_qkv, 3 * d);
    A(&l.w_o, (size_t)d * d); A(&l.b_o, d);
    A(&l.w_fc1, (size_t)cfg->ffn * d); A(&l.b_fc1, cfg->ffn);
    A(&l.w_fc2, (size_t)d * cfg->ffn); A(&l.b_fc2, d);
    l.gru_w = l.gru_b = l.gru_const = l.gru_w2 = l.gru_b2 = nullptr;
    if (cfg->wavlm_rel_bias) { A(&l.gru_w, 8 * hd); A(&l.gru_b, 8); A(&l.gru_const, cfg->heads); A(&l.gru_w2, 2 * hd); A(&l.gru_b2, 2); }
  }
  A(&h->fin_g, d); A(&h->fin_b, d);
  if (cfg->arch == SERENC_ARCH_W2V) {
    const int C = cfg->conv_dim;
    if (C != CONV0_C) { st = SERENC_ERR_INVALID; serenc::set_error("conv_dim %d unsupported (512)", C); }
    if (cfg->pos_conv_groups <= 0 || d % cfg->pos_conv_groups || (d / cfg->pos_conv_groups) % 8 || cfg->pos_conv_kernel < 1) {
      st = SERENC_ERR_INVALID; serenc::set_error("unsupported positional conv configuration");
    }
    if (!st) {
      A(&h->conv0_w, (size_t)C * W2V_K[0]);
      for (int i = 0; i < 7; ++i) {
        if (i > 0) A(&h->conv_w[i], (size_t)C * W2V_K[i] * C);
        A(&h->conv_b[i], C); A(&h->conv_g[i], C); A(&h->conv_be[i], C);
      }
      A(&h->fp_g, C); A(&h->fp_be, C); A(&h->fp_w, (size_t)d * C); A(&h->fp_b, d);
      h->pos_cg = d / cfg->pos_conv_groups;
      h->pos_cg_pad = ceil_div(h->pos_cg, 64) * 64;
      A(&h->pos_w, (size_t)d * cfg->pos_conv_kernel * h->pos_cg_pad); A(&h->pos_b, d);
      if (cfg->wavlm_rel_bias) A(&h->btab, (size_t)cfg->heads * (2 * WAVLM_MAXD - 1));
    }
  } else if (cfg->arch == SERENC_ARCH_WHISPER) {
    h->mel_pad = ceil_div(cfg->n_mels, 64) * 64;
    A(&h->wc1, (size_t)d * 3 * h->mel_pad); A(&h->bc1, d);
    A(&h->wc2, (size_t)d * 3 * d); A(&h->bc2, d);
    A(&h->pos_emb, (size_t)cfg->max_source_positions * d);
    if (cfg->max_source_positions != 1500 || d % 64) { st = SERENC_ERR_INVALID; serenc::set_error("whisper: max_source_positions must be 1500"); }
  } else if (cfg->arch == SERENC_ARCH_TEXT) {
    if (hd != 64 || cfg->vocab_size < 1 || cfg->max_positions < 2 || cfg->type_vocab_size < 1 || cfg->pad_token_id < 0 ||
        cfg->pad_token_id >= cfg->max_positions) {
      st = SERENC_ERR_INVALID; serenc::set_error("text encoder: head_dim must be 64 and vocab / position / type sizes positive");
    } else {
      A(&h->emb_word, (size_t)cfg->vocab_size * d); A(&h->emb_pos, (size_t)cfg->max_positions * d); A(&h->emb_type, (size_t)cfg->type_vocab_size * d);
    }
  } else {
    st = SERENC_ERR_INVALID;
    serenc::set_error("unknown arch %d", cfg->arch);
  }
  if (!st) {
    auto attr = [&](cudaError_t e) { if (e != cudaSuccess && !st) { st = SERENC_ERR_CUDA; serenc::set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); } };
    attr(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<64>::SMEM_BYTES));
    attr(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<64>::SMEM_BYTES));
    attr(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128>::SMEM_BYTES));
    attr(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128>::SMEM_BYTES));
    attr(cudaFuncSetAttribute(gemm_bf16_tcgen05_2cta_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg<0>::SMEM_BYTES));
    attr(cudaFuncSetAttribute(gemm_bf16_tcgen05_2cta_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg<0>::SMEM_BYTES));
    attr(cudaFuncSetAttribute(gemm_bf16_tcgen05_2cta_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg<2>::SMEM_BYTES));
#ifdef SERENC_AB_ARMS
    if (!st) st = hd == 64 ? set_attn_attr<64>() : (hd == 80 ? set_attn_attr<80>() : set_attn_attr<120>());
#endif
    attr(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_LIMIT));
    attr(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_FIXED));
    attr(cudaFuncSetAttribute(attention_tc_wide_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, FawCfg<64>::SMEM_BYTES));
    attr(cudaFuncSetAttribute(attention_tc_wide_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, FawCfg<80>::SMEM_BYTES));
    attr(cudaFuncSetAttribute(attention_tc_wide_kernel<120>, cudaFuncAttributeMaxDynamicSharedMemorySize, FawCfg<120>::SMEM_BYTES));
#ifdef SERENC_AB_ARMS
    attr(cudaFuncSetAttribute(attention_tc_kernel<true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_LIMIT));
    attr(cudaFuncSetAttribute(attention_tc_kernel<false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_FIXED));
    attr(cudaFuncSetAttribute(attention_tc_kernel<true, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_LIMIT));
    attr(cudaFuncSetAttribute(attention_tc_kernel<false, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_FIXED));
    attr(cudaFuncSetAttribute(attention_tc_kernel<true, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_LIMIT));
    attr(cudaFuncSetAttribute(attention_tc_kernel<false, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_FIXED));
    attr(cudaFuncSetAttribute(attention_tc_kernel<true, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_LIMIT));
    attr(cudaFuncSetAttribute(attention_tc_kernel<false, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_FIXED));
    attr(cudaFuncSetAttribute(attention_tc_wide_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FawCfg<64>::SMEM_BYTES));
    attr(cudaFuncSetAttribute(attention_tc_wide_kernel<80, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FawCfg<80>::SMEM_BYTES));
    attr(cudaFuncSetAttribute(attention_tc_wide_kernel<120, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FawCfg<120>::SMEM_BYTES));
    attr(cudaFuncSetAttribute(attention_tc_v3_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_LIMIT / 2));
    attr(cudaFuncSetAttribute(attention_tc_v3_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fa3Cfg<64>::SMEM_FIXED));
    attr(cudaFuncSetAttribute(attention_tc_v3_kernel<64, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fa3Cfg<64>::SMEM_FIXED));
    attr(cudaFuncSetAttribute(attention_tc_v3_kernel<80, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fa3Cfg<80>::SMEM_FIXED));
    attr(cudaFuncSetAttribute(attention_tc_v3_kernel<120, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fa3Cfg<120>::SMEM_FIXED));
    attr(cudaFuncSetAttribute(attention_tc_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_LIMIT));
    attr(cudaFuncSetAttribute(attention_tc_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM_FIXED));
#endif
    attr(cudaFuncSetAttribute(conv0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C0T_SMEM));
    attr(cudaFuncSetAttribute(posconv_tcgen05_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
    attr(cudaFuncSetAttribute(posconv_tcgen05_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
  }
  if (st) {
    serenc_destroy(h);
    *out = nullptr;
    return st;
  }
  return 0;
}

extern "C" int serenc_destroy(serenc_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return 0;
}

// conv weight [Cout, Cin, k] fp32 -> [Cout, k * cin_pad] bf16, K index = tap * cin_pad + c (zero padded)
static std::vector<bf16> pack_conv(const float* w, int cout, int cin, int k, int cin_pad) {
  std::vector<bf16> o((size_t)cout * k * cin_pad, __float2bfloat16(0.f));
  for (int n = 0; n < cout; ++n)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < k; ++t)
        o[((size_t)n * k + t) * cin_pad + c] = __float2bfloat16(w[((size_t)n * cin + c) * k + t]);
  return o;
}
static std::vector<bf16> to_bf16(const float* w, size_t n) {
  std::vector<bf16> o(n);
  for (size_t i = 0; i < n; ++i) o[i] = __float2bfloat16(w[i]);
  return o;
}

extern "C" int serenc_load_tensor(serenc_handle* h, const char* name, const float* data, const int64_t* shape, int ndim) {
  if (!h || !name || !data || !shape) SERENC_FAIL(SERENC_ERR_INVALID, "null argument");
  if (h->finalized) SERENC_FAIL(SERENC_ERR_STATE, "handle already finalized");
  SERENC_CUDA_OK(cudaSetDevice(h->device));
  const serenc_config& c = h->cfg;
  const int d = c.hidden, hd = h->head_dim;
  const int64_t n = shape_numel(shape, ndim);
  const std::string nm(name);
  auto expect = [&](int64_t want) -> int {
    if (n != want) SERENC_FAIL(SERENC_ERR_INVALID, "tensor %s: %lld elements, expected %lld", name, (long long)n, (long long)want);
    return 0;
  };
  int li = -1, ci = -1;
  char sub[64] = "";
  int st = SERENC_ERR_INVALID;
  bool known = true;

  if (sscanf(name, "layer%d.%63s", &li, sub) == 2) {
    if (li < 0 || li >= c.layers) SERENC_FAIL(SERENC_ERR_INVALID, "tensor %s: layer index out of range", name);
    LayerW& l = h->L[li];
    const std::string s(sub);
    if (s == "ln1.weight") { SERENC_TRY(expect(d)); st = upload_f32(l.ln1_g, data, n); }
    else if (s == "ln1.bias") { SERENC_TRY(expect(d)); st = upload_f32(l.ln1_b, data, n); }
    else if (s == "ln2.weight") { SERENC_TRY(expect(d)); st = upload_f32(l.ln2_g, data, n); }
    else if (s == "ln2.bias") { SERENC_TRY(expect(d)); st = upload_f32(l.ln2_b, data, n); }
    else if (s == "q.weight" || s == "k.weight" || s == "v.weight") {
      SERENC_TRY(expect((int64_t)d * d));
      const int slot = s[0] == 'q' ? 0 : (s[0] == 'k' ? 1 : 2);
      st = upload_bf16(l.w_qkv + (size_t)slot * d * d, to_bf16(data, n));
    } else if (s == "q.bias" || s == "k.bias" || s == "v.bias") {
      SERENC_TRY(expect(d));
      const int slot = s[0] == 'q' ? 0 : (s[0] == 'k' ? 1 : 2);
      st = upload_f32(l.b_qkv + (size_t)slot * d, data, n);
    } else if (s == "o.weight") { SERENC_TRY(expect((int64_t)d * d)); st = upload_bf16(l.w_o, to_bf16(data, n)); }
    else if (s == "o.bias") { SERENC_TRY(expect(d)); st = upload_f32(l.b_o, data, n); }
    else if (s == "fc1.weight") { SERENC_TRY(expect((int64_t)c.ffn * d)); st = upload_bf16(l.w_fc1, to_bf16(data, n)); }
    else if (s == "fc1.bias") { SERENC_TRY(expect(c.ffn)); st = upload_f32(l.b_fc1, data, n); }
    else if (s == "fc2.weight") { SERENC_TRY(expect((int64_t)c.ffn * d)); st = upload_bf16(l.w_fc2, to_bf16(data, n)); }
    else if (s == "fc2.bias") { SERENC_TRY(expect(d)); st = upload_f32(l.b_fc2, data, n); }
    else if (c.wavlm_rel_bias && s == "gru.weight") {
      SERENC_TRY(expect(8 * hd));
      st = upload_f32(l.gru_w, data, n);
      // the gate only uses view(.., 2, 4).sum(-1) of the 8 outputs (HF modeling_wavlm.py:170-172): two hd-vectors
      std::vector<float> w2((size_t)2 * hd);
      for (int g2 = 0; g2 < 2; ++g2)
        for (int k = 0; k < hd; ++k)
          w2[(size_t)g2 * hd + k] = (data[(size_t)(4 * g2) * hd + k] + data[(size_t)(4 * g2 + 1) * hd + k]) +
                                    (data[(size_t)(4 * g2 + 2) * hd + k] + data[(size_t)(4 * g2 + 3) * hd + k]);
      if (st == 0) st = upload_f32(l.gru_w2, w2.data(), w2.size());
    }
    else if (c.wavlm_rel_bias && s == "gru.bias") {
      SERENC_TRY(expect(8));
      st = upload_f32(l.gru_b, data, n);
      const float b2[2] = {(data[0] + data[1]) + (data[2] + data[3]), (data[4] + data[5]) + (data[6] + data[7])};
      if (st == 0) st = upload_f32(l.gru_b2, b2, 2);
    }
    else if (c.wavlm_rel_bias && s == "gru.const") { SERENC_TRY(expect(c.heads)); st = upload_f32(l.gru_const, data, n); }
    else known = false;
  } else if (nm == "final_ln.weight") { SERENC_TRY(expect(d)); st = upload_f32(h->fin_g, data, n); }
  else if (nm == "final_ln.bias") { SERENC_TRY(expect(d)); st = upload_f32(h->fin_b, data, n); }
  else if (c.arch == SERENC_ARCH_W2V && sscanf(name, "conv%d.%63s", &ci, sub) == 2) {
    if (ci < 0 || ci > 6) SERENC_FAIL(SERENC_ERR_INVALID, "tensor %s: conv index out of range", name);
    const std::string s(sub);
    const int C = c.conv_dim;
    if (s == "weight") {
      if (ci == 0) { SERENC_TRY(expect((int64_t)C * W2V_K[0])); st = upload_f32(h->conv0_w, data, n); }
      else { SERENC_TRY(expect((int64_t)C * C * W2V_K[ci])); st = upload_bf16(h->conv_w[ci], pack_conv(data, C, C, W2V_K[ci], C)); }
    } else if (s == "bias") { SERENC_TRY(expect(C)); st = upload_f32(h->conv_b[ci], data, n); }
    else if (s == "ln.weight") { SERENC_TRY(expect(C)); st = upload_f32(h->conv_g[ci], data, n); }
    else if (s == "ln.bias") { SERENC_TRY(expect(C)); st = upload_f32(h->conv_be[ci], data, n); }
    else known = false;
  } else if (c.arch == SERENC_ARCH_W2V && nm == "featproj.ln.weight") { SERENC_TRY(expect(c.conv_dim)); st = upload_f32(h->fp_g, data, n); }
  else if (c.arch == SERENC_ARCH_W2V && nm == "featproj.ln.bias") { SERENC_TRY(expect(c.conv_dim)); st = upload_f32(h->fp_be, data, n); }
  else if (c.arch == SERENC_ARCH_W2V && nm == "featproj.weight") { SERENC_TRY(expect((int64_t)d * c.conv_dim)); st = upload_bf16(h->fp_w, to_bf16(data, n)); }
  else if (c.arch == SERENC_ARCH_W2V && nm == "featproj.bias") { SERENC_TRY(expect(d)); st = upload_f32(h->fp_b, data, n); }
  else if (c.arch == SERENC_ARCH_W2V && nm == "posconv.weight") {
    SERENC_TRY(expect((int64_t)d * h->pos_cg * c.pos_conv_kernel));
    st = upload_bf16(h->pos_w, pack_conv(data, d, h->pos_cg, c.pos_conv_kernel, h->pos_cg_pad));
  } else if (c.arch == SERENC_ARCH_W2V && nm == "posconv.bias") { SERENC_TRY(expect(d)); st = upload_f32(h->pos_b, data, n); }
  else if (c.arch == SERENC_ARCH_W2V && c.wavlm_rel_bias && nm == "rel_attn_embed") {
    SERENC_TRY(expect((int64_t)c.num_buckets * c.heads));
    h->rel_embed_host.assign(data, data + n);
    st = 0;
  } else if (c.arch == SERENC_ARCH_WHISPER && nm == "conv1.weight") {
    SERENC_TRY(expect((int64_t)d * c.n_mels * 3));
    st = upload_bf16(h->wc1, pack_conv(data, d, c.n_mels, 3, h->mel_pad));
  } else if (c.arch == SERENC_ARCH_WHISPER && nm == "conv1.bias") { SERENC_TRY(expect(d)); st = upload_f32(h->bc1, data, n); }
  else if (c.arch == SERENC_ARCH_WHISPER && nm == "conv2.weight") {
    SERENC_TRY(expect((int64_t)d * d * 3));
    st = upload_bf16(h->wc2, pack_conv(data, d, d, 3, d));
  } else if (c.arch == SERENC_ARCH_WHISPER && nm == "conv2.bias") { SERENC_TRY(expect(d)); st = upload_f32(h->bc2, data, n); }
  else if (c.arch == SERENC_ARCH_WHISPER && nm == "embed_positions") { SERENC_TRY(expect((int64_t)c.max_source_positions * d)); st = upload_f32(h->pos_emb, data, n); }
  else if (c.arch == SERENC_ARCH_TEXT && nm == "embed.word") { SERENC_TRY(expect((int64_t)c.vocab_size * d)); st = upload_f32(h->emb_word, data, n); }
  else if (c.arch == SERENC_ARCH_TEXT && nm == "embed.position") { SERENC_TRY(expect((int64_t)c.max_positions * d)); st = upload_f32(h->emb_pos, data, n); }
  else if (c.arch == SERENC_ARCH_TEXT && nm == "embed.type") { SERENC_TRY(expect((int64_t)c.type_vocab_size * d)); st = upload_f32(h->emb_type, data, n); }
  else if (c.arch == SERENC_ARCH_WHISPER && nm == "mel_filters") {
    SERENC_TRY(expect((int64_t)LM_BINS * c.n_mels));
    h->mel_filters_host.assign(data, data + n);
    st = 0;
  } else known = false;

  if (!known) SERENC_FAIL(SERENC_ERR_INVALID, "unknown tensor name '%s' for this architecture", name);
  if (st == 0) h->loaded.insert(nm);
  return st;
}

extern "C" int serenc_finalize(serenc_handle* h) {
  if (!h) SERENC_FAIL(SERENC_ERR_INVALID, "null handle");
  if (h->finalized) return 0;
  SERENC_CUDA_OK(cudaSetDevice(h->device));
  const serenc_config& c = h->cfg;
  std::vector<std::string> req;
  char buf[96];
  for (int i = 0; i < c.layers; ++i) {
    static const char* per[] = {"ln1.weight", "ln1.bias", "ln2.weight", "ln2.bias", "q.weight", "q.bias", "k.weight",
                                "v.weight", "v.bias", "o.weight", "o.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"};
    for (const char* s : per) { snprintf(buf, sizeof(buf), "layer%d.%s", i, s); req.push_back(buf); }
    if (c.arch != SERENC_ARCH_WHISPER) { snprintf(buf, sizeof(buf), "layer%d.k.bias", i); req.push_back(buf); }  // Whisper's k_proj has no bias
    if (c.wavlm_rel_bias) {
      for (const char* s : {"gru.weight", "gru.bias", "gru.const"}) { snprintf(buf, sizeof(buf), "layer%d.%s", i, s); req.push_back(buf); }
    }
  }
  req.push_back("final_ln.weight"); req.push_back("final_ln.bias");
  if (c.arch == SERENC_ARCH_W2V) {
    for (int i = 0; i < 7; ++i) {
      snprintf(buf, sizeof(buf), "conv%d.weight", i); req.push_back(buf);
      if (!c.conv_group_norm || i == 0) {
        snprintf(buf, sizeof(buf), "conv%d.ln.weight", i); req.push_back(buf);
        snprintf(buf, sizeof(buf), "conv%d.ln.bias", i); req.push_back(buf);
      }
      if (c.conv_bias) { snprintf(buf, sizeof(buf), "conv%d.bias", i); req.push_back(buf); }
    }
    for (const char* s : {"featproj.weight", "featproj.bias", "posconv.weight", "posconv.bias"}) req.push_back(s);
    if (!c.no_feat_proj_ln) { req.push_back("featproj.ln.weight"); req.push_back("featproj.ln.bias"); }
    if (c.wavlm_rel_bias) req.push_back("rel_attn_embed");
  } else if (c.arch == SERENC_ARCH_TEXT) {
    for (const char* s : {"embed.word", "embed.position", "embed.type"}) req.push_back(s);
  } else {
    for (const char* s : {"conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "embed_positions", "mel_filters"}) req.push_back(s);
  }
  for (const auto& r : req)
    if (!h->loaded.count(r)) SERENC_FAIL(SERENC_ERR_STATE, "finalize: tensor '%s' was never loaded", r.c_str());

  if (c.arch == SERENC_ARCH_W2V && c.wavlm_rel_bias) {
    // bias_h[delta] = rel_attn_embed[bucket(delta), h]  (HF compute_bias, modeling_wavlm.py:243-251)
    const int W = 2 * WAVLM_MAXD - 1;
    std::vector<float> tab((size_t)c.heads * W);
    for (int dlt = -(WAVLM_MAXD - 1); dlt <= WAVLM_MAXD - 1; ++dlt) {
      const int bk = serenc_wavlm_bucket(dlt, c.num_buckets, c.max_distance);
      for (int hh = 0; hh < c.heads; ++hh) tab[(size_t)hh * W + dlt + WAVLM_MAXD - 1] = h->rel_embed_host[(size_t)bk * c.heads + hh];
    }
    SERENC_TRY(upload_f32(h->btab, tab.data(), tab.size()));
  }
  if (c.arch == SERENC_ARCH_WHISPER) {
    std::vector<float> hann(LM_NFFT), ct(LM_NFFT), stb(LM_NFFT);
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < LM_NFFT; ++i) hann[i] = (float)(0.5 - 0.5 * cos(2.0 * PI * i / LM_NFFT));  // torch.hann_window(400), periodic
    // twiddles of the folded DFT: row n - 1 (n = 1..99), column t = thread of logmel_power_kernel, bin k = 2 (t % 128) + t / 128;
    // the angle is reduced exactly (n k mod 400) before the double-precision cos / sin
    std::vector<float> tw((size_t)99 * 256 * 2, 0.f);
    for (int n = 1; n < 100; ++n)
      for (int t = 0; t < 256; ++t) {
        const int k = 2 * (t & 127) + (t >> 7);
        if (k >= LM_BINS) continue;
        const int m = (n * k) % LM_NFFT;
        tw[((size_t)(n - 1) * 256 + t) * 2 + 0] = (float)cos(2.0 * PI * m / LM_NFFT);
        tw[((size_t)(n - 1) * 256 + t) * 2 + 1] = (float)sin(2.0 * PI * m / LM_NFFT);
      }
    std::vector<int32_t> ptr(c.n_mels + 1, 0), bin;
    std::vector<float> w;
    for (int m = 0; m < c.n_mels; ++m) {
      for (int k = 0; k < LM_BINS; ++k) {
        const float v = h->mel_filters_host[(size_t)k * c.n_mels + m];
        if (v != 0.f) { bin.push_back(k); w.push_back(v); }
      }
      ptr[m + 1] = (int32_t)bin.size();
    }
    SERENC_TRY(dev_alloc(h, &h->hann, LM_NFFT));
    { float* twp = nullptr; SERENC_TRY(dev_alloc(h, &twp, tw.size())); h->twid = reinterpret_cast<float2*>(twp); }
    SERENC_TRY(dev_alloc(h, &h->mel_ptr, ptr.size())); SERENC_TRY(dev_alloc(h, &h->mel_bin, bin.size())); SERENC_TRY(dev_alloc(h, &h->mel_w, w.size()));
    SERENC_TRY(upload_f32(h->hann, hann.data(), LM_NFFT)); SERENC_TRY(upload_f32(reinterpret_cast<float*>(h->twid), tw.data(), tw.size()));
    SERENC_CUDA_OK(cudaMemcpy(h->mel_ptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice));
    SERENC_CUDA_OK(cudaMemcpy(h->mel_bin, bin.data(), bin.size() * 4, cudaMemcpyHostToDevice));
    SERENC_TRY(upload_f32(h->mel_w, w.data(), w.size()));
  }
  SERENC_CUDA_OK(cudaDeviceSynchronize());
  h->finalized = true;
  return 0;
}

// =================================================================================================
// shared transformer stack
// =================================================================================================
namespace {

struct EmitCtx {
  uint64_t mask;
  int reduce;
  int n_sel;
  int sel = 0;
  bool first = true;
  float* frames_out;
  float* pooled_out;
  float* acc;  // REDUCE_MEAN / _WEIGHTED accumulator ([sumT, d]); == frames_out when that is given
  const float* weights = nullptr;   // host [n_sel]: SERENC_REDUCE_WEIGHTED; nullptr = 1 / n_sel each (the mean)
  bool pooled_linear = false;       // reducing, pooled output only: every selected state is pooled straight into pooled_out
  int64_t sumT;
  int d;
  int batch;
  const int32_t* frame_off_dev;
  const int32_t* n_keep_dev;
};

int pool_launch(serenc_handle* h, const float* x, int64_t sumT, int d, int batch, const int32_t* foff, const int32_t* n_keep, float* out, cudaStream_t st,
                float scale = 1.f, int accumulate = 0) {
  const dim3 grid(ceil_div(d, 128), batch);
  ProfScope ps(h, SERENC_PROF_POOL, 1, 0.0, (double)sumT * d * 4 + (double)batch * d * 4, st);
  masked_mean_pool_kernel<<<grid, 256, 0, st>>>(x, d, foff, n_keep, out, scale, accumulate);
  SERENC_CUDA_OK(cudaGetLastError());
  return 0;
}

int emit_hidden(serenc_handle* h, EmitCtx& e, int idx, const float* src, cudaStream_t st) {
  if (!((e.mask >> idx) & 1ull)) return 0;
  const int64_t n = e.sumT * e.d;
  if (e.reduce == SERENC_REDUCE_NONE) {
    if (e.frames_out)
      SERENC_CUDA_OK(cudaMemcpyAsync(e.frames_out + (int64_t)e.sel * n, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (e.pooled_out)
      SERENC_TRY(pool_launch(h, src, e.sumT, e.d, e.batch, e.frame_off_dev, e.n_keep_dev, e.pooled_out + (int64_t)e.sel * e.batch * e.d, st));
  } else {
    const float wgt = e.weights ? e.weights[e.sel] : 1.0f / (float)e.n_sel;
    if (e.pooled_linear) {
      // mean over layers and mean over frames commute: one read of the hidden state, no [sum_T, d] accumulator traffic
      SERENC_TRY(pool_launch(h, src, e.sumT, e.d, e.batch, e.frame_off_dev, e.n_keep_dev, e.pooled_out, st, wgt, e.first ? 0 : 1));
    } else {
      const int64_t n4 = n / 4;
      ProfScope ps(h, SERENC_PROF_POOL, 1, 0.0, (double)n * (e.first ? 8 : 12), st);
      accum_scaled_kernel<<<(unsigned)ceil_div64(n4, 256), 256, 0, st>>>(e.acc, src, n4, wgt, e.first ? 1 : 0);
      SERENC_CUDA_OK(cudaGetLastError());
    }
    e.first = false;
  }
  e.sel++;
  return 0;
}

struct StackBufs {
  float* x;     // [sumT, d] fp32 residual stream (in/out)
  float* xf;    // [sumT, d] fp32 final-LN output
  bf16* hln;    // [sumT, d]
  bf16* qkv;    // [sumT, 3d]
  bf16* att;    // [sumT, d]
  bf16* ffn;    // [sumT, ffn]
  float* gate = nullptr;  // [sumT, heads] WavLM gate of the current layer
  const int32_t* key_len = nullptr;  // [batch] text encoder: keys (non-pad tokens) per sequence; nullptr = every row is a key
};

// gate of layer li's attention, to be produced by the LayerNorm that writes its input (tcgen05 attention path only)
LnGate ln_gate_for(const serenc_handle* h, const StackBufs& b, int li) {
  LnGate g;
  if (h->cfg.wavlm_rel_bias && h->head_dim == 64 && !h->force_mma_sync_attn && li < h->cfg.layers && b.gate) {
    const LayerW& l = h->L[li];
    g.w2 = l.gru_w2; g.b2 = l.gru_b2; g.gconst = l.gru_const; g.out = b.gate;
  }
  return g;
}

// Pre-LN ("stable layer norm") encoder stack + final LayerNorm, emitting the selected hidden states.
// WavLMEncoderLayerStableLayerNorm / Wav2Vec2EncoderLayerStableLayerNorm / WhisperEncoderLayer
// (HF modeling_wavlm.py:339-373, :450-522; modeling_whisper.py:361-414).
int run_stack(serenc_handle* h, const StackBufs& b, int64_t sumT, int batch, int tmax, const int32_t* frame_off_dev,
              double attn_flops, EmitCtx& e, cudaStream_t st) {
  const serenc_config& c = h->cfg;
  const int d = c.hidden;
  SERENC_TRY(emit_hidden(h, e, 0, b.x, st));
  for (int li = 0; li < c.layers; ++li) {
    const LayerW& l = h->L[li];
    const LnGate lg = ln_gate_for(h, b, li);
    SERENC_TRY((launch_ln_t<float, bf16, false>(h, b.x, d, b.hln, d, l.ln1_g, l.ln1_b, sumT, d, nullptr, nullptr, c.layer_norm_eps, st, nullptr, lg)));
    {
      GemmCall g = linear_call(b.hln, sumT, d, l.w_qkv, 3 * d);
      g.bias = l.b_qkv; g.out_bf16 = b.qkv; g.ld_bf16 = 3 * d; g.prof_cls = SERENC_PROF_GEMM_QKV;
      SERENC_TRY(launch_gemm(h, g, st));
    }
    {
      AttnParams p;
      p.qkv = b.qkv; p.ld_qkv = 3 * d; p.d = d; p.frame_off = frame_off_dev; p.out = b.att;
      p.scale = 1.0f / sqrtf((float)h->head_dim);
      p.hln = b.hln; p.gru_w = l.gru_w; p.gru_b = l.gru_b; p.gru_const = l.gru_const; p.btab = h->btab; p.gate = b.gate; p.key_len = b.key_len;
      p.gate_ready = lg.out != nullptr;
      SERENC_TRY(launch_attn(h, p, c.wavlm_rel_bias != 0, tmax, batch, sumT, attn_flops, st));
    }
    {
      GemmCall g = linear_call(b.att, sumT, d, l.w_o, d);
      g.bias = l.b_o; g.resid = b.x; g.out_f32 = b.x; g.ld_f32 = d; g.prof_cls = SERENC_PROF_GEMM_OUT;
      SERENC_TRY(launch_gemm(h, g, st));
    }
    SERENC_TRY((launch_ln_t<float, bf16, false>(h, b.x, d, b.hln, d, l.ln2_g, l.ln2_b, sumT, d, nullptr, nullptr, c.layer_norm_eps, st)));
    {
      GemmCall g = linear_call(b.hln, sumT, d, l.w_fc1, c.ffn);
      g.bias = l.b_fc1; g.act = 1; g.out_bf16 = b.ffn; g.ld_bf16 = c.ffn; g.prof_cls = SERENC_PROF_GEMM_FC1;
      SERENC_TRY(launch_gemm(h, g, st));
    }
    {
      GemmCall g = linear_call(b.ffn, sumT, c.ffn, l.w_fc2, d);
      g.bias = l.b_fc2; g.resid = b.x; g.out_f32 = b.x; g.ld_f32 = d; g.prof_cls = SERENC_PROF_GEMM_FC2;
      SERENC_TRY(launch_gemm(h, g, st));
    }
    if (li + 1 < c.layers) SERENC_TRY(emit_hidden(h, e, li + 1, b.x, st));
  }
  SERENC_TRY((launch_ln_t<float, float, false>(h, b.x, d, b.xf, d, h->fin_g, h->fin_b, sumT, d, nullptr, nullptr, c.layer_norm_eps, st)));
  SERENC_TRY(emit_hidden(h, e, c.layers, b.xf, st));
  if (e.reduce != SERENC_REDUCE_NONE && e.pooled_out && e.n_sel > 0 && !e.pooled_linear)
    SERENC_TRY(pool_launch(h, e.acc, sumT, d, batch, frame_off_dev, e.n_keep_dev, e.pooled_out, st));
  return 0;
}

// Post-LN encoder stack (do_stable_layer_norm = false; the base-size checkpoints): encoder LayerNorm BEFORE the
// layers, each layer  x = LN1(x + attn(x));  x = LN2(x + ffn(x)),  no LayerNorm after the last layer.
// Wav2Vec2Encoder / Wav2Vec2EncoderLayer, WavLMEncoder / WavLMEncoderLayer (HF modeling_wav2vec2.py:634-713,
// modeling_wavlm.py:296-337, :375-448).  The LayerNorms run in place on the fp32 stream and also write the bf16
// copy the next GEMM reads, so the stream is read once per norm.
int run_stack_post_ln(serenc_handle* h, const StackBufs& b, int64_t sumT, int batch, int tmax, const int32_t* frame_off_dev,
                      double attn_flops, EmitCtx& e, cudaStream_t st) {
  const serenc_config& c = h->cfg;
  const int d = c.hidden;
  SERENC_TRY((launch_ln_t<float, float, false>(h, b.x, d, b.x, d, h->fin_g, h->fin_b, sumT, d, nullptr, nullptr, c.layer_norm_eps, st, b.hln, ln_gate_for(h, b, 0))));
  SERENC_TRY(emit_hidden(h, e, 0, b.x, st));
  for (int li = 0; li < c.layers; ++li) {
    const LayerW& l = h->L[li];
    {
      GemmCall g = linear_call(b.hln, sumT, d, l.w_qkv, 3 * d);
      g.bias = l.b_qkv; g.out_bf16 = b.qkv; g.ld_bf16 = 3 * d; g.prof_cls = SERENC_PROF_GEMM_QKV;
      SERENC_TRY(launch_gemm(h, g, st));
    }
    {
      AttnParams p;
      p.qkv = b.qkv; p.ld_qkv = 3 * d; p.d = d; p.frame_off = frame_off_dev; p.out = b.att;
      p.scale = 1.0f / sqrtf((float)h->head_dim);
      p.hln = b.hln; p.gru_w = l.gru_w; p.gru_b = l.gru_b; p.gru_const = l.gru_const; p.btab = h->btab; p.gate = b.gate; p.key_len = b.key_len;
      p.gate_ready = ln_gate_for(h, b, li).out != nullptr;
      SERENC_TRY(launch_attn(h, p, c.wavlm_rel_bias != 0, tmax, batch, sumT, attn_flops, st));
    }
    {
      GemmCall g = linear_call(b.att, sumT, d, l.w_o, d);
      g.bias = l.b_o; g.resid = b.x; g.out_f32 = b.x; g.ld_f32 = d; g.prof_cls = SERENC_PROF_GEMM_OUT;
      SERENC_TRY(launch_gemm(h, g, st));
    }
    SERENC_TRY((launch_ln_t<float, float, false>(h, b.x, d, b.x, d, l.ln1_g, l.ln1_b, sumT, d, nullptr, nullptr, c.layer_norm_eps, st, b.hln)));
    {
      GemmCall g = linear_call(b.hln, sumT, d, l.w_fc1, c.ffn);
      g.bias = l.b_fc1; g.act = 1; g.out_bf16 = b.ffn; g.ld_bf16 = c.ffn; g.prof_cls = SERENC_PROF_GEMM_FC1;
      SERENC_TRY(launch_gemm(h, g, st));
    }
    {
      GemmCall g = linear_call(b.ffn, sumT, c.ffn, l.w_fc2, d);
      g.bias = l.b_fc2; g.resid = b.x; g.out_f32 = b.x; g.ld_f32 = d; g.prof_cls = SERENC_PROF_GEMM_FC2;
      SERENC_TRY(launch_gemm(h, g, st));
    }
    SERENC_TRY((launch_ln_t<float, float, false>(h, b.x, d, b.x, d, l.ln2_g, l.ln2_b, sumT, d, nullptr, nullptr, c.layer_norm_eps, st, b.hln, ln_gate_for(h, b, li + 1))));
    SERENC_TRY(emit_hidden(h, e, li + 1, b.x, st));
  }
  if (e.reduce != SERENC_REDUCE_NONE && e.pooled_out && e.n_sel > 0 && !e.pooled_linear)
    SERENC_TRY(pool_launch(h, e.acc, sumT, d, batch, frame_off_dev, e.n_keep_dev, e.pooled_out, st));
  return 0;
}

int popcount64(uint64_t v) { int n = 0; while (v) { n += (int)(v & 1); v >>= 1; } return n; }

int make_emit(EmitCtx* e, uint64_t layer_mask, int reduce, const float* layer_weights, float* frames_out, float* pooled_out,
              float* acc_ws, int64_t sumT, int d, int batch, const int32_t* foff_dev, const int32_t* n_keep_dev) {
  if (reduce != SERENC_REDUCE_NONE && reduce != SERENC_REDUCE_MEAN && reduce != SERENC_REDUCE_WEIGHTED)
    SERENC_FAIL(SERENC_ERR_INVALID, "unknown reduce mode %d", reduce);
  if (reduce == SERENC_REDUCE_WEIGHTED && !layer_weights) SERENC_FAIL(SERENC_ERR_INVALID, "SERENC_REDUCE_WEIGHTED needs layer_weights");
  e->mask = layer_mask; e->reduce = reduce; e->n_sel = popcount64(layer_mask);
  e->weights = reduce == SERENC_REDUCE_WEIGHTED ? layer_weights : nullptr;
  e->frames_out = frames_out; e->pooled_out = pooled_out;
  e->acc = (reduce != SERENC_REDUCE_NONE && frames_out) ? frames_out : acc_ws;
  e->pooled_linear = reduce != SERENC_REDUCE_NONE && !frames_out && pooled_out;
  e->sumT = sumT; e->d = d; e->batch = batch; e->frame_off_dev = foff_dev; e->n_keep_dev = n_keep_dev;
  return 0;
}

// ---- wav2vec2-family plan ----
struct W2VPlan {
  int batch = 0;
  std::vector<int32_t> T[7];
  std::vector<int32_t> r6;      // slot row offset at layer 6
  std::vector<int32_t> foff;    // [B+1]
  int64_t r6_total = 0;
  int64_t sumT = 0;
  int tmax = 0;
  int64_t rows[7];              // rows of every conv output buffer
  int pad = 0;                  // pos-conv padding
  int64_t rgap = 0, mpos = 0;
};

int make_w2v_plan(const serenc_handle* h, const int32_t* len, int batch, W2VPlan* p) {
  if (batch <= 0) SERENC_FAIL(SERENC_ERR_INVALID, "batch must be positive");
  p->batch = batch;
  for (int k = 0; k < 7; ++k) p->T[k].resize(batch);
  p->r6.resize(batch);
  p->foff.resize(batch + 1);
  int64_t r6 = 0, sum = 0;
  int tmax = 0;
  for (int b = 0; b < batch; ++b) {
    int64_t n = len[b];
    for (int k = 0; k < 7; ++k) {
      n = n < W2V_K[k] ? 0 : (n - W2V_K[k]) / W2V_S[k] + 1;
      p->T[k][b] = (int32_t)n;
    }
    if (n < 1) SERENC_FAIL(SERENC_ERR_INVALID, "utterance %d has %d samples: shorter than the 400-sample receptive field (0 frames)", b, (int)len[b]);
    p->r6[b] = (int32_t)r6;
    p->foff[b] = (int32_t)sum;
    r6 += n + 2;
    sum += n;
    if (n > tmax) tmax = (int)n;
  }
  p->foff[batch] = (int32_t)sum;
  p->r6_total = r6;
  p->sumT = sum;
  p->tmax = tmax;
  for (int k = 0; k < 7; ++k) p->rows[k] = r6 << (6 - k);
  if (p->rows[0] >= (int64_t)1 << 31) SERENC_FAIL(SERENC_ERR_INVALID, "batch too large (conv0 rows overflow int32)");
  p->pad = h->cfg.pos_conv_kernel / 2;
  p->rgap = sum + (int64_t)p->pad * (batch + 1);
  p->mpos = p->rgap - p->pad;
  return 0;
}

struct W2VWs {
  Conv0Utt* utts; float2* stats; int32_t* foff; int32_t* r6; int32_t *fp_gather, *gap_row, *pos_rowmap;
  int32_t* c0_tiles;   // [B+1] prefix sum of the 128-frame conv0 tiles of every utterance (conv0_tc_kernel)
  double2* gn_partial; float2* gn_affine; int gn_tiles;
  bf16 *cbuf0, *cbuf1; bf16* featln; bf16* posin;
  StackBufs sb; float* acc;
  size_t bytes;
};

void carve_w2v(const serenc_handle* h, const W2VPlan& p, void* base, W2VWs* w) {
  const serenc_config& c = h->cfg;
  const int d = c.hidden, C = c.conv_dim;
  Carver cv(base);
  w->utts = cv.take<Conv0Utt>(p.batch);
  w->stats = cv.take<float2>(p.batch);
  w->foff = cv.take<int32_t>(p.batch + 1);
  w->r6 = cv.take<int32_t>(p.batch);
  w->c0_tiles = cv.take<int32_t>(p.batch + 1);
  w->fp_gather = cv.take<int32_t>(p.sumT);
  w->gap_row = cv.take<int32_t>(p.sumT);
  w->pos_rowmap = cv.take<int32_t>(p.mpos);
  w->gn_tiles = 0; w->gn_partial = nullptr; w->gn_affine = nullptr;
  if (c.conv_group_norm) {
    int slot_max = 0;
    for (int b = 0; b < p.batch; ++b) slot_max = slot_max > ((p.T[6][b] + 2) << 6) ? slot_max : ((p.T[6][b] + 2) << 6);
    w->gn_tiles = ceil_div(slot_max, CONV0_TILE);
    w->gn_partial = cv.take<double2>((size_t)p.batch * w->gn_tiles * CONV0_C);
    w->gn_affine = cv.take<float2>((size_t)p.batch * CONV0_C);
  }
  w->cbuf0 = cv.take<bf16>((size_t)p.rows[0] * C);  // conv0, conv2, conv4, conv6 outputs
  w->cbuf1 = cv.take<bf16>((size_t)p.rows[1] * C);  // conv1, conv3, conv5 outputs
  w->featln = cv.take<bf16>((size_t)p.sumT * C);
  w->posin = cv.take<bf16>((size_t)p.rgap * c.pos_conv_groups * h->pos_cg_pad);
  w->sb.x = cv.take<float>((size_t)p.sumT * d);
  w->sb.xf = cv.take<float>((size_t)p.sumT * d);
  w->sb.hln = cv.take<bf16>((size_t)p.sumT * d);
  w->sb.qkv = cv.take<bf16>((size_t)p.sumT * 3 * d);
  w->sb.att = cv.take<bf16>((size_t)p.sumT * d);
  w->sb.gate = c.wavlm_rel_bias ? cv.take<float>((size_t)p.sumT * c.heads) : nullptr;
  w->sb.ffn = cv.take<bf16>((size_t)p.sumT * c.ffn);
  w->acc = cv.take<float>((size_t)p.sumT * d);
  w->bytes = cv.used();
}

// small per-call metadata goes through a thread-local pinned staging buffer
struct Staging {
  void* host = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;
  bool pending = false;
};
// one per (thread, device): the event belongs to the device that was current when it was created, and a thread may drive
// replicas on several devices (tests/test_gpu_round2.py::test_second_replica_on_another_device)
constexpr int SERENC_MAX_DEVICES = 64;
thread_local Staging g_stages[SERENC_MAX_DEVICES];
thread_local int g_stage_dev = 0;   // device of the reservation in flight (stage_reserve .. stage_commit)

// `st` is the stream the staged bytes will be copied on. While that stream is being CAPTURED into a CUDA graph the
// shared staging buffer must not be used: the copy node re-reads its host source at every replay, and the event
// synchronisation below is illegal inside a capture. A captured call gets host memory of its own instead (plain
// malloc: cudaMallocHost is itself prohibited under the default capture mode; kept for the life of the process
// because a graph may be replayed at any time), and no event is recorded.
std::mutex g_graph_stage_mu;
std::vector<void*> g_graph_stage;   // host buffers owned by captured graphs
thread_local bool g_stage_captured = false;

int stage_reserve(size_t bytes, void** out, cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  SERENC_CUDA_OK(cudaStreamIsCapturing(st, &cs));
  g_stage_captured = (cs != cudaStreamCaptureStatusNone);
  if (g_stage_captured) {
    void* p = malloc(bytes ? bytes : 1);
    if (!p) SERENC_FAIL(SERENC_ERR_CUDA, "out of host memory staging %zu bytes for a captured call", bytes);
    std::lock_guard<std::mutex> lk(g_graph_stage_mu);
    g_graph_stage.push_back(p);
    *out = p;
    return 0;
  }
  int dev = 0;
  SERENC_CUDA_OK(cudaGetDevice(&dev));   // the entry point made the handle's device current (check_ready)
  if (dev < 0 || dev >= SERENC_MAX_DEVICES) SERENC_FAIL(SERENC_ERR_INVALID, "device index %d out of range", dev);
  g_stage_dev = dev;
  Staging& s = g_stages[dev];
  if (s.pending) { SERENC_CUDA_OK(cudaEventSynchronize(s.ev)); s.pending = false; }
  if (bytes > s.cap) {
    if (s.host) cudaFreeHost(s.host);
    s.host = nullptr; s.cap = 0;
    const size_t cap = (bytes + 65535) & ~size_t(65535);
    SERENC_CUDA_OK(cudaMallocHost(&s.host, cap));
    s.cap = cap;
  }
  if (!s.ev) SERENC_CUDA_OK(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  *out = s.host;
  return 0;
}
int stage_commit(cudaStream_t st) {
  if (g_stage_captured) return 0;
  SERENC_CUDA_OK(cudaEventRecord(g_stages[g_stage_dev].ev, st));
  g_stages[g_stage_dev].pending = true;
  return 0;
}

// stage a [batch] UttSpan table (sample ranges only) and copy it to `dst_dev`
int upload_spans(const int64_t* sample_start, const int32_t* sample_len, int batch, UttSpan* dst_dev, cudaStream_t st) {
  void* hs;
  SERENC_TRY(stage_reserve(sizeof(UttSpan) * (size_t)batch, &hs, st));
  UttSpan* hu = reinterpret_cast<UttSpan*>(hs);
  for (int b = 0; b < batch; ++b) {
    hu[b].sample_start = sample_start[b];
    hu[b].sample_len = sample_len[b];
    hu[b].T0 = 0; hu[b].row0 = 0; hu[b].slot = 0; hu[b].pad_ = 0;
  }
  SERENC_CUDA_OK(cudaMemcpyAsync(dst_dev, hu, sizeof(UttSpan) * (size_t)batch, cudaMemcpyHostToDevice, st));
  return stage_commit(st);
}

}  // namespace

// =================================================================================================
// wav2vec2 / HuBERT / WavLM
// =================================================================================================
extern "C" int serenc_w2v_workspace_bytes(const serenc_handle* h, const int32_t* sample_len, int batch, size_t* out_bytes) {
  if (!h || !sample_len || !out_bytes) SERENC_FAIL(SERENC_ERR_INVALID, "null argument");
  if (h->cfg.arch != SERENC_ARCH_W2V) SERENC_FAIL(SERENC_ERR_INVALID, "not a wav2vec2-family handle");
  W2VPlan p;
  SERENC_TRY(make_w2v_plan(h, sample_len, batch, &p));
  W2VWs w;
  carve_w2v(h, p, nullptr, &w);
  *out_bytes = w.bytes + 256;
  return 0;
}

static int wav_normalize_impl(serenc_handle* h, const float* wav_dev, const int64_t* sample_start, const int32_t* sample_len,
                              int batch, float* out_dev, int64_t out_stride, int32_t out_len, void* stream);
extern "C" int serenc_wav_normalize(serenc_handle* h, const float* wav_dev, const int64_t* sample_start, const int32_t* sample_len,
                                    int batch, float* out_dev, int64_t out_stride, int32_t out_len, void* stream) {
  return guarded(h, [&] { return wav_normalize_impl(h, wav_dev, sample_start, sample_len, batch, out_dev, out_stride, out_len, stream); });
}
static int wav_normalize_impl(serenc_handle* h, const float* wav_dev, const int64_t* sample_start, const int32_t* sample_len,
                              int batch, float* out_dev, int64_t out_stride, int32_t out_len, void* stream) {
  SERENC_TRY(check_ready(h, -1));
  if (!wav_dev || !sample_start || !sample_len || !out_dev || batch <= 0) SERENC_FAIL(SERENC_ERR_INVALID, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // tiny device scratch (span table + stats) from the stream-ordered pool
  void* dscratch;
  SERENC_CUDA_OK(cudaMallocAsync(&dscratch, (size_t)batch * (sizeof(UttSpan) + sizeof(float2)) + 64, st));
  UttSpan* d_utts = reinterpret_cast<UttSpan*>(dscratch);
  float2* d_stats = reinterpret_cast<float2*>(d_utts + batch);
  SERENC_TRY(upload_spans(sample_start, sample_len, batch, d_utts, st));
  wav_stats_kernel<<<batch, 1024, 0, st>>>(wav_dev, 0, d_utts, d_stats);
  SERENC_CUDA_OK(cudaGetLastError());
  const dim3 grid(ceil_div(out_len, 1024) < 64 ? ceil_div(out_len, 1024) : 64, batch);
  wav_normalize_kernel<<<grid, 256, 0, st>>>(wav_dev, d_utts, d_stats, out_dev, out_stride, out_len);
  SERENC_CUDA_OK(cudaGetLastError());
  SERENC_CUDA_OK(cudaFreeAsync(dscratch, st));
  return 0;
}

static int encode_w2v_impl(serenc_handle* h, const serenc_w2v_call& a) {
  SERENC_TRY(check_ready(h, SERENC_ARCH_W2V));
  const void* wav_dev = a.wav_dev;
  const int64_t* sample_start = a.sample_start;
  const int32_t* sample_len = a.sample_len;
  const int batch = a.batch, normalize = a.normalize, reduce = a.reduce;
  uint64_t layer_mask = a.layer_mask;
  float* frames_out_dev = a.frames_out_dev;
  float* pooled_out_dev = a.pooled_out_dev;
  int64_t* frame_offsets_out = a.frame_offsets_out;
  void* workspace_dev = a.workspace_dev;
  const size_t workspace_bytes = a.workspace_bytes;
  void* stream = a.stream;
  if (!wav_dev || !sample_start || !sample_len || !workspace_dev) SERENC_FAIL(SERENC_ERR_INVALID, "null argument");
  if (a.wav_dtype != SERENC_WAV_F32 && a.wav_dtype != SERENC_WAV_I16) SERENC_FAIL(SERENC_ERR_INVALID, "unknown wav_dtype %d", a.wav_dtype);
  const int wav_i16 = a.wav_dtype == SERENC_WAV_I16;
  if (a.extract_features_out_dev && h->cfg.no_feat_proj_ln)
    SERENC_FAIL(SERENC_ERR_INVALID, "extract_features needs a feature-projection LayerNorm (HubertModel returns none)");
  const serenc_config& c = h->cfg;
  const int d = c.hidden, C = c.conv_dim;
  if (c.layers < 63) layer_mask &= ((1ull << (c.layers + 1)) - 1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  W2VPlan p;
  SERENC_TRY(make_w2v_plan(h, sample_len, batch, &p));
  W2VWs w;
  carve_w2v(h, p, workspace_dev, &w);
  if (w.bytes > workspace_bytes) SERENC_FAIL(SERENC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  if (frame_offsets_out) for (int b = 0; b <= batch; ++b) frame_offsets_out[b] = p.foff[b];

  // ---- per-utterance tables -> device ----
  int c0_total_tiles = 0;
  {
    const size_t sz_u = sizeof(Conv0Utt) * batch, sz_f = 4 * (size_t)(batch + 1), sz_r = 4 * (size_t)batch;
    void* hs;
    SERENC_TRY(stage_reserve(sz_u + 2 * sz_f + sz_r + 64, &hs, st));
    Conv0Utt* hu = reinterpret_cast<Conv0Utt*>(hs);
    for (int b = 0; b < batch; ++b) {
      hu[b].sample_start = sample_start[b];
      hu[b].sample_len = sample_len[b];
      hu[b].T0 = p.T[0][b];
      hu[b].row0 = (int64_t)p.r6[b] << 6;
      hu[b].slot = (p.T[6][b] + 2) << 6;
      hu[b].pad_ = 0;
    }
    int32_t* hf = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(hs) + sz_u);
    memcpy(hf, p.foff.data(), sz_f);
    int32_t* hr = hf + batch + 1;
    memcpy(hr, p.r6.data(), sz_r);
    int32_t* ht = hr + batch;
    ht[0] = 0;
    for (int b = 0; b < batch; ++b) ht[b + 1] = ht[b] + ceil_div((p.T[6][b] + 2) << 6, C0T_BM);
    c0_total_tiles = ht[batch];
    SERENC_CUDA_OK(cudaMemcpyAsync(w.c0_tiles, ht, sz_f, cudaMemcpyHostToDevice, st));
    SERENC_CUDA_OK(cudaMemcpyAsync(w.utts, hu, sz_u, cudaMemcpyHostToDevice, st));
    SERENC_CUDA_OK(cudaMemcpyAsync(w.foff, hf, sz_f, cudaMemcpyHostToDevice, st));
    SERENC_CUDA_OK(cudaMemcpyAsync(w.r6, hr, sz_r, cudaMemcpyHostToDevice, st));
    SERENC_TRY(stage_commit(st));
    const int64_t nthreads = p.sumT > p.mpos ? p.sumT : p.mpos;
    ProfScope ps(h, SERENC_PROF_MISC, 1, 0.0, 0.0, st);
    w2v_plan_kernel<<<(unsigned)ceil_div64(nthreads, 256), 256, 0, st>>>(w.foff, w.r6, batch, p.pad, p.sumT, p.mpos, w.fp_gather, w.gap_row, w.pos_rowmap);
    SERENC_CUDA_OK(cudaGetLastError());
  }

  // ---- feature encoder: conv0 (+norm stats) then conv1..6 as implicit GEMMs, each followed by LN + GELU ----
  {
    if (normalize) {
      double nsamp = 0;
      for (int b = 0; b < batch; ++b) nsamp += sample_len[b];
      ProfScope ps(h, SERENC_PROF_MISC, 1, 0.0, 8.0 * nsamp, st);
      wav_stats_kernel<<<batch, 1024, 0, st>>>(wav_dev, wav_i16, w.utts, w.stats);
      SERENC_CUDA_OK(cudaGetLastError());
    }
    int slot_max = 0;
    for (int b = 0; b < batch; ++b) slot_max = slot_max > ((p.T[6][b] + 2) << 6) ? slot_max : ((p.T[6][b] + 2) << 6);
    const dim3 grid(ceil_div(slot_max, CONV0_TILE), batch);
    double t0sum = 0, nsamp0 = 0;
    for (int b = 0; b < batch; ++b) { t0sum += p.T[0][b]; nsamp0 += sample_len[b]; }
    const float2* stp = normalize ? w.stats : nullptr;
    const float* b0 = c.conv_bias ? h->conv_b[0] : nullptr;
    if (!c.conv_group_norm) {
      // LayerNorm variant: the 10-tap dot products on the tensor cores, LayerNorm + GELU straight out of TMEM (conv0_tc.cuh)
      ProfScope ps(h, SERENC_PROF_CONV0, 1, 2.0 * t0sum * CONV0_C * CONV0_K, (wav_i16 ? 2.0 : 4.0) * nsamp0 + 2.0 * t0sum * CONV0_C, st);
      const int ctas = c0_total_tiles < h->num_sms ? c0_total_tiles : h->num_sms;
      conv0_tc_kernel<<<ctas, C0T_THREADS, C0T_SMEM, st>>>(wav_dev, wav_i16, w.utts, w.c0_tiles, batch, stp, h->conv0_w, b0, h->conv_g[0], h->conv_be[0], w.cbuf0);
      SERENC_CUDA_OK(cudaGetLastError());
    } else {
      // GroupNorm(512 groups) on conv0: statistics over each utterance's valid frames, then recompute + apply
      ProfScope ps(h, SERENC_PROF_CONV0, 3, 4.0 * t0sum * CONV0_C * CONV0_K, 8.0 * nsamp0 + 2.0 * t0sum * CONV0_C, st);
      conv0_kernel<1><<<grid, 256, 0, st>>>(wav_dev, wav_i16, w.utts, stp, h->conv0_w, b0, nullptr, nullptr, w.cbuf0, w.gn_partial, nullptr, w.gn_tiles);
      SERENC_CUDA_OK(cudaGetLastError());
      conv0_gn_finalize_kernel<<<dim3(CONV0_C / 128, batch), 128, 0, st>>>(w.gn_partial, w.utts, w.gn_tiles, h->conv_g[0], h->conv_be[0], w.gn_affine);
      SERENC_CUDA_OK(cudaGetLastError());
      conv0_kernel<2><<<grid, 256, 0, st>>>(wav_dev, wav_i16, w.utts, stp, h->conv0_w, b0, nullptr, nullptr, w.cbuf0, nullptr, w.gn_affine, w.gn_tiles);
      SERENC_CUDA_OK(cudaGetLastError());
    }
  }
  bf16* cin = w.cbuf0;
  bf16* cout = w.cbuf1;
  for (int k = 1; k < 7; ++k) {
    GemmCall g;
    g.A = cin; g.a_cols = C; g.a_rows = p.rows[k - 1]; g.a_ld = C; g.a_stride = W2V_S[k]; g.a_kpt = C / GEMM_BK; g.taps = W2V_K[k];
    g.M = p.rows[k]; g.W = h->conv_w[k]; g.w_rows = C; g.n_per_group = C; g.groups = 1;
    g.bias = c.conv_bias ? h->conv_b[k] : nullptr;
    g.out_bf16 = cout; g.ld_bf16 = C;
    g.prof_cls = SERENC_PROF_GEMM_CONV;
    g.act = c.conv_group_norm ? 1 : 0;   // "group" encoders: conv -> GELU (no norm after layer 0), fused in the epilogue
    { double tk = 0; for (int b = 0; b < batch; ++b) tk += p.T[k][b]; g.alg_flops = 2.0 * tk * C * C * W2V_K[k]; }
    SERENC_TRY(launch_gemm(h, g, st));
    if (!c.conv_group_norm)
      SERENC_TRY((launch_ln_t<bf16, bf16, true>(h, cout, C, cout, C, h->conv_g[k], h->conv_be[k], p.rows[k], C, nullptr, nullptr, 1e-5f, st)));
    bf16* t = cin; cin = cout; cout = t;
  }
  // ---- feature projection: LN(512) over the valid frames (gathered into the packed layout) -> Linear(512 -> d) ----
  if (!c.no_feat_proj_ln && a.extract_features_out_dev) {
    // HF's `extract_features` (the normed 512-d features, modeling_wavlm.py:93-105) in fp32 + the bf16 GEMM operand
    SERENC_TRY((launch_ln_t<bf16, float, false>(h, cin, C, a.extract_features_out_dev, C, h->fp_g, h->fp_be, p.sumT, C, w.fp_gather, nullptr,
                                                c.layer_norm_eps, st, w.featln)));
  } else if (!c.no_feat_proj_ln) {
    SERENC_TRY((launch_ln_t<bf16, bf16, false>(h, cin, C, w.featln, C, h->fp_g, h->fp_be, p.sumT, C, w.fp_gather, nullptr, c.layer_norm_eps, st)));
  } else {
    ProfScope ps(h, SERENC_PROF_MISC, 1, 0.0, (double)p.sumT * C * 4, st);
    const int64_t nthr = p.sumT * (C / 8);
    gather_rows_bf16_kernel<<<(unsigned)ceil_div64(nthr, 256), 256, 0, st>>>(cin, C, w.featln, C, p.sumT, C / 8, w.fp_gather);
    SERENC_CUDA_OK(cudaGetLastError());
  }
  {
    GemmCall g = linear_call(w.featln, p.sumT, C, h->fp_w, d);
    g.bias = h->fp_b; g.out_f32 = w.sb.x; g.ld_f32 = d;
    SERENC_TRY(launch_gemm(h, g, st));
  }
  // ---- positional conv embedding: x += gelu(grouped_conv(x) + b) ----
  {
    const int64_t ld_pos = (int64_t)c.pos_conv_groups * h->pos_cg_pad;
    {
      const int64_t nthr = p.sumT * (d / 4);
      ProfScope ps(h, SERENC_PROF_MISC, 2, 0.0, (double)p.sumT * d * 6, st);
      SERENC_CUDA_OK(cudaMemsetAsync(w.posin, 0, (size_t)p.rgap * ld_pos * sizeof(bf16), st));
      scatter_posconv_in_kernel<<<(unsigned)ceil_div64(nthr, 256), 256, 0, st>>>(w.sb.x, p.sumT, d, h->pos_cg, h->pos_cg_pad, w.gap_row, w.posin, ld_pos);
      SERENC_CUDA_OK(cudaGetLastError());
    }
    GemmCall g;
    g.A = w.posin; g.a_cols = ld_pos; g.a_rows = p.rgap; g.a_ld = ld_pos; g.a_stride = 1; g.a_kpt = h->pos_cg_pad / GEMM_BK;
    g.taps = c.pos_conv_kernel; g.a_group_stride = h->pos_cg_pad; g.a_cg_valid = h->pos_cg;
    g.M = p.mpos; g.W = h->pos_w; g.w_rows = d; g.n_per_group = h->pos_cg; g.groups = c.pos_conv_groups;
    g.bias = h->pos_b; g.act = 1; g.resid = w.sb.x; g.out_f32 = w.sb.x; g.ld_f32 = d; g.rowmap = w.pos_rowmap;
    g.prof_cls = SERENC_PROF_GEMM_POSCONV;
    g.alg_flops = 2.0 * (double)p.sumT * d * h->pos_cg * c.pos_conv_kernel;
    SERENC_TRY(launch_gemm(h, g, st));
  }
  // ---- transformer stack ----
  EmitCtx e;
  SERENC_TRY(make_emit(&e, layer_mask, reduce, a.layer_weights, frames_out_dev, pooled_out_dev, w.acc, p.sumT, d, batch, w.foff, nullptr));
  double attn_flops = 0.0;
  for (int b = 0; b < batch; ++b) attn_flops += 4.0 * (double)p.T[6][b] * p.T[6][b] * d;
  if (c.post_layer_norm) SERENC_TRY(run_stack_post_ln(h, w.sb, p.sumT, batch, p.tmax, w.foff, attn_flops, e, st));
  else SERENC_TRY(run_stack(h, w.sb, p.sumT, batch, p.tmax, w.foff, attn_flops, e, st));
  return 0;
}

extern "C" int serenc_encode_w2v_ex(serenc_handle* h, const serenc_w2v_call* call) {
  if (!call) SERENC_FAIL(SERENC_ERR_INVALID, "null argument");
  return guarded(h, [&] { return encode_w2v_impl(h, *call); });
}

extern "C" int serenc_encode_w2v(serenc_handle* h, const float* wav_dev, const int64_t* sample_start, const int32_t* sample_len,
                                 int batch, int normalize, uint64_t layer_mask, int reduce, float* frames_out_dev,
                                 float* pooled_out_dev, int64_t* frame_offsets_out, void* workspace_dev, size_t workspace_bytes,
                                 void* stream) {
  serenc_w2v_call a;
  memset(&a, 0, sizeof(a));
  a.wav_dev = wav_dev; a.wav_dtype = SERENC_WAV_F32; a.batch = batch; a.sample_start = sample_start; a.sample_len = sample_len;
  a.normalize = normalize; a.reduce = reduce; a.layer_mask = layer_mask; a.frames_out_dev = frames_out_dev;
  a.pooled_out_dev = pooled_out_dev; a.frame_offsets_out = frame_offsets_out; a.workspace_dev = workspace_dev;
  a.workspace_bytes = workspace_bytes; a.stream = stream;
  return serenc_encode_w2v_ex(h, &a);
}

static int unpack_rows_impl(serenc_handle* h, const float* packed_dev, const int64_t* frame_offsets, int batch,
                            int32_t t_max, int32_t d, float* out_dev, void* stream) {
  SERENC_TRY(check_ready(h, -1));
  if (!packed_dev || !frame_offsets || !out_dev || batch <= 0 || t_max <= 0 || d <= 0 || d % 4) SERENC_FAIL(SERENC_ERR_INVALID, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int32_t* d_off;
  SERENC_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&d_off), 4 * (size_t)(batch + 1), st));
  void* hs;
  SERENC_TRY(stage_reserve(4 * (size_t)(batch + 1), &hs, st));
  for (int b = 0; b <= batch; ++b) reinterpret_cast<int32_t*>(hs)[b] = (int32_t)frame_offsets[b];
  SERENC_CUDA_OK(cudaMemcpyAsync(d_off, hs, 4 * (size_t)(batch + 1), cudaMemcpyHostToDevice, st));
  SERENC_TRY(stage_commit(st));
  const dim3 grid(ceil_div(d / 4, 128), t_max, batch);
  unpack_frames_kernel<<<grid, 128, 0, st>>>(packed_dev, d, d_off, t_max, out_dev);
  SERENC_CUDA_OK(cudaGetLastError());
  SERENC_CUDA_OK(cudaFreeAsync(d_off, st));
  return 0;
}

extern "C" int serenc_unpack_rows(serenc_handle* h, const float* packed_dev, const int64_t* frame_offsets, int batch,
                                  int32_t t_max, int32_t cols, float* out_dev, void* stream) {
  return guarded(h, [&] { return unpack_rows_impl(h, packed_dev, frame_offsets, batch, t_max, cols, out_dev, stream); });
}
extern "C" int serenc_unpack_frames(serenc_handle* h, const float* packed_dev, const int64_t* frame_offsets, int batch,
                                    int32_t t_max, float* out_dev, void* stream) {
  if (!h) SERENC_FAIL(SERENC_ERR_INVALID, "null handle");
  return serenc_unpack_rows(h, packed_dev, frame_offsets, batch, t_max, h->cfg.hidden, out_dev, stream);
}

// =================================================================================================
// Whisper
// =================================================================================================
static int logmel_impl(serenc_handle* h, const void* wav_dev, int wav_dtype, const int64_t* sample_start, const int32_t* sample_len,
                       int batch, float* mel_out_dev, void* scratch_dev, void* stream) {
  SERENC_TRY(check_ready(h, SERENC_ARCH_WHISPER));
  if (!wav_dev || !sample_start || !sample_len || !mel_out_dev || !scratch_dev || batch <= 0) SERENC_FAIL(SERENC_ERR_INVALID, "bad argument");
  if (wav_dtype != SERENC_WAV_F32 && wav_dtype != SERENC_WAV_I16) SERENC_FAIL(SERENC_ERR_INVALID, "unknown wav_dtype %d", wav_dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // scratch layout: [batch] ordered-max words, then the span table
  UttSpan* d_utts = reinterpret_cast<UttSpan*>(reinterpret_cast<uint8_t*>(scratch_dev) + (((size_t)batch * 4 + 63) & ~size_t(63)));
  SERENC_TRY(upload_spans(sample_start, sample_len, batch, d_utts, st));
  uint32_t* umax = reinterpret_cast<uint32_t*>(scratch_dev);
  SERENC_CUDA_OK(cudaMemsetAsync(umax, 0, 4 * (size_t)batch, st));
  LogmelTables tb;
  tb.hann = h->hann; tb.twid = h->twid;
  tb.mel_ptr = h->mel_ptr; tb.mel_bin = h->mel_bin; tb.mel_w = h->mel_w; tb.n_mels = h->cfg.n_mels;
  const dim3 grid(ceil_div(LM_FRAMES, LM_FR), batch);
  ProfScope ps(h, SERENC_PROF_LOGMEL, 3, 0.0, (double)batch * (4.0 * LM_NSAMP + 4.0 * h->cfg.n_mels * LM_FRAMES), st);
  logmel_power_kernel<<<grid, LM_THREADS, 0, st>>>(wav_dev, wav_dtype == SERENC_WAV_I16, d_utts, tb, mel_out_dev, umax);
  SERENC_CUDA_OK(cudaGetLastError());
  const int64_t per_utt = (int64_t)h->cfg.n_mels * LM_FRAMES;
  logmel_finalize_kernel<<<dim3(64, batch), 256, 0, st>>>(mel_out_dev, umax, per_utt);
  SERENC_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int serenc_logmel_ex(serenc_handle* h, const void* wav_dev, int wav_dtype, const int64_t* sample_start, const int32_t* sample_len,
                                int batch, float* mel_out_dev, void* scratch_dev, void* stream) {
  return guarded(h, [&] { return logmel_impl(h, wav_dev, wav_dtype, sample_start, sample_len, batch, mel_out_dev, scratch_dev, stream); });
}
extern "C" int serenc_logmel(serenc_handle* h, const float* wav_dev, const int64_t* sample_start, const int32_t* sample_len,
                             int batch, float* mel_out_dev, void* scratch_dev, void* stream) {
  return serenc_logmel_ex(h, wav_dev, SERENC_WAV_F32, sample_start, sample_len, batch, mel_out_dev, scratch_dev, stream);
}

namespace {
struct WhisperWs {
  int32_t *map1, *map2, *foff, *n_keep;
  bf16 *melrows, *c1;
  StackBufs sb; float* acc;
  size_t bytes;
};
void carve_whisper(const serenc_handle* h, int batch, void* base, WhisperWs* w) {
  const serenc_config& c = h->cfg;
  const int d = c.hidden;
  const int64_t sumT = (int64_t)batch * 1500;
  Carver cv(base);
  w->map1 = cv.take<int32_t>((size_t)batch * 3002);
  w->map2 = cv.take<int32_t>((size_t)batch * 1501);
  w->foff = cv.take<int32_t>(batch + 1);
  w->n_keep = cv.take<int32_t>(batch);
  w->melrows = cv.take<bf16>((size_t)batch * 3002 * h->mel_pad);
  w->c1 = cv.take<bf16>((size_t)batch * 3002 * d);
  w->sb.x = cv.take<float>((size_t)sumT * d);
  w->sb.xf = cv.take<float>((size_t)sumT * d);
  w->sb.hln = cv.take<bf16>((size_t)sumT * d);
  w->sb.qkv = cv.take<bf16>((size_t)sumT * 3 * d);
  w->sb.att = cv.take<bf16>((size_t)sumT * d);
  w->sb.ffn = cv.take<bf16>((size_t)sumT * c.ffn);
  w->acc = cv.take<float>((size_t)sumT * d);
  w->bytes = cv.used();
}
}  // namespace

extern "C" int serenc_whisper_workspace_bytes(const serenc_handle* h, int batch, size_t* out_bytes) {
  if (!h || !out_bytes || batch <= 0) SERENC_FAIL(SERENC_ERR_INVALID, "bad argument");
  if (h->cfg.arch != SERENC_ARCH_WHISPER) SERENC_FAIL(SERENC_ERR_INVALID, "not a Whisper handle");
  WhisperWs w;
  carve_whisper(h, batch, nullptr, &w);
  *out_bytes = w.bytes + 256;
  return 0;
}

static int encode_whisper_impl(serenc_handle* h, const float* mel_dev, int batch, uint64_t layer_mask, int reduce,
                               const float* layer_weights, const int32_t* n_keep, float* frames_out_dev, float* pooled_out_dev,
                               void* workspace_dev, size_t workspace_bytes, void* stream) {
  SERENC_TRY(check_ready(h, SERENC_ARCH_WHISPER));
  if (!mel_dev || !workspace_dev || batch <= 0) SERENC_FAIL(SERENC_ERR_INVALID, "bad argument");
  const serenc_config& c = h->cfg;
  const int d = c.hidden;
  if (c.layers < 63) layer_mask &= ((1ull << (c.layers + 1)) - 1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  WhisperWs w;
  carve_whisper(h, batch, workspace_dev, &w);
  if (w.bytes > workspace_bytes) SERENC_FAIL(SERENC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  const int64_t sumT = (int64_t)batch * 1500;

  {
    void* hs;
    SERENC_TRY(stage_reserve(4 * (size_t)(2 * batch + 1), &hs, st));
    int32_t* hf = reinterpret_cast<int32_t*>(hs);
    for (int b = 0; b <= batch; ++b) hf[b] = b * 1500;
    for (int b = 0; b < batch; ++b) hf[batch + 1 + b] = n_keep ? (n_keep[b] < 1 ? 1 : (n_keep[b] > 1500 ? 1500 : n_keep[b])) : 1500;
    SERENC_CUDA_OK(cudaMemcpyAsync(w.foff, hf, 4 * (size_t)(batch + 1), cudaMemcpyHostToDevice, st));
    SERENC_CUDA_OK(cudaMemcpyAsync(w.n_keep, hf + batch + 1, 4 * (size_t)batch, cudaMemcpyHostToDevice, st));
    SERENC_TRY(stage_commit(st));
    ProfScope ps(h, SERENC_PROF_MISC, 1, 0.0, 0.0, st);
    whisper_plan_kernel<<<(unsigned)ceil_div64((int64_t)batch * 3002, 256), 256, 0, st>>>(batch, w.map1, w.map2);
    SERENC_CUDA_OK(cudaGetLastError());
  }
  // conv stem (HF modeling_whisper.py:619-625): gelu(conv1 k3 p1) -> gelu(conv2 k3 s2 p1) -> + positions
  {
    ProfScope ps(h, SERENC_PROF_MISC, 3, 0.0, (double)batch * 3000 * c.n_mels * 6 + (double)batch * 3002 * d * 2, st);
    SERENC_CUDA_OK(cudaMemsetAsync(w.melrows, 0, (size_t)batch * 3002 * h->mel_pad * sizeof(bf16), st));
    SERENC_CUDA_OK(cudaMemsetAsync(w.c1, 0, (size_t)batch * 3002 * d * sizeof(bf16), st));
    const dim3 grid(ceil_div(LM_FRAMES, 32), ceil_div(c.n_mels, 32), batch);
    mel_to_rows_kernel<<<grid, 256, 0, st>>>(mel_dev, c.n_mels, w.melrows, h->mel_pad);
    SERENC_CUDA_OK(cudaGetLastError());
  }
  {
    GemmCall g;
    g.prof_cls = SERENC_PROF_GEMM_CONV;
    g.alg_flops = 2.0 * batch * 3000.0 * d * c.n_mels * 3;
    g.A = w.melrows; g.a_cols = h->mel_pad; g.a_rows = (int64_t)batch * 3002; g.a_ld = h->mel_pad; g.a_stride = 1;
    g.a_kpt = h->mel_pad / GEMM_BK; g.taps = 3;
    g.M = (int64_t)batch * 3002; g.W = h->wc1; g.w_rows = d; g.n_per_group = d; g.groups = 1;
    g.bias = h->bc1; g.act = 1; g.out_bf16 = w.c1; g.ld_bf16 = d; g.rowmap = w.map1;
    SERENC_TRY(launch_gemm(h, g, st));
  }
  {
    const int64_t per_utt4 = (int64_t)1500 * d / 4;
    {
      ProfScope ps(h, SERENC_PROF_MISC, 1, 0.0, (double)batch * 1500 * d * 4, st);
      broadcast_rows_kernel<<<dim3(128, batch), 256, 0, st>>>(h->pos_emb, per_utt4, w.sb.x);
      SERENC_CUDA_OK(cudaGetLastError());
    }
    GemmCall g;
    g.prof_cls = SERENC_PROF_GEMM_CONV;
    g.alg_flops = 2.0 * batch * 1500.0 * d * d * 3;
    g.A = w.c1; g.a_cols = d; g.a_rows = (int64_t)batch * 3002; g.a_ld = d; g.a_stride = 2; g.a_kpt = d / GEMM_BK; g.taps = 3;
    g.M = (int64_t)batch * 1501; g.W = h->wc2; g.w_rows = d; g.n_per_group = d; g.groups = 1;
    g.bias = h->bc2; g.act = 1; g.resid = w.sb.x; g.out_f32 = w.sb.x; g.ld_f32 = d; g.rowmap = w.map2;
    SERENC_TRY(launch_gemm(h, g, st));
  }
  EmitCtx e;
  SERENC_TRY(make_emit(&e, layer_mask, reduce, layer_weights, frames_out_dev, pooled_out_dev, w.acc, sumT, d, batch, w.foff, w.n_keep));
  SERENC_TRY(run_stack(h, w.sb, sumT, batch, 1500, w.foff, 4.0 * 1500.0 * 1500.0 * d * batch, e, st));
  return 0;
}
extern "C" int serenc_encode_whisper_ex(serenc_handle* h, const float* mel_dev, int batch, uint64_t layer_mask, int reduce,
                                        const float* layer_weights, const int32_t* n_keep, float* frames_out_dev,
                                        float* pooled_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  return guarded(h, [&] {
    return encode_whisper_impl(h, mel_dev, batch, layer_mask, reduce, layer_weights, n_keep, frames_out_dev, pooled_out_dev, workspace_dev,
                               workspace_bytes, stream);
  });
}
extern "C" int serenc_encode_whisper(serenc_handle* h, const float* mel_dev, int batch, uint64_t layer_mask, int reduce,
                                     const int32_t* n_keep, float* frames_out_dev, float* pooled_out_dev, void* workspace_dev,
                                     size_t workspace_bytes, void* stream) {
  return serenc_encode_whisper_ex(h, mel_dev, batch, layer_mask, reduce, nullptr, n_keep, frames_out_dev, pooled_out_dev, workspace_dev,
                                  workspace_bytes, stream);
}

// =================================================================================================
// text encoder (RoBERTa): embeddings + post-LN stack
// =================================================================================================
namespace {
// x[row] = word[ids[row]] + position[pos_id] + type[0];  pos_id = pad + 1 + t for the valid_len leading tokens, pad after
// (HF RobertaEmbeddings.forward + create_position_ids_from_input_ids, modeling_roberta.py). One warp per row.
__global__ void __launch_bounds__(256) text_embed_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ valid_len,
                                                         int seq_len, int64_t rows, int d, int vocab, int max_pos, int pad,
                                                         const float* __restrict__ word, const float* __restrict__ pos,
                                                         const float* __restrict__ type, float* __restrict__ x) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(row / seq_len), t = (int)(row - (int64_t)b * seq_len);
  int id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);   // memory safety only: the host side rejects out-of-range ids
  int pid = t < valid_len[b] ? pad + 1 + t : pad;
  pid = pid >= max_pos ? max_pos - 1 : pid;
  const float4* w4 = reinterpret_cast<const float4*>(word + (int64_t)id * d);
  const float4* p4 = reinterpret_cast<const float4*>(pos + (int64_t)pid * d);
  const float4* t4 = reinterpret_cast<const float4*>(type);
  float4* o4 = reinterpret_cast<float4*>(x + row * d);
  for (int c = lane; c < d / 4; c += 32) {
    const float4 a = __ldg(w4 + c), e = __ldg(t4 + c), q = __ldg(p4 + c);
    // HF: inputs_embeds + token_type_embeddings, then + position_embeddings
    o4[c] = make_float4((a.x + e.x) + q.x, (a.y + e.y) + q.y, (a.z + e.z) + q.z, (a.w + e.w) + q.w);
  }
}

struct TextWs {
  int32_t *foff, *klen;
  StackBufs sb; float* acc;
  size_t bytes;
};
void carve_text(const serenc_handle* h, int batch, int seq_len, void* base, TextWs* w) {
  const serenc_config& c = h->cfg;
  const int d = c.hidden;
  const int64_t sumT = (int64_t)batch * seq_len;
  Carver cv(base);
  w->foff = cv.take<int32_t>(batch + 1);
  w->klen = cv.take<int32_t>(batch);
  w->sb.x = cv.take<float>((size_t)sumT * d);
  w->sb.xf = nullptr;
  w->sb.hln = cv.take<bf16>((size_t)sumT * d);
  w->sb.qkv = cv.take<bf16>((size_t)sumT * 3 * d);
  w->sb.att = cv.take<bf16>((size_t)sumT * d);
  w->sb.ffn = cv.take<bf16>((size_t)sumT * c.ffn);
  w->acc = cv.take<float>((size_t)sumT * d);
  w->bytes = cv.used();
}
}  // namespace

extern "C" int serenc_text_workspace_bytes(const serenc_handle* h, int batch, int seq_len, size_t* out_bytes) {
  if (!h || !out_bytes || batch <= 0 || seq_len <= 0) SERENC_FAIL(SERENC_ERR_INVALID, "bad argument");
  if (h->cfg.arch != SERENC_ARCH_TEXT) SERENC_FAIL(SERENC_ERR_INVALID, "not a text-encoder handle");
  TextWs w;
  carve_text(h, batch, seq_len, nullptr, &w);
  *out_bytes = w.bytes + 256;
  return 0;
}

static int encode_text_impl(serenc_handle* h, const int32_t* input_ids_dev, const int32_t* valid_len, int batch, int seq_len,
                            uint64_t layer_mask, int reduce, const float* layer_weights, float* frames_out_dev,
                            float* pooled_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  SERENC_TRY(check_ready(h, SERENC_ARCH_TEXT));
  if (!input_ids_dev || !valid_len || !workspace_dev || batch <= 0 || seq_len <= 0) SERENC_FAIL(SERENC_ERR_INVALID, "bad argument");
  const serenc_config& c = h->cfg;
  const int d = c.hidden;
  if (seq_len + c.pad_token_id + 1 > c.max_positions)
    SERENC_FAIL(SERENC_ERR_INVALID, "sequence length %d exceeds the position table (%d rows, first position id %d)", seq_len, c.max_positions, c.pad_token_id + 1);
  if (c.layers < 63) layer_mask &= ((1ull << (c.layers + 1)) - 1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TextWs w;
  carve_text(h, batch, seq_len, workspace_dev, &w);
  if (w.bytes > workspace_bytes) SERENC_FAIL(SERENC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  const int64_t sumT = (int64_t)batch * seq_len;
  double attn_flops = 0.0;
  {
    void* hs;
    SERENC_TRY(stage_reserve(4 * (size_t)(2 * batch + 1), &hs, st));
    int32_t* hf = reinterpret_cast<int32_t*>(hs);
    for (int b = 0; b <= batch; ++b) hf[b] = b * seq_len;
    for (int b = 0; b < batch; ++b) {
      if (valid_len[b] < 1 || valid_len[b] > seq_len)
        SERENC_FAIL(SERENC_ERR_INVALID, "sequence %d: valid_len %d outside [1, %d]", b, (int)valid_len[b], seq_len);
      hf[batch + 1 + b] = valid_len[b];
      attn_flops += 4.0 * (double)seq_len * valid_len[b] * d;
    }
    SERENC_CUDA_OK(cudaMemcpyAsync(w.foff, hf, 4 * (size_t)(batch + 1), cudaMemcpyHostToDevice, st));
    SERENC_CUDA_OK(cudaMemcpyAsync(w.klen, hf + batch + 1, 4 * (size_t)batch, cudaMemcpyHostToDevice, st));
    SERENC_TRY(stage_commit(st));
  }
  {
    ProfScope ps(h, SERENC_PROF_MISC, 1, 0.0, (double)sumT * d * 12, st);
    text_embed_kernel<<<(unsigned)ceil_div64(sumT, 8), 256, 0, st>>>(input_ids_dev, w.klen, seq_len, sumT, d, c.vocab_size, c.max_positions,
                                                                    c.pad_token_id, h->emb_word, h->emb_pos, h->emb_type, w.sb.x);
    SERENC_CUDA_OK(cudaGetLastError());
  }
  w.sb.key_len = w.klen;
  EmitCtx e;
  SERENC_TRY(make_emit(&e, layer_mask, reduce, layer_weights, frames_out_dev, pooled_out_dev, w.acc, sumT, d, batch, w.foff, w.klen));
  // RobertaEmbeddings.LayerNorm is the stack's leading LayerNorm (loaded as final_ln.*); RobertaLayer is post-LN
  SERENC_TRY(run_stack_post_ln(h, w.sb, sumT, batch, seq_len, w.foff, attn_flops, e, st));
  return 0;
}
extern "C" int serenc_encode_text(serenc_handle* h, const int32_t* input_ids_dev, const int32_t* valid_len, int batch, int seq_len,
                                  uint64_t layer_mask, int reduce, const float* layer_weights, float* frames_out_dev,
                                  float* pooled_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  return guarded(h, [&] {
    return encode_text_impl(h, input_ids_dev, valid_len, batch, seq_len, layer_mask, reduce, layer_weights, frames_out_dev, pooled_out_dev,
                            workspace_dev, workspace_bytes, stream);
  });
}

// =================================================================================================
// diagnostic op-level entry points
// =================================================================================================
extern "C" int serenc_op_gemm(serenc_handle* h, const void* a, int64_t m, int64_t k, int64_t a_row_stride, const void* wt, int64_t n,
                              const float* bias, const float* resid, int act, float* out_f32, void* out_bf16, void* stream) {
  SERENC_TRY(check_ready(h, -1));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GemmCall g;
  if (a_row_stride == k || a_row_stride <= 0) {
    g = linear_call(reinterpret_cast<const bf16*>(a), m, (int)k, reinterpret_cast<const bf16*>(wt), (int)n);
  } else {
    // conv view: input rows of C = gcd-compatible channels; k = taps * C, a_row_stride = s * C, s in {1, 2}
    int taps = 0, C = 0, s = 0;
    for (int tps = 2; tps <= 4 && !taps; ++tps)
      for (int ss = 1; ss <= 2; ++ss)
        if (k % tps == 0 && (k / tps) * ss == a_row_stride) { taps = tps; C = (int)(k / tps); s = ss; break; }
    if (!taps || C % GEMM_BK) SERENC_FAIL(SERENC_ERR_INVALID, "op_gemm: cannot interpret k=%lld stride=%lld as a conv", (long long)k, (long long)a_row_stride);
    g.A = reinterpret_cast<const bf16*>(a); g.a_cols = C; g.a_rows = (m - 1) * s + taps; g.a_ld = C; g.a_stride = s;
    g.a_kpt = C / GEMM_BK; g.taps = taps; g.M = m; g.W = reinterpret_cast<const bf16*>(wt); g.w_rows = n; g.n_per_group = (int)n; g.groups = 1;
  }
  g.bias = bias; g.resid = resid; g.act = act; g.out_f32 = out_f32; g.ld_f32 = n; g.out_bf16 = reinterpret_cast<bf16*>(out_bf16); g.ld_bf16 = n;
  return launch_gemm(h, g, st);
}

extern "C" int serenc_op_gemm_grouped(serenc_handle* h, const void* x, int64_t rows, int groups, int cg_pad, int taps, const void* wt,
                                      int n_per_group, const float* bias, int act, float* out_f32, void* stream) {
  SERENC_TRY(check_ready(h, -1));
  if (cg_pad % GEMM_BK) SERENC_FAIL(SERENC_ERR_INVALID, "cg_pad must be a multiple of 64");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GemmCall g;
  const int64_t ld = (int64_t)groups * cg_pad;
  g.A = reinterpret_cast<const bf16*>(x); g.a_cols = ld; g.a_rows = rows; g.a_ld = ld; g.a_stride = 1; g.a_kpt = cg_pad / GEMM_BK; g.taps = taps;
  g.a_group_stride = cg_pad; g.M = rows - taps + 1; g.W = reinterpret_cast<const bf16*>(wt); g.w_rows = (int64_t)groups * n_per_group;
  g.n_per_group = n_per_group; g.groups = groups; g.bias = bias; g.act = act; g.out_f32 = out_f32; g.ld_f32 = (int64_t)groups * n_per_group;
  return launch_gemm(h, g, st);
}

extern "C" int serenc_op_layernorm(serenc_handle* h, const float* x, int64_t rows, int cols, const float* gamma, const float* beta,
                                   float eps, int gelu, float* out_f32, void* out_bf16, void* stream) {
  SERENC_TRY(check_ready(h, -1));
  if (!ln_width_ok(cols)) SERENC_FAIL(SERENC_ERR_INVALID, "layernorm: unsupported width %d", cols);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_f32) {
    if (gelu) SERENC_FAIL(SERENC_ERR_INVALID, "layernorm: gelu variant writes bf16 only");
    SERENC_TRY((launch_ln_t<float, float, false>(h, x, cols, out_f32, cols, gamma, beta, rows, cols, nullptr, nullptr, eps, st)));
  }
  if (out_bf16) {
    bf16* o = reinterpret_cast<bf16*>(out_bf16);
    if (gelu) SERENC_TRY((launch_ln_t<float, bf16, true>(h, x, cols, o, cols, gamma, beta, rows, cols, nullptr, nullptr, eps, st)));
    else SERENC_TRY((launch_ln_t<float, bf16, false>(h, x, cols, o, cols, gamma, beta, rows, cols, nullptr, nullptr, eps, st)));
  }
  return 0;
}

extern "C" int serenc_op_attention(serenc_handle* h, const void* qkv, const int64_t* frame_offsets, int batch, int wavlm, int layer,
                                   const void* hln, void* out, void* scratch_dev, void* stream) {
  SERENC_TRY(check_ready(h, -1));
  if (!qkv || !frame_offsets || !out || !scratch_dev || batch <= 0) SERENC_FAIL(SERENC_ERR_INVALID, "bad argument");
  if (wavlm && (!h->cfg.wavlm_rel_bias || layer < 0 || layer >= h->cfg.layers || !hln)) SERENC_FAIL(SERENC_ERR_INVALID, "wavlm attention needs a WavLM handle, a layer and hln");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int d = h->cfg.hidden;
  int32_t* d_off = reinterpret_cast<int32_t*>(scratch_dev);
  void* hs;
  SERENC_TRY(stage_reserve(4 * (size_t)(batch + 1), &hs, st));
  int tmax = 0;
  for (int b = 0; b <= batch; ++b) {
    reinterpret_cast<int32_t*>(hs)[b] = (int32_t)frame_offsets[b];
    if (b && frame_offsets[b] - frame_offsets[b - 1] > tmax) tmax = (int)(frame_offsets[b] - frame_offsets[b - 1]);
  }
  SERENC_CUDA_OK(cudaMemcpyAsync(d_off, hs, 4 * (size_t)(batch + 1), cudaMemcpyHostToDevice, st));
  SERENC_TRY(stage_commit(st));
  AttnParams p;
  p.qkv = reinterpret_cast<const bf16*>(qkv); p.ld_qkv = 3 * d; p.d = d; p.frame_off = d_off; p.out = reinterpret_cast<bf16*>(out);
  p.scale = 1.0f / sqrtf((float)h->head_dim);
  p.hln = reinterpret_cast<const bf16*>(hln);
  const LayerW& l = h->L[wavlm ? layer : 0];
  p.gru_w = l.gru_w; p.gru_b = l.gru_b; p.gru_const = l.gru_const; p.btab = h->btab;
  if (wavlm) SERENC_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&p.gate), sizeof(float) * (size_t)frame_offsets[batch] * h->cfg.heads, st));
  const int rc = launch_attn(h, p, wavlm != 0, tmax, batch, frame_offsets[batch], 0.0, st);
  if (wavlm) SERENC_CUDA_OK(cudaFreeAsync(p.gate, st));
  return rc;
}

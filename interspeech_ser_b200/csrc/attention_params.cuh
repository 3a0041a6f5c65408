// Parameter block shared by the attention kernels (attention_tc.cuh, attention_tc_wide.cuh and the A/B arm attention.cuh).
#pragma once
#include "common.cuh"

namespace serenc {

constexpr int WAVLM_MAXD = 1024;  // bias table covers delta in [-(MAXD-1), MAXD-1]; clamped beyond (buckets saturate at 778)

struct AttnParams {
  const bf16* qkv;   // [rows, ld_qkv]: q | k | v, each d wide, head h at columns h*HD
  int64_t ld_qkv;
  int d;             // model width (= H * HD)
  const int32_t* frame_off;  // [B+1]
  const int32_t* key_len = nullptr;  // [B] keys that take part (text encoder: the non-pad tokens; every row is still a query); nullptr = all rows
  bf16* out;         // [rows, d]
  float scale;       // head_dim^-0.5
  // WavLM only
  const bf16* hln;        // [rows, d] layer input (post-LN) the gate is computed from
  const float* gru_w;     // [8, HD]
  const float* gru_b;     // [8]
  const float* gru_const; // [H]
  const float* btab;      // [H, 2*WAVLM_MAXD-1]
  // tcgen05 kernel only
  float* gate = nullptr;       // [rows, heads] WavLM gate (LayerNorm epilogue or wavlm_gate_kernel writes it, attention_tc_kernel reads it)
  bool gate_ready = false;     // the LayerNorm that produced hln already filled `gate`
  int heads = 0, batch = 0;
  int ntile = 0;               // 128-query tiles per utterance = ceil(tmax / 128)
  int nwin = 0;                // stride of one bias-window buffer (entries)
  long long* trace = nullptr;  // debug (serenc_debug_gemm_trace): per-CTA clock stamps
};

}  // namespace serenc

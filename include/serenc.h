/* libserenc — C ABI of the B200-native speech-SSL encoder forward (embedding extraction hot path).
 *
 * This is the drop-in boundary for the ONE path of AI-Unicamp/interspeech_ser that this repository
 * accelerates: the encoder forward that preprocessing/preprocess_speech.py:48-67 and
 * preprocessing/preprocess_whisper.py:48-73 delegate to HuggingFace transformers
 *     processor(y, sampling_rate=16000, return_tensors="pt", padding=True)
 *     model(input_values=, attention_mask=, output_hidden_states=True)         (wav2vec2 / HuBERT / WavLM)
 *     processor(y, ...)["input_features"]; model.encoder(input_features, output_hidden_states=True)   (Whisper)
 * followed by hidden-state selection / mean-of-last-4 (preprocess_speech.py:56-67) and, in the
 * consumers, masked mean pooling over frames (lora_wavlm/model.py:189-195).
 *
 * Conventions
 *  - plain pointers and sizes only; every `*_dev` pointer is CUDA device memory on the handle's device,
 *    every other pointer is host memory. The caller owns inputs, outputs and the workspace; the library
 *    owns weights and immutable tables. No allocation happens inside the encode calls.
 *  - every entry point returns 0 (SERENC_OK) or a negative serenc_status; serenc_last_error() returns a
 *    thread-local human-readable message for the last failure on the calling thread.
 *  - a handle is immutable after serenc_finalize(); concurrent encode calls from several threads are safe
 *    when each call uses its own stream and workspace (the reference drives one model from 4 threads,
 *    preprocess_speech.py:120-122).
 *  - there is no CPU fallback: without a CUDA device of compute capability 10.x serenc_create fails.
 *  - kernel faults are sticky: once an entry point has seen a CUDA error that invalidates the context (illegal
 *    address, launch failure, ... - the errors CUDA itself reports on every later call) the handle is POISONED:
 *    every further call on it returns SERENC_ERR_CUDA with the original message, serenc_is_poisoned() returns 1,
 *    and only serenc_destroy() is useful. Kernels run asynchronously, so a fault surfaces at the next call on
 *    the handle or at serenc_sync(), which callers should use where the reference would have read a result.
 */
#ifndef SERENC_H_
#define SERENC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct serenc_handle serenc_handle;

typedef enum {
  SERENC_OK = 0,
  SERENC_ERR_INVALID = -1,   /* bad argument / unsupported configuration (HF would raise ValueError)   */
  SERENC_ERR_CUDA = -2,      /* CUDA runtime / driver failure (sticky: the handle is poisoned)          */
  SERENC_ERR_STATE = -3,     /* call order: tensor missing, handle not finalized, ...                   */
  SERENC_ERR_WORKSPACE = -4, /* workspace too small                                                     */
  SERENC_ERR_NO_DEVICE = -5  /* no sm_100 device: this library has no fallback path                      */
} serenc_status;

typedef enum {
  SERENC_ARCH_W2V = 0,     /* Wav2Vec2Model / HubertModel / WavLMModel (HF modeling_wav2vec2.py, modular_hubert.py, modeling_wavlm.py) */
  SERENC_ARCH_WHISPER = 1, /* WhisperEncoder (HF modeling_whisper.py:541-647) */
  SERENC_ARCH_TEXT = 2     /* RobertaModel: token/position/type embeddings + post-LN encoder (HF modeling_roberta.py;
                              preprocessing/preprocess_roberta.py:48-74 of the reference) */
} serenc_arch;

typedef enum {
  SERENC_REDUCE_NONE = 0,    /* emit every selected hidden state                                             */
  SERENC_REDUCE_MEAN = 1,    /* emit the mean over the selected hidden states (preprocess_speech.py:56-63)   */
  SERENC_REDUCE_WEIGHTED = 2 /* emit sum_i w_i * hidden_states[sel_i] with caller-supplied weights (the softmax-
                                weighted layer sum of lora_wavlm/model.py:164-181; the caller applies the softmax) */
} serenc_reduce;

typedef enum {
  SERENC_WAV_F32 = 0, /* float32 samples (what librosa.load returns, preprocess_speech.py:47)                */
  SERENC_WAV_I16 = 1  /* int16 PCM as stored in the WAV file; the kernels apply librosa's x / 32768 on load,  */
                      /* bit-identical to the float path at half the host decode work and H2D bytes          */
} serenc_wav_dtype;

/* Architecture constants (the fields of the HF config.json this path depends on). */
typedef struct {
  int32_t arch;                 /* serenc_arch */
  int32_t hidden;               /* hidden_size / d_model                  */
  int32_t layers;               /* num_hidden_layers / encoder_layers      */
  int32_t heads;                /* num_attention_heads; head_dim in {64, 80, 120} */
  int32_t ffn;                  /* intermediate_size / encoder_ffn_dim     */
  int32_t conv_dim;             /* W2V: 512 (all 7 feature-encoder layers) */
  int32_t conv_bias;            /* W2V: config.conv_bias                   */
  int32_t wavlm_rel_bias;       /* W2V: 1 = WavLM gated relative position bias */
  int32_t num_buckets;          /* WavLM: 320                              */
  int32_t max_distance;         /* WavLM: 800                              */
  int32_t pos_conv_kernel;      /* W2V: num_conv_pos_embeddings (128)      */
  int32_t pos_conv_groups;      /* W2V: num_conv_pos_embedding_groups (16) */
  int32_t n_mels;               /* Whisper: 128                            */
  int32_t max_source_positions; /* Whisper: 1500                           */
  float layer_norm_eps;         /* 1e-5                                    */
  int32_t conv_group_norm;      /* W2V: 1 = feat_extract_norm "group" (base checkpoints): GroupNorm on conv0 only */
  int32_t post_layer_norm;      /* W2V: 1 = do_stable_layer_norm false (base checkpoints): post-LN encoder layers  */
  int32_t no_feat_proj_ln;      /* W2V: 1 = feature projection without LayerNorm (HuBERT-base)                     */
  int32_t vocab_size;           /* TEXT: rows of the word embedding (roberta: 50265)                               */
  int32_t max_positions;        /* TEXT: rows of the position embedding (roberta: 514)                             */
  int32_t type_vocab_size;      /* TEXT: rows of the token-type embedding (roberta: 1)                             */
  int32_t pad_token_id;         /* TEXT: padding_idx; position ids start at pad_token_id + 1 (roberta: 1)          */
  int32_t reserved[1];
} serenc_config;

/* ---- lifecycle ------------------------------------------------------------------------------------ */

/* Replaces AutoModel.from_pretrained(...).eval().to(device) (preprocess_speech.py:112-114): creates an empty
 * model on CUDA device `device`. Weights are then supplied tensor by tensor and frozen by serenc_finalize. */
int serenc_create(const serenc_config* cfg, int device, serenc_handle** out);
int serenc_destroy(serenc_handle* h);

/* Upload one fp32 tensor by canonical name (see INTEGRATION.md for the name table, e.g.
 * "conv3.weight", "featproj.ln.bias", "posconv.weight" (weight-norm already folded), "layer7.q.weight",
 * "layer0.gru.weight", "rel_attn_embed", "final_ln.weight", "embed_positions", "mel_filters").
 * Layout conversion (bf16 cast, QKV concatenation, conv tap-major packing) happens here. */
int serenc_load_tensor(serenc_handle* h, const char* name, const float* data, const int64_t* shape, int ndim);

/* Checks that every tensor the configuration needs was loaded and builds the derived tables
 * (WavLM bucket->Toeplitz bias vectors, Hann/DFT twiddles, mel filterbank CSR). */
int serenc_finalize(serenc_handle* h);

const char* serenc_last_error(void);
const char* serenc_version(void);

/* Waits for everything enqueued on `stream` and reports (and records, see "sticky" above) a kernel fault. */
int serenc_sync(serenc_handle* h, void* stream);
/* 1 once a sticky CUDA error was observed through this handle, else 0. */
int serenc_is_poisoned(const serenc_handle* h);

/* ---- wav2vec2 / HuBERT / WavLM -------------------------------------------------------------------- */

/* Frames produced for `n_samples` input samples (HF _get_feat_extract_output_lengths,
 * modeling_wavlm.py:640-659); <= 0 when the input is shorter than the 400-sample receptive field. */
int64_t serenc_w2v_num_frames(int64_t n_samples);

/* WavLMAttention._relative_positions_bucket (HF modeling_wavlm.py:253-271) for delta = key - query position.
 * Pure host function; the per-head bias vectors are built from it in serenc_finalize. */
int serenc_wavlm_bucket(int delta, int num_buckets, int max_distance);

/* Bytes of workspace serenc_encode_w2v needs for these utterance lengths. */
int serenc_w2v_workspace_bytes(const serenc_handle* h, const int32_t* sample_len, int batch, size_t* out_bytes);

/* Replaces Wav2Vec2FeatureExtractor.__call__ (HF feature_extraction_wav2vec2.py:99-236, zero_mean_unit_var_norm
 * :77-97): per-utterance (x - mean) / sqrt(var + 1e-7) over the valid samples, right padding written as 0.
 * Utterance b reads wav_dev[sample_start[b] .. +sample_len[b]) and writes out_dev[b*out_stride .. +out_len). */
int serenc_wav_normalize(serenc_handle* h, const float* wav_dev, const int64_t* sample_start,
                         const int32_t* sample_len, int batch, float* out_dev, int64_t out_stride, int32_t out_len,
                         void* stream);

/* Replaces model(input_values, attention_mask, output_hidden_states=True) + layer selection + pooling
 * (preprocess_speech.py:48-67). Utterance b is wav_dev[sample_start[b] .. +sample_len[b]) (padded batches
 * and packed batches are both expressible). Frames are produced in a packed layout: utterance b owns rows
 * frame_offsets[b] .. frame_offsets[b+1] of every [sum_T, hidden] matrix.
 *   normalize   1: apply the feature extractor's normalisation inside the first conv kernel; 0: input_values
 *               are already normalised (HF model semantics).
 *   layer_mask  bit i selects hidden_states[i], i in [0, layers]; [0] = encoder input after the positional
 *               conv, [layers] = after the final LayerNorm (HF hidden_states tuple semantics).
 *   reduce      SERENC_REDUCE_NONE: frames_out_dev is [n_sel, sum_T, hidden], pooled_out_dev [n_sel, batch, hidden]
 *               (ascending layer index); SERENC_REDUCE_MEAN: [sum_T, hidden] and [batch, hidden].
 *   frames_out_dev / pooled_out_dev   fp32, either may be NULL. pooled = masked mean over the valid frames.
 *   frame_offsets_out                 host, [batch + 1], may be NULL.
 */
int serenc_encode_w2v(serenc_handle* h, const float* wav_dev, const int64_t* sample_start, const int32_t* sample_len,
                      int batch, int normalize, uint64_t layer_mask, int reduce, float* frames_out_dev,
                      float* pooled_out_dev, int64_t* frame_offsets_out, void* workspace_dev, size_t workspace_bytes,
                      void* stream);

/* The same call with every optional input / output (a zero-initialised struct with the fields of serenc_encode_w2v
 * set behaves exactly like it):
 *   wav_dtype                 serenc_wav_dtype of wav_dev (sample_start counts samples either way)
 *   layer_weights             host, [number of selected layers], ascending layer index; SERENC_REDUCE_WEIGHTED only
 *   extract_features_out_dev  fp32 [sum_T, conv_dim]: the feature projection's LayerNorm output, HF's
 *                             `extract_features` (modeling_wavlm.py:93-105); NULL = not wanted. Needs a
 *                             feature-projection LayerNorm (HubertModel returns no extract_features). */
typedef struct {
  const void* wav_dev;
  int32_t wav_dtype;
  int32_t batch;
  const int64_t* sample_start;
  const int32_t* sample_len;
  int32_t normalize;
  int32_t reduce;
  uint64_t layer_mask;
  const float* layer_weights;
  float* frames_out_dev;
  float* pooled_out_dev;
  float* extract_features_out_dev;
  int64_t* frame_offsets_out;
  void* workspace_dev;
  size_t workspace_bytes;
  void* stream;
} serenc_w2v_call;
int serenc_encode_w2v_ex(serenc_handle* h, const serenc_w2v_call* call);

/* packed [sum_T, hidden] -> HF-shaped padded [batch, t_max, hidden] (pad frames = 0). */
int serenc_unpack_frames(serenc_handle* h, const float* packed_dev, const int64_t* frame_offsets, int batch,
                         int32_t t_max, float* out_dev, void* stream);
/* The same for rows of any width (`cols` % 4 == 0), e.g. extract_features [sum_T, 512]. */
int serenc_unpack_rows(serenc_handle* h, const float* packed_dev, const int64_t* frame_offsets, int batch,
                       int32_t t_max, int32_t cols, float* out_dev, void* stream);

/* ---- Whisper -------------------------------------------------------------------------------------- */

/* Replaces WhisperFeatureExtractor.__call__ -> _torch_extract_fbank_features (HF feature_extraction_whisper.py
 * :135-164, :296-320): pad/truncate to 480000 samples, log-mel, per-utterance dynamic-range clamp.
 * mel_out_dev is [batch, n_mels, 3000] fp32. scratch_dev: at least 64 + 40*batch bytes. */
int serenc_logmel(serenc_handle* h, const float* wav_dev, const int64_t* sample_start, const int32_t* sample_len,
                  int batch, float* mel_out_dev, void* scratch_dev, void* stream);

/* serenc_logmel for int16 PCM or float32 samples (wav_dtype = serenc_wav_dtype). */
int serenc_logmel_ex(serenc_handle* h, const void* wav_dev, int wav_dtype, const int64_t* sample_start,
                     const int32_t* sample_len, int batch, float* mel_out_dev, void* scratch_dev, void* stream);

int serenc_whisper_workspace_bytes(const serenc_handle* h, int batch, size_t* out_bytes);

/* Replaces model.encoder(input_features, output_hidden_states=True) + selection + crop + pooling
 * (preprocess_whisper.py:57-76). mel_dev is [batch, n_mels, 3000]; every utterance yields 1500 frames.
 * n_keep (host, [batch], may be NULL) = frames that count for pooling (the script keeps
 * min(ceil(len/320), 1280) frames, preprocess_whisper.py:49-50,75-76). Outputs as serenc_encode_w2v with
 * sum_T = 1500 * batch. */
int serenc_encode_whisper(serenc_handle* h, const float* mel_dev, int batch, uint64_t layer_mask, int reduce,
                          const int32_t* n_keep, float* frames_out_dev, float* pooled_out_dev, void* workspace_dev,
                          size_t workspace_bytes, void* stream);

/* serenc_encode_whisper + SERENC_REDUCE_WEIGHTED: layer_weights is host, [number of selected layers]. */
int serenc_encode_whisper_ex(serenc_handle* h, const float* mel_dev, int batch, uint64_t layer_mask, int reduce,
                             const float* layer_weights, const int32_t* n_keep, float* frames_out_dev,
                             float* pooled_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- text encoder (RoBERTa) ------------------------------------------------------------------------ */

int serenc_text_workspace_bytes(const serenc_handle* h, int batch, int seq_len, size_t* out_bytes);

/* Replaces RobertaModel(**encoding, output_hidden_states=True) + layer selection (preprocess_roberta.py:48-70):
 * input_ids_dev is [batch, seq_len] int32 (the tokenizer pads to max_length = 80, :49-55); valid_len (host,
 * [batch]) = number of leading non-pad tokens of each row, i.e. attention_mask.sum(-1) of a right-padded batch.
 * Position ids are pad_token_id + 1 + t for t < valid_len and pad_token_id after (HF
 * create_position_ids_from_input_ids), token types are 0, keys >= valid_len are masked out of every attention.
 * ALL seq_len rows of every sequence are produced, pad positions included, as the reference saves them:
 * frames_out_dev is [n_sel, batch * seq_len, hidden] (or [batch * seq_len, hidden] when reducing); hidden_states[0]
 * is the embedding output after its LayerNorm. pooled_out_dev = mean over the valid_len leading rows. */
int serenc_encode_text(serenc_handle* h, const int32_t* input_ids_dev, const int32_t* valid_len, int batch, int seq_len,
                       uint64_t layer_mask, int reduce, const float* layer_weights, float* frames_out_dev,
                       float* pooled_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- launch accounting / per-kernel-class device timing (used by bench.py for the roofline numbers) ---- */

typedef enum {
  SERENC_PROF_GEMM_LINEAR = 0, /* other Linear layers (feature projection) (tcgen05)           */
  SERENC_PROF_GEMM_CONV = 1,   /* conv1..6 / Whisper conv stem as implicit GEMMs (tcgen05)     */
  SERENC_PROF_GEMM_POSCONV = 2,/* grouped positional conv as implicit GEMM (tcgen05)           */
  SERENC_PROF_ATTENTION = 3,
  SERENC_PROF_LAYERNORM = 4,
  SERENC_PROF_CONV0 = 5,
  SERENC_PROF_POOL = 6,        /* hidden-state accumulation + masked mean pooling               */
  SERENC_PROF_LOGMEL = 7,
  SERENC_PROF_MISC = 8,
  SERENC_PROF_GEMM_QKV = 9,    /* fused q|k|v projection, bias, bf16 out                        */
  SERENC_PROF_GEMM_OUT = 10,   /* attention out projection, bias + fp32 residual                */
  SERENC_PROF_GEMM_FC1 = 11,   /* FFN up projection, bias + exact GELU, bf16 out                */
  SERENC_PROF_GEMM_FC2 = 12,   /* FFN down projection, bias + fp32 residual                     */
  SERENC_PROF_NUM_CLASSES = 13
} serenc_prof_class;

/* Kernel launches (incl. memset nodes) issued through this handle since creation. */
int64_t serenc_launch_count(const serenc_handle* h);
/* enable != 0: bracket every launch with CUDA events on its stream (single-threaded use); clears old records. */
int serenc_set_profiling(serenc_handle* h, int enable);
/* Synchronises and sums, per class, device milliseconds / algorithmic FLOPs / algorithmic bytes / launches. */
int serenc_get_profile(serenc_handle* h, int n_classes, double* ms, double* flops, double* bytes, int64_t* launches);

/* Debug: per-tile clock64 stamps of CTA 0 of the CTA-pair GEMM are written to dev_buf ([tiles][8] int64); NULL disables. */
int serenc_debug_gemm_trace(serenc_handle* h, void* dev_buf);

/* ---- diagnostic entry points (op-level known-answer tests; not needed by an integrator) ------------ */

/* out = epilogue(A[M,K] * W[N,K]^T): bf16 operands, fp32 accumulate. a_row_stride (elements) may be smaller
 * than K (overlapping rows = implicit-GEMM conv). act: 0 none, 1 exact GELU. Any of bias/resid/out_f32/out_bf16
 * may be NULL. */
int serenc_op_gemm(serenc_handle* h, const void* a_bf16_dev, int64_t m, int64_t k, int64_t a_row_stride,
                   const void* w_bf16_dev, int64_t n, const float* bias_dev, const float* resid_dev, int act,
                   float* out_f32_dev, void* out_bf16_dev, void* stream);

/* Grouped temporal conv as a rank-3 implicit GEMM: x is [rows, groups*cg_pad] bf16 (cg_pad % 64 == 0),
 * w is [groups*n_per_group, taps*cg_pad] bf16, out[m, g*n_per_group + j] = sum_{tap,c} x[m+tap, g*cg_pad+c] * w[..]. */
int serenc_op_gemm_grouped(serenc_handle* h, const void* x_bf16_dev, int64_t rows, int groups, int cg_pad, int taps,
                           const void* w_bf16_dev, int n_per_group, const float* bias_dev, int act,
                           float* out_f32_dev, void* stream);

int serenc_op_layernorm(serenc_handle* h, const float* x_dev, int64_t rows, int cols, const float* gamma_dev,
                        const float* beta_dev, float eps, int gelu, float* out_f32_dev, void* out_bf16_dev,
                        void* stream);

/* qkv: [sum_T, 3*hidden] bf16; out: [sum_T, hidden] bf16. wavlm != 0 uses layer `layer`'s gate weights and the
 * handle's bias table; hln_dev = the layer input the gate is computed from. */
int serenc_op_attention(serenc_handle* h, const void* qkv_bf16_dev, const int64_t* frame_offsets, int batch,
                        int wavlm, int layer, const void* hln_bf16_dev, void* out_bf16_dev, void* scratch_dev,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SERENC_H_ */

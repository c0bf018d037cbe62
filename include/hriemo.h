/* hriemo.h — C ABI of libhriemo_b200.so: the B200 (sm_100a) kernels behind the
 * HRI-EMO fusion-and-decode forward path.
 *
 * The reference (Makiato1999/HRI-EMO) is pure Python/PyTorch; its "FFI" for this
 * path is the set of torch.nn calls made by models/*.py.  Each entry point below
 * names the reference call group it replaces (paths relative to the reference
 * repository root).  The Python nn.Modules in hri-emo_b200/models/ are the only
 * intended callers (ctypes binding in hri-emo_b200/hriemo/lib.py; the binding a
 * reference maintainer would add is shown in INTEGRATION.md).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless stated otherwise; the caller owns
 *    all buffers; the library never allocates caller-visible memory, never
 *    synchronises the host, and enqueues on the `stream` handle it is given
 *    (a cudaStream_t passed as void*; NULL = legacy default stream);
 *  - "bf16" is a 16-bit bfloat16, "f32" IEEE binary32; matrices are row-major
 *    with an explicit leading dimension in ELEMENTS;
 *  - masks are uint8 (torch.bool storage), 1 = PAD (ignored key), like the
 *    reference's key_padding_mask (models/cross_modal_block_tacfn.py:66-67);
 *  - return value: 0 = ok, negative = error; hriemo_last_error() gives the text
 *    of the last error raised on the calling thread.  No C++ exception crosses
 *    the boundary.
 */
#ifndef HRIEMO_H_
#define HRIEMO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRIEMO_VERSION 100 /* 0.1.0 */

enum {
  HRIEMO_OK = 0,
  HRIEMO_ERR_INVALID = -1, /* bad shape / alignment / argument */
  HRIEMO_ERR_CUDA = -2,    /* a CUDA runtime / driver call failed */
  HRIEMO_ERR_DEVICE = -3   /* not an sm_100 device */
};

int hriemo_version(void);
const char* hriemo_last_error(void);
/* Number of kernels this library has launched from the calling process so far
 * (monotonic; bench.py reports the delta over its timed region). */
int64_t hriemo_launch_count(void);

/* ---------------------------------------------------------------- cast ----
 * fp32 features -> bf16 operand, zero-padding each row from `cols` to `ld_out`
 * (TMA needs 16-byte row pitches; MOSEI d_audio=74 / d_text=300 are padded).
 * Replaces the implicit autocast of models/mosei_fusion_with_emotion_decoder.py:64-66
 * inputs and the `.float()` feature load (scripts/fusion/train_fusion_seq_level_decoder.py:151).
 */
int hriemo_cast_f32_to_bf16(const float* in, int64_t ld_in, void* out_bf16, int64_t ld_out,
                            int64_t rows, int32_t cols, void* stream);

/* ---------------------------------------------------------------- GEMM ----
 * out = epilogue(A[M,K] . W[N,K]^T + bias) on tcgen05 tensor cores (TMA-fed,
 * fp32 accumulation in TMEM).  W is exactly nn.Linear.weight ([out,in]).
 * Replaces every nn.Linear / MHA in-projection / out-projection on the path:
 *   models/cross_modal_block_tacfn.py:24-52 (MHA in/out projections, ffn_a, ffn_t),
 *   models/emotion_decoder.py:14,20,25-27, models/mosei_fusion_with_emotion_decoder.py:41-42.
 */
enum {
  HRIEMO_EPI_BIAS = 0,           /* out bf16 = acc + bias                                   */
  HRIEMO_EPI_BIAS_RELU = 1,      /* out bf16 = relu(acc + bias)            (FFN first half) */
  HRIEMO_EPI_BIAS_RESID = 2,     /* out bf16 = acc + bias + resid(bf16)    (pre-LayerNorm)  */
  HRIEMO_EPI_BIAS_RESID_F32 = 3, /* out f32  = acc + bias + resid(f32)     (decoder stream) */
  HRIEMO_EPI_BIAS_F32 = 5,       /* out f32  = acc + bias   (4 is retired: a transposed-V epilogue) */
  HRIEMO_EPI_BIAS_MASK = 6       /* out bf16 = resid(bf16) > 0 ? acc + bias : 0: the input gradient of a Linear whose INPUT
                                    was relu(.) -- resid is that post-ReLU tensor (models/cross_modal_block_tacfn.py:46
                                    under loss.backward()) */
};

typedef struct hriemo_gemm_args {
  const void* A;     /* bf16 [M,K]   */
  int64_t lda;       /* multiple of 8 */
  const void* W;     /* bf16 [N,K]   */
  int64_t ldw;       /* multiple of 8 */
  const float* bias; /* f32 [N] or NULL */
  int64_t M;
  int32_t N;         /* multiple of 32 */
  int32_t K;         /* multiple of 8  */
  int32_t epilogue;  /* HRIEMO_EPI_*   */
  int32_t cta_pair;  /* 0 = library picks; 1 = one-CTA tiles (128 x BN); 2 = CTA-pair tiles
                        (256 x 256, tcgen05 cta_group::2; needs N % 256 == 0) */
  void* out;         /* bf16 or f32, see epilogue */
  int64_t ldo;
  const void* resid; /* bf16 or f32 [M,N], RESID modes only */
  int64_t ldr;
  /* Fused LayerNorm (all optional, bf16-output epilogues only).  It replaces the separate
   * nn.LayerNorm pass after each residual (models/cross_modal_block_tacfn.py:81,92,105,106,118,119):
   * the pre-LayerNorm sum x is what lives in HBM, and LN is applied where x is consumed.
   *  - a_stats/a_colsum: the rows of A are pre-LayerNorm; W must already be W*gamma (columnwise),
   *    bias must be b + W.beta, a_colsum[n] = sum_k W[n,k]; then
   *      out = rstd[m] * (acc - mean[m] * a_colsum[n]) + bias[n]  ==  LN(A) . W_orig^T + b.
   *  - resid_stats/resid_gamma/resid_beta: the residual is pre-LayerNorm; LN(resid) is added.
   *  - stats_out: [ceil(N/64)][M][2] f32, per 64-column slab (sum, sum of squares) of the fp32
   *    outputs of each row; hriemo_ln_stats_finalize turns them into (mean, rstd). */
  const float* a_stats;     /* [M][2] (mean, rstd) */
  const float* a_colsum;    /* [N] */
  const float* resid_stats; /* [M][2] (mean, rstd) */
  const float* resid_gamma; /* [N] */
  const float* resid_beta;  /* [N] */
  float* stats_out;
} hriemo_gemm_args;

int hriemo_gemm_bf16(const hriemo_gemm_args* args, void* stream);

/* ----------------------------------------------------------- attention ----
 * O = softmax(Q K^T * scale + key_padding) V per (utterance, head), flash-style
 * on tcgen05 (S and O accumulate in TMEM, online softmax in registers).
 * Replaces the scaled_dot_product_attention inside nn.MultiheadAttention at
 *   models/cross_modal_block_tacfn.py:74-80, 85-91, 98-104, 111-117 and
 *   models/cross_modal_block.py:56-59, 64-67.
 * Q rows are (b, t_q) with row pitch ldq, head h at columns [h*dh, (h+1)*dh);
 * K and V likewise over t_k (row pitches ldk, ldv): all three may be column
 * slices of one packed [Q|K|V] projection output, exactly PyTorch's layout.
 * A fully masked key row yields NaN, as torch.softmax does in the reference.
 */
typedef struct hriemo_attn_args {
  const void* q;  int64_t ldq;   /* bf16 */
  const void* k;  int64_t ldk;   /* bf16 */
  const void* v;  int64_t ldv;   /* bf16 */
  const uint8_t* key_pad;        /* [B, Tk] 1 = PAD, or NULL */
  void* out;      int64_t ldo;   /* bf16 [B*Tq, H*dh] */
  int32_t B, H, Tq, Tk, dh;      /* dh in {32, 64, 96, 128} */
  float scale;                   /* 1/sqrt(dh) */
  /* Optional, with key_pad: kv_steps[b] = number of 64-key tiles to process for utterance b (from
   * hriemo_attention_kv_steps).  Trailing tiles that hold only PAD keys are skipped; the result is
   * bit-identical (a masked key contributes exactly 0).  This is the padded-tensor form of the
   * reference's collate (scripts/fusion/train_fusion_seq_level_decoder.py:191-232: zero-pad to the
   * batch maximum + True=PAD mask) without paying for the padding in the attention. */
  const int32_t* kv_steps;
  /* Tq <= 128 with an even H runs two heads per work item (one per query tile of the CTA) instead of leaving
   * the second tile idle; same result bit for bit.  Non-zero switches that off (A/B measurements, tests). */
  int32_t no_head_pairs;
  /* Optional: lse[b, h, t_q] = ln sum_k exp(scale * q.k) over the unmasked keys (f32 [B, H, Tq]; -inf when every key
   * is masked) -- the per-row statistic a backward pass needs to rebuild the probabilities. */
  float* lse;
  /* Training only: dropout on the attention probabilities (nn.MultiheadAttention(dropout=p) of the reference).  drop_p8
   * = round(256 p) (0 = off), drop_scale = 1 / (1 - drop_p8 / 256), drop_key = the stream key of this attention
   * (csrc/dropout.cuh: probability (b, h, q, k) is kept iff byte (k & 3) of drop_word(drop_key_bh(key, b * H + h), q,
   * k >> 2) >= drop_p8).  The row sums (and lse) are those of the undropped probabilities, as in torch. */
  uint32_t drop_p8;
  uint32_t drop_key;
  float drop_scale;
} hriemo_attn_args;

/* steps[b] = (index of the last valid key of utterance b) / 64 + 1, or 1 when every key is PAD. */
int hriemo_attention_kv_steps(const uint8_t* key_pad, int32_t B, int32_t Tk, int32_t* steps, void* stream);

int hriemo_attention_bf16(const hriemo_attn_args* args, void* stream);

/* Head-averaged attention probabilities mean_h softmax(...)[B,Tq,Tk] (f32) — the
 * `need_weights=True` / return_attention=True side output of
 * models/cross_modal_block_tacfn.py:79,90,103,116 and models/emotion_decoder.py:53.
 * V is not needed.  Interpretability path; CUDA-core kernel, not tuned. */
int hriemo_attention_probs(const void* q, int64_t ldq, const void* k, int64_t ldk,
                           const uint8_t* key_pad, float* probs, int32_t B, int32_t H, int32_t Tq,
                           int32_t Tk, int32_t dh, float scale, void* stream);

/* Small-query attention for the emotion decoder (N_q learned queries over Tk
 * keys; models/emotion_decoder.py:42 and :48-54).  K and V are row-major bf16
 * ([B*Tk, ld], head h at columns h*dh..).  One CTA per utterance, K/V head
 * slices staged through shared memory.  probs (optional, f32 [B,Nq,Tk]) receives
 * the head-averaged weights. */
int hriemo_small_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                           int64_t ldv, const uint8_t* key_pad, void* out_bf16, int64_t ldo,
                           float* probs, int32_t B, int32_t H, int32_t Nq, int32_t Tk, int32_t dh,
                           float scale, void* stream);

/* ----------------------------------------------------------- LayerNorm ----
 * y = LayerNorm(x) over the last dim (biased variance, eps inside the sqrt),
 * x bf16 or f32 (already holding residual + sub-layer output), y written as
 * bf16 and/or f32 (either may be NULL).  Replaces nn.LayerNorm at
 *   models/cross_modal_block_tacfn.py:81,92,105,106,118,119, models/emotion_decoder.py:43,55,59,
 *   models/fusion_classifier.py:73.
 */
int hriemo_layernorm(const void* x, int32_t x_is_f32, int64_t ldx, const float* gamma,
                     const float* beta, float eps, void* y_bf16, float* y_f32, int64_t ldy,
                     int64_t rows, int32_t d, void* stream);

/* (sum, sum of squares) partials written by hriemo_gemm_bf16 (stats_out) -> stats[m] = (mean,
 * rstd = 1/sqrt(biased variance + eps)) over d columns. */
int hriemo_ln_stats_finalize(const float* partials, int32_t n_slabs, int64_t rows, int32_t d, float eps,
                             float* stats, void* stream);

/* One-time operand preparation for a GEMM that consumes LN(x) given pre-LayerNorm x (a_stats mode):
 * w_folded[n,k] = bf16(W[n,k] * gamma[k]),  colsum[n] = sum_k float(w_folded[n,k]),
 * bias_folded[n] = bias[n] + sum_k W[n,k] * beta[k]   (bias may be NULL = 0). */
int hriemo_fold_ln_weight(const float* W, int64_t ldw, const float* gamma, const float* beta,
                          const float* bias, void* w_folded_bf16, int64_t ldo, float* colsum,
                          float* bias_folded, int32_t N, int32_t K, void* stream);

/* -------------------------------------------------------------- β-gate ----
 * models/beta_gate_tacfn.py:79-84 (and models/beta_gate.py:81-82 with
 * apply_ln = 0): pooled[b,:] = masked_mean_t(LN(x[b,t,:])).  x bf16 [B,T,d].
 */
/* pre_gamma/pre_beta (optional): x is itself a pre-LayerNorm tensor; LN_pre is applied to each row
 * first (the encoder's last LayerNorm, models/cross_modal_block_tacfn.py:106,119, fused here).
 * pre_stats (optional, [B*T][2] = (mean, rstd) as written by hriemo_ln_stats_finalize) spares the
 * kernel the statistics of LN_pre. */
int hriemo_ln_masked_mean(const void* x_bf16, int64_t ldx, const float* gamma, const float* beta,
                          float eps, int32_t apply_ln, const uint8_t* pad, float* pooled,
                          int64_t ld_pooled, int32_t B, int32_t T, int32_t d, const float* pre_gamma,
                          const float* pre_beta, const float* pre_stats, void* stream);

/* models/beta_gate_tacfn.py:87-89: g = [a, t, |a-t|, a*t]  (f32 [B,4d]). */
int hriemo_gate_input(const float* a_pool, const float* t_pool, float* g, int32_t B, int32_t d,
                      void* stream);

/* Small fp32 GEMM on CUDA cores: out = act(A[M,K] . W[N,K]^T + bias), act in
 * {0 none, 1 relu, 2 sigmoid}; act | 4: ReLU is applied to A as it is read.  Used where the reference result feeds a
 * bit-sensitive decision (gate MLP models/beta_gate_tacfn.py:62-66,92; emotion
 * head models/emotion_decoder.py:155; classifier models/fusion_classifier.py:74-77). */
int hriemo_sgemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                     float* out, int64_t ldo, int64_t M, int32_t N, int32_t K, int32_t act,
                     void* stream);

/* models/beta_gate_tacfn.py:95-116: beta[b] = mean_d w[b,:];
 * h[b,t,:] = w[b,:]*LN_a(a[b,t,:]) + (1-w[b,:])*LN_t(t[b,t,:]) for t < L.
 * With apply_ln = 0 and w_is_scalar = 1 it is the legacy scalar gate
 * (models/beta_gate.py:103-112; beta_out then just copies w).  pre_gamma_x / pre_beta_x (optional):
 * that stream is a pre-LayerNorm tensor and LN_pre is applied to its rows first; pre_stats_a is indexed
 * [b*T_a + t], pre_stats_t [b*L + t] (optional, as for hriemo_ln_masked_mean).
 * a is [B,T_a,d] bf16 (row pitch lda; only the first L rows are read), t is [B,L,d]. */
int hriemo_gate_blend(const void* a_bf16, int64_t lda, int32_t T_a, const void* t_bf16, int64_t ldt,
                      const float* gamma_a, const float* beta_a, const float* gamma_t,
                      const float* beta_t, float eps, int32_t apply_ln, const float* w,
                      int32_t w_is_scalar, void* h_bf16, float* h_f32, int64_t ldh, float* beta_out,
                      int32_t B, int32_t L, int32_t d, const float* pre_gamma_a, const float* pre_beta_a,
                      const float* pre_gamma_t, const float* pre_beta_t, const float* pre_stats_a,
                      const float* pre_stats_t, void* stream);

/* Post-path outputs of the inference script (scripts/infer/mosei_eval_infer.py:237-270,
 * scripts/analysis/mosei_summary_metrics.py:51): probs = sigmoid(logits) [B, n_classes] f32 and
 * decisions = probs >= thresholds[class] (uint8; thresholds NULL = 0.5 for every class).
 * Either output may be NULL. */
int hriemo_emotion_outputs(const float* logits, const float* thresholds, float* probs, uint8_t* decisions,
                           int64_t B, int32_t n_classes, void* stream);

/* ------------------------------------------------- length-bucketed staging ----
 * The step BEFORE the path.  The reference's collate zero-pads every
 * utterance to the batch maximum and marks the tail True = PAD
 * (scripts/fusion/train_fusion_seq_level_decoder.py:191-232, scripts/infer/mosei_eval_infer.py:128-147).
 * Padded rows never reach logits / beta / z, so hriemo.pipeline sorts utterances by valid length
 * and runs slabs trimmed to their own maximum; these entry points build and un-build such slabs.
 */
/* lens[b] = (index of the last entry of pad[b, :] that is 0) + 1, or 0 when every entry is PAD. */
int hriemo_mask_lengths(const uint8_t* pad, int32_t B, int32_t T, int32_t* lens, void* stream);

/* out[i, t, :] = bf16(in[utt[i], t, :]) for t < min(T_in, T_out); rows T_in <= t < T_out and columns
 * cols <= c < ld_out are zero.  in is [B, T_in, ld_in] f32 (in_is_f32 = 1) or bf16; utt is int32 [n]. */
int hriemo_gather_utterances_bf16(const void* in, int32_t in_is_f32, int64_t ld_in, int32_t T_in,
                                  const int32_t* utt, void* out_bf16, int64_t ld_out, int32_t n,
                                  int32_t T_out, int32_t cols, void* stream);

/* out[i, t] = pad[utt[i], t] for t < min(T_in, T_out), 1 (PAD) for T_in <= t < T_out. */
int hriemo_gather_masks(const uint8_t* pad, int32_t T_in, const int32_t* utt, uint8_t* out, int32_t n,
                        int32_t T_out, void* stream);

/* out[utt[i], :] = in[i, :] (f32, `cols` contiguous elements per row): a slab's logits / beta / z
 * go back to the utterances' original positions. */
int hriemo_scatter_rows_f32(const float* in, const int32_t* utt, float* out, int64_t n, int64_t cols,
                            void* stream);

/* HOST function: every pointer is a HOST pointer, no device work, no stream.  Same gather + trim as
 * hriemo_gather_utterances_bf16 for a padded fp32 batch that lives in host memory, written as bf16 into
 * (pinned) staging memory by `n_threads` C++ threads: dst[i, t, :] = bf16_rne(src[utt[i], t, :]) for
 * t < min(lens[i], T_in, T_out), zero elsewhere.  utt NULL = identity, lens NULL = T_in for everyone.
 * n_src = utterances the source batch holds: every utt[i] (or i) must lie in [0, n_src), checked before any
 * row is touched (HRIEMO_ERR_INVALID otherwise).  Rounding is round-to-nearest-even like the GPU cast
 * (AVX-512 integer path with non-temporal stores where the CPU has it), so results are bit-identical to it. */
int hriemo_host_pack_bf16(const float* src, int64_t ld_src, int64_t T_in, int64_t cols, const int32_t* utt,
                          const int32_t* lens, void* dst_bf16, int64_t ld_dst, int64_t T_out, int64_t n,
                          int64_t n_src, int32_t n_threads);

/* ---------------------------------------------------- packed feature shards ----
 * The on-disk side of the step before the path.  The reference keeps one torch
 * pickle per utterance and modality, {"hidden": [L,d] f32, "attention_mask": [L]}
 * (scripts/iemocap_feature_extraction_seq_level/extract_audio_feats_wavlm_seq.py:118-135, read back one by
 * one at scripts/fusion/train_fusion_seq_level_decoder.py:139-156).  A shard (format HRIEMOS1, written by
 * hri-emo_b200/hriemo/shards.py, layout in hri-emo_b200/csrc/host_shard.cpp) packs many utterances into one
 * file: valid rows back to back (bf16 or f32), per-row PAD bytes, an index of lengths.  These HOST functions
 * mmap a shard and copy slabs of utterances into caller-provided (pinned) buffers as the zero-padded
 * [n, T, d] tensors + True = PAD masks the reference's collate would have produced
 * (scripts/fusion/train_fusion_seq_level_decoder.py:191-232).  No device work, no stream. */
typedef struct hriemo_shard_info_t {
  int64_t n_utt, rows_a, rows_t, meta_bytes;
  int32_t d_a, d_t;
  int32_t dtype;                /* 1 = bf16, 2 = f32 */
  int32_t max_len_a, max_len_t; /* longest stored utterance per modality */
} hriemo_shard_info_t;

int hriemo_shard_open(const char* path, void** handle);
int hriemo_shard_info(void* handle, hriemo_shard_info_t* info);
/* len_a / len_t: int32 [n_utt], rows stored per utterance (= index of the last valid position + 1). */
int hriemo_shard_lengths(void* handle, int32_t* len_a, int32_t* len_t);
/* The writer's JSON metadata (uids, labels, provenance), meta_bytes bytes, not NUL-terminated. */
int hriemo_shard_meta(void* handle, char* dst, int64_t dst_bytes);
/* Utterances utt[0..n) (or first..first+n when utt is NULL) as dst_a [n, T_a, d_a] / dst_t [n, T_t, d_t] of the
 * shard's dtype (rows past an utterance's length are zero; longer utterances are cut at T) and
 * mask_a [n, T_a] / mask_t [n, T_t] (1 = PAD).  Any of the four outputs may be NULL.  n_threads C++ threads. */
int hriemo_shard_read(void* handle, const int64_t* utt, int64_t first, int64_t n, int32_t T_a, int32_t T_t,
                      void* dst_a, void* dst_t, uint8_t* mask_a, uint8_t* mask_t, int32_t n_threads);
int hriemo_shard_close(void* handle);

/* ------------------------------------------------- loss and optimizer ----
 * The parts of the reference's training step that are not the model's backward pass
 * (scripts/fusion/train_fusion_seq_level_decoder.py:318-335, setup :405-416).  Parameters, gradients and
 * the two AdamW moments are flat fp32 arenas (the module's tensors are views into them), so the global norm
 * is one reduction, the update one launch and the data-parallel exchange one all-reduce.  The model's
 * backward kernels built so far follow below. */
/* loss_out[0] = BCEWithLogitsLoss(mean)(logits [B,C], labels [B,C]) - beta_weight * mean(beta * (1 - beta));
 * d_logits [B,C] / d_beta [B] (optional) receive d loss / d logits and d loss / d beta. */
int hriemo_bce_beta_loss(const float* logits, const float* labels, const float* beta, float beta_weight,
                         int64_t B, int32_t C, float* loss_out, float* d_logits, float* d_beta, void* stream);
/* clip_grad_norm_(max_norm) over a flat gradient arena, without touching the gradients and without a host
 * sync: out2[0] = total L2 norm, out2[1] = min(1, max_norm / (total + 1e-6)); hriemo_adamw_step reads
 * out2 + 1 as its grad_scale.  workspace: hriemo_grad_norm_workspace_bytes() bytes, 8-byte aligned. */
int64_t hriemo_grad_norm_workspace_bytes(void);
int hriemo_grad_norm_clip(const float* grads, int64_t n, float max_norm, void* workspace, float* out2, void* stream);
/* torch.optim.AdamW step `step` (1-based) with decoupled weight decay over flat arenas; grad_scale (device
 * pointer or NULL) multiplies every gradient first; params_bf16 (optional) receives the updated parameters as
 * bf16 (the GEMM operand copy).  Hyper-parameters are doubles: the derived constants (1 - beta, 1 - lr*wd,
 * bias corrections) are formed in double like PyTorch does and only then rounded to the kernel's fp32. */
int hriemo_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      int32_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                      const float* grad_scale, void* params_bf16, void* stream);

/* Weight and bias gradient of y = x W^T + b (what loss.backward() accumulates into nn.Linear.weight.grad / .bias.grad,
 * scripts/fusion/train_fusion_seq_level_decoder.py:332):  dW[N,K] = dY[M,N]^T . X[M,K]  (f32, row pitch K),
 * db[N] = column sums of dY (optional), both bf16 operands consumed as they lie in HBM (MN-major tcgen05 operands,
 * split over the rows, partial tiles summed in a fixed order).  accumulate != 0 adds to what dW / db hold.
 * N and K multiples of 128.  workspace: hriemo_linear_wgrad_workspace_bytes(M, N, K) bytes, 16-byte aligned.
 * The input gradient dX = dY . W is hriemo_gemm_bf16 on the transposed weight (hriemo_transpose_bf16). */
int64_t hriemo_linear_wgrad_workspace_bytes(int64_t M, int32_t N, int32_t K);
int hriemo_linear_wgrad_bf16(const void* dY, int64_t lddy, const void* X, int64_t ldx, int64_t M, int32_t N, int32_t K,
                             float* dW, float* db, int32_t accumulate, void* workspace, void* stream);
/* out[c, r] = in[r, c] for a bf16 matrix [rows, cols] (leading dimensions in elements). */
int hriemo_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int32_t rows, int32_t cols, void* stream);

/* Backward of nn.LayerNorm over the last dim (bf16 activations, fp32 parameter gradients): given the layer's INPUT x
 * (the pre-LayerNorm sum the forward keeps) and dy, writes dx (bf16) and dgamma / dbeta [d] (f32; accumulate != 0 adds
 * to what they hold).  Row statistics are recomputed from x.  d <= 1024, multiple of 8.
 * workspace: hriemo_layernorm_backward_workspace_bytes(rows, d) bytes. */
int64_t hriemo_layernorm_backward_workspace_bytes(int64_t rows, int32_t d);
int hriemo_layernorm_backward(const void* x_bf16, int64_t ldx, const void* dy_bf16, int64_t lddy, const float* gamma,
                              float eps, void* dx_bf16, int64_t lddx, float* dgamma, float* dbeta, int32_t accumulate,
                              void* workspace, int64_t rows, int32_t d, void* stream);
/* Backward of ReLU from its OUTPUT h (what the forward keeps): dx = dy where h > 0, else 0 (bf16 [rows, cols]). */
int hriemo_relu_backward_bf16(const void* dy, int64_t lddy, const void* h, int64_t ldh, void* dx, int64_t lddx,
                              int64_t rows, int32_t cols, void* stream);

/* Backward of hriemo_small_attention (the decoder's attention, models/emotion_decoder.py:42, :48-54): given d_out
 * [B*Nq, H*dh] writes dq [B*Nq, H*dh], dk and dv [B*Tk, H*dh] (bf16; the probabilities are rebuilt from q and k).
 * At most 8 queries. */
int hriemo_small_attention_backward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                    const void* d_out, int64_t lddo, const uint8_t* key_pad, void* dq, int64_t lddq,
                                    void* dk, int64_t lddk, void* dv, int64_t lddv, int32_t B, int32_t H, int32_t Nq,
                                    int32_t Tk, int32_t dh, float scale, void* stream);

/* ---- backward of the fp32 parts between the loss and the encoder outputs: the gate (models/beta_gate_tacfn.py:79-116)
 * and the decoder's head and queries (models/emotion_decoder.py:127, :153-155).  All reductions in a fixed order. */
/* Backward of y = x W^T + b in fp32 (the gate MLP, the Linear(d,1) head; forward: hriemo_sgemm_f32): any of
 * dX [M,K] = dY . W, dW [N,K] = dY^T . X (row pitch K), db [N] = column sums of dY may be NULL; accumulate != 0 adds to
 * what dW / db hold. */
int hriemo_linear_backward_f32(const float* dY, int64_t lddy, const float* X, int64_t ldx, const float* W, int64_t ldw,
                               int64_t M, int32_t N, int32_t K, float* dX, int64_t lddx, float* dW, float* db,
                               int32_t accumulate, void* stream);
/* dx = dy * act'(y) from the activation's OUTPUT y (HRIEMO_ACT_RELU / HRIEMO_ACT_SIGMOID / NONE), n fp32 elements. */
int hriemo_act_backward_f32(const float* dy, const float* y, float* dx, int64_t n, int32_t act, void* stream);
/* out[c] (+)= sum over the rows of x [rows, cols] (bf16, or f32 when x_is_f32): the gradient of the emotion queries
 * (models/emotion_decoder.py:127 broadcasts them over the batch) from a [B, N_e*d] view. */
int hriemo_sum_rows(const void* x, int32_t x_is_f32, int64_t ldx, float* out, int64_t rows, int32_t cols,
                    int32_t accumulate, void* stream);
/* Backward of hriemo_gate_input (g = [a, t, |a-t|, a*t], :87-89): dg [B,4d] -> da_pool, dt_pool [B,d]. */
int hriemo_gate_input_backward(const float* dg, const float* a_pool, const float* t_pool, float* da_pool,
                               float* dt_pool, int32_t B, int32_t d, void* stream);
/* inv_counts[b] = 1 / max(1, number of non-PAD positions) (masked_mean's denominator, :20-24); pad NULL -> 1 / T. */
int hriemo_mask_inv_counts(const uint8_t* pad, int32_t B, int32_t T, float* inv_counts, void* stream);
/* Gradient of the gate vector through the blend h = w*na[:, :L] + (1-w)*nt and beta = mean_d(w) (:95, :113-116):
 * dw[b,c] = sum_l dh[b,l,c] (na[b,l,c] - nt[b,l,c]) + dbeta[b] / d.  dh, nt: bf16 [B*L, d]; na: bf16 [B*T_a, d]
 * (the LayerNorm-ed streams); dbeta [B] may be NULL. */
int hriemo_gate_blend_backward_w(const void* dh, int64_t lddh, const void* na, int64_t ldna, int32_t T_a, const void* nt,
                                 int64_t ldnt, const float* dbeta, float* dw, int32_t B, int32_t L, int32_t d,
                                 void* stream);
/* Gradient w.r.t. one LayerNorm-ed stream [B,T,d] of the gate: the blend's share on its first L rows (coefficient w, or
 * 1 - w with one_minus) plus the masked mean's share dpool[b] * inv_counts[b] on its non-PAD rows.  dn: bf16 [B*T, d]. */
int hriemo_gate_stream_grad(const void* dh, int64_t lddh, int32_t L, const float* w, int32_t one_minus,
                            const float* dpool, const uint8_t* pad, const float* inv_counts, void* dn, int64_t lddn,
                            int32_t B, int32_t T, int32_t d, void* stream);

/* Backward of hriemo_attention_bf16 (the encoder's attention) for the training step: from the forward's operands, its
 * output `out`, its log-sum-exp `lse` and the output gradient d_out, writes dq [B*Tq, H*dh], dk and dv [B*Tk, H*dh]
 * (bf16; PAD keys receive zeros).  P is rebuilt as exp(scale * q.k - lse); dsum [B, H, Tq] (f32) is scratch that
 * receives rowsum(d_out o out).  Deterministic (no atomics): one pass owns dK / dV per key tile, one owns dQ per query
 * tile.  impl: 0 (default) / 3 = tcgen05 + TMEM + TMA (csrc/attention_bwd_tc.cu: 128-row tiles, S | dP side by side in
 * tensor memory, P / dS written back over them as the A operand of the accumulating MMAs, step rows read as MN-major
 * shared-memory operands); 1 = fp32 FMA loops (slow; validation); 2 = the first warp-level mma.sync form; 4 = the
 * mma.sync form with ldmatrix(.trans) fragments (the round-1 kernel, kept for A/B measurements). */
typedef struct hriemo_attn_bwd_args {
  const void* q;      int64_t ldq;    /* bf16, as in the forward */
  const void* k;      int64_t ldk;
  const void* v;      int64_t ldv;
  const uint8_t* key_pad;             /* [B, Tk] 1 = PAD, or NULL */
  const void* out;    int64_t ldo;    /* bf16 [B*Tq, H*dh]: the forward's output */
  const void* d_out;  int64_t lddo;   /* bf16 [B*Tq, H*dh] */
  const float* lse;                   /* f32 [B, H, Tq] from the forward */
  float* dsum;                        /* f32 [B, H, Tq] scratch */
  void* dq;           int64_t lddq;   /* bf16 outputs */
  void* dk;           int64_t lddk;
  void* dv;           int64_t lddv;
  int32_t B, H, Tq, Tk, dh;           /* dh in {32, 64, 96, 128} */
  float scale;
  int32_t impl;
  /* Optional, with key_pad: kv_steps[b] from hriemo_attention_kv_steps -- tiles of trailing PAD keys are neither
   * visited by the dQ pass nor computed by the dK / dV pass (they receive zeros); same result (impl 0 only). */
  const int32_t* kv_steps;
  /* The forward's dropout on the probabilities (hriemo_attn_args.drop_*): the same mask is recomputed (impl 0 only). */
  uint32_t drop_p8;
  uint32_t drop_key;
  float drop_scale;
} hriemo_attn_bwd_args;
int hriemo_attention_backward_bf16(const hriemo_attn_bwd_args* args, void* stream);

/* ------------------------------------------------- "tf32-class" precision mode ----
 * north_star: logits within 1e-4 of the reference's fp32 forward.  The tensor cores multiply bf16 here, so an fp32
 * operand travels as two bf16 numbers (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits; TF32 keeps 10) and a product
 * as three bf16 products accumulated in fp32: x.w ~= hi.hi + lo.hi + hi.lo.  Laid out along K -- activations as
 * [hi | lo | hi], weights as [hi | hi | lo] -- that is ONE hriemo_gemm_bf16 launch at 3 K with an f32 epilogue
 * (HRIEMO_EPI_BIAS_F32 / HRIEMO_EPI_BIAS_RESID_F32).  These entry points are the passes around it (f32 in, f32 out);
 * hriemo/precise.py holds the schedule.  Replaces the same reference lines as the bf16 entry points above. */

/* x f32 [rows, K] (row pitch ldx) -> out bf16 [rows, 3 * Kp], Kp = K rounded up to 8 (zero padded), row pitch ldo.
 * pattern 0 (activations): [hi | lo | hi]; pattern 1 (weights): [hi | hi | lo].  relu != 0: max(x, 0) first (the
 * FFN's activation, models/cross_modal_block_tacfn.py:46). */
int hriemo_split3(const float* x, int64_t ldx, void* out_bf16, int64_t ldo, int64_t rows, int32_t K, int32_t pattern,
                  int32_t relu, void* stream);

/* softmax(q k^T * scale + key padding) v in fp32 on the CUDA cores; q [B*Tq, >= H*dh], k / v [B*Tk, >= H*dh] (column
 * slices of a packed projection are fine; dh % 4 == 0, K rows 16-byte aligned), out f32 [B*Tq, H*dh] (pitch ldo).
 * probs: NULL, or f32 [B, Tq, Tk] ZEROED by the caller, receives the head-averaged probabilities (need_weights of
 * nn.MultiheadAttention; accumulated with atomics).  A fully masked row gives NaN, as torch.softmax does. */
int hriemo_attention_f32(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                         const uint8_t* key_pad, float* out, int64_t ldo, float* probs, int32_t B, int32_t H,
                         int32_t Tq, int32_t Tk, int32_t dh, float scale, void* stream);

/* models/beta_gate_tacfn.py:6-24 on an f32 [B, T, d] tensor: pooled [B, d] = masked mean over time. */
int hriemo_masked_mean_f32(const float* x, const uint8_t* pad, float* pooled, int32_t B, int32_t T, int32_t d,
                           void* stream);

/* models/beta_gate_tacfn.py:95-116 on f32 tensors: h [B, L, d] = w * a[:, :L] + (1 - w) * t, beta [B] = mean_d(w);
 * a holds T_a >= L rows per utterance, t holds L, w is [B, d]. */
int hriemo_gate_blend_f32(const float* a, int32_t T_a, const float* t, const float* w, float* h, float* beta, int32_t B,
                          int32_t L, int32_t d, void* stream);

/* ---- dropout of the training step (csrc/dropout.cuh: counter-based masks, nothing stored).
 * hriemo_dropout: out = (keep ? x * scale : 0) [+ resid], x / resid / out all f32 (is_f32) or all bf16, [rows, cols] with
 * cols % 4 == 0; p8 = round(256 p), scale = 1 / (1 - p8 / 256), key = the stream key of the site.  The forward calls it on
 * a sub-layer's output with the residual (nn.Dropout at models/cross_modal_block_tacfn.py:81-119,
 * models/emotion_decoder.py:43-59), the backward on the gradient with the same key.
 * hriemo_dropout_mask: the keep mask itself as bytes (tests; the definition of the mapping): rows_per_stream = 0 -> one
 * stream of `rows` rows under `key`; > 0 -> consecutive blocks of rows_per_stream rows are the streams drop_key_bh(key, s)
 * (the probabilities of (utterance, head) s = b * H + h: rows = B * H * Tq, rows_per_stream = Tq, cols = Tk). */
int hriemo_dropout(const void* x, int32_t is_f32, int64_t ldx, const void* resid, int64_t ldr, void* out, int64_t ldo,
                   int64_t rows, int32_t cols, uint32_t p8, float scale, uint32_t key, void* stream);
int hriemo_dropout_mask(uint8_t* out, int64_t rows, int32_t cols, uint32_t key, uint32_t p8, int64_t rows_per_stream,
                        void* stream);

/* The decoder's attention with dropout on the probabilities (training; nn.MultiheadAttention(dropout=p) at
 * models/emotion_decoder.py:14, 20): hriemo_small_attention / _backward plus (drop_p8, drop_key, drop_scale) as in
 * hriemo_attn_args; streams are (utterance, head), rows the queries. */
int hriemo_small_attention_dropout(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                   const uint8_t* key_pad, void* out_bf16, int64_t ldo, int32_t B, int32_t H, int32_t Nq,
                                   int32_t Tk, int32_t dh, float scale, uint32_t drop_p8, uint32_t drop_key, float drop_scale,
                                   void* stream);
int hriemo_small_attention_backward_dropout(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                            const void* d_out, int64_t lddo, const uint8_t* key_pad, void* dq, int64_t lddq,
                                            void* dk, int64_t lddk, void* dv, int64_t lddv, int32_t B, int32_t H, int32_t Nq,
                                            int32_t Tk, int32_t dh, float scale, uint32_t drop_p8, uint32_t drop_key,
                                            float drop_scale, void* stream);

/* Mean over time (UNMASKED) of an f32 [B,L,d] tensor — models/fusion_classifier.py:145. */
int hriemo_mean_over_time(const float* x, float* out, int32_t B, int32_t L, int32_t d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HRIEMO_H_ */

/*
 * scp_b200.h -- C ABI of libscp_b200.so: the SpeechCLIP+ data-parallel training hot path on B200 (sm_100a).
 *
 * Drop-in boundary.  The reference (ShampooWang/SpeechCLIP_plus) is pure Python: its "plugin API" for this path
 * is getattr-by-name on Python modules (avssl/model/kwClip.py:84, avssl/model/kw_branches.py:75-91) plus the
 * WeightedSumLayer constructor (avssl/module/speech_encoder_plus.py:218-220, :472-476).  It has no FFI of its
 * own; the entry points below are what a ctypes binding for the three hot-path classes binds, one group per
 * reference class.  The Python mirror of the reference interface lives in speechclip_plus_b200/ and calls
 * these through ctypes (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - every function returns 0 (SCP_OK) or a negative SCP_ERR_* code; scp_last_error_string() explains it.
 *   - never throws, never allocates device memory, never synchronises the device or the stream:
 *     scratch space is caller-provided (query the size with the matching *_workspace_bytes function).
 *   - all data pointers are DEVICE pointers valid on the stream's device unless marked "host".
 *   - sizes are int64_t element counts; strides are in elements; the last dimension is contiguous.
 *   - kernels are launched on `stream` (a cudaStream_t / CUstream); re-entrant, no global mutable state
 *     except the lazily resolved driver entry point for tensor-map encoding and one lazily created helper stream
 *     (+ two events) per device: scp_vq_fwd forks its arg-max kernel onto it and joins it back before returning, so
 *     every result is ordered on `stream` as usual (event record / wait only: valid under CUDA-graph capture).
 *   - there is no CPU fallback: on a device that is not sm_100 the functions return SCP_ERR_ARCH.
 */
#ifndef SCP_B200_H_
#define SCP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* scp_stream_t; /* == cudaStream_t */

enum {
  SCP_OK = 0,
  SCP_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size, misaligned pointer) */
  SCP_ERR_UNSUPPORTED = -2, /* shape / dtype combination this build has no kernel for */
  SCP_ERR_WORKSPACE = -3,   /* workspace too small */
  SCP_ERR_CUDA = -4,        /* a CUDA runtime / driver call failed (launch error, tensor-map encode ...) */
  SCP_ERR_ARCH = -5         /* device is not compute capability 10.x */
};

enum { SCP_F32 = 0, SCP_F16 = 1, SCP_BF16 = 2 };

#define SCP_MAX_LAYERS 32
#define SCP_MAX_PACKED 16
#define SCP_MAX_MASKED 8

/* ---- library ------------------------------------------------------------------------------------------- */
int scp_version(void);                         /* MAJOR*10000 + MINOR*100 + PATCH */
const char* scp_last_error_string(int code);   /* static string for `code`; detail of the last failure on this thread */
int scp_num_launches(void);                    /* kernels launched by this library in this process (bench accounting) */

/* ---- S1: upstream-feature fusion -- replaces WeightedSumLayer.forward (avssl/module/weighted_sum.py:26-45) and the
 *      per-layer rescale of its caller (FairseqSpeechEncoder_Hubert.forward, avssl/module/speech_encoder_plus.py:572-592) -- */
/* Normalisation applied to every layer BEFORE the weighted sum:
 *   SCP_NORM_NONE       plain sum                                                   (weighted_sum.py:43)
 *   SCP_NORM_LAYERNORM  non-affine LayerNorm over D, eps                             (weighted_sum.py:41-42; normalize_type "s3prl")
 *   SCP_NORM_L2_FRAME   x / (||x||_2 + 1e-8) per frame                               (speech_encoder_plus.py:576-583, "method1")
 *   SCP_NORM_UTT_MEAN   x / mean_t ||x_t||_2 per (layer, utterance)                  (speech_encoder_plus.py:584-590, "method2");
 *                       needs utt_scale (L,B) f32 = 1/mean filled by scp_wsum_utt_scale on the same layers           */
enum { SCP_NORM_NONE = 0, SCP_NORM_LAYERNORM = 1, SCP_NORM_L2_FRAME = 2, SCP_NORM_UTT_MEAN = 3 };

/* y[b,t,:] = sum_l softmax(weights)_l * norm(x_l[b,t,:]).  layer_ptrs: HOST array of L device pointers; every layer
 * is addressed as x_l[b*stride_b + t*stride_t + d] (the reference hands over (T,B,D) storage viewed as (B,T,D):
 * avssl/module/speech_encoder_plus.py:596-599).  y is (B,T,D) contiguous.  utt_scale: nullable unless
 * norm_mode == SCP_NORM_UTT_MEAN. */
int scp_wsum_fwd(const void* const* layer_ptrs, int L, int64_t B, int64_t T, int64_t D,
                 int64_t stride_b, int64_t stride_t, int dtype_in,
                 const float* weights, int norm_mode, float eps, const float* utt_scale,
                 void* y, int dtype_out, scp_stream_t stream);

/* Statistics pre-pass of SCP_NORM_UTT_MEAN: utt_scale[l*B + b] = 1 / mean_t ||x_l[b,t,:]||_2 (one extra read of the
 * layers; deterministic). */
int scp_wsum_utt_scale(const void* const* layer_ptrs, int L, int64_t B, int64_t T, int64_t D,
                       int64_t stride_b, int64_t stride_t, int dtype_in, float* utt_scale, scp_stream_t stream);

size_t scp_wsum_bwd_workspace_bytes(int L, int64_t B, int64_t T, int64_t D);

/* Backward of the above.  d_weights[l] = d(loss)/d(weights_l) (softmax backward included).  g_layers: nullable HOST
 * array of L device pointers to (B,T,D)-contiguous fp32 buffers receiving d(loss)/d(x_l) (only needed when the
 * upstream encoder is trainable: avssl/module/speech_encoder_plus.py:416-446; not available for SCP_NORM_UTT_MEAN). */
int scp_wsum_bwd(const void* const* layer_ptrs, int L, int64_t B, int64_t T, int64_t D,
                 int64_t stride_b, int64_t stride_t, int dtype_in,
                 const float* weights, int norm_mode, float eps, const float* utt_scale,
                 const void* g_y, int dtype_g, float* d_weights, void* const* g_layers,
                 void* workspace, size_t workspace_bytes, scp_stream_t stream);

/* ---- N1: keyword batch-norm prologue -- replaces Kw_BatchNorm.forward / Kw_BatchNorm_dynamic.forward
 *      (avssl/module/speechclip_c_modules/kw_bn.py:97-164, :216-228), the step between the keyword projection and the VQ
 *      (GeneralBranch.project_feats_to_CLIPspace, avssl/model/kw_branches.py:143-156) ------------------------------- */
/* x, y, g_y, g_x: (M,D) f32 row-major, rows m = b*K + k.  Row m belongs to statistics group m % n_groups:
 *   n_groups = 1  batchnorm_type "same" and Kw_BatchNorm_dynamic (one BatchNorm1d(D) over all keyword rows),
 *   n_groups = K  batchnorm_type "eachKw" (one set of statistics per keyword slot).
 * gamma / beta / running_mean / running_var are addressed IN PLACE as p[g*gstride + d*dstride]:
 *   BatchNorm1d(D): (0,1);  BatchNorm1d(D*K) of the `parallel` variant (kw_bn.py:119-127): (1,K);  K stacked layers: (D,1).
 * row_valid: nullable (M,) bytes; rows with 0 are excluded from the statistics and passed through unchanged (the
 * seq_lens branch, kw_bn.py:137-159).  training != 0: batch statistics (biased variance), running statistics updated
 * with `momentum` (unbiased variance) like torch.nn.BatchNorm1d; training == 0: running statistics are used.
 * save_mean / save_rstd: (n_groups, D) f32 contiguous, kept for the backward pass. */
size_t scp_kwbn_workspace_bytes(int64_t M, int64_t D, int n_groups);
int scp_kwbn_fwd(const float* x, int64_t M, int64_t D, int n_groups, int64_t gstride, int64_t dstride,
                 const uint8_t* row_valid, const float* gamma, const float* beta,
                 float* running_mean, float* running_var, int training, float momentum, float eps,
                 float* y, float* save_mean, float* save_rstd,
                 void* workspace, size_t workspace_bytes, scp_stream_t stream);
/* g_gamma / g_beta (nullable) are written with the parameter addressing above. */
int scp_kwbn_bwd(const float* g_y, const float* x, int64_t M, int64_t D, int n_groups, int64_t gstride, int64_t dstride,
                 const uint8_t* row_valid, const float* gamma, const float* save_mean, const float* save_rstd,
                 int training, float* g_x, float* g_gamma, float* g_beta,
                 void* workspace, size_t workspace_bytes, scp_stream_t stream);

/* ---- S2: keyword vector quantiser -- replaces GeneralBranch.get_keyword_cosine_score + SimpleVectorQuantizer.forward
 *      + the lookup matmul (avssl/model/kw_branches.py:158-197, avssl/module/speechclip_c_modules/my_vector_quantizer.py:64-165) -- */

/* One-time (per table version) preparation of the frozen CLIP token table E (V,D) fp32 row-major:
 *   table_hat  (Vp,D)  f16 : rows normalised to unit L2 norm (eps 1e-8), zero rows for v >= V, Vp = round_up(V,256)
 *   table_hat_t(D,Vp)  f16 : its transpose (K-major operand of the backward GEMM)
 *   table_norm (Vp,)   f32 : max(||e_v||, 1e-8)
 *   table_mean (D+1,)  f32 : mean_v e_v (centring vector of the backward pass); element [D] = max_v ||e_v||          */
int64_t scp_vq_padded_vocab(int64_t V);
int scp_vq_prepare_table(const float* table, int64_t V, int64_t D,
                         void* table_hat, void* table_hat_t, float* table_norm, float* table_mean,
                         scp_stream_t stream);

/* The workspace holds the per-chunk / per-group maxima scanned by the exact arg-max, the split row statistics and --
 * unless the environment sets SCP_VQ_COLSUM=0 (older two-sweep mode) -- an (Mp, Vp) fp16 scratch of e^c from which avg_probs
 * is reduced in one HBM pass: 2*Mp*Vp bytes (202 MB at M = 2048, V = 49408). */
size_t scp_vq_fwd_workspace_bytes(int64_t M, int64_t V, int64_t D);

/* Forward.  kw (M,D) fp32 = keyword vectors in CLIP space, M = B*K rows ordered (b,k).
 *   idx        (M,)  int64 : arg-max code per row, first maximum wins, masked columns excluded
 *                            (bit-exact w.r.t. an exact evaluation of the cosine; see DESIGN.md "exact arg-max")
 *   keywords   (M,D) f32   : E[idx]  (value of subword_prob @ E, kw_branches.py:195)
 *   row_stats  (M,4) f32   : {lse at temperature 1, lse at temperature tau, entropy at temperature 1, 1/max(||kw||,1e-8)}
 *   code_hist  (Vp,) f32   : histogram of idx (counts)
 *   avg_probs  (Vp,) f32   : mean_m softmax(cos[m,:]) at temperature 1  (my_vector_quantizer.py:102); nullable -> skipped
 *   metrics    (3+K,) f32  : {code_perplexity, prob_perplexity, diversity_loss, ent_per_t[0..K)}
 *   kw_hat     (Mp,D) f16  : normalised keywords kept for the backward pass, Mp = round_up(M,128)
 * masked_cols: HOST array of n_masked (<= SCP_MAX_MASKED) column ids that receive -inf (prob_msk). */
int scp_vq_fwd(const float* kw, int64_t M, int64_t K, int64_t V, int64_t D,
               const void* table_hat, const float* table_norm, const float* table,
               const int32_t* masked_cols, int n_masked, const float* tau,
               int64_t* idx, float* keywords, float* row_stats, float* code_hist, float* avg_probs,
               float* metrics, void* kw_hat,
               void* workspace, size_t workspace_bytes, scp_stream_t stream);

/* Forward that keeps what the backward pass would otherwise recompute.  saved_probs (nullable -> exactly scp_vq_fwd),
 * (Mp, Vp) fp16, scp_vq_saved_probs_bytes(M, V) bytes, takes the place of the (M,V) scratch inside the workspace (use
 * scp_vq_fwd_save_workspace_bytes) and receives the soft-max numerators at temperature tau,
 *   P''[m,v] = exp((cos[m,v] - 1)/tau + 10)      (= e^{cos/tau} at tau = 0.1; in [e^-10, e^10] for every tau; 0 for masked /
 *   padding columns),
 * i.e. softmax_tau(cos[m,:]) up to the row's normaliser (kw_branches.py:158-179 + my_vector_quantizer.py:130-136); avg_probs
 * and the exact arg-max are derived from the same buffer (e^cos = P''^tau e^{1 - 10 tau}).  Outputs are identical to
 * scp_vq_fwd within the stated tolerances.  The buffer must stay untouched until scp_vq_bwd_saved has consumed it.  Meant for
 * tau >= 0.1 (every shipped recipe): the cosines span 2/tau nats of exponent and fp16 offers ~20 with the factor 2 of head
 * room Q'' needs; below 0.1 the numerators of strongly negative cosines (cos < 1 - 26.6 tau) underflow, which the gradient
 * tolerates (they carry no probability mass at temperature tau) but avg_probs, derived from them at temperature 1, does not
 * (the Python wrapper selects scp_vq_fwd / scp_vq_bwd then). */
size_t scp_vq_saved_probs_bytes(int64_t M, int64_t V);
size_t scp_vq_fwd_save_workspace_bytes(int64_t M, int64_t V, int64_t D);   /* saved_probs != NULL: no (M,V) scratch inside */
int scp_vq_fwd_save(const float* kw, int64_t M, int64_t K, int64_t V, int64_t D,
                    const void* table_hat, const float* table_norm, const float* table,
                    const int32_t* masked_cols, int n_masked, const float* tau,
                    int64_t* idx, float* keywords, float* row_stats, float* code_hist, float* avg_probs,
                    float* metrics, void* kw_hat, void* saved_probs,
                    void* workspace, size_t workspace_bytes, scp_stream_t stream);

size_t scp_vq_bwd_workspace_bytes(int64_t M, int64_t V, int64_t D);

/* Backward of the straight-through estimator (training mode):
 *   g_p = g_keywords E^T ; g_c = p_tau (g_p - <p_tau,g_p>) / tau ; g_khat = g_c Ehat ;
 *   g_kw = (g_khat - <g_khat,khat> khat) / ||kw||          
 * g_tau (nullable, (1,) f32) receives d(loss)/d(tau) for a learnable temperature (finite closed form). */
int scp_vq_bwd(const float* g_keywords, const float* kw, int64_t M, int64_t V, int64_t D,
               const void* kw_hat, const void* table_hat, const void* table_hat_t,
               const float* table_norm, const float* table_mean, const float* row_stats,
               const int32_t* masked_cols, int n_masked, const float* tau,
               float* g_kw, float* g_tau,
               void* workspace, size_t workspace_bytes, scp_stream_t stream);

/* Same gradient from the numerators saved by scp_vq_fwd_save (saved_probs nullable -> scp_vq_bwd): only g . E^T is formed
 * on the tensor cores (no k . E^T product, no exponentials: 6 instead of 8 M V D executed FLOP over forward + backward ... 4
 * instead of 8 in this call) and the output GEMM reads P'' in place.  Falls back to the recompute path when g_tau is
 * requested (a learnable temperature needs the logits) and for single-tile problems. */
/* 1 when scp_vq_bwd_saved would use the saved numerators for this shape (at least two 128-row tiles): callers
 * that cannot profit should call scp_vq_fwd (its avg_probs pass is cheaper than the one of scp_vq_fwd_save). */
int scp_vq_bwd_saved_available(int64_t M, int64_t V, int64_t D);
size_t scp_vq_bwd_saved_workspace_bytes(int64_t M, int64_t V, int64_t D, int want_tau);   /* want_tau: g_tau != NULL */
int scp_vq_bwd_saved(const float* g_keywords, const float* kw, int64_t M, int64_t V, int64_t D,
                     const void* kw_hat, const void* table_hat, const void* table_hat_t,
                     const float* table_norm, const float* table_mean, const float* row_stats,
                     const int32_t* masked_cols, int n_masked, const float* tau, const void* saved_probs,
                     float* g_kw, float* g_tau,
                     void* workspace, size_t workspace_bytes, scp_stream_t stream);

/* Dense-input form of SimpleVectorQuantizer.forward (my_vector_quantizer.py:64-165) for callers that already hold the
 * (M,V) score matrix x (fp32, row pitch ldx).  Masks x IN PLACE like the reference (:78-79).  subword_prob (nullable,
 * (M,V) f32 contiguous) receives the forward VALUE of `hard + p - p.detach()` / `hard`, i.e. the one-hot; with
 * `training` = 3 (bit 0: training mode, bit 1: the module's hard = False) it receives softmax(x / tau) itself (:130-131).
 * scp_vq_dense_bwd is the backward of both (the straight-through estimator's gradient IS the soft-max gradient).  Outputs as
 * in scp_vq_fwd (row_stats[.,3] unused; code_hist / avg_probs have V entries). */
size_t scp_vq_dense_workspace_bytes(int64_t M, int64_t V);
int scp_vq_dense_fwd(float* x, int64_t M, int64_t K, int64_t V, int64_t ldx,
                     const int32_t* masked_cols, int n_masked, const float* tau, int training,
                     int64_t* idx, float* row_stats, float* code_hist, float* avg_probs, float* metrics,
                     float* subword_prob, void* workspace, size_t workspace_bytes, scp_stream_t stream);
/* g_x (M,V contiguous) = p_tau * (g_p - <p_tau, g_p>) / tau for the dense form (g_p = d loss / d subword_prob, pitch ldg);
 * g_tau nullable. */
int scp_vq_dense_bwd(const float* x_masked, const float* g_p, int64_t M, int64_t V, int64_t ldx, int64_t ldg,
                     const float* row_stats, const float* tau, float* g_x, float* g_tau, scp_stream_t stream);

/* ---- N3: text-transformer input splice -- replaces the prologue of ClipModel.encode_keywords
 *      (avssl/module/clip_official.py:240-267: id tensor, token-embedding lookup, the per-sample Python splice loop
 *      :261-265, positional embedding) and get_keypadding_mask (avssl/util/data_utils.py:6-22) --------------------- */
/* x[b,l,:] = src(b,l) + pos_emb[l,:],  src = E[sot] (l = 0) | keywords[b,l-1] (1 <= l <= n_b) | E[eot] (l = n_b+1) |
 * E[0] (else);  n_b = kw_num[b] (device int64, nullable) or fixed_num, clamped to [0, min(Kmax, L-2)].
 * keywords (B,Kmax,D) f32; table (V,D), pos_emb (L,D) and x (B,L,D) share `dtype` (keywords are rounded to it first,
 * as the reference's slice-assignment does).  eot_index (nullable, (B,) int64) = n_b + 1 (clip_official.py:274-277). */
int scp_kw_splice_fwd(const float* keywords, const int64_t* kw_num, int64_t fixed_num,
                      const void* table, const void* pos_emb, int dtype,
                      int64_t B, int64_t Kmax, int64_t D, int64_t L, int64_t sot_id, int64_t eot_id,
                      void* x, int64_t* eot_index, scp_stream_t stream);
/* g_keywords[b,j,:] = g_x[b,1+j,:] for j < n_b, 0 otherwise (token table and positional embedding are frozen:
 * clip_official.py:110-123). */
int scp_kw_splice_bwd(const void* g_x, int dtype, const int64_t* kw_num, int64_t fixed_num,
                      int64_t B, int64_t Kmax, int64_t D, int64_t L, float* g_keywords, scp_stream_t stream);
/* mask[b,j] = (j >= lens[b]) as bytes (True = padding). */
int scp_keypadding_mask(const int64_t* lens, int64_t B, int64_t max_len, uint8_t* mask, scp_stream_t stream);

/* ---- N4: CIF down-sampler -- replaces CIF.integrate_and_fire (avssl/module/cif.py:157-311), which produces the
 *      dynamic-length keyword sequence of the "+" branches ------------------------------------------------------------ */
/* Step 1: csum (B,S) = SEQUENTIAL fp32 cumulative sum of alpha (B,S); feat_len[b] = clip(floor(csum[b,S-1]/threshold), 1,
 * max_len) (cif.py:183-188, MAX_FEAT_LEN = 75).  The caller then reads T = max_b feat_len (the output shape depends on
 * it, as in the reference). */
int scp_cif_plan(const float* alpha, int64_t B, int64_t S, float threshold, int max_len,
                 float* csum, int64_t* feat_len, scp_stream_t stream);
/* Step 2: out (B, T+1, C) f32 = integrate-and-fire of x (B,S,C) f32 (cif.py:191-243; row T collects the tail); every
 * output row is written exactly once.  fire_mask (nullable, (B,S) bytes) = fire_num > 0 (:207); tail_w (nullable, (B,)) =
 * weight that reached the row just past feat_len[b] (:251-258). */
int scp_cif_fire_fwd(const float* x, const float* alpha, const float* csum, const int64_t* feat_len,
                     int64_t B, int64_t S, int64_t C, float threshold, int64_t T,
                     float* out, uint8_t* fire_mask, float* tail_w, scp_stream_t stream);
/* Inference tail handling (cif.py:246-296) on out (B,T1,C), T1 = T+1: utterances with tail_w >= firing_threshold fire
 * once more (row feat_len[b] scaled by threshold/tail_w), feat_len_new = min(feat_len + extend, max_len), rows >=
 * feat_len_new erased. */
int scp_cif_tail(float* out, int64_t B, int64_t T1, int64_t C, const int64_t* feat_len, const float* tail_w,
                 float threshold, float firing_threshold, int max_len, int64_t* feat_len_new, scp_stream_t stream);
/* Backward of steps 2(+3): g_out (B,T_out,C) is the gradient of the returned slice out[:, :T_out]; keep_len / scale_row /
 * tail_w (nullable) describe the tail handling of the forward (feat_len_new, feat_len, tail_w).  g_x (B,S,C), g_alpha (B,S)
 * (through right_weight = csum - right_idx*threshold and left_weight = alpha - right_weight - ..., :209-224).
 * workspace: 2*B*S floats. */
int scp_cif_fire_bwd(const float* g_out, int64_t T_out, const float* x, const float* alpha, const float* csum,
                     int64_t B, int64_t S, int64_t C, float threshold, int64_t T,
                     const int64_t* keep_len, const int64_t* scale_row, const float* tail_w, float firing_threshold,
                     float* g_x, float* g_alpha, void* workspace, size_t workspace_bytes, scp_stream_t stream);

/* ---- N0 + G0: L2-normalise the loss features and pack them into the all-gather send buffer
 *      (avssl/model/kwClip.py:857, :905-907, :913-915; gather point kwClip.py:149-169) -------------------- */
/* packed layout per rank: n_feats blocks of (n,D) f32 followed by n int64 ids.  feats: HOST array of device ptrs.
 * inv_norms (n_feats,n) f32 receives 1/||f|| for the backward of the normalisation. */
size_t scp_pack_bytes(int n_feats, int64_t n, int64_t D);
int scp_l2norm_pack(const void* const* feats, int n_feats, int64_t n, int64_t D, int dtype_in,
                    const int64_t* ids, void* packed_out, float* inv_norms, scp_stream_t stream);
/* backward of f/||f||: g_f = (g_n - <g_n,fhat> fhat) * inv_norm, all (n,D) f32 */
int scp_l2norm_bwd(const float* g_n, const float* f_hat, const float* inv_norm, int64_t n, int64_t D,
                   float* g_f, scp_stream_t stream);

/* ---- S3: masked in-batch InfoNCE -- replaces MaskedContrastiveLoss.forward (avssl/module/losses.py:185-245) ---- */
size_t scp_nce_workspace_bytes(int64_t N, int64_t D);

/* A, Bm (N,D) f32, ids (N,) int64 or NULL.  logit scale = exp(*log_scale) when log_scale != NULL, else fixed_scale
 * (losses.py:219-222).  Rows/columns [row_begin,row_end) are this rank's shard: the loss is over all N rows (every rank
 * computes the same value), gradients are produced for the local rows only.
 *   loss    (1,) f32, lse_row (N,) f32 = log sum_j mask*exp(S_ij), lse_col (N,) f32 = log sum_i mask*exp(S_ij) */
int scp_nce_fwd(const float* A, const float* Bm, const int64_t* ids, int64_t N, int64_t D,
                const float* log_scale, float fixed_scale, float margin, int dcl, int a2b, int b2a,
                int prepare_bwd /* also stage the transposed operands of the backward GEMM in `workspace` */,
                float* loss, float* lse_row, float* lse_col,
                void* workspace, size_t workspace_bytes, scp_stream_t stream);

/* Sharded forward for the one-process-per-GPU layout (option B): A, Bm, ids hold the GATHERED
 * global batch; this rank evaluates only the denominators of its own samples -- rows [row_begin,row_end) of A against
 * every row of Bm and rows [row_begin,row_end) of Bm against every row of A: 2 n N D instead of 2 N^2 D multiply-adds.
 *   stats_local (3, n) f32: [0] = log sum_j mask*exp(S_ij) for the local rows i, [1] = log sum_i mask*exp(S_ij) for the
 *   local columns j, [2] = <A_i, Bm_i>.
 * The ranks all-gather stats_local (12 bytes per sample) into stats_all (world, 3, n) and call scp_nce_loss_from_stats,
 * which returns the loss over all N samples and the contiguous (N,) lse_row / lse_col vectors that scp_nce_bwd reads.
 * With world = 1 the pair equals scp_nce_fwd. */
int scp_nce_fwd_local(const float* A, const float* Bm, const int64_t* ids, int64_t N, int64_t D,
                      const float* log_scale, float fixed_scale, float margin, int dcl,
                      int64_t row_begin, int64_t row_end, int prepare_bwd, float* stats_local,
                      void* workspace, size_t workspace_bytes, scp_stream_t stream);
int scp_nce_loss_from_stats(const float* stats_all, int world, int64_t n_local, const float* log_scale,
                            float fixed_scale, float margin, int a2b, int b2a,
                            float* loss, float* lse_row, float* lse_col, scp_stream_t stream);

/* dA, dB: (row_end-row_begin, D) f32 gradients of g_loss*loss w.r.t. A[row_begin:row_end], Bm[row_begin:row_end]
 * (dB nullable).  d_log_scale (nullable, (1,)): this shard's PART of the gradient w.r.t. the log-scale parameter -- the
 * sum of G (.) S over rows [row_begin,row_end) and all columns; the parts of all shards add up to the full gradient
 * (which is what DDP's gradient all-reduce produces; see ddp_grad_scale). */
int scp_nce_bwd(const float* A, const float* Bm, const int64_t* ids, int64_t N, int64_t D,
                const float* log_scale, float fixed_scale, float margin, int dcl, int a2b, int b2a,
                const float* lse_row, const float* lse_col, const float* g_loss,
                int64_t row_begin, int64_t row_end,
                int fwd_state_valid /* `workspace` is the untouched workspace of scp_nce_fwd(prepare_bwd=1) on the same inputs */,
                float* dA, float* dB, float* d_log_scale,
                void* workspace, size_t workspace_bytes, scp_stream_t stream);

/* ---- tail of the data-parallel step: packed gradient exchange + fused Adam for the path's own trainable tensors
 *      (weightedsum_layer.weights, criterion.temperature, ...; avssl/model/kwClip.py:636-668: ONE Adam group, yaml
 *      audio_encoder.optim: lr 1e-4, weight_decay 1e-6).  The reference's DataParallel reduce_add_coalesced +
 *      torch.optim.Adam become: scp_grad_pack -> one NCCL all-reduce of `packed` (issued by the caller) ->
 *      scp_adam_packed.  grads / params: HOST arrays of n <= SCP_MAX_PACKED device pointers (f32), sizes: HOST array of
 *      element counts; packed, exp_avg, exp_avg_sq: (sum sizes) f32; a NULL grads[i] packs zeros. ---------------- */
int scp_grad_pack(const float* const* grads, const int64_t* sizes, int n, float scale, float* packed,
                  scp_stream_t stream);
/* torch.optim.Adam semantics (betas, eps, L2 weight_decay, bias correction; no amsgrad) applied in place to every
 * registered tensor; g = grad_scale * packed.  step: device int64 counter, read and incremented by the kernel (a
 * replayed CUDA graph therefore advances it).  lr_device (nullable): device f32 learning rate written by the caller's
 * scheduler; otherwise `lr`. */
int scp_adam_packed(float* const* params, const int64_t* sizes, int n, const float* packed_grads,
                    int n_shards /* packed_grads holds n_shards buffers (one per rank, as gathered), summed in rank order */,
                    int64_t shard_stride /* elements between them */, float grad_scale,
                    float* exp_avg, float* exp_avg_sq, int64_t* step, const float* lr_device, float lr, float beta1,
                    float beta2, float eps, float weight_decay, scp_stream_t stream);

/* ---- one-shot all-gather over NVLink peer memory for the path's small exchanges (G0 and friends): every rank pushes its
 *      payload into the gather buffer of every peer with plain stores on peer-mapped pointers, raises a flag there and
 *      waits for the flags of all peers -- one launch + one collect launch instead of an NCCL collective (25-50 us of
 *      latency each at 8 ranks inside a CUDA graph).  The caller allocates the same buffer of scp_p2p_buffer_bytes() on
 *      every rank as peer-accessible (symmetric) memory, zeroes it once, and passes the DEVICE array of the world base
 *      pointers (peer_bufs_device[r] = rank r's buffer as mapped into this process).  state: 2 x u32, zero-initialised,
 *      private to this rank.  out: (world, nbytes) local result.  nbytes: multiple of 16, <= nbytes_capacity. ---------- */
size_t scp_p2p_buffer_bytes(int world, size_t nbytes_capacity);
int scp_p2p_allgather(const void* src, size_t nbytes, void* const* peer_bufs_device, void* local_buf, int rank, int world,
                      size_t nbytes_capacity, uint32_t* state, void* out, scp_stream_t stream);
/* Same exchange, but `out` is SEGMENT-major: the payload is n_seg (<= 4) segments of seg_bytes[i] bytes (HOST array, each a
 * multiple of 16, adding up to nbytes); out holds for every segment the world copies in rank order, i.e. segment i is the
 * contiguous block [world * sum(seg_bytes[0..i)), + world * seg_bytes[i]) -- the gathered features and ids come out as
 * ready-to-use (world * n, D) / (world * n,) arrays. */
int scp_p2p_allgather_segments(const void* src, size_t nbytes, void* const* peer_bufs_device, void* local_buf, int rank, int world,
                               size_t nbytes_capacity, uint32_t* state, const int64_t* seg_bytes, int n_seg, void* out,
                               scp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SCP_B200_H_ */

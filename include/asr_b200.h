/*
 * asr_b200.h — C ABI of libasr_b200.so: the B200-native acoustic hot path of
 * igormq/asr-study (MFCC/log-mel -> stacked BiLSTM -> CTC loss/grad + decode).
 *
 * The reference has no FFI: its extension surface is Python duck typing over
 * Keras/TF ops.  Each entry point below is what a binding for that path would
 * call; the reference site it replaces is cited as file:line (relative to the
 * reference repo).  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - every function returns 0 on success, a negative asr_status on failure;
 *     asr_last_error() returns a thread-local message for the last failure.
 *   - all data pointers are CALLER-OWNED DEVICE pointers unless the name ends
 *     in _host; `stream` is a cudaStream_t passed as void*.
 *   - no hidden allocation on the hot path: scratch is passed in; sizes come
 *     from the *_workspace_bytes queries.  Plans own small constant tables.
 *   - re-entrant; host threads may call concurrently on different streams.  No
 *     process-global switches: engine / scheduling choices are explicit flag
 *     arguments (ASR_GEMM_*, ASR_LSTM_*), never environment variables.
 *   - internal activation layout is TIME-MAJOR [T, N, *] (what TF's CTC ops
 *     consume after the reference's own transpose, core/ctc_utils.py:39,69).
 */
#ifndef ASR_B200_H
#define ASR_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ASR_OK = 0,
  ASR_ERR_INVALID = -1,      /* bad argument / unsupported configuration   */
  ASR_ERR_CUDA = -2,         /* a CUDA runtime call failed                  */
  ASR_ERR_UNSUPPORTED = -3,  /* valid in the reference, not built here yet  */
  ASR_ERR_TIMEOUT = -4       /* a persistent kernel's watchdog fired        */
} asr_status;

const char* asr_last_error(void);
int32_t asr_version(void);
/* number of kernels launched by this library in this process (for bench.py's
 * gpu_launches claim) */
int64_t asr_launch_count(void);

/* ------------------------------------------------------------------------- *
 * K1  fused MFCC / log-mel front end
 * replaces preprocessing/audio.py:41-75 (Feature.__call__, _standarize),
 *          :223-253 (FBank._call), :339-367 (MFCC._call), :419-442 (LogFbank)
 *          preprocessing/audio_utils.py:17-50,98-120,143-173
 * ------------------------------------------------------------------------- */
typedef struct {
  float fs;             /* 16000                          audio.py:28  */
  float win_len;        /* 0.025 s                        audio.py:180 */
  float win_step;       /* 0.010 s                                     */
  int32_t num_filt;     /* 40                                          */
  int32_t nfft;         /* 512 (only value built)                      */
  float low_freq;       /* 20                                          */
  float high_freq;      /* 7800                                        */
  float pre_emph;       /* 0.97                                        */
  int32_t kind;         /* 0 = mfcc, 1 = logfbank, 2 = fbank (linear)  */
  int32_t num_cep;      /* 13                              audio.py:324 */
  int32_t cep_lifter;   /* 22                                          */
  int32_t append_energy;/* mfcc: c0 <- log(energy+eps); logfbank: extra col */
  int32_t d, dd;        /* deltas / delta-deltas                       */
  int32_t mean_norm, var_norm; /* per-utterance CMVN       audio.py:70-75 */
  float eps;            /* 1e-8                                        */
  int32_t stride;       /* feats[::stride]                 audio.py:82  */
  int32_t num_context;  /* +-context frames, audio.py:88-150 (0 = off)  */
} asr_mfcc_config;

typedef struct asr_mfcc_plan asr_mfcc_plan;

int32_t asr_mfcc_plan_create(const asr_mfcc_config* cfg, asr_mfcc_plan** out);
void    asr_mfcc_plan_destroy(asr_mfcc_plan* plan);
int32_t asr_mfcc_num_feats(const asr_mfcc_plan* plan);
/* frames for a clip of `num_samples` after stride (audio_utils.py:30-33) */
int32_t asr_mfcc_num_frames(const asr_mfcc_plan* plan, int64_t num_samples);
/* bytes of zero-initialised device scratch for a batch of n utterances */
size_t  asr_mfcc_workspace_bytes(const asr_mfcc_plan* plan, int32_t n);
/* same, for plans with num_context > 0 (they stage the un-normalised [n, t_max, F] features in the workspace);
 * equals asr_mfcc_workspace_bytes (rounded up) when num_context == 0 */
size_t  asr_mfcc_workspace_bytes_ex(const asr_mfcc_plan* plan, int32_t n, int32_t t_max);
/*
 * pcm      f32 [sum samples], utterance i = pcm[offsets[i] .. offsets[i+1])
 * offsets  i64 [n+1] (device)
 * out      f32 [n, t_max, F] (time_major=0, the DatasetIterator batch contract,
 *          datasets/dataset_generator.py:223-235) or [t_max, n, F] (time_major=1);
 *          frames >= out_len[i] are zero-filled ('post' padding)
 * out_len  i32 [n]
 * ws       workspace, zeroed once by the caller (the kernel leaves it zeroed)
 */
int32_t asr_mfcc_forward(const asr_mfcc_plan* plan, const float* pcm,
                         const int64_t* offsets, int32_t n, int32_t t_max,
                         float* out, int32_t* out_len, int32_t time_major,
                         void* ws, void* stream);
/* host convenience for Feature.__call__ on ONE ndarray (audio.py:60-65):
 * copies in, runs asr_mfcc_forward, copies out.  out_host f32 [T, F]. */
int32_t asr_mfcc_forward_host(const asr_mfcc_plan* plan, const float* pcm_host,
                              int64_t num_samples, float* out_host);

/* ------------------------------------------------------------------------- *
 * K2/K5  TN GEMM on tcgen05:  C[M,N] (+)= A[M,K] * B[N,K]^T
 * replaces K.dot(x*B_W, W) hoisted out of the step (core/layers.py:439), the
 * TimeDistributed(Dense) (core/models.py:71,278) and the autodiff dW/dU/dX.
 * A, B: 16-bit (dtype_in: 0 = fp16, 1 = bf16), K-major, lda/ldb in elements
 * (multiples of 8, 16-byte aligned rows).  C: dtype_out 0 = fp32, 1 = fp16,
 * 2 = bf16, row-major ldc.  bias (f32 [N]) optional.  accumulate != 0 adds
 * into fp32 C.  lda < K is allowed: row m of A then starts lda elements after
 * row m-1 and the rows overlap (the convolution view of the conv section below).
 * ------------------------------------------------------------------------- */
int32_t asr_gemm_tn(int32_t dtype_in, int32_t dtype_out, int32_t M, int32_t N,
                    int32_t K, const void* A, int64_t lda, const void* B,
                    int64_t ldb, void* C, int64_t ldc, const float* bias,
                    float alpha, int32_t accumulate, void* stream);
/* same, with scheduling hints.  ASR_GEMM_BACKGROUND: the GEMM runs on a side stream beside a persistent
 * recurrence (e.g. dW/dU of layer l during the BPTT of layer l-1) and only gets the SMs that kernel leaves
 * idle; the non-persistent tiling is used so its CTAs are scheduled one by one as SMs free up.
 * ASR_GEMM_TILE128: pin the non-persistent 128x128 tiling (96 KB of shared memory per CTA, two CTAs per SM), e.g.
 * when the GEMM shares SMs with a beam search.  Both engines are tcgen05; a shape neither takes is an error. */
#define ASR_GEMM_BACKGROUND 1
#define ASR_GEMM_TILE128 2
int32_t asr_gemm_tn_ex(int32_t dtype_in, int32_t dtype_out, int32_t M, int32_t N,
                       int32_t K, const void* A, int64_t lda, const void* B,
                       int64_t ldb, void* C, int64_t ldc, const float* bias,
                       float alpha, int32_t accumulate, int32_t flags, void* stream);

/* ------------------------------------------------------------------------- *
 * K3/K4  persistent BiLSTM recurrence (both directions in one launch)
 * replaces LSTM.step under K.rnn / Bidirectional (core/layers.py:432-469;
 * core/models.py:68-70, 261-271) and its autodiff.  Keras-1 semantics: gate
 * order i,f,c,o; hard_sigmoid gates; tanh cell; h0=c0=0; no masking; the
 * reverse direction consumes t = T-1..0.
 * ------------------------------------------------------------------------- */
typedef struct {
  int32_t T, N, H;          /* see asr_lstm_persistent_supported() for the shapes the engines take */
  int32_t training;         /* 1: save gates/c (and transposed copies)        */
  const float* zx;          /* f32 [T, N, 2, 4H]  x_t*W  (no bias)            */
  const float* bias;        /* f32 [2, 4H]                                    */
  const float* U;           /* f32 [2, H, 4H]  master recurrent weights       */
  const void*  U16;         /* fp16 [2, 4H, H]  U^T (tensor-core operand)     */
  void*  h16;               /* fp16 [T, N, 2H]  layer output fwd|bwd          */
  void*  hT16;              /* bf16 [2H, T*N]   transposed copy (training)    */
  float* h32;               /* f32 [T, N, 2H]   optional fp32 output          */
  float* gates;             /* f32 [T, N, 2, 4H] activated i,f,g,o (training) */
  float* cell;              /* f32 [T, N, 2, H]  c_t             (training)   */
  int32_t* flags;           /* device scratch, asr_lstm_flags_bytes(), zeroed */
  const float* mask_u;      /* f32 [2, N, H] variational-dropout mask B_U (already /(1-p)) applied to
                               h_{t-1} in the recurrence (core/layers.py:438), or NULL.  When set,
                               hT16 holds h*mask (the operand of dU); h16/h32 stay unmasked.        */
  /* fused dropout outputs (engines that report asr_lstm_fuses_masks() == 1; ignored otherwise):          */
  const float* mask_next;   /* f32 [2, N, 2H]  B_W of the NEXT layer's fwd | bwd LSTM, or NULL             */
  void*  hm16;              /* fp16 [2, T*N, 2H]  h * mask_next[i]: the next layer's projection operands    */
  void*  hmT16;             /* bf16 [2, 2H, T*N]  transposed copies: the next layer's dW operands (training)*/
  void*  hT16u;             /* bf16 [2H, T*N]  unmasked transposed copy (the Dense dW operand), or NULL     */
  /* element-wise brsmv1 switches on the tensor-core engine (asr_lstm_fuses_variants() == 1; core/layers.py:441-467):  */
  const float* mi;          /* f32 [3, 2, 4H]  alpha | beta1 | beta2 of multiplicative integration, or NULL  */
  float* uh;                /* f32 [T, N, 2, 4H]  raw recurrent product, saved when training with mi         */
  float  zoneout;           /* level in [0,1): zoneout on h and c (0 = off)                                  */
  const float* zmask;       /* f32 [2 (h|c), 2, T, H]  train-phase keep masks, NULL = inference blend 1-level */
  /* 16-bit storage + TMA staging (lstm_tc4.cu; asr_lstm_fp16_storage() == 1): when zx16 is set it replaces zx, and
   * training saves gates16 / cell16 instead of gates / cell.  Halves the recurrences' HBM traffic; fp32 arithmetic. */
  const void* zx16;         /* fp16 [T, N, 2, 4H]                                                            */
  void*  gates16;           /* fp16 [T, N, 2, 4H]                                                            */
  void*  cell16;            /* fp16 [T, N, 2, H]                                                             */
  int32_t opts;             /* ASR_LSTM_* flags                                                              */
} asr_lstm_fwd_args;

/* opts bits of both argument records */
#define ASR_LSTM_SHARED_SM 1   /* do not reserve the whole SM's shared memory: other kernels' CTAs (e.g. a beam search
                                  on a second stream) may then share SMs with the recurrence                        */
#define ASR_LSTM_PIN_FP32  2   /* run the exact fp32 CUDA-core engine (the parity reference; N <= 32, small H)       */
#define ASR_LSTM_GROUP16   4   /* 16-sample CTA groups even when 8-sample groups fit one cooperative wave           */

typedef struct {
  int32_t T, N, H;
  const float* dh;          /* f32 [T, N, 2H]  dL/d(layer output)             */
  const float* gates;       /* saved by forward                               */
  const float* cell;
  const float* U;           /* f32 [2, H, 4H]                                 */
  const void*  U16;         /* bf16 [2, H, 4H] (tensor-core operand)          */
  void*  dz16;              /* bf16 [T, N, 2, 4H]                             */
  void*  dzT16;             /* bf16 [2*4H, T*N]                               */
  float* dz32;              /* f32 [T, N, 2, 4H] optional                     */
  float* dbias;             /* f32 [2, 4H]  (overwritten)                     */
  int32_t* flags;
  const float* mask_u;      /* same mask as in the forward call, or NULL      */
  /* fused dropout inputs (asr_lstm_fuses_masks() == 1): dL/d(output) = dh * mask_dh[0] + dh2 * mask_dh[1] */
  const float* dh2;         /* f32 [T, N, 2H]  second partial (the bwd LSTM of the layer above), or NULL    */
  const float* mask_dh;     /* f32 [2, N, 2H]  B_W of the layer above (fwd | bwd), or NULL = ones            */
  /* element-wise switches (see asr_lstm_fwd_args); with mi, dz16 / dzT16 carry dL/d(zx) (dW, dX) and            */
  const float* mi;          /* duhT16 carries dL/d(uh) (dU); dmi receives the alpha|beta1|beta2 gradients    */
  const float* zx;          /* f32 [T, N, 2, 4H]  the forward call's zx (mi)                                 */
  const float* uh;          /* f32 [T, N, 2, 4H]  saved by the forward call (mi)                             */
  float* dmi;               /* f32 [3, 2, 4H]  (overwritten)                                                 */
  void*  duhT16;            /* bf16 [2*4H, T*N]                                                              */
  float  zoneout;
  const float* zmask;
  const void* gates16;      /* fp16 copies saved by the forward call with zx16 set (replace gates / cell)     */
  const void* cell16;
  int32_t opts;             /* ASR_LSTM_* flags                                                              */
} asr_lstm_bwd_args;

/* 1 when the engine asr_lstm_forward/backward would select for this shape implements the fused dropout
 * fields above (mask_next/hm16/hmT16/hT16u, dh2/mask_dh); 0 = the caller must mask with asr_mask_cast /
 * asr_mask_combine instead. */
int32_t asr_lstm_fuses_masks(int32_t T, int32_t N, int32_t H, int32_t opts);
/* 1 when that engine also implements the element-wise switches (mi / zoneout fields); 0 = use the general cell. */
int32_t asr_lstm_fuses_variants(int32_t T, int32_t N, int32_t H, int32_t opts);
/* 1 when the 16-bit-storage / TMA engine (zx16, gates16, cell16) takes this shape */
int32_t asr_lstm_fp16_storage(int32_t T, int32_t N, int32_t H, int32_t opts);
/* 1 when asr_lstm_forward/backward take this shape (a persistent engine exists for it); 0 = use the general-cell
 * entry points below (any N, H <= 1024).  The tensor-core engine is instantiated for H in {128, 256, 384, 512, 640,
 * 768, 832, 896} and any N that is a multiple of 8 (batches wider than one cooperative wave run as several
 * launches over consecutive batch groups).  Other widths up to 896, e.g. H = 800 of BASELINE config 4, are meant to
 * be run zero-padded at the next instantiated width (W = U = b = 0 for the extra units keeps them at exactly 0 in
 * both passes; asr-study_b200/engine.py does this at parameter load / export), ragged batches padded with zero
 * utterances and zero dlogits rows. */
int32_t asr_lstm_persistent_supported(int32_t T, int32_t N, int32_t H, int32_t training, int32_t opts);

size_t  asr_lstm_flags_bytes(void);
int32_t asr_lstm_forward(const asr_lstm_fwd_args* a, void* stream);
int32_t asr_lstm_backward(const asr_lstm_bwd_args* a, void* stream);

/* ------------------------------------------------------------------------- *
 * K3/K4 (general cell)  LSTM.step with the brsmv1 switches
 * replaces core/layers.py:432-469 with layer normalisation (:407-430,
 * core/layers_utils.py:16-19), multiplicative integration (:441-443) and
 * zoneout (:457-467, layers_utils.py:34-42).  fp32 throughout; takes the same
 * argument records as the default engine (zx = K.dot(x*B_W, W) WITHOUT bias).
 * Every vector is f32 [2, 4H] (fwd|bwd direction) unless noted; a NULL group
 * switches the feature off.
 * ------------------------------------------------------------------------- */
typedef struct {
  const float* mi_alpha;    /* z = alpha*Wx*Uh + beta1*Uh + beta2*Wx + b             */
  const float* mi_beta1;
  const float* mi_beta2;
  const float* ln_gain_uh;  /* LN on K.dot(h*B_U, U)                                  */
  const float* ln_bias_uh;
  const float* ln_gain_wx;  /* LN on K.dot(x*B_W, W)                                  */
  const float* ln_bias_wx;
  const float* ln_gain_c;   /* f32 [2, H]  LN on the new cell (h = o * tanh(LN(c)))   */
  const float* ln_bias_c;
  float ln_eps;             /* 1e-5 in the reference                                  */
  float zoneout_h;          /* level in [0,1); 0 = off                                */
  float zoneout_c;
  const float* zmask_h;     /* f32 [2, T, H] keep masks of the train phase (one per time step,
                               shared by the batch), NULL = inference blend with (1 - level)   */
  const float* zmask_c;
} asr_lstm_variant;

typedef struct {            /* parameter gradients, overwritten; same shapes as above */
  float* mi_alpha;
  float* mi_beta1;
  float* mi_beta2;
  float* ln_gain_uh;
  float* ln_bias_uh;
  float* ln_gain_wx;
  float* ln_bias_wx;
  float* ln_gain_c;
  float* ln_bias_c;
} asr_lstm_variant_grads;

/* forward: needs a->U (fp32) and a->h32 or a->h16; training also fills gates/cell and
 * uh_raw f32 [T, N, 2, 4H] (the pre-LN recurrent product the backward pass re-normalises). */
int32_t asr_lstm_cell_forward(const asr_lstm_fwd_args* a, const asr_lstm_variant* v,
                              float* uh_raw, void* stream);
/* backward: a->dz32 receives dL/d(zx) (the W-side gradient: dW, dX), duh f32 [T, N, 2, 4H]
 * receives dL/d(uh_raw) (the U-side gradient: dU); a->dbias and g->* are overwritten.
 * a->dz16 / a->dzT16 are not written by this engine. */
int32_t asr_lstm_cell_backward(const asr_lstm_bwd_args* a, const asr_lstm_variant* v,
                               const float* zx, const float* uh_raw, float* duh,
                               const asr_lstm_variant_grads* g, void* stream);

/* ------------------------------------------------------------------------- *
 * K6  CTC loss + gradient      replaces tf.nn.ctc_loss, core/ctc_utils.py:68-70
 * logits f32 [T, N, C] time-major; softmax applied inside; blank = C-1 in the
 * reference.  labels: flat i32 + offsets i32 [N+1].  loss f32 [N];
 * grad f32 [T, N, C] = grad_scale * dloss_n/dlogits (0 for t >= in_len[n]).
 * ------------------------------------------------------------------------- */
size_t  asr_ctc_workspace_bytes(int32_t T, int32_t N, int32_t max_label_len);
int32_t asr_ctc_loss_grad(const float* logits, int32_t T, int32_t N, int32_t C,
                          const int32_t* in_len, const int32_t* labels,
                          const int32_t* label_off, int32_t max_label_len,
                          int32_t blank, float grad_scale, float* loss,
                          float* grad, void* ws, void* stream);

/* K7  best-path decode         replaces tf.nn.ctc_greedy_decoder, ctc_utils.py:42
 * out_labels i32 [N, T] (-1 padded, like to_dense, core/layers_utils.py:54-57),
 * out_len i32 [N]. */
int32_t asr_ctc_greedy(const float* logits, int32_t T, int32_t N, int32_t C,
                       const int32_t* in_len, int32_t blank, int32_t merge_repeated,
                       int32_t* out_labels, int32_t* out_len, void* stream);

/* K10 label error rate         replaces tf.edit_distance(hyp, truth, normalize=True), core/metrics.py:4-8
 * hyp i32 [N, hyp_stride] (-1 padded: what K7 / K8 emit) with hyp_len i32 [N] (or NULL: count the leading labels >= 0),
 * truth i32 flat + truth_off i32 [N+1] (the sparse-label contract of K6), max_truth_len >= every truth length.
 * out f32 [N] = Levenshtein distance (/ truth length when normalize; empty truth: 0 for an empty hyp, +inf otherwise). */
int32_t asr_edit_distance(const int32_t* hyp, int32_t N, int32_t hyp_stride, const int32_t* hyp_len,
                          const int32_t* truth, const int32_t* truth_off, int32_t max_truth_len,
                          int32_t normalize, float* out, void* stream);

/* K8  prefix beam search       replaces tf.nn.ctc_beam_search_decoder
 * (top_paths = 1), core/ctc_utils.py:44-50; utils/core_utils.py:70-71 */
size_t  asr_ctc_beam_workspace_bytes(int32_t T, int32_t N, int32_t C, int32_t beam_width);
int32_t asr_ctc_beam(const float* logits, int32_t T, int32_t N, int32_t C,
                     const int32_t* in_len, int32_t blank, int32_t beam_width,
                     int32_t merge_repeated, int32_t* out_labels,
                     int32_t* out_len, void* ws, void* stream);

/* ------------------------------------------------------------------------- *
 * K9  global-norm clip + Adam / SGD-momentum   replaces Keras-1 optimizers via
 * train.py:133-137.  One flat fp32 bucket.  decay_mask (u8 [n] or NULL) marks
 * elements that carry the l2(weight_decay) regulariser (core/models.py:263-264,
 * 279): g <- grad_scale*g + 2*wd*p for those.  asr_grad_sqnorm writes
 * sum(g^2) (f64) to *sqnorm; asr_adam_step reads it for the clip.
 * ------------------------------------------------------------------------- */
int32_t asr_grad_sqnorm(const float* grad, const float* param,
                        const uint8_t* decay_mask, int64_t n, float grad_scale,
                        float weight_decay, double* sqnorm, void* stream);
int32_t asr_adam_step(float* param, const float* grad, float* m, float* v,
                      const uint8_t* decay_mask, int64_t n, float grad_scale,
                      float weight_decay, const double* sqnorm, float clipnorm,
                      float lr, float beta1, float beta2, float eps,
                      int32_t step, void* stream);
int32_t asr_sgd_step(float* param, const float* grad, float* mom,
                     const uint8_t* decay_mask, int64_t n, float grad_scale,
                     float weight_decay, const double* sqnorm, float clipnorm,
                     float lr, float momentum, void* stream);

/* ------------------------------------------------------------------------- *
 * layout / precision helpers used between the kernels above
 * ------------------------------------------------------------------------- */
/* dst16[r, c] = cast(src[r, c]) for r<rows, c<cols; zero-fills c in [cols, min(ld_dst, roundup8(cols))).
 * dtype 0 = fp16, 1 = bf16 */
int32_t asr_cast_rows(const float* src, int64_t ld_src, void* dst16, int64_t ld_dst,
                      int64_t rows, int32_t cols, int32_t dtype, void* stream);
/* dst16[c, r] = cast(src[r, c])  (transpose), ld_dst >= rows */
int32_t asr_cast_transpose(const float* src, int64_t ld_src, void* dst16, int64_t ld_dst,
                           int64_t rows, int32_t cols, int32_t dtype, void* stream);
/* the same two operations for a whole list of tensors in one launch (the per-step 16-bit operand copies of every weight) */
typedef struct {
  const float* src; int64_t ld_src;
  void* dst16;      int64_t ld_dst;
  int64_t rows;     int32_t cols;
  int32_t dtype;            /* 0 = fp16, 1 = bf16, 16 = fp16 rounding residual                 */
  int32_t transpose;        /* 0: asr_cast_rows semantics, 1: asr_cast_transpose semantics, 2: rows without the K-padding
                               fill (a column sub-block next to another job's: jobs of one launch run concurrently) */
} asr_cast_job;
int32_t asr_cast_batch(const asr_cast_job* jobs_host, int32_t n_jobs, void* stream);
/* variational-dropout operand views (core/layers.py:439: x * B_W[0], mask constant over time):
 * rows are time-major r = t*n_batch + n; mask f32 [n_batch, cols] (already scaled by 1/(1-p)).
 * src_dtype: 0 = fp16, 1 = bf16, 2 = fp32.  dst16[r, c] = cast(src[r, c] * mask[r % n_batch, c]);
 * transpose != 0 writes dst16[c, r] instead (ld_dst >= rows).  Zero-fills the K padding like asr_cast_rows. */
int32_t asr_mask_cast(const void* src, int32_t src_dtype, int64_t ld_src, const float* mask, int32_t n_batch,
                      void* dst16, int32_t dtype, int64_t ld_dst, int64_t rows, int32_t cols,
                      int32_t transpose, void* stream);
/* out[r, c] = a[r, c] * mask_a[r % n_batch, c] + b[r, c] * mask_b[r % n_batch, c]   (dX of the two directions) */
int32_t asr_mask_combine(const float* a, const float* b, const float* mask_a, const float* mask_b,
                         int32_t n_batch, float* out, int64_t rows, int32_t cols, void* stream);
/* out = (a + b) * mask[r % n_batch, c]  (b, mask optional; out may alias a): residual merge(mode='sum')
 * (core/models.py:273-274) and the element-wise input Dropout (:257-258, n_batch = rows) with its backward. */
int32_t asr_add_mask(const float* a, const float* b, const float* mask, int64_t n_batch,
                     float* out, int64_t rows, int32_t cols, void* stream);
/* variational-dropout masks (core/layers.py:306-339 under Keras-1 K.dropout): out[i] = keep ? 1/(1-p) : 0,
 * keep ~ Bernoulli(1-p) from a counter-based generator keyed by (seed, offset + i): one launch per step. */
int32_t asr_dropout_mask(float* out, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream);
/* same generator with a caller-chosen keep value: out[i] = keep ? keep_value : 0.  Zoneout keep masks (0 / 1,
 * core/layers_utils.py:34-42) and the element-wise input Dropout (0 / 1/(1-p), core/models.py:257-258). */
int32_t asr_bernoulli_mask(float* out, int64_t n, float p, float keep_value, uint64_t seed, uint64_t offset,
                           void* stream);
/* GaussianNoise(std) (core/models.py:67,251; train phase only): x[r, c] += std * N(0,1) for r < rows, c < cols, row
 * stride ld, from the same counter-based generator (Box-Muller). */
int32_t asr_add_gaussian_noise(float* x, int64_t rows, int32_t cols, int64_t ld, float stdv, uint64_t seed,
                               uint64_t offset, void* stream);
/* out[c] = sum_r src[r, c]  (fp32; bias gradients) */
int32_t asr_colsum(const float* src, int64_t ld, int64_t rows, int32_t cols,
                   float* out, void* stream);

/* ------------------------------------------------------------------------- *
 * Convolutional front end of BASELINE configs[3] (DeepSpeech2-style 2 x Conv in front of the BiLSTM stack; NOT in
 * the reference — README.md:118 lists it as future work — so its semantics are this header's: cross-correlation with
 * zero padding, bias, clipped ReLU min(max(z, 0), clip)).  Implicit GEMM without a patch matrix (csrc/conv.cu):
 * activations are fp16 [N, t_padded, F * C], batch-major, the T valid frames of an utterance at rows [pt, pt + T) between
 * zero rows; the convolution window of output frame (n, t') is kt * F * C CONTIGUOUS values starting at element
 * (n * t_padded + st * t') * F * C, i.e. row n * rows + t' of a matrix with row stride st * F * C < its row length —
 * asr_gemm_tn takes that view as its A operand (lda < K); the frequency taps are folded into a banded weight matrix.
 * These entry points are the layout kernels around the GEMMs.
 * ------------------------------------------------------------------------- */
typedef struct {
  int32_t T, N, F, C;       /* input frames, utterances, frequency bins, channels (st * F * C must be a multiple of 8) */
  int32_t kt, kf;           /* kernel extent in time / frequency                   */
  int32_t st, sf;           /* strides                                              */
  int32_t pt, pf;           /* zero padding on both sides                           */
} asr_conv_geom;

typedef struct {
  int32_t t_out, f_out;     /* output frames / bins                                                              */
  int32_t rows;             /* GEMM rows per utterance (>= t_out, multiple of 8); rows t' >= t_out are scratch     */
  int32_t t_padded;         /* frames per utterance of the padded input = st * rows (>= T + 2 pt)                 */
  int32_t k, k_padded;      /* kt * F * C and its padding to a multiple of 8                                      */
} asr_conv_plan;

int32_t asr_conv_plan_for(const asr_conv_geom* geom, asr_conv_plan* plan);
/* x f32 [T, N, F * C] time-major (what the feature kernel emits) -> xp16 fp16 [N, t_padded, F * C] rows [pt, pt + T); the
 * caller zeroes the buffer once (plus k_padded elements of slack behind it), nothing writes the padding rows */
int32_t asr_conv_pack(const float* x, const asr_conv_geom* geom, void* xp16, void* stream);
/* w f32 [c_out, kt, kf, C], b f32 [c_out] -> wt16 fp16 [f_out * c_out, ldw]: the forward GEMM's B operand
 * wt[(f', co), (dkt, f, c)] = w[co, dkt, f - sf * f' + pf, c] (0 outside the kernel and in the K padding);
 * w2_16 (optional) bf16 [F * C, kt * f_out * c_out]: the input-gradient GEMM's B operand (taps mirrored in time);
 * bias_t f32 [f_out * c_out]: the bias tiled over the output bins (the GEMM's bias argument) */
int32_t asr_conv_toeplitz(const float* w, const float* b, const asr_conv_geom* geom, int32_t c_out, void* wt16,
                          int64_t ldw, void* w2_16, float* bias_t, void* stream);
/* z f32 [N * rows, f_out * c_out] (GEMM output) -> y = min(max(z, 0), clip) for t' < t_out: fp16 into rows
 * [y_row0, y_row0 + t_out) of a [N, y_rows, f_out * c_out] buffer (the next layer's padded input) and / or f32 time-major
 * [t_out, N, f_out * c_out] (the first BiLSTM's input) */
int32_t asr_conv_act(const float* z, const asr_conv_geom* geom, int32_t c_out, float clip, void* y16, int32_t y_rows,
                     int32_t y_row0, float* y32_tm, void* stream);
/* g = gout * [0 < y < clip] in GEMM-row space (0 for t' >= t_out): g16 bf16 [N * rows, Wo], gT16 bf16 [Wo, N * rows],
 * g32 f32 [N * rows, Wo] (each optional; Wo = f_out * c_out).  gout f32: element (n, t') at row
 * t' * g_t_stride + n * g_n_stride + g_row0 (time-major: N, 1, 0; batch-major padded: 1, t_padded, pt); y16 as in asr_conv_act */
int32_t asr_conv_act_backward(const float* gout, int64_t g_t_stride, int64_t g_n_stride, int64_t g_row0, const void* y16,
                              int32_t y_rows, int32_t y_row0, const asr_conv_geom* geom, int32_t c_out, float clip,
                              void* g16, void* gT16, float* g32, void* stream);
/* out16 bf16 [k, ldT]: out[m, r] = xp16_flat[r * st * F * C + m] — the overlapping view transposed, the K-major operand
 * of the weight-gradient GEMM dwt [Wo, k] = gT16 . out16^T */
int32_t asr_conv_unfold_t(const void* xp16, const asr_conv_geom* geom, void* out16, int64_t ldT, void* stream);
/* dwt f32 [Wo, ld] (gradient of the banded matrix), colsum f32 [Wo] (column sums of g32) -> dw f32 [c_out, kt, kf, C], db f32 [c_out] */
int32_t asr_conv_toeplitz_grad(const float* dwt, int64_t ld, const float* colsum, const asr_conv_geom* geom, int32_t c_out,
                               float* dw, float* db, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ASR_B200_H */

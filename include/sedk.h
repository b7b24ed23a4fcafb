/* libsedk - C ABI of the B200-native sound-event-detection hot path (drop-in for DCASE-REPO/DESED_task).
 *
 * The reference has no FFI of its own (it is pure Python over PyTorch / torchaudio); these entry points are
 * what a maintainer binds (ctypes stub in INTEGRATION.md) behind the reference's Python call sites.  Every
 * function cites the reference interface it replaces (paths relative to the upstream repo).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless its name ends in _host or the comment says otherwise;
 *   - caller owns all buffers; no hidden allocation, no hidden synchronisation; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), so calls are CUDA-graph capturable;
 *   - return value: SEDK_OK (0) or a negative error code; sedk_last_error() gives the message (thread-local);
 *   - activations are channels-last: [B, T, F, C] ("T" = time frames = conv H, "F" = mel bins = conv W);
 *   - sm_100a only.  There is no CPU fallback anywhere.
 */
#ifndef SEDK_H_
#define SEDK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SEDK_API __attribute__((visibility("default")))
#else
#define SEDK_API
#endif

#define SEDK_OK 0
#define SEDK_ERR_INVALID (-1)     /* bad argument (shape, alignment, null pointer)            */
#define SEDK_ERR_CUDA (-2)        /* a CUDA runtime call / kernel launch failed               */
#define SEDK_ERR_UNSUPPORTED (-3) /* configuration outside what the sm_100a kernels implement */

#define SEDK_MAX_CONV 8
#define SEDK_MAX_GRU_LAYERS 4

SEDK_API const char* sedk_last_error(void);
SEDK_API int sedk_version(void);
/* compute capability of the current device as major*10+minor (100 on B200); <0 on error */
SEDK_API int sedk_device_cc(void);
/* The convolutions with 32/64/128 channels run on tcgen05 (TMEM accumulators, TMA-fed) in the TF32 precision mode; this
 * switch (default on; env SEDK_DISABLE_TCGEN05=1 turns it off) selects the legacy mma.sync kernels instead - kept for
 * A/B parity tests. */
SEDK_API int sedk_set_tcgen05(int on);
SEDK_API int sedk_get_tcgen05(void);
/* H = 128 GRU recurrence as a 2-CTA cluster (cs = 2) instead of one CTA per (row, direction) (cs = 1, default) */
SEDK_API int sedk_set_gru_cluster(int cs);
/* Named integer switches selecting between kernel variants (A/B measurements, parity tests).  Unset options take their
 * default from the environment variable SEDK_<NAME> (upper case), else the built-in default.  Known names:
 *   "gru_v3"      1 (default): H = 128 recurrence, third generation (csrc/gru3.cu): 8 warps, every W_hh weight in registers,
 *                 octet-per-4-units layout (4 LDS.128 of h per thread and step, transposing-butterfly reductions); 2: the
 *                 16-warp variant of the same layout; 0: fall through to "gru_v2"
 *   "pdl"         0 (default): plain stream order; 1: the kernels on the step's dependency chain are launched with
 *                 programmatic stream serialisation (programmatic dependent launch: kernel k + 1 is scheduled while kernel k
 *                 drains and blocks in griddepcontrol.wait until k has completed).  Parity-tested; measured slower on B200
 *                 (2.205 vs 2.188 ms per supervised step: the waiting CTAs take SM slots from the parallel graph branches)
 *   "logmel_v2"   1 (default): second-generation front end (csrc/logmel2.cu); 0: first generation (csrc/logmel.cu)
 *   "gru_v2"      1 (default): H = 128 recurrence with the quad-per-unit layout (shuffle reductions, one barrier per
 *                 step); 0: first-generation kernel (row x k-segment layout, partial sums through shared memory)
 *   "bnglu_small" 1 (default): register-resident warp-autonomous BN+GLU+pool kernels for 16 / 32 channels; 0: tiled kernel
 *   "bnglu_tc5"   1 (default): tcgen05 / TMEM / TMA BN+GLU+pool kernels for the 128-channel layers (TF32 mode)
 *   "gemm_tc5"    1 (default): tcgen05 GEMMs for the GRU input projections and input gradients (TF32 mode)
 *   "side_stream" 1 (default): weight-gradient GEMMs and weight packs on a forked side stream
 *   "conv_pair"   0 (default): 1 runs the 16 <-> 32 channel convolutions on tcgen05 through a paired-pixel view (TF32 mode);
 *                 measured slower than the halo-staged mma.sync kernel on B200, kept as a parity-tested experiment */
SEDK_API int sedk_set_option(const char* name, int value);
SEDK_API int sedk_get_option(const char* name, int dflt);
/* number of kernels this library has launched (or captured into a CUDA graph) so far in this process */
SEDK_API long long sedk_launch_count(void);
/* Optional eager-mode kernel timing: between sedk_profile_enable(1) and sedk_profile_report every launcher is bracketed by
 * CUDA events on its stream; the report is text, one line per kernel family: "<name> <launches> <total_ms>". */
SEDK_API int sedk_profile_enable(int on);
SEDK_API int sedk_profile_report(char* buf, int buflen);

/* ------------------------------------------------------------------------------------------------------------
 * Front end: waveform -> (log-)mel.  Replaces torchaudio MelSpectrogram + AmplitudeToDB as built at
 * recipes/dcase2023_task4_baseline/local/sed_trainer.py:79-91 and called at :282 / :253-264 (take_log).
 * n_fft = win_length = 2048 (fixed by the kernel), center=True, reflect padding, onesided, power=1, HTK fb.
 *
 * Tables (device, built once by the host module):
 *   window[2048]; tw2048[1024] float2 = exp(-2 pi i k/2048); tw32x32[32*32] float2 = exp(-2 pi i n2*k1/1024)
 *   laid out [k1][n2]; sparse filterbank: fb_start[n_mels], fb_len[n_mels], fb_off[n_mels] (int32) and
 *   fb_w[sum len] (the non-zero run of column m of torchaudio's fb[1025, n_mels]).
 * wave  [B, L] contiguous.  out element (b, m, t) at out + b*out_sb + m*out_sm + t*out_st,
 *   t in [0, 1 + L/hop).  log_mode 0: linear-amplitude mel;  1: 20*log10(max(x, amin)) clamped to [db_lo, db_hi].
 * minmax: optional [B][2] uint32 (order-preserving encoding, see sedk_minmax_*): per-clip min / max of what was
 *   written; must be initialised with sedk_minmax_init.  May be NULL.
 */
typedef struct {
    const float* window;
    const float* tw2048;  /* float2[1024] */
    const float* tw32x32; /* float2[1024] */
    const int32_t* fb_start;
    const int32_t* fb_len;
    const int32_t* fb_off;
    const float* fb_w;
    int32_t n_mels;
    int32_t hop;
} sedk_mel_tables;

SEDK_API int sedk_logmel_fwd(const float* wave, int B, int L, const sedk_mel_tables* tab, float* out, int64_t out_sb,
                    int64_t out_sm, int64_t out_st, int log_mode, float amin, float db_lo, float db_hi,
                    uint32_t* minmax, void* stream);

/* Same front end on 16-bit PCM (the format the datasets are stored in): every sample is used as x / 32768, exactly what
 * torchaudio.load hands the reference (desed_task/dataio/datasets.py:24-74), so the result is bit-identical to
 * sedk_logmel_fwd on the normalised fp32 waveform while the H2D copy and the HBM read halve. */
SEDK_API int sedk_logmel_fwd_i16(const int16_t* wave, int B, int L, const sedk_mel_tables* tab, float* out, int64_t out_sb,
                        int64_t out_sm, int64_t out_st, int log_mode, float amin, float db_lo, float db_hi,
                        uint32_t* minmax, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Elementwise feature ops.
 */
/* minmax[b] = {ord(+inf), ord(-inf)} */
SEDK_API int sedk_minmax_init(uint32_t* minmax, int B, void* stream);
/* decode to float pairs {min, max} */
SEDK_API int sedk_minmax_decode(const uint32_t* minmax, float* out, int B, void* stream);

/* take_log (sed_trainer.py:253-264) fused with mixup on linear mel (desed_task/data_augm.py:31-37, called at
 * sed_trainer.py:296-301) and the per-clip min/max of TorchScaler (desed_task/utils/scaler.py:114-120):
 *   v = perm ? c[b]*x[b] + (1-c[b])*x[perm[b]] : x[b];  out[b] = log_mode ? clamp(20*log10(max(v,amin))) : v
 * x/out: [B, n] contiguous per clip.  perm: int64[B] device or NULL; coef: float[B] device or NULL (c=1).
 * minmax: optional, as above.  */
SEDK_API int sedk_feat_mix_log(const float* x, const int64_t* perm, const float* coef, float* out, int B, int64_t n,
                      int log_mode, float amin, float db_lo, float db_hi, uint32_t* minmax, void* stream);

/* TorchScaler instance/minmax apply (scaler.py:114-120): out = (x-min)/(max-min+eps)*2-1 */
SEDK_API int sedk_minmax_scale(const float* x, float* out, const uint32_t* minmax, int B, int64_t n, float eps, void* stream);
/* per-clip mean and (unbiased) std over n elements (scaler.py:107-112): stats[b] = {mean, std} */
SEDK_API int sedk_instance_stats(const float* x, float* stats, int B, int64_t n, void* stream);
/* out = (x - a[b or 0]) * s[b or 0] ... generic affine used by the mean/standard/dataset scaler modes:
 * out[b,i] = (x[b,i] - sub[b*sub_sb + i*sub_si]) * mul[b*mul_sb + i*mul_si] */
SEDK_API int sedk_affine_bcast(const float* x, float* out, const float* sub, int64_t sub_sb, int64_t sub_si, const float* mul,
                      int64_t mul_sb, int64_t mul_si, int B, int64_t n, void* stream);

/* label mixup (data_augm.py:38-44): soft: clamp(c*y + (1-c)*y[perm], 0, 1); hard: clamp(y + y[perm], 0, 1) */
SEDK_API int sedk_label_mix(const float* y, const int64_t* perm, const float* coef, float* out, int B, int64_t n, int hard,
                   void* stream);
/* frame_shift (data_augm.py:7-16): out[b, r, (j + shift[b]) mod n_cols] = x[b, r, j];  shift int32[B] device */
SEDK_API int sedk_roll_last(const float* x, float* out, const int32_t* shift, int B, int rows, int cols, void* stream);
/* add_noise (data_augm.py:56-77): out = x + noise * std(x[b]) / 10^(snr_db[b]/20); stats from sedk_instance_stats */
SEDK_API int sedk_add_noise(const float* x, const float* noise, const float* snr_db, const float* stats, float* out, int B,
                   int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimiser / mean teacher: one flat multi-tensor kernel.
 * update_ema (sed_trainer.py:187-199) + torch.optim.Adam (train_sed.py:199-201).  Order inside the kernel follows
 * the reference step order under PL 1.9 (EMA of the *current* weights first, then the Adam update):
 *   if (ema)  ema = ema_alpha*ema + (1-ema_alpha)*p
 *   if (do_adam) { g *= grad_scale; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 *                  p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps) }         bc_i = 1 - b_i^step
 */
SEDK_API int sedk_adam_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n, int do_adam, float lr,
                  float beta1, float beta2, float eps, int step, float ema_alpha, float grad_scale, void* stream);
/* Same kernel with the per-step scalars read from DEVICE memory so that a captured CUDA graph can be replayed with new
 * values: hyper[4] = { lr / (1 - beta1^step), 1 / sqrt(1 - beta2^step), ema_alpha, grad_scale }. */
SEDK_API int sedk_adam_ema_dev(float* p, const float* g, float* m, float* v, float* ema, int64_t n, int do_adam, float beta1,
                      float beta2, float eps, const float* hyper, void* stream);
/* Data-parallel step without NCCL: gradient all-reduce (SUM) over NVLink / NVSwitch peer memory fused with the update above
 * in ONE launch (csrc/nvls.cu).  There is no reference counterpart (train_sed.py:269-276 refuses > 1 GPU); it replaces
 * `dist.all_reduce(flat_grad)` + sedk_adam_ema_dev of this library's own data-parallel path.
 *   g_peers[world]    the flat gradient buffer of EVERY rank as mapped into this process (g_peers[rank] is the local one,
 *                     n floats each, 16-byte aligned) - a symmetric allocation, e.g. torch.distributed._symmetric_memory;
 *   g_mc              the multicast address covering all of them (NVLS: the switch reduces on multimem.ld_reduce and
 *                     replicates on multimem.st), or NULL: peers are read / written one by one (P2P);
 *   flag_peers[world] per-rank flag blocks of sedk_nvls_flag_bytes() bytes, zero before the first call, same mapping rule.
 * On return (stream order) every rank's gradient buffer holds the sum - each element reduced exactly once, so replicas stay
 * bit-identical - and, with do_adam = 1, p / m / v / ema have been updated as by sedk_adam_ema_dev (hyper[3] = 1 / world
 * folds the mean).  do_adam = 0: all-reduce only (gradient clipping needs the global norm first).  Every rank must make
 * the same sequence of calls; world <= 8 (one NVSwitch domain).  Capturable in a CUDA graph. */
SEDK_API int64_t sedk_nvls_flag_bytes(void);
/* A barrier wait inside sedk_allreduce_adam_nvls gives up after 10 s (a peer process died): returns 1 once if that happened
 * since the last call (synchronises the device), 0 otherwise.  After a fault the buffers of this process are invalid. */
SEDK_API int sedk_nvls_fault(void);
/* diagnostic (option "nvls_debug" = 1): device-clock stamps (ns) of CTA 0 in the last launch - start, after barrier A,
 * after phase 1, after barrier B, end */
SEDK_API int sedk_nvls_debug_stamps(uint64_t* out5);
SEDK_API int sedk_allreduce_adam_nvls(float* p, float* m, float* v, float* ema, int64_t n, int do_adam, float beta1,
                             float beta2, float eps, const float* hyper, void* g_mc, void* const* g_peers,
                             void* const* flag_peers, int rank, int world, void* stream);
/* *counter += inc (one thread); pair with sedk_crnn_plan.seed_dev */
SEDK_API int sedk_bump_counter(uint64_t* counter, uint64_t inc, void* stream);
/* SpecAugment / dropstep span draws on the device (CRNN.apply_specaugment, desed_task/nnet/CRNN.py:207-219 and :288-293;
 * torchaudio mask_along_axis_iid semantics: value = u*param, start = floor(u'*(size - value)), end = start + floor(value)):
 * out int32 [B][4] = {start_a, end_a, start_b, end_b}; param < 1 disables an axis.  Philox keyed by (seed + *seed_dev,
 * stream_id, example) so that a captured CUDA graph draws fresh spans on every replay. */
SEDK_API int sedk_mask_spans(int32_t* out, int B, int size_a, int param_a, int size_b, int param_b, uint64_t seed,
                    const uint64_t* seed_dev, uint64_t stream_id, void* stream);
/* sum of squares of g into out[0] (double), for gradient clipping (2024 recipe gradient_clip 5.0) */
SEDK_API int sedk_sumsq(const float* g, int64_t n, double* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Median-filter post-processing: scipy.ndimage.median_filter(scores[T, C], (k, 1)), mode='reflect'
 * (recipes/dcase2023_task4_baseline/local/utils.py:58; per class: desed_task/utils/postprocess.py:5-17).
 * scores/out: element (b, c, t) at base + b*sb + c*sc + t*st.  win: int32[C] device window per class (1..31).
 */
SEDK_API int sedk_median_filter(const float* scores, float* out, int B, int C, int T, int64_t sb, int64_t sc, int64_t st,
                       int64_t ob, int64_t oc, int64_t ot, const int32_t* win, void* stream);

/* Threshold + run-length event decoding of (median-filtered) frame scores on the device: replaces the per-clip D2H,
 * `c_scores > c_th` and ManyHotEncoder.decode_strong loop of batched_decode_preds
 * (recipes/dcase2023_task4_baseline/local/utils.py:64-71; desed_task/utils/encoder.py:189-211, whose
 * DecisionEncoder.find_contiguous_regions comes from dcase_util: regions [onset, offset) of consecutive true frames).
 * scores element (b, c, t) at base + b*sb + c*sc + t*st; thresholds float[n_th] (device); n_frames int32[B] (device) = true
 * length of each clip in frames (pad_indx), or NULL for T.
 * Rows are ordered (threshold, clip, class) - the order in which the reference appends to its per-threshold DataFrames.
 *   offsets int32[n_th*B*C + 1]: on return the exclusive prefix sum of the per-row event counts, offsets[rows] = total
 *   events  int32[capacity][2] : {onset_frame, offset_frame} of event e of row r at events[offsets[r] + e]
 * Events beyond `capacity` are dropped (the total still reports all of them, so the caller can re-run with a larger buffer). */
SEDK_API int sedk_decode_events(const float* scores, int B, int C, int T, int64_t sb, int64_t sc, int64_t st,
                       const float* thresholds, int n_th, const int32_t* n_frames, int32_t* offsets, int32_t* events,
                       int capacity, void* stream);

/* Embedding storage format.  The 2024 recipe stores BEATs frame embeddings as fp32 [768, 496] per clip
 * (1.52 MB; recipes/dcase2024_task4_baseline/extract_embeddings.py:48-53, desed_task/dataio/datasets.py:221-228) and the
 * CRNN pools them to its 156 frames in every forward (CRNN.py:280-283).  sedk_pool_embeddings does that aggregation once
 * (mode 0: adaptive_avg_pool1d, 1: nearest-exact interpolate; the arithmetic of the fusion kernel, so the fp32 output is
 * what the forward would compute) and writes [B, E, T] as fp32 or bf16 (240 KB per clip); sedk_bf16_to_f32 restores the
 * working precision.  A pre-pooled tensor goes through the unchanged fusion path (pooling T -> T frames is the identity). */
SEDK_API int sedk_pool_embeddings(const float* emb, void* out, int B, int E, int Te, int T, int mode, int out_bf16, void* stream);
SEDK_API int sedk_bf16_to_f32(const void* in_bf16, float* out, int64_t n, void* stream);

/* Strong-label encoding on the device: replaces the per-item pandas loop of
 * ManyHotEncoder.encode_strong_df (desed_task/utils/encoder.py:80-171, called from desed_task/dataio/datasets.py:187-237).
 * events int32 [n][4] = {clip, class, onset_frame, offset_frame} grouped by clip in their original order (frames computed by
 * the caller exactly as the reference does: int(_time_to_frame(onset)), int(ceil(_time_to_frame(offset)))); values float[n]
 * (the `confidence` column) or NULL for 1; clip_offsets int32 [B + 1].  labels [B, C, T] is overwritten: zero, then every
 * event in order labels[b][class][onset:offset] = value (later events win, like the reference's sequential assignment). */
SEDK_API int sedk_encode_strong(const int32_t* events, const float* values, const int32_t* clip_offsets, float* labels, int B,
                       int C, int T, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * CRNN (desed_task/nnet/CRNN.py, CNN.py, RNN.py) - whole-network forward / backward on a caller-provided plan.
 */
typedef struct {
    /* geometry */
    int32_t cin, cout;      /* channels                                                        */
    int32_t T, F;           /* conv input = output spatial size (3x3, pad 1, stride 1)         */
    int32_t pt, pf;         /* AvgPool2d kernel (= stride), CNN.py:96-98                       */
    /* parameters (reference layouts, desed_task state_dict) */
    const float* w;         /* conv weight [cout, cin, 3, 3]                                   */
    const float* b;         /* conv bias [cout]                                                */
    const float* gamma;     /* batchnorm weight [cout]                                         */
    const float* beta;      /* batchnorm bias [cout]                                           */
    float* running_mean;    /* [cout] (updated in training)                                    */
    float* running_var;     /* [cout]                                                          */
    int64_t* num_batches;   /* scalar int64, may be NULL                                       */
    const float* glu_w;     /* GLU linear weight [cout, cout]                                  */
    const float* glu_b;     /* [cout]                                                          */
    /* gradients (same layouts), NULL when not training */
    float *gw, *gb, *ggamma, *gbeta, *gglu_w, *gglu_b;
    /* workspace */
    float* wpack;           /* [2][9][cout][cin]: fwd pack [tap][co][ci] then dgrad pack [tap][ci][co]; layers with 16
                               channels on one side need 4x that size (paired-pixel packs of the tcgen05 path)        */
    float* gwpack;          /* [9][cout][cin] wgrad accumulator                                */
    float* z;               /* conv output (pre-BN) [B,T,F,cout]                               */
    float* gy;              /* grad wrt BN output / conv output (in place) [B,T,F,cout]        */
    float* out;             /* pooled block output [B,T/pt,F/pf,cout]                          */
    float* gout;            /* grad wrt out                                                    */
    double* stats;          /* [4][cout]: sum z, sum z^2, sum gy, sum gy*zhat                   */
    float* bn;              /* [4][cout]: scale, shift, mean, invstd (saved for backward)      */
    /* optional, cout in {64, 128} layers only (NULL: the mma.sync BN+GLU kernels are used): workspace of the tcgen05
     * BN+GLU path */
    float* glu_pack;        /* [3*128*128 + 128]: BN-folded gate weight, transposed gate weight (both 128 x 128, block-
                               diagonal for cout = 64), folded bias, weight-gradient scratch                           */
    float* lin;             /* [B,T,F,cout]: gate pre-activation saved by the forward; the backward overwrites it with
                               g_lin (the gate-output gradient), the operand of the gate weight-gradient GEMM          */
} sedk_conv_layer;

typedef struct {
    int32_t in_dim, hidden;
    /* nn.GRU parameters, [dir] = forward, reverse (RNN.py:19-30) */
    const float* w_ih[2];   /* [3H, in]  */
    const float* w_hh[2];   /* [3H, H]   */
    const float* b_ih[2];   /* [3H]      */
    const float* b_hh[2];   /* [3H]      */
    float *gw_ih[2], *gw_hh[2], *gb_ih[2], *gb_hh[2];
    /* workspace, per direction */
    float* gi[2];           /* [B,T,3H] input projections (fwd) / dgi (bwd, in place)          */
    float* gates[2];        /* [B,T,4H]: r, z, n, hn_pre (saved for backward)                  */
    float* hprev[2];        /* [B,T,H] h_{t-1} in processing order                             */
    float* dghn[2];         /* [B,T,H]                                                          */
    float* out;             /* [B,T,2H] layer output                                           */
    float* gout;            /* [B,T,2H] grad wrt out                                           */
} sedk_gru_layer;

typedef struct {
    int32_t B;
    int32_t n_mels, n_frames;       /* model input [B, n_mels, n_frames]                       */
    int32_t n_conv, n_gru;
    int32_t nclass;
    int32_t training;               /* 1: batch-stat BN + dropout + saves for backward         */
    int32_t precision;              /* 0: TF32 tensor cores; 1: 3xTF32 (fp32-equivalent)       */
    int32_t bn_eval;                /* 1: BatchNorm uses (and does not update) the running statistics although training = 1:
                                       `freeze_bn` (CRNN.py:308-323) and autograd through an eval-mode forward */
    int32_t activation;             /* CNN.py:81-88: 0 "glu", 1 "cg" (ContextGating), 2 "relu", 3 "leakyrelu" (slope 0.2) */
    float dropout_p;                /* CNN.py:90-91, CRNN.py:103                               */
    float bn_eps, bn_momentum;      /* CNN.py:76 (1e-3, 0.99)                                  */
    uint64_t seed;                  /* dropout Philox seed for this forward                    */
    const uint64_t* seed_dev;       /* optional device counter added to `seed` inside the kernels (lets a captured CUDA
                                       graph draw fresh dropout masks on every replay); NULL: unused           */
    /* input: log-mel (un-scaled) with strides; scaler + specaugment are fused into the first conv load */
    const float* x;
    int64_t x_sb, x_sm, x_st;
    const uint32_t* minmax;         /* per-clip {min,max} (instance/minmax scaler); NULL: x is already scaled */
    float scaler_eps;
    const int32_t* specaug;         /* int32 [B][4] = f_start, f_end, t_start, t_end or NULL (CRNN.py:207-219) */
    float* x0;                      /* [B, n_frames, n_mels] workspace: scaled + masked first-layer input (saved for bwd) */
    sedk_conv_layer conv[SEDK_MAX_CONV];
    sedk_gru_layer gru[SEDK_MAX_GRU_LAYERS];
    /* optional embedding fusion (aggregation_type="pool1d", CRNN.py:280-294) */
    const float* emb;               /* [B, emb_dim, emb_T] or NULL                             */
    int32_t emb_dim, emb_T;
    int32_t emb_mode;               /* aggregation_type: 0 "pool1d" (CRNN.py:280-283), 1 "interpolate" (nearest-exact, :270-278) */
    const float* cat_w;             /* [nb, nb+emb_dim]                                        */
    const float* cat_b;
    float *gcat_w, *gcat_b;
    float* cat_in;                  /* [B,T',nb+emb_dim] workspace (after dropout)             */
    float* fused;                   /* [B,T',nb] workspace                                     */
    float* gfused;
    const int32_t* dropstep;        /* int32 [B][4] x_start,x_end,e_start,e_end or NULL.  Without embeddings (emb NULL) the
                                       x span + dropout are applied to the CNN output (CRNN.py:295-301; needs cat_in
                                       [B,T',nb] and gfused) */
    /* heads (CRNN.py:152-178) */
    const float* dense_w;           /* [C, 2H] */
    const float* dense_b;
    const float* soft_w;            /* [C, 2H] */
    const float* soft_b;
    float *gdense_w, *gdense_b, *gsoft_w, *gsoft_b;
    const uint8_t* classes_mask;    /* [B, C] (1 = valid class) or NULL                        */
    float* rnn_drop;                /* [B,T',2H] workspace: post-RNN dropout output            */
    float* grnn_drop;
    float* strong;                  /* out: [B, C, T'] */
    float* weak;                    /* out: [B, C]     */
    float* sof;                     /* [B, T', C] workspace: class-softmax attention (unclamped) */
    float* hsum;                    /* [B, 2, C] workspace: sum_t strong*att, sum_t att        */
    float* gstrong;                 /* in (backward): [B, C, T'] */
    float* gweak;                   /* in (backward): [B, C]     */
    /* Optional bulk zeroing.  When the caller lays out every `stats` array of the conv layers contiguously it passes that
     * region as zero_fwd: the forward clears it with ONE memset (both halves: the backward half is only written by the
     * backward that follows).  Likewise zero_bwd covers every gradient / gradient-accumulator buffer of the plan (gw, gb,
     * ggamma, gbeta, gglu_*, gwpack, GRU and head gradients): one memset at the start of the backward replaces ~35.
     * NULL: each buffer is cleared individually (the default). */
    void* zero_fwd;
    int64_t zero_fwd_bytes;
    void* zero_bwd;
    int64_t zero_bwd_bytes;
    /* Optional: 304 doubles (SEDK_L0_SUMS).  When given, and the first block is 1 -> 16 filters / "glu" / pooling (2, 2)
     * (the shipped configs; CNN.py:66-98 with i = 0), that block runs with its convolution output recomputed from the
     * one-channel input instead of stored: conv[0].z and conv[0].gy are not touched, and the BatchNorm backward + conv
     * weight gradient are evaluated in closed form from sums accumulated here (csrc/layer0.cu).  Must be zero before a
     * training forward: place it inside zero_fwd when zero_fwd is used (checked), else the forward clears it itself.
     * NULL: the convolution output goes through HBM (conv0 + BN/GLU kernels). */
    double* l0_sums;
} sedk_crnn_plan;
#define SEDK_L0_SUMS 304

SEDK_API int sedk_crnn_forward(const sedk_crnn_plan* plan, void* stream);
SEDK_API int sedk_crnn_backward(const sedk_crnn_plan* plan, void* stream);
/* The backward in pieces, for callers that overlap a collective with it.  phases is a bit mask: 1 = heads + BiGRU
 * (+ embedding fusion); 4 = conv layers [3, n_conv); 8 = conv layers [0, 3); 2 = 4 | 8 = the whole CNN.  Pieces must run in
 * that order after one forward; every gradient of the parameters a call covered is final when its work completes.
 * phases = 3 or 15: everything (= sedk_crnn_backward). */
SEDK_API int sedk_crnn_backward_phase(const sedk_crnn_plan* plan, int phases, void* stream);
SEDK_API int sedk_sizeof_crnn_plan(void);

/* Losses of SEDTask4.training_step (sed_trainer.py:309-342): BCE(strong rows [0,n_strong)) + BCE(weak rows
 * [n_strong, n_strong+n_weak)) + weight * (MSE(strong, teacher) + MSE(weak, teacher)); teacher pointers may be NULL.
 * labels [B,C,T'], labels_weak [n_weak, C].  `losses` is a 16-float device buffer: on return losses[0..8) =
 * {total, bce_strong, bce_weak, mse_strong, mse_weak, bce_strong_teacher, bce_weak_teacher, cons_weight}
 * (losses[8..16) is scratch); gstrong / gweak (may be NULL) receive d total / d strong, d total / d weak. */
SEDK_API int sedk_sed_loss(const float* strong, const float* weak, const float* t_strong, const float* t_weak,
                  const float* labels, const float* labels_weak, int B, int C, int T, int n_strong, int n_weak,
                  float cons_weight, float* losses, float* gstrong, float* gweak, void* stream);

/* Same, with the consistency weight read from device memory (CUDA-graph replay with a new ramp value each step). */
SEDK_API int sedk_sed_loss_dev(const float* strong, const float* weak, const float* t_strong, const float* t_weak,
                      const float* labels, const float* labels_weak, int B, int C, int T, int n_strong, int n_weak,
                      const float* cons_weight_dev, float* losses, float* gstrong, float* gweak, void* stream);

/* General form.  cons_row0: the consistency term covers rows [cons_row0, B) only - 0 in the 2023 recipe, `indx_maestro` in the
 * 2024 one (`mask_unlabeled`, recipes/dcase2024_task4_baseline/local/sed_trainer_pretrained.py:343-346,408-415); its mean is
 * taken over those rows.  cons_kind: 0 = MSELoss, 1 = BCELoss(student, teacher) (`self_sup_loss: bce`,
 * recipes/dcase2023_task4_baseline/local/sed_trainer.py:96-100).  cons_weight_dev (device float) overrides cons_weight
 * when not NULL. */
SEDK_API int sedk_sed_loss_ex(const float* strong, const float* weak, const float* t_strong, const float* t_weak,
                     const float* labels, const float* labels_weak, int B, int C, int T, int n_strong, int n_weak,
                     int cons_row0, int cons_kind, float cons_weight, const float* cons_weight_dev, float* losses,
                     float* gstrong, float* gweak, void* stream);

/* Stand-alone entry points of the two convolution building blocks (channels-last tensors, 3x3 / pad 1 / stride 1), used by the
 * unit tests and micro-benchmarks; inside the network they are driven through sedk_crnn_forward / backward.
 *   sedk_conv3x3   : out[B,T,F,cout] = bias + conv(in[B,T,F,cin], wpack[9][cout][cin]);  stats (double[2*cout], may be NULL)
 *                    accumulates per-channel sum and sum of squares (BatchNorm statistics).
 *   sedk_conv_wgrad: gwpack[9][cout][cin] += sum_pixels gz[pix][cout] * x[pix + tap][cin]   (gwpack pre-zeroed by the caller) */
SEDK_API int sedk_conv3x3(const float* in, const float* wpack, const float* bias, float* out, double* stats, int B, int T,
                 int F, int cin, int cout, int precision, void* stream);
SEDK_API int sedk_conv_wgrad(const float* x, const float* gz, float* gwpack, int B, int T, int F, int cin, int cout,
                    int precision, void* stream);

/* plain GEMM building block (TF32 / 3xTF32 mma): C[M,N] = alpha*op(A)op(B) + beta*C + bias[n]
 * transA 0: A is [M,K] (lda), 1: A is [K,M];  transB 0: B is [K,N] (ldb), 1: B is [N,K]. */
SEDK_API int sedk_gemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* Bm,
              int ldb, float beta, float* C, int ldc, const float* bias, int precision, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEDK_H_ */

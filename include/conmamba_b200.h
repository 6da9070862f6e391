/*
 * conmamba_b200.h - C ABI of the B200-native ConMamba hot path.
 *
 * One shared library (libconmamba_b200.so, built by __graft_entry__.build()) exports the entry points
 * below.  Every entry point
 *   - takes a POD argument block of raw DEVICE pointers, element strides and sizes (no torch types),
 *   - enqueues sm_100a kernels on the CUDA stream passed as `stream` (a cudaStream_t; NULL = default),
 *   - never allocates, never synchronises the host, keeps no global state (thread-safe),
 *   - returns 0 on success, a positive cudaError_t from the launch, or a negative CM_ERR_* code when
 *     the arguments describe something this library does not implement.  There is no CPU fallback.
 *
 * The reference (mattmireles/Mamba-ASR) has no FFI of its own: its hot path calls two pybind11 extension
 * modules that are pip dependencies (requirement.txt:7-8).  Each entry point cites the reference call site
 * whose native callee it replaces; INTEGRATION.md shows the ctypes binding and the drop-in Python modules.
 *
 * Layout conventions
 *   cm_tensor3 describes a (batch, channel-or-state, time) tensor by three ELEMENT strides, so both the
 *   reference's time-contiguous (B, D, L) tensors (sl == 1) and this library's preferred channel-last
 *   (B, L, D) buffers (sd == 1) are accepted without copies.  Channel-last is the fast path: one warp
 *   lane per channel makes every global access of the scan a coalesced row segment.
 */
#ifndef CONMAMBA_B200_H_
#define CONMAMBA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CM_ABI_VERSION 17

/* element type of activations (u, delta, z, B, C, x, out, and their gradients) */
enum { CM_F32 = 0, CM_BF16 = 1, CM_F16 = 2 };

/* negative return codes */
enum {
  CM_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, bad enum */
  CM_ERR_UNSUPPORTED = -2   /* valid request outside the implemented envelope (e.g. dstate > 16) */
};

#define CM_MAX_DSTATE 16
#define CM_SCAN_CKPT_STEPS 8   /* forward saves the recurrent state every 8 processed steps */
#define CM_CONV_MAX_WIDTH 4

/* flags */
#define CM_FLAG_DELTA_SOFTPLUS 1u /* delta = softplus(delta + delta_bias)  (selective_scan_interface.py:109-112) */
#define CM_FLAG_SILU 1u           /* conv: apply SiLU                       (selective_scan_interface.py:182)    */

typedef struct {
  void* ptr;      /* device pointer, NULL = tensor absent */
  int64_t sb;     /* batch stride, elements */
  int64_t sd;     /* channel (or state) stride, elements */
  int64_t sl;     /* time stride, elements */
} cm_tensor3;

/* ------------------------------------------------------------------------------------------------------
 * Selective scan.  Replaces selective_scan_cuda.fwd / .bwd of mamba-ssm 1.1.3.post1
 * (reference call sites: modules/mamba/selective_scan_interface.py:42, :67, :218, :252), and additionally
 * fuses what modules/mamba/bimamba.py:223-253 does around two such calls: the reversed-time second
 * direction (no torch.flip), the shared z gate and the 0.5*(fwd + bwd) output add.
 *
 * Per direction r, channel d, state n, in fp32:
 *   Delta = softplus(delta + delta_bias[d])                          (if CM_FLAG_DELTA_SOFTPLUS)
 *   h     = exp(Delta * A[d,n]) * h + Delta * u * B[n]               time ascending, or descending if reverse
 *   y_r   = sum_n C[n] * h[n] + Dskip[d] * u
 * out_pre = sum_r y_r ;  out = out_scale * out_pre * silu(z)   (no gate if z.ptr == NULL)
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t reverse;          /* 0: time 0..L-1 ; 1: time L-1..0 */
  int32_t bc_const;         /* 0: B, C are (batch, dstate, time) shared by all channels;
                               1: B, C are per-channel fp32 constants (dim, dstate) regardless of `dtype`:
                                  stride sb = channel, sd = state, sl unused */
  cm_tensor3 u;             /* (batch, dim, time) */
  cm_tensor3 delta;         /* (batch, dim, time) */
  cm_tensor3 Bm, Cm;        /* see bc_const */
  const float* A;           /* (dim, dstate) fp32 */
  int64_t A_sd, A_sn;
  const float* Dskip;       /* (dim) fp32 or NULL */
  const float* delta_bias;  /* (dim) fp32 or NULL */
  float* ckpt;              /* NULL (inference) or fp32 checkpoints [batch][dim][cm_scan_num_ckpt()][dstate16] */
  int64_t ckpt_sb, ckpt_sd; /* strides (floats) of the batch / channel index; checkpoint j of a row starts at j*16 */
  float* last_state;        /* NULL or (batch, dim, dstate) fp32: state after the last processed step */
  int64_t ls_sb, ls_sd, ls_sn;
} cm_scan_dir;

typedef struct {
  int32_t batch, dim, seqlen, dstate;
  int32_t ndir;             /* 1 or 2 */
  int32_t dtype;            /* CM_F32 / CM_BF16 / CM_F16 */
  uint32_t flags;
  float out_scale;          /* 1.0, or 0.5 for bimamba if_devide_out (bimamba.py:250-253) */
  int32_t lanes_per_channel;/* 0 = choose; 1, 2 or 4 lanes cooperate on one channel's 16 states */
  int32_t reserved;
  cm_scan_dir dir[2];
  cm_tensor3 z;             /* gate, shared by both directions; ptr NULL = none */
  cm_tensor3 out;           /* gated output */
  cm_tensor3 out_pre;       /* optional pre-gate sum over directions (saved for backward when z is given) */
  void* workspace;          /* optional device scratch (16-byte aligned) of cm_scan_fwd_workspace_bytes() bytes: lets an */
  int64_t workspace_bytes;  /* inference launch on few long sequences run as parallel time windows; NULL / 0 = never   */
} cm_scan_fwd_args;

/* number of fp32 [16] checkpoints per (batch, channel, direction) row that forward writes and backward reads */
int cm_scan_num_ckpt(int32_t seqlen, int32_t ndir);
/* channels handled by one warp (= the slab width of the dB/dC partial sums) for a lanes_per_channel choice */
int cm_scan_slab_channels(int32_t lanes_per_channel);
/* the lanes_per_channel the library would pick for this problem when the caller passes 0 */
int cm_scan_pick_lanes(int32_t batch, int32_t dim, int32_t ndir);
/* the same for cm_scan_bwd (its register / shared-memory footprint favours 2 lanes per channel) */
int cm_scan_pick_lanes_bwd(int32_t batch, int32_t dim, int32_t ndir);

/* Scratch bytes with which cm_scan_fwd would split THIS launch into time windows (chunk-parallel scan: per-window
 * summaries, a serial combine over windows, then every window from its incoming state), or 0 when the whole-sequence
 * launch already fills the GPU or checkpoints are requested.  The library never allocates: the caller provides it. */
int64_t cm_scan_fwd_workspace_bytes(const cm_scan_fwd_args* args);
int cm_scan_fwd(const cm_scan_fwd_args* args, void* stream);

typedef struct {
  cm_scan_dir in;           /* same tensors as forward (ckpt required, last_state ignored) */
  cm_tensor3 du;            /* (batch, dim, time) grad of u */
  cm_tensor3 ddelta;        /* (batch, dim, time) grad of the raw (pre-softplus) delta */
  float* dBC_part;          /* bc_const == 0: fp32 [batch][n_slab][seqlen][32] partial sums over a slab's channels,
                               columns 0..15 = dB[n], 16..31 = dC[n];  reduce with cm_reduce_dbc().
                               bc_const == 1: fp32 [batch][dim][32] per-row sums over time */
  float* dA_part;           /* fp32 [batch][dim][16] per-row sums over time; reduce over batch with cm_reduce_rows() */
  float* dD_part;           /* fp32 [batch][dim] or NULL */
  float* dbias_part;        /* fp32 [batch][dim] or NULL */
} cm_scan_bwd_dir;

typedef struct {
  int32_t batch, dim, seqlen, dstate;
  int32_t ndir;
  int32_t dtype;
  uint32_t flags;
  float out_scale;
  int32_t lanes_per_channel;
  int32_t reserved;
  cm_scan_bwd_dir dir[2];
  cm_tensor3 z;             /* or NULL */
  cm_tensor3 out_pre;       /* required when z is given */
  cm_tensor3 dout;          /* grad of the gated output */
  cm_tensor3 dz;            /* grad of z (written when z is given) */
} cm_scan_bwd_args;

/* Channels per slab of dBC_part for THIS launch (a pure function of the argument block, pointers included - fill everything
 * but dBC_part first): 128 when the lane-per-channel TMA kernel applies (scan_bwd_lc.cu), else cm_scan_slab_channels(lanes).
 * n_slab = ceil(dim / slab). */
int cm_scan_bwd_slab_channels(const cm_scan_bwd_args* args);
int cm_scan_bwd(const cm_scan_bwd_args* args, void* stream);

/* dB[b,n,l] = sum_slab part[b][slab][l][n], dC likewise with column 16+n; written in `dtype` through strides. */
int cm_reduce_dbc(const float* part, int32_t batch, int32_t n_slab, int32_t seqlen, int32_t dstate,
                  int32_t dtype, cm_tensor3 dB, cm_tensor3 dC, void* stream);
/* out[c] = sum_r part[r*cols + c]   (fp32; deterministic order) - dA, dD, d(delta_bias), conv dweight/dbias */
int cm_reduce_rows(const float* part, int64_t rows, int64_t cols, float* out, void* stream);
/* the same for up to CM_REDUCE_MAX_JOBS independent (part, rows, cols, out) jobs in ONE launch */
#define CM_REDUCE_MAX_JOBS 8
typedef struct {
  const float* part;
  float* out;
  int64_t rows, cols;
} cm_reduce_job;
int cm_reduce_multi(const cm_reduce_job* jobs, int32_t njobs, void* stream);
/* Up to CM_REDUCE_BATCH_MAX jobs of any shapes in ONE launch (a 1-D grid; every CTA looks its job up): what a host that
 * queues the partial sums of a whole backward pass calls once at its end (kernels.reduce_many with deferral on; replaces the
 * per-operator reducer launches and the .sum(0) of split-K weight-gradient products).  `stride` = floats between two rows
 * of `part` (>= cols): a job may sum a column block of a wider partial buffer.  Rows are added in index order per column
 * within fixed groups: deterministic. */
#define CM_REDUCE_BATCH_MAX 64
typedef struct {
  const float* part;
  float* out;
  int64_t rows, cols, stride;
} cm_reduce_job2;
int cm_reduce_batch(const cm_reduce_job2* jobs, int32_t njobs, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Depthwise causal conv1d (+ SiLU).  Replaces causal_conv1d_cuda.causal_conv1d_fwd / _bwd of
 * causal-conv1d 1.1.3.post1 (reference call sites: selective_scan_interface.py:182, :244, :286;
 * bimamba.py:282-287), and fuses the two BiMamba-v2 directions: one read of x produces
 *   causal      out_f[l] = act(bias_f + sum_k w_f[k] * x[l-(W-1)+k])
 *   anticausal  out_b[l] = act(bias_b + sum_k w_b[k] * x[l+(W-1)-k])     (= flip -> conv1d_b -> flip)
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t anticausal;
  int32_t reserved;
  const float* weight;      /* (dim, width) fp32 contiguous */
  const float* bias;        /* (dim) fp32 or NULL */
  cm_tensor3 out;           /* forward: output.  backward: grad of the output (input) */
  float* dweight_part;      /* backward: fp32 [n_part][dim][width] partial sums; reduce with cm_reduce_rows() */
  float* dbias_part;        /* backward: fp32 [n_part][dim] or NULL */
} cm_conv_dir;

typedef struct {
  int32_t batch, dim, seqlen, width;
  int32_t ndir;
  int32_t dtype;
  uint32_t flags;
  int32_t reserved;
  cm_tensor3 x;             /* (batch, dim, time) input */
  cm_tensor3 dx;            /* backward only: grad of x, summed over directions */
  cm_conv_dir dir[2];
} cm_conv_args;

int cm_conv_fwd(const cm_conv_args* args, void* stream);
/* number of partial rows n_part that cm_conv_bwd writes per direction for this problem */
int cm_conv_num_part(int32_t batch, int32_t seqlen);
int cm_conv_bwd(const cm_conv_args* args, void* stream);

/* single-token update (causal_conv1d_update, reference call site bimamba.py:335-341):
 * rolls conv_state (batch, dim, width) left by one, appends x (batch, dim), returns act(bias + <state, w>) */
int cm_conv_update(const void* x, void* conv_state, const float* weight, const float* bias, void* out,
                   int32_t batch, int32_t dim, int32_t width, int32_t dtype, uint32_t flags, void* stream);

/* single-token selective-state update for incremental decoding.  Replaces `selective_state_update` of mamba-ssm
 * 1.1.3.post1 (reference call site modules/mamba/bimamba.py:354-356; the torch fallback the reference runs without it is
 * bimamba.py:345-352):  dt = softplus(dt + dt_bias); state = state*exp(dt*A) + (dt*B)*x  (in place);
 * out = (<state, C> + D*x) * silu(z). */
typedef struct {
  int32_t batch, dim, dstate;
  int32_t dtype;            /* CM_* of x, dt, z, Bm, Cm, out */
  int32_t state_dtype;      /* CM_* of state */
  uint32_t flags;           /* CM_FLAG_DELTA_SOFTPLUS */
  void* state;              /* (batch, dim, dstate) contiguous, updated in place */
  const void* x;            /* (batch, dim), unit channel stride, row stride x_sb elements */
  const void* dt;           /* (batch, dim) */
  const void* z;            /* (batch, dim) or NULL */
  const void* Bm;           /* (batch, dstate) */
  const void* Cm;           /* (batch, dstate) */
  void* out;                /* (batch, dim) */
  int64_t x_sb, dt_sb, z_sb, b_sb, c_sb, out_sb;
  const float* A;           /* (dim, dstate) fp32 contiguous */
  const float* Dskip;       /* (dim) fp32 or NULL */
  const float* dt_bias;     /* (dim) fp32 or NULL */
} cm_ssm_step_args;

int cm_ssm_step(const cm_ssm_step_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Fbank tail.  Replaces the torch op chain behind speechbrain.lobes.features.Fbank after the STFT
 * (reference call sites train_CTC.py:285, train_S2S.py:349; YAML hparams/CTC/conmamba_large.yaml:322-326):
 *   power = re^2 + im^2 ; mel = power @ fbank ; db = 10*log10(max(mel, 1e-10)) ;
 *   db = max(db, max_over_utterance(db) - top_db)
 * cm_fbank_logmel does power -> mel -> dB and the per-utterance running max in one pass over the STFT;
 * cm_fbank_floor applies the top_db floor.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, frames, nbins, nmels;
  const float* stft;        /* complex64 as float pairs; element (b, f, t) at b*s_b + f*s_f + t*s_t (complex units) */
  int64_t s_b, s_f, s_t;
  const float* fbank;       /* (nbins, nmels) fp32 row-major triangular filterbank */
  float* out;               /* (batch, frames, nmels) fp32 contiguous */
  float* utt_max;           /* (batch) fp32, must be pre-filled with -inf; receives max dB per utterance */
  float amin;               /* 1e-10 */
  float multiplier;         /* 10 */
  float db_offset;          /* multiplier*log10(max(amin, ref_value)) = 0 */
  float top_db;             /* 80 */
} cm_fbank_args;

int cm_fbank_logmel(const cm_fbank_args* args, void* stream);
int cm_fbank_floor(const cm_fbank_args* args, void* stream);

/* The whole front-end in one kernel: windowed DFT -> power -> mel -> dB + per-utterance max,
 * straight from the samples.  Replaces torch.stft (cuFFT) + cm_fbank_logmel for the transform sizes of the reference YAMLs
 * (n_fft 400 and 512: hparams/CTC/conmamba_large.yaml:103-105, hparams/S2S/conformer_small.yaml:148); the complex STFT is
 * never written to memory.  Framing follows torch.stft(center=True, pad_mode="constant"): frame t covers samples
 * [t*hop - n_fft/2, t*hop + n_fft/2), zeros outside the utterance; `window` is the analysis window zero-padded (centred) to
 * n_fft.  Follow with cm_fbank_floor on (out, utt_max). */
typedef struct {
  int32_t batch, n_samples, frames, nmels;
  int32_t n_fft, hop;
  const float* wav;         /* (batch, n_samples) fp32, row stride wav_sb elements */
  int64_t wav_sb;
  const float* window;      /* (n_fft) fp32 */
  const float* fbank;       /* (n_fft/2 + 1, nmels) fp32 row-major triangular filterbank */
  const int32_t* band;      /* (nmels, 2) int32: [first, one-past-last) non-zero bin of each filter */
  float* out;               /* (batch, frames, nmels) fp32 contiguous */
  float* utt_max;           /* (batch) fp32, must be pre-filled with -inf */
  float amin, multiplier, db_offset, top_db;
} cm_fbank_wav_args;

int cm_fbank_wav_supported(int32_t n_fft);   /* 1 for the transform sizes cm_fbank_wav_logmel implements */
int cm_fbank_wav_logmel(const cm_fbank_wav_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension (the six LayerNorms around each Mamba block,
 * reference modules/Conmamba.py:595-621, 638-649).  y = (x - mean) * rstd * gamma + beta, statistics in fp32.
 * dy / y share y_dtype; dx has x_dtype.  Backward writes cm_layernorm_num_part(rows) partial rows of dgamma / dbeta
 * (fp32, [n_part][cols]) to be summed with cm_reduce_multi.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t rows;
  int32_t cols;             /* <= 1024 */
  int32_t x_dtype, y_dtype;
  float eps;
  const void* x;  int64_t x_stride;      /* row strides in elements; unit column stride */
  void* y;        int64_t y_stride;      /* forward output */
  const float* gamma;                    /* (cols) fp32 or NULL */
  const float* beta;                     /* (cols) fp32 or NULL */
  float* mean;                           /* (rows) fp32: written by forward, read by backward */
  float* rstd;                           /* (rows) fp32 */
  const void* dy; int64_t dy_stride;     /* backward: grad of y */
  void* dx;       int64_t dx_stride;     /* backward: grad of x */
  float* dgamma_part;
  float* dbeta_part;
  int32_t act;                           /* 0, or CM_LN_OUT_GELU: y = gelu(LayerNorm(x)) (exact erf GELU; the LayerNorm -> GELU
                                            pair after the depthwise conv of the convolution module, modules/Conmamba.py:
                                            292-301).  Backward then takes dy for the activated output and needs beta.
                                            Only for even cols / strides and 8-byte aligned pointers (else UNSUPPORTED). */
  int32_t n_part;                        /* v17 (was reserved): partial rows allocated for dgamma_part / dbeta_part =
                                            cm_layernorm_num_part2(rows, cols); 0 = cm_layernorm_num_part(rows) */
} cm_layernorm_args;
#define CM_LN_OUT_GELU 1

int cm_layernorm_num_part(int64_t rows);
/* v17: partial rows of cm_layernorm_bwd for this row width; pass the value as args.n_part.  For cols % 4 == 0 the backward
 * then runs the 16-byte-access kernels of cm_add_ln_bwd (rows staged through shared memory by bulk copies for cols <= 256). */
int cm_layernorm_num_part2(int64_t rows, int32_t cols);
int cm_layernorm_fwd(const cm_layernorm_args* args, void* stream);
int cm_layernorm_bwd(const cm_layernorm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Residual add + dropout + LayerNorm in one pass (reference modules/Conmamba.py:638-649:
 * every sub-block of a ConMamba layer ends in  s = a + alpha * dropout(b)  and the next starts with  y = LayerNorm(s)).
 *   forward : s = a + alpha * keep/(1-p) * b ; y = (s - mean) * rstd * gamma + beta ; writes s, y, mean, rstd, mask
 *   backward: t = LayerNorm'(dy; s) + ds ; da = t ; db = alpha * keep/(1-p) * t ; dgamma / dbeta partial rows
 *             (cm_add_ln_num_part(rows, cols) rows each, summed by cm_reduce_multi)
 * a, s, ds, da share a_dtype; b, db share b_dtype; y, dy share y_dtype.  Supported (a, b, y): (f32, bf16, bf16),
 * (f32, bf16, f32), (f32, f32, f32), (bf16, bf16, bf16), (bf16, bf16, f32); cols even and <= 1024, even strides.
 * The keep mask is a counter-based hash of (*seed, call_id, row, column); p_drop > 0 with b non-NULL applies dropout.  The
 * mask is either stored as one byte per element (mask non-NULL) or regenerated by backward from the key forward wrote.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t rows;
  int32_t cols;
  int32_t a_dtype, b_dtype, y_dtype;
  float eps, alpha, p_drop;       /* p_drop in [0, 1); the mask pointer decides whether dropout is applied */
  uint32_t call_id;               /* distinguishes call sites that share a seed */
  int32_t reserved;
  const int64_t* seed;            /* device scalar (read at kernel time: graph replays see host-side advances) or NULL */
  const void* a;  int64_t a_stride;
  const void* b;  int64_t b_stride;   /* NULL: s = a */
  void* s;        int64_t s_stride;   /* forward: output (may be NULL if the caller never needs s); backward: input */
  void* y;        int64_t y_stride;
  uint8_t* mask;                      /* (rows, cols) contiguous; NULL = no dropout */
  const float* gamma;
  const float* beta;
  float* mean;
  float* rstd;
  const void* dy; int64_t dy_stride;
  const void* ds; int64_t ds_stride;  /* gradient reaching s from its other consumers, or NULL */
  void* da;       int64_t da_stride;
  void* db;       int64_t db_stride;  /* NULL when b was NULL */
  float* dgamma_part;
  float* dbeta_part;
  uint32_t* key;                      /* v12: forward writes the 32-bit mask key here (if non-NULL); with mask NULL and
                                         p_drop > 0 backward regenerates the keep bits from it instead of reading a mask */
  float* dbsum_part;                  /* v12, backward, optional: cm_add_ln_num_part(rows, cols) partial rows of the column sums
                                         of db (the bias gradient of the Linear that produced b); cols % 4 == 0 layouts only */
} cm_add_ln_args;

/* 1 when cm_add_ln_bwd can also produce dbsum_part for this geometry (the quad-vectorised kernels: cols % 4 == 0, row strides
 * % 4 == 0, 16-byte aligned fp32 / 8-byte aligned 16-bit rows) */
int cm_add_ln_dbsum_supported(int32_t cols, int64_t min_stride);
/* v12: partial rows of cm_add_ln_bwd's dgamma_part / dbeta_part / dbsum_part (was cm_layernorm_num_part) */
int cm_add_ln_num_part(int64_t rows, int32_t cols);

int cm_add_ln_fwd(const cm_add_ln_args* args, void* stream);
int cm_add_ln_bwd(const cm_add_ln_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * GELU (exact, erf) + dropout in one pass over a flat tensor (the activation and the
 * Dropout between the two Linears of speechbrain's PositionalwiseFeedForward, reference modules/Conmamba.py:595-621).
 *   forward : y = keep/(1-p) * gelu(x), mask (1 byte per element, NULL = no dropout) written for backward
 *   backward: dx = keep/(1-p) * gelu'(x) * dy
 * n must be a multiple of 8, pointers 16-byte aligned (mask 8).  Mask = hash of (*seed, call_id, element index).
 * ---------------------------------------------------------------------------------------------------- */
int cm_gelu_dropout_fwd(const void* x, void* y, uint8_t* mask, int64_t n, int32_t dtype, float p_drop,
                        const int64_t* seed, uint32_t call_id, void* stream);
int cm_gelu_dropout_bwd(const void* x, const void* dy, const uint8_t* mask, void* dx, int64_t n, int32_t dtype,
                        float p_drop, void* stream);

/* Struct form (ABI v12).  The dropout mask need not be stored: forward writes its 32-bit mask key to *key and backward
 * regenerates the keep bits from it (mask NULL); with cols > 0 (cols % 8 == 0, 256 % (cols / 8) == 0, n % cols == 0:
 * cm_act_colsum_supported) backward also writes cm_act_num_part(n) partial rows of the column sums of dx over the
 * (n / cols, cols) matrix - the bias gradient of the Linear that produced x - for cm_reduce_multi. */
typedef struct {
  const void* x;
  void* y;                  /* forward output */
  const void* dy;           /* backward */
  void* dx;
  uint8_t* mask;            /* optional byte mask (forward writes, backward reads); NULL: regenerate from *key */
  const int64_t* seed;      /* forward: device seed (NULL: a fixed default) */
  uint32_t* key;            /* forward writes the mask key; backward reads it when mask is NULL */
  uint32_t call_id;
  int32_t dtype;
  float p_drop;
  int32_t cols;             /* backward: > 0 = also column sums of dx */
  int64_t n;
  float* colsum_part;       /* [cm_act_num_part(n)][cols] fp32 */
  uint8_t* keep_bits;       /* optional, n / 8 bytes: bit i of byte v = element 8 v + i is kept.  Forward writes it (next to
                               whatever else it stores), backward reads it in preference to mask / key: one byte per eight
                               elements instead of four hashes */
} cm_act_args;

int cm_act_colsum_supported(int64_t n, int32_t cols);
int cm_act_num_part(int64_t n);
int cm_gelu_dropout_fwd_v2(const cm_act_args* args, void* stream);
int cm_gelu_dropout_bwd_v2(const cm_act_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Tall-skinny weight-gradient GEMM  C[M, N] = sum_r A[r, m] * B[r, n]  (A^T B): the gradients of the Mamba block's skinny
 * projections, d dt_proj.weight = ddelta^T x_dbl[:, :R] and d x_proj.weight = dx_dbl^T conv1d_out (reference
 * modules/mamba/selective_scan_interface.py:277-283), whose reduction dimension is batch * L.
 * A (rows, M) and B (rows, N) are 16-bit row-major with leading dimensions lda, ldb (elements, multiples of 8; 16-byte
 * aligned bases); M a multiple of 8, N a multiple of 8 and <= 64.  Writes cm_tsmm_num_part(rows, M) partial blocks
 * part[chunk][M][N] (fp32); sum them with cm_reduce_multi (rows = chunks, cols = M * N).
 * ---------------------------------------------------------------------------------------------------- */
int cm_tsmm_num_part(int64_t rows, int32_t M);
int cm_tsmm(const void* A, int64_t lda, const void* B, int64_t ldb, float* part, int64_t rows, int32_t M, int32_t N,
            int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Depthwise conv1d over time (the kernel_size = 31 convolution of the ConMamba
 * convolution module, reference modules/Conmamba.py:281-290, nn.Conv1d(C, C, K, padding, groups=C)).
 *   y[b,l,c] = bias[c] + sum_k weight[c,k] * x[b, l - pad_left + k, c]      (zero outside [0, L))
 * pad_left = (K-1)/2 for "same" padding, K-1 for the causal variant.  K in {3, 7, 15, 31}.
 * Backward-data is the same entry point with flip = 1, pad_left' = K-1-pad_left, bias = NULL, x = grad of y.
 * cm_dwconv_bwd_weight writes cm_dwconv_num_part() partial rows: dweight_part [n_part][dim][K], dbias_part
 * [n_part][dim] (fp32); sum them with cm_reduce_multi.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, dim, seqlen, ksize;
  int32_t pad_left;
  int32_t dtype;
  int32_t flip;             /* 1: use weight[c][K-1-k] (backward-data) */
  int32_t reserved;
  cm_tensor3 x;             /* (batch, dim, time) input (forward) / grad of y (backward-data) */
  cm_tensor3 y;             /* output (forward) / grad of x (backward-data) */
  const float* weight;      /* (dim, K) fp32 contiguous */
  const float* bias;        /* (dim) fp32 or NULL */
  cm_tensor3 dy;            /* backward-weight: grad of y */
  float* dweight_part;      /* backward-weight outputs */
  float* dbias_part;        /* or NULL */
} cm_dwconv_args;

int cm_dwconv_num_part(int32_t batch, int32_t seqlen, int32_t ksize);
int cm_dwconv_fwd(const cm_dwconv_args* args, void* stream);
int cm_dwconv_bwd_weight(const cm_dwconv_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Column sums of a (rows, cols) matrix = bias gradient of a Linear / pointwise conv (SURVEY.md section 8(f) rank 2;
 * the six biased Linear layers around each Mamba block, reference modules/Conmamba.py:595-621).
 * Writes cm_colsum_num_part(rows) partial rows part[n_part][cols] (fp32); sum them with cm_reduce_multi.
 * cols and row_stride must be even (pair accesses): CM_ERR_UNSUPPORTED otherwise.
 * ---------------------------------------------------------------------------------------------------- */
int cm_colsum_num_part(int64_t rows);
int cm_colsum(const void* x, int64_t rows, int32_t cols, int64_t row_stride, int32_t dtype, float* part, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * LayerNorm + activation in one pass: the two conv blocks of the reference's
 * ConvolutionFrontEnd (hparams/CTC/conmamba_large.yaml:187-194: conv 3x3 stride 2 -> LayerNorm([F', C]) -> LeakyReLU; rows
 * of F'*C = 2560 and 640 elements at the BASELINE shapes, the conv bias folded in as pre_bias), and the LayerNorm ->
 * activation after the depthwise conv of the ConMamba convolution module (modules/Conmamba.py:297-303).
 *   x' = x + pre_bias[col % pre_bias_n]                                   (pre_bias NULL: x' = x)
 *   forward : y = act((x' - mean) * rstd * gamma + beta)                  statistics of x' in fp32
 *             act = LeakyReLU(slope) (slope = 1: plain LayerNorm) or exact (erf) GELU
 *   backward: dx (= grad of x'), and cm_ln_act_num_part(rows, cols) partial rows of dgamma / dbeta ([n_part][cols] fp32,
 *             summed with cm_reduce_multi); the pre-activation is recomputed from x, mean, rstd.
 * x, y, dy, dx are dense (rows, cols) matrices of one dtype; cols a multiple of 4, <= 2560; 16-byte aligned bases;
 * pre_bias_n a multiple of 4 that divides cols.
 * ---------------------------------------------------------------------------------------------------- */
enum { CM_LN_ACT_LEAKY_RELU = 0, CM_LN_ACT_GELU = 1 };
typedef struct {
  int64_t rows;
  int32_t cols;
  int32_t dtype;
  float eps;
  float slope;              /* LeakyReLU negative slope (torch default 0.01) */
  int32_t act;              /* CM_LN_ACT_* */
  int32_t pre_bias_n;       /* period of pre_bias in columns */
  const float* pre_bias;    /* (pre_bias_n) fp32 or NULL */
  const void* x;
  void* y;                  /* forward output */
  const float* gamma;       /* (cols) fp32 */
  const float* beta;        /* (cols) fp32 */
  float* mean;              /* (rows) fp32: written by forward, read by backward */
  float* rstd;              /* (rows) fp32 */
  const void* dy;           /* backward: grad of y */
  void* dx;                 /* backward: grad of x */
  float* dgamma_part;
  float* dbeta_part;
} cm_ln_act_args;

int cm_ln_act_num_part(int64_t rows, int32_t cols);
int cm_ln_act_fwd(const cm_ln_act_args* args, void* stream);
int cm_ln_act_bwd(const cm_ln_act_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Optimizer step on flat fp32 buffers (the optimizer of the training recipes, reference
 * train_CTC.py:716-717 with hparams/CTC/conmamba_large.yaml:91, 248-252: torch.optim.AdamW + max_grad_norm clipping).
 *   cm_sumsq_partial: part[i] = sum of g^2 over CTA i's share (cm_optim_num_part(n) fp32 values, fixed order)
 *   cm_adamw_step   : norm = sqrt(sum part) * grad_scale ; c = (max_grad_norm > 0 ? min(1, max_grad_norm/(norm+1e-6)) : 1)
 *                     * grad_scale ; g' = c*g ; m = b1*m + (1-b1)*g' ; v = b2*v + (1-b2)*g'^2 ;
 *                     p = p*(1 - lr*weight_decay) - (lr/bias_corr1) * m / (sqrt(v)/sqrt(bias_corr2) + eps)
 *                     (bias_corr = 1 - beta^t).  sumsq_part NULL: no clipping, c = grad_scale.  p_bf16 non-NULL: also writes
 *                     the bf16 copy of p.  n must be a multiple of 4, all pointers 16-byte aligned.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  float* p;                 /* (n) parameters, updated in place */
  const float* g;           /* (n) gradients (summed over ranks when grad_scale = 1/world_size) */
  float* m;                 /* (n) first moment */
  float* v;                 /* (n) second moment */
  void* p_bf16;             /* (n) bf16 copy of p, or NULL */
  const float* sumsq_part;  /* cm_sumsq_partial output, or NULL */
  float* norm_out;          /* optional device scalar: the gradient norm that was clipped against */
  int64_t n;
  int32_t n_part;
  float lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2, max_grad_norm, grad_scale;
} cm_adamw_args;

int cm_optim_num_part(int64_t n);
int cm_sumsq_partial(const float* g, int64_t n, float* part, void* stream);
int cm_adamw_step(const cm_adamw_args* args, void* stream);

/* library identification: returns CM_ABI_VERSION; writes the compiled-for arch (e.g. 100) to *sm_arch if non-NULL */
int cm_version(int32_t* sm_arch);
/* ------------------------------------------------------------------------------------------------------
 * CTC loss and its gradient in one launch (the `ctc_loss` stand-in of the layer shell;
 * reference call site train_CTC.py:297-302 -> speechbrain.nnet.losses.ctc_loss -> torch.nn.functional.ctc_loss).
 * Per utterance b: nll[b] = -log p(targets_b | log_probs_b) (+inf if no alignment exists) and, if `grad` is given,
 * grad[b, t, c] = d nll[b] / d log_probs[b, t, c] (0 for t >= input_lengths[b]); reduction and zero_infinity are the
 * caller's.  One CTA per utterance runs the alpha and the beta recursion concurrently; up to 255 labels per utterance.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, max_time, classes, max_target;
  int32_t blank, ws_states;         /* ws_states = 2 * max_target + 1 */
  const float* log_probs;           /* (batch, time, class) fp32, unit class stride */
  int64_t lp_sb, lp_st;
  const int64_t* targets;           /* (batch, max_target) int64, row stride tg_sb */
  int64_t tg_sb;
  const int64_t* input_lengths;     /* (batch) or NULL = max_time */
  const int64_t* target_lengths;    /* (batch) or NULL = max_target */
  float* nll;                       /* (batch) fp32 */
  float* grad;                      /* (batch, time, class) fp32 or NULL */
  int64_t g_sb, g_st;
  float* workspace;                 /* cm_ctc_workspace_floats() floats: alpha and beta */
} cm_ctc_args;

int64_t cm_ctc_workspace_floats(int32_t batch, int32_t max_time, int32_t max_target);
int cm_ctc_loss(const cm_ctc_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * First block of the CNN front-end in one kernel each way (the layer shell; reference
 * speechbrain ConvolutionFrontEnd block 1 as configured by hparams/CTC/conmamba_large.yaml:187-199, reached from
 * train_CTC.py:288):
 *   y = LeakyReLU(LayerNorm_[F', C](Conv2d(1 -> C, 3 x 3, stride 2, zero padding 1)(in) + bias)),  T' = (frames-1)/2+1,
 *   F' = (feats-1)/2+1.  in: dense (batch, frames, feats); y, dy: dense (batch, T', F', C); weight (C, 3, 3) fp32 (= torch's
 *   (C, 1, 3, 3)); gamma / beta (F', C) fp32; mean / rstd (batch * T') fp32 written by forward, read by backward.
 *   backward recomputes the conv output from `in` and writes cm_stem_num_part(batch, frames) partial rows of dgamma / dbeta
 *   ([n_part][F' * C]), dweight ([n_part][C * 9]) and dbias ([n_part][C]) for cm_reduce_multi; no input gradient.
 * Envelope (cm_stem_supported): C % 4 == 0, C <= 128, 512 % C == 0, feats <= 168, F' * C <= 2560; 16-byte aligned y, dy,
 * gamma, beta, bias and partial buffers.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, frames, feats, channels;
  int32_t in_dtype, out_dtype;
  float eps, slope;
  const void* in;
  const float* weight;
  const float* bias;         /* (C) fp32 or NULL */
  const float* gamma;
  const float* beta;
  void* y;                   /* forward output */
  float* mean;
  float* rstd;
  const void* dy;            /* backward: grad of y (out_dtype) */
  float* dgamma_part;
  float* dbeta_part;
  float* dweight_part;
  float* dbias_part;
} cm_stem_args;

int cm_stem_supported(int32_t feats, int32_t channels);
int cm_stem_num_part(int32_t batch, int32_t frames);
int cm_stem_fwd(const cm_stem_args* args, void* stream);
int cm_stem_bwd(const cm_stem_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * GLU over the last dimension of a channel-last tensor (the gate of the ConMamba convolution module, reference
 * modules/Conmamba.py:268-279: Conv1d(C, 2C, 1) -> nn.GLU):
 *   forward   y[r, c]  = h[r, c] * sigmoid(h[r, dim + c])                                   h: rows x 2*dim, y: rows x dim
 *   backward  dh[r, c] = dy[r, c] * s,  dh[r, dim + c] = dy[r, c] * h[r, c] * s * (1 - s),  s = sigmoid(h[r, dim + c])
 * Strides are in elements per row; dim and the strides multiples of 8, pointers 16-byte aligned (else CM_ERR_UNSUPPORTED).
 * ---------------------------------------------------------------------------------------------------- */
int cm_glu_fwd(const void* h, void* y, int64_t rows, int32_t dim, int64_t h_stride, int64_t y_stride, int32_t dtype, void* stream);
int cm_glu_bwd(const void* h, const void* dy, void* dh, int64_t rows, int32_t dim, int64_t h_stride, int64_t dy_stride,
               int64_t dh_stride, int32_t dtype, void* stream);

/* sizeof() of the argument structs, for binding self-checks: 0 cm_tensor3, 1 cm_scan_dir, 2 cm_scan_fwd_args,
 * 3 cm_scan_bwd_dir, 4 cm_scan_bwd_args, 5 cm_conv_dir, 6 cm_conv_args, 7 cm_fbank_args, 8 cm_reduce_job,
 * 9 cm_layernorm_args, 10 cm_dwconv_args, 11 cm_ssm_step_args, 12 cm_add_ln_args, 13 cm_ln_act_args, 14 cm_adamw_args,
 * 15 cm_fbank_wav_args, 16 cm_ctc_args, 17 cm_stem_args, 18 cm_act_args, 19 cm_reduce_job2 */
int cm_abi_sizeof(int32_t which);

#ifdef __cplusplus
}
#endif
#endif /* CONMAMBA_B200_H_ */

/*
 * gd_b200.h — C-ABI of the B200 (sm_100a) gesture-DDPM sampling kernels.
 *
 * The reference (wubowen416/Speech-driven-Gesture-Generation-...) is pure Python and has no
 * FFI of its own; the boundary a maintainer binds is therefore the set of torch ops its hot
 * loop issues.  Each entry point below names the reference call site it replaces
 * (paths relative to the reference repo root).  All pointers are raw CUDA device pointers
 * unless stated otherwise, `stream` is a `cudaStream_t` passed as `void*`, every function
 * returns 0 on success or a negative gd_status and records a message readable through
 * gd_last_error().  Nothing here allocates device memory or synchronises the device; the
 * caller owns every buffer.  Not thread-safe per stream.
 *
 * Conventions
 *   - activations are row-major "token rows": row = clip * tokens_per_clip + token
 *   - GEMM inputs are bf16, accumulation is fp32 (tcgen05.mma kind::f16 into TMEM)
 *   - the denoise-step index t is read by the kernels from a device int (`step_ptr`) so that a
 *     captured CUDA graph of one step can be replayed for every t (1000 replays = one chain)
 */
#ifndef GD_B200_H
#define GD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GD_ABI_VERSION 4

enum gd_status {
    GD_OK = 0,
    GD_ERR_INVALID = -1, /* bad argument (shape / alignment / null pointer) */
    GD_ERR_CUDA = -2,    /* a CUDA runtime / driver call failed            */
    GD_ERR_ARCH = -3     /* device is not sm_100                           */
};

enum gd_act { GD_ACT_NONE = 0, GD_ACT_RELU2 = 1, GD_ACT_SILU = 2 };

int gd_abi_version(void);
const char* gd_last_error(void);
/* Number of kernels launched by this library since load (bench.py's `gpu_launches`). */
uint64_t gd_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * gd_linear_bf16:  out = act(A · Wᵀ + bias + rowbias[(row % period) + offset]) + residual
 *
 * Replaces every nn.Linear on the path: Q/K/V `models/modules/transformer.py:51,57`, attention
 * output `transformer.py:73,118`, FeedForward `transformer.py:146-154` (act = GD_ACT_RELU2 is
 * SquaredReLU `transformer.py:8-16`), emb_x/emb_mem `models/nn.py:189-190,393-394` with the
 * PositionalEncoding add `transformer.py:176-180` folded in as `rowbias`, blend_layer
 * `models/model.py:79,105`, the timestep MLP `nn.py:41-46` (GD_ACT_SILU) and the residual adds
 * `nn.py:99,103,110,118,123,162,167,172`.
 *
 *   A        bf16 [M, K], row stride lda (elements, multiple of 8)
 *   W        bf16 [N, K], row stride ldw  — nn.Linear.weight layout, K-major
 *   K % 64 == 0, N % 64 == 0 (pad d_pose 123/126 -> 128 on the host)
 *   bias     fp32 [N] or NULL
 *   rowbias  fp32 [*, N] or NULL; row r adds rowbias[((r % rowbias_period) + rowbias_offset) * N + col]
 *   residual fp32 [M, N] row stride ldr, or NULL (may alias out_f32)
 *   out_f32  fp32 [M, N] row stride ldo_f32, or NULL
 *   out_bf16 bf16 [M, N] row stride ldo_bf16, or NULL   (at least one output required)
 * ------------------------------------------------------------------------------------------ */
typedef struct gd_linear_desc {
    const void* A;
    const void* W;
    int32_t M, N, K;
    int32_t lda, ldw;
    const float* bias;
    const float* rowbias;
    int32_t rowbias_period;
    int32_t rowbias_offset;
    const float* residual;
    int32_t ldr;
    int32_t act;
    float* out_f32;
    int32_t ldo_f32;
    void* out_bf16;
    int32_t ldo_bf16;
    int32_t max_ctas;   /* ABI 4: > 0 caps the persistent grid (SMs this launch may occupy) so that kernels of two concurrent
                           streams share the GPU by SM count instead of queueing behind each other; 0 = every SM */
} gd_linear_desc;

int gd_linear_bf16(const gd_linear_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * DDPM ancestral update  (models/modules/gaussian_diffusion.py:287-292 _predict_xstart_from_eps,
 * :207-232 q_posterior_mean_variance, :300-329 p_sample, optional in-paint blend
 * models/generator.py:255-281), for the step t = *step_ptr:
 *
 *   x0   = A[t]*x - B[t]*eps            A = sqrt_recip_alphas_cumprod, B = sqrt_recipm1_alphas_cumprod
 *   x0   = (1-f[j])*m[n,j]*seed[n,j,c] + f[j]*m[n,j]*x0 + (1-m[n,j])*x0      (only if inpaint_seed)
 *   mean = C1[t]*x0 + C2[t]*x           posterior_mean_coef1/2
 *   x'   = mean + (t != 0) * sigma[t] * noise[t]     sigma = exp(0.5*posterior_log_variance_clipped)
 *
 * Every product/sum is rounded separately in fp32 (no FMA contraction), which is the order the
 * reference's elementwise torch ops use.  x / noise / eps_out are (N, C, T) with T contiguous
 * (the reference's boundary layout, models/model.py:12-15).
 * ------------------------------------------------------------------------------------------ */
typedef struct gd_ddpm_desc {
    float* x;                 /* (N, C, T) fp32, updated in place                          */
    const float* noise_tape;  /* (n_steps, N, C, T) fp32 indexed by t; may be NULL (=0)     */
    const float* coef_A;      /* [n_steps] each, fp32                                       */
    const float* coef_B;
    const float* coef_C1;
    const float* coef_C2;
    const float* sigma;
    const int32_t* step_ptr;  /* device int: current t                                      */
    int32_t n_clips, C, T;
    float* eps_out;           /* optional (N, C, T): predicted eps (parity harness)         */
    float* x0_out;            /* optional (N, C, T): pred_x_start after in-painting         */
    void* xa_bf16;            /* optional bf16 [N*T, ld_xa]: x' transposed to token rows    */
    int32_t ld_xa;            /*   (next step's emb_x GEMM operand; cols >= C written as 0) */
    const float* inpaint_seed;   /* optional (N, T, C) fp32                                 */
    const float* inpaint_mask;   /* (N, T) fp32, required with inpaint_seed                 */
    const float* inpaint_factor; /* [T] fp32 trans_factor ramp, required with inpaint_seed  */
    float clip_x0;            /* > 0: clamp x0 to [-clip, clip]; 0 = off (reference has none) */
    const float* xa_add;      /* optional (N, C, T) fp32 added to x' before the bf16 cast into xa_bf16: the
                                 loop-invariant `proj([inpaint_pose*mask | mask])` that Speech2GestureModelInpaint
                                 adds to its input (models/model.py:155-166)                                        */
    float* mean_out;          /* ABI 4, optional (N, C, T): posterior mean (`mean` of p_mean_variance, :276-285)      */
    float* raw_x0_out;        /* ABI 4, optional (N, C, T): x0 BEFORE the in-paint blend (`raw_x_start`, :254)        */
    const int32_t* aux_step_ptr; /* ABI 4, optional device int: when non-NULL and >= 0, the four optional outputs
                                 (eps_out, x0_out, mean_out, raw_x0_out) are written ONLY in the step whose t equals it
                                 (a whole-chain graph needs them for its last step only); NULL or < 0: every step       */
} gd_ddpm_desc;

/* Standalone update from an eps tensor laid out (N, C, T). */
int gd_ddpm_update(const gd_ddpm_desc* u, const float* eps, void* stream);

/* Final projection (out_layers Linear, models/nn.py:211-214,423-426) with the update above fused
 * into the GEMM epilogue: eps never round-trips through HBM.  d->N must be 128 >= u->C, the
 * outputs/residual/rowbias of `d` are ignored, d->M == n_clips*T. */
int gd_linear_ddpm(const gd_linear_desc* d, const gd_ddpm_desc* u, void* stream);

/* ------------------------------------------------------------------------------------------
 * gd_layernorm: nn.LayerNorm([D]) eps 1e-5 (models/nn.py:70-84,141-147,212,424),
 * fp32 rows in, bf16 rows out (the following GEMM's A operand).  D in {256, 512}.
 * ------------------------------------------------------------------------------------------ */
int gd_layernorm(const float* x, int32_t ldx, const float* gamma, const float* beta, void* out_bf16, int32_t ldo,
                 int32_t M, int32_t D, float eps, void* stream);
/* ABI 4: the same with a second parameter set for rows >= split_row (gamma2 / beta2 may be NULL = one set): the tedexp joint
 * attention updates pose rows and memory rows in one GEMM, and they feed norm_ff / norm_ff_mem (models/nn.py:116-123). */
int gd_layernorm_split(const float* x, int32_t ldx, const float* gamma, const float* beta, const float* gamma2,
                       const float* beta2, int32_t split_row, void* out_bf16, int32_t ldo, int32_t M, int32_t D, float eps,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * gd_linear_resid_ln: residual GEMM with the LayerNorm that follows it fused into the epilogue:
 *
 *     H   += A · Wᵀ + bias              (d->out_f32 == d->residual: the fp32 residual stream, in place)
 *     xn   = LayerNorm(H) * gamma + beta  -> bf16
 *
 * Replaces `x = x + self.dropout(attn(...))` / `x = x + self.dropout(ff(...))` together with the next
 * `self.norm_*(x)` of models/nn.py:97-124,158-173 (out-projection `transformer.py:118` or FeedForward layer2
 * `transformer.py:153`, then nn.LayerNorm([d]) eps 1e-5).  d->N must be the model width (256 or 512) so that a CTA
 * (pair) owns complete rows.  Rows >= split_row are normalised with (gamma2, beta2): the tedexp joint attention
 * updates pose rows and memory rows in one GEMM but they feed norm_ff / norm_ff_mem (nn.py:116-123).
 * ------------------------------------------------------------------------------------------ */
typedef struct gd_ln_desc {
    const float* gamma;   /* [N] */
    const float* beta;    /* [N] */
    const float* gamma2;  /* optional second parameter set for rows >= split_row */
    const float* beta2;
    int32_t split_row;
    void* out_bf16;       /* bf16 [M, N], row stride ldo */
    int32_t ldo;
    float eps;
} gd_ln_desc;

int gd_linear_resid_ln(const gd_linear_desc* d, const gd_ln_desc* ln, void* stream);

/* ------------------------------------------------------------------------------------------
 * gd_linear_ln_bf16 (ABI 4): LayerNorm as the PROLOGUE of the GEMM that consumes it,
 *
 *     out_bf16 = act( LayerNorm(H; gamma, beta, eps) · Wᵀ + bias )
 *
 * Replaces `self.norm_*(x)` + the fused Q|K|V projection / the first FeedForward layer (models/nn.py:97-124,158-173 ->
 * transformer.py:51,57,146-152).  d->A is the fp32 residual stream H [M, K] (row stride d->lda in fp32 elements), K must be
 * the model width (256 or 512), N a multiple of 128; bias required; bf16 output only.  The normalised rows are produced
 * inside the kernel, bit-identical to gd_layernorm, and never reach HBM.
 * ------------------------------------------------------------------------------------------ */
int gd_linear_ln_bf16(const gd_linear_desc* d, const float* gamma, const float* beta, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * gd_dconv_attention: MultiDConvHeadAttention core (models/modules/transformer.py:88-126):
 * per (clip, head): Q,K,V = depth-wise conv3 over tokens (SpatialDepthWiseConv :19-44, zero "same"
 * padding, taps shared by all heads) of the already-projected rows, S = softmax_keys(QKᵀ·scale),
 * O = S·V.  The token sequence of a clip is the concatenation of up to two row segments
 * (tedexp joint attention over [x ; memory], models/nn.py:105-113); the conv runs across the seam.
 *   q/k/v[s]   bf16, row (clip*rows[s] + i) at ptr + row*ld, head h at columns [h*d_k, (h+1)*d_k)
 *   out[s]     bf16, same row structure as q (row clip*q_rows[s] + i); out[1] may be NULL (halo segment, see below)
 *   conv_*     fp32 [d_k, 3] taps and [d_k] bias per projection
 * d_k in {32, 64}; total keys per clip <= 160.
 * ------------------------------------------------------------------------------------------ */
typedef struct gd_attn_desc {
    const void* q[2];
    int32_t q_rows[2];
    int32_t q_ld[2];
    const void* k[2];
    const void* v[2];
    int32_t kv_rows[2];
    int32_t kv_ld[2];
    void* out[2];
    int32_t out_ld[2];
    const float* conv_wq;
    const float* conv_bq;
    const float* conv_wk;
    const float* conv_bk;
    const float* conv_wv;
    const float* conv_bv;
    int32_t n_clips, heads, d_k;
    float scale;
    /* ABI 4.  q_clip_stride[s] > 0: consecutive clips of query segment s are that many rows apart (default q_rows[s]).
     * A query segment with out[s] == NULL is a conv HALO: its rows take part in the depth-wise conv of the
     * concatenated query sequence but get no attention output.  The last tedexp layer needs it: only the pose rows are
     * read after the joint attention (models/nn.py:445-447), but the conv3 of the last pose frame reaches memory row 0
     * (models/nn.py:105-113, transformer.py:19-44), so q[1] = memory row 0 of every clip, q_rows[1] = 1,
     * q_clip_stride[1] = memory rows per clip, out[1] = NULL. */
    int32_t q_clip_stride[2];
    int32_t max_ctas_sms; /* ABI 4: > 0 sizes the persistent grid for that many SMs (see gd_linear_desc.max_ctas); 0 = all */
} gd_attn_desc;

int gd_dconv_attention(const gd_attn_desc* d, void* stream);
/* Same, but q/k/v rows are fp32 (the fp32-activation parity path); ld in fp32 elements. */
int gd_dconv_attention_f32in(const gd_attn_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Step-dependent row scatter.  The only t-dependent conditioning is the timestep token
 * (models/model.py:50-52,90-92 -> memory row 0).  Its image under the loop-invariant linear maps is
 * tabulated once per chain as table[t, :]; each step copies row t into `row_index` of every
 * clip's block:   dst[(clip*rows_per_clip + row_index)*ld + j] = table[t*width + j].
 * Optionally first restores dst from `init` (tedexp: the memory stream is rewritten by the
 * layers, so each step starts again from emb_mem(speech)+PE, models/nn.py:433-442).
 * ------------------------------------------------------------------------------------------ */
int gd_scatter_step_row_f32(float* dst, const float* init, const float* table, const int32_t* step_ptr,
                            int32_t n_clips, int32_t rows_per_clip, int32_t row_index, int32_t width, int32_t ld,
                            void* stream);
int gd_scatter_step_row_bf16(void* dst, const void* table, const int32_t* step_ptr, int32_t n_clips,
                             int32_t rows_per_clip, int32_t row_index, int32_t width, int32_t ld, void* stream);

/* (N, C, T) fp32 -> bf16 token rows [N*T, ld] (cols >= C zeroed): builds the first emb_x operand from x_T.
 * `add` (optional, (N, C, T) fp32) is added before the cast (see gd_ddpm_desc.xa_add). */
int gd_pack_pose_rows(const float* x, void* xa_bf16, int32_t n_clips, int32_t C, int32_t T, int32_t ld, void* stream);
int gd_pack_pose_rows_add(const float* x, const float* add, void* xa_bf16, int32_t n_clips, int32_t C, int32_t T,
                          int32_t ld, void* stream);

/* fp32 -> bf16 row conversion (weights / conditioning repack): dst[r*ldd + c] = src[r*lds + c], c < cols;
 * columns [cols, cols_padded) are zero-filled. */
int gd_cast_rows_bf16(const float* src, int32_t lds, void* dst, int32_t ldd, int32_t rows, int32_t cols,
                      int32_t cols_padded, void* stream);

/* ------------------------------------------------------------------------------------------
 * Speech encoder (once per clip): HA2GSpeechEncoder `models/modules/ha2g/speech_encoder.py:37-61` over the ResNetSE-34
 * trunk `models/modules/ha2g/model/ResNetSE34V2.py:118-189` with SEBasicBlock / SELayer `.../ResNetBlocks.py:7-37,81-96`.
 *
 * Feature maps are channel-last bf16 "pixel rows" over a zero-bordered grid: row = image*(grid_h*grid_w) + y*grid_w + x,
 * c channels contiguous - one bf16 plane [c] (`split` = 0) or two planes [hi(c) | lo(c)] with value = hi + lo
 * (`split` = 1, ~16 mantissa bits).  The border pixels are never written, so a buffer zeroed once keeps supplying the
 * convolutions' zero padding.
 *
 * gd_conv_taps_bf16: nn.Conv2d (+ bias) [+ ReLU] + eval-mode BatchNorm2d as one implicit GEMM on the tensor cores:
 *
 *   acc[row, co] = sum_tap sum_{k < k_per_tap} W[co, tap*k_per_tap + k] * in[row + tap_shift[tap], k mod in_ld]   (rows outside read 0)
 *   v            = (relu ? max(acc + bias, 0) : acc + bias) * scale + shift                  (scale/shift = folded BatchNorm)
 *
 * stored as bf16 for the pixels with y0 <= y <= y1, x0 <= x <= x1 whose (y - y0, x - x0) are multiples of `stride`, at
 * output row  image*out_img_stride + ((y-y0)/stride)*out_y_stride + ((x-x0)/stride)*out_x_stride + out_offset.
 * A 3x3 "same" convolution on a bordered grid is tap_shift = (ky-1)*grid_w + (kx-1) over the interior; a stride-2
 * convolution is the same accumulation keeping every other pixel (ResNetSE34V2.py:96-112 `_make_layer`).
 *
 * Split precision (the default of the encoder): a feature map row holds two bf16 planes [hi(c) | lo(c)], value = hi + lo
 * (~16 mantissa bits).  The K walk of a tap wraps around the input row (k mod in_ld), so with k_per_tap = 3c the operand
 * sequence is hi, lo, hi; the host packs W per tap as [Whi | Whi | Wlo] and the GEMM accumulates
 * hi*Whi + lo*Whi + hi*Wlo in fp32 - the bf16x3 product, whose dropped term lo*Wlo is ~2^-18 relative.  With split_out
 * the epilogue stores channels [0, c_store) as [hi(c_store) | lo(c_store)].  Plain bf16: in_ld = k_per_tap = c_in.
 * `walk` selects how the K dimension of a tap is laid out (same arithmetic, fewer operand loads):
 *   0  generic: k-th element of a tap reads input column k mod in_ld (W as described above);
 *   1  rows [hi(32) | lo(32)] (in_ld 64, k_per_tap 128): W per tap = [Whi|Whi | Wlo|0]; the input tile is loaded once and
 *      multiplied by both 64-wide W blocks;
 *   2  rows [hi(c) | lo(c)], c % 64 == 0 (k_per_tap = in_ld = 2c): W per tap and 64-channel block = [Whi(64) | Wlo(64)]; the hi
 *      and lo tiles of the block are loaded once and the three products hi*Whi, lo*Whi, hi*Wlo are formed from them.
 * ------------------------------------------------------------------------------------------ */
#define GD_CONV_MAX_TAPS 9
typedef struct gd_conv_desc {
    const void* in;      /* bf16 [n_images*grid_h*grid_w, in_ld]                                */
    const void* W;       /* bf16 [c_out, n_taps*k_per_tap]                                      */
    int32_t n_images, grid_h, grid_w;
    int32_t in_ld;       /* input row width (elements), multiple of 64                          */
    int32_t k_per_tap;   /* K elements per tap, multiple of 64                                  */
    int32_t c_out;       /* GEMM N: multiple of 64                                              */
    int32_t n_taps;
    int32_t tap_shift[GD_CONV_MAX_TAPS];
    const float* bias;   /* [c_out] or NULL                                                     */
    const float* scale;  /* [c_out]                                                             */
    const float* shift;  /* [c_out]                                                             */
    int32_t relu;
    int32_t y0, y1, x0, x1, stride;
    void* out;           /* bf16 rows of out_ld elements                                        */
    int32_t out_ld;
    int32_t out_img_stride, out_y_stride, out_x_stride, out_offset;
    int32_t c_store;     /* output channels stored (multiple of 32, <= c_out)                   */
    int32_t split_out;   /* 0: [c_store] bf16;  1: [hi(c_store) | lo(c_store)]                  */
    int32_t walk;        /* K layout of a tap: 0 generic wrap, 1 / 2 split rows with operand reuse */
} gd_conv_desc;

int gd_conv_taps_bf16(const gd_conv_desc* d, void* stream);

/* Mel front end of HA2GSpeechEncoder (speech_encoder.py:18-27,50; PreEmphasis ha2g/model/utils.py:22-37;
 * torchaudio MelSpectrogram(16 kHz, n_fft 1024, hop 512, 128 HTK mel bins, power 2, centre reflect padding)):
 *   y[t] = wav[t] - preemph * wav[t-1] (reflect pad 1);  S = |STFT(y, periodic Hann `window`[1024])|^2;
 *   mel[clip, m, frame] = sum_k S[k, frame] * fb[k, m] + add_eps          frames = wav_len/512 + 1
 * twiddle: fp32 [512][2] = (cos, -sin)(2*pi*k/1024); fb: fp32 [513, 128] (`mel_scale.fb`); fb_range: int32 [128][2] =
 * first / last frequency bin with a non-zero weight per mel bin.  Fixed n_fft / hop / bins as in the reference. */
int gd_mel_power(const float* wav, int32_t n_clips, int32_t wav_len, const float* window, const float* twiddle,
                 const float* fb, const int32_t* fb_range, float preemph, float add_eps, float* mel, void* stream);

/* nn.InstanceNorm1d(128) (speech_encoder.py:28,51): every row of `len` values (one mel bin of one clip over time) is
 * normalised in place with its own mean and biased variance. */
int gd_instance_norm_rows(float* x, int32_t rows, int32_t len, float eps, void* stream);

/* Stem: conv1 (1 -> c_real channels, 3x3, bias) + ReLU + BatchNorm (ResNetSE34V2.py:127-129) on the normalised mel image
 * mel fp32 (n_images, H, W) -> bf16 pixel rows on the (H+2) x (W+2) bordered grid, channels [c_real, c_pad) written as 0.
 *   w fp32 [c_real, 9], bias / scale / shift fp32 [c_real]. */
int gd_speech_stem(const float* mel, const float* w, const float* bias, const float* scale, const float* shift,
                   void* out_bf16, int32_t n_images, int32_t H, int32_t W, int32_t c_real, int32_t c_pad, int32_t split,
                   void* stream);

/* SELayer gate (ResNetBlocks.py:81-96): gate[img, c] = sigmoid(W2 · relu(W1 · mean_pixels(y[img]) + b1) + b2).
 * y: pixel rows on a bordered grid (border = 0), `c` stored channels of which the first c_real are real;
 * w1 fp32 [c_hidden, c_real], w2 fp32 [c_real, c_hidden]; gate fp32 [n_images, c] (padding channels get 0).
 * The pixel sum is taken in fixed slices that depend on the grid only: the gate of a clip does not depend on the batch
 * it is in.  `scratch`: at least gd_se_gate_scratch_bytes() bytes, zeroed once by the caller (the kernel leaves its
 * counters at zero; they sit in a fixed-size block at the start, so one scratch sized for the largest batch serves calls
 * with any smaller n_images), not shared with a concurrently running gd_se_gate. */
int64_t gd_se_gate_scratch_bytes(int32_t n_images, int32_t grid_h, int32_t grid_w, int32_t c);
int gd_se_gate(const void* y_bf16, int32_t n_images, int32_t grid_h, int32_t grid_w, int32_t c, int32_t split, int32_t c_real,
               int32_t c_hidden, const float* w1, const float* b1, const float* w2, const float* b2, float* gate,
               void* scratch, int64_t scratch_bytes, void* stream);

/* Block tail (ResNetBlocks.py:30-36): out = relu(gate[img, c] * y + residual) on the interior pixels of the grid. */
int gd_se_residual_relu(const void* y_bf16, const void* residual_bf16, const float* gate, void* out_bf16,
                        int32_t n_images, int32_t grid_h, int32_t grid_w, int32_t c, int32_t split, void* stream);

/* nn.PixelShuffle(r) (ResNetSE34V2.py:167-168,179-180) from a bordered grid (H+2) x (W+2) with c_in channels to an
 * unbordered (H*r) x (W*r) grid with c_in/r^2 real channels zero-padded to c_out_pad. */
int gd_pixel_shuffle_rows(const void* in_bf16, void* out_bf16, int32_t n_images, int32_t H, int32_t W, int32_t c_in,
                          int32_t r, int32_t c_out_pad, int32_t split, void* stream);

/* *step_ptr += delta (one thread) — closes a denoise step inside a captured graph. */
int gd_step_add(int32_t* step_ptr, int32_t delta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GD_B200_H */

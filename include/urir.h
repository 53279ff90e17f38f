/* liburir -- C-ABI of the B200-native unet-rir hot path (sm_100a only).
 *
 * The reference (igmsalinas/unet-rir) is pure Python/TensorFlow and has no FFI of its own
 * (SURVEY.md 8b); every device op it triggers is a TF/cuDNN/librosa library call. This header is
 * therefore the boundary a maintainer of the reference would bind (ctypes stub shown in
 * INTEGRATION.md) to replace those calls one for one. Each entry point cites the reference
 * call site(s) it stands in for (paths relative to the reference root).
 *
 * Conventions (all entry points):
 *   - return 0 on success, < 0 on error (URIR_ERR_*); urir_last_error() gives the thread-local
 *     message. Shapes that the library does not support fail loudly; there is no CPU fallback.
 *   - never allocate, never synchronise, never own memory; all pointers are DEVICE pointers
 *     unless the argument name ends in _host; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), so calls are CUDA-graph capturable.
 *   - activations are NHWC. A tensor may live inside a wider buffer (skip-concat slices):
 *     `*_ld` is the number of elements per pixel of the holding buffer, `*_coff` the channel
 *     offset of the tensor inside it.
 *   - conv weights are passed in two bf16 layouts produced by urir_weight_prep():
 *       w_ck : [R*S][C][K]  (== Keras HWIO, and == Keras Conv2DTranspose HWOI of the
 *                             transposed layer)
 *       w_kc : [R*S][K][C]
 */
#ifndef URIR_H_
#define URIR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define URIR_VERSION 100

#define URIR_OK           0
#define URIR_ERR_ARG     -1   /* bad / unsupported shape or argument           */
#define URIR_ERR_CUDA    -2   /* CUDA runtime / driver error                   */
#define URIR_ERR_UNSUP   -3   /* requested implementation cannot run the shape */

#define URIR_F32   0
#define URIR_BF16  1

#define URIR_IMPL_AUTO   0    /* tcgen05 implicit GEMM when the shape qualifies, else SIMT */
#define URIR_IMPL_SIMT   1    /* CUDA-core direct convolution                              */
#define URIR_IMPL_TC     2    /* tcgen05 implicit GEMM or URIR_ERR_UNSUP                   */
#define URIR_IMPL_HALO   3    /* persistent halo-tile tcgen05 kernel (stride-1 fprop/dgrad)
                                 or URIR_ERR_UNSUP; AUTO picks it for the wide resolutions  */

#define URIR_IMPL_DEEP   4    /* persistent padded-sequence tcgen05 kernel of the deep stride-1 3x3 layers (>= 128
                                 channels either side) or URIR_ERR_UNSUP; AUTO prefers it wherever it applies          */

#define URIR_ACT_NONE    0
#define URIR_ACT_SIGMOID 1
#define URIR_ACT_RELU    2    /* inference: BatchNorm folded into weights / bias, ReLU in the epilogue */

/* One strided convolution y[N,P,Q,K] = conv(x[N,H,W,C], w[R,S,C,K]) (+bias), TF "SAME"
 * geometry given explicitly. Conv2DTranspose layers are described by the strided conv they are
 * the input-gradient of (x = the LARGE tensor). */
typedef struct urir_conv_desc {
    int32_t N, H, W, C;         /* x dims                                             */
    int32_t K, R, S;            /* filter count and window                            */
    int32_t stride;             /* 1 or 2, both axes                                  */
    int32_t pad_top, pad_left;  /* leading pads; trailing pads are implied by P, Q    */
    int32_t P, Q;               /* y spatial dims                                     */
    int32_t x_ld, x_coff;       /* x buffer: elements per pixel, channel offset       */
    int32_t y_ld, y_coff;       /* y buffer: elements per pixel, channel offset       */
    int32_t x_dtype, y_dtype;   /* URIR_F32 | URIR_BF16                               */
    int32_t impl;               /* URIR_IMPL_*                                        */
    int32_t act;                /* fprop only: URIR_ACT_* applied after bias          */
    int32_t accumulate;         /* fprop/dgrad: out += result instead of out = result;
                                   wgrad: dw += result (the caller zeroed dw, e.g. one memset of the
                                   flat gradient buffer) instead of dw = result                        */
} urir_conv_desc;

typedef struct urir_stft_desc {
    int32_t n_fft, win_length, hop_length;  /* dataset.py:62-64 : 256, 128, 64          */
    int32_t n_samples;                      /* 9600 = 0.2 s x 48 kHz (dataset.py:66-67)  */
    int32_t n_bins, n_frames;               /* 129, 151 (postprocess.py:54)              */
    int32_t H_pad, W_pad;                   /* 144, 160 (dataset.py:70)                  */
    int32_t pad_mode;                       /* 0 = constant (librosa>=0.10), 1 = reflect */
    int32_t remove_mean;                    /* Loader.load mean removal preprocess.py:56 */
    int32_t normalized;                     /* 1: spec holds Normalizer-normalised amp/phase
                                               (preprocess.py:26-41); 0: raw |S| and angle(S)  */
} urir_stft_desc;

/* ---- library state ------------------------------------------------------------------- */
int         urir_version(void);
const char* urir_last_error(void);
/* number of kernels launched by this library since load; kind 0 = all, 1 = tcgen05 only */
long long   urir_launch_count(int kind);
/* which kernel family served the convolution calls: urir_family_calls(f) = successful urir_conv2d_* calls dispatched to
 * family f since load; urir_family_name(f) its name. Parity tests use these to prove that the dispatch they checked
 * against the oracle is the one the benchmark times. */
#define URIR_FAM_SIMT           0   /* CUDA-core direct convolution (cross-check path)                      */
#define URIR_FAM_IGEMM          1   /* one-tile-per-CTA tcgen05 implicit GEMM (fprop / dgrad)              */
#define URIR_FAM_HALO           2   /* persistent halo-tile tcgen05 kernel, stride-1 fprop / dgrad         */
#define URIR_FAM_HALO_S2_FPROP  3   /* the same kernel over the four parity planes of a stride-2 input     */
#define URIR_FAM_HALO_UP2       4   /* stride-2 dgrad / Conv2DTranspose forward as one 2x2 problem         */
#define URIR_FAM_THIN_GEMM      5   /* 2-channel stem fprop / head dgrad                                   */
#define URIR_FAM_HEAD_FPROP     6   /* 2-channel 6x6 sigmoid head                                          */
#define URIR_FAM_WGRAD_TC       7   /* split-pixel tcgen05 weight gradient                                 */
#define URIR_FAM_WGRAD_HALO     8   /* persistent halo weight gradient, stride 1                           */
#define URIR_FAM_WGRAD_HALO_S2  9   /* persistent halo weight gradient, stride 2                           */
#define URIR_FAM_THIN_WGRAD    10   /* 2-channel stem / head weight gradient                               */
#define URIR_FAM_DEEP          11   /* persistent padded-sequence tcgen05 kernel of the deep (<= 36x40) layers */
#define URIR_FAM_COUNT         12
long long   urir_family_calls(int family);
const char* urir_family_name(int family);
/* programmatic dependent launch of the library's kernels (default on; URIR_NO_PDL=1 starts with it off):
 * a kernel's prologue may overlap its predecessor's tail. Returns the previous setting. Turn it off to time
 * kernels one by one with events. */
int         urir_set_pdl(int enabled);
/* deterministic mode (default off; URIR_DETERMINISTIC=1 starts with it on): every cross-CTA floating-point reduction
 * (BatchNorm statistics from the conv epilogues, BatchNorm backward sums, weight-gradient splits, bias gradients,
 * embedding gradient, loss scalars) is committed in a fixed CTA order instead of by unordered atomics, so two runs
 * on the same inputs are bit-identical. Slower (the commits serialise); meant for parity tests. Returns the previous
 * setting. Takes effect for kernels launched (or captured into a graph) afterwards. */
int         urir_set_deterministic(int enabled);

/* ---- convolutions: replace tf.keras Conv2D / Conv2DTranspose and their autodiff ------- */
/* Conv2D forward (dl_models/u_net.py:269-276, 366, 248, 262). Also Conv2DTranspose's
 * input-gradient. `stats` (optional, fp32[2*K], caller-zeroed) receives per-channel sum and
 * sum of squares of the output for the BatchNormalization that follows (u_net.py:367-368). */
int urir_conv2d_fprop(const urir_conv_desc* d, const void* x, const void* w_ck, const void* w_kc,
                      const float* bias, void* y, float* stats, void* stream);
/* Conv2D input-gradient (tape.gradient, amp_phase_trainer.py:138) and Conv2DTranspose forward
 * (dl_models/u_net.py:297-304): dx[N,H,W,C] = sum dy[N,P,Q,K] * w (+bias).
 * `stats` (optional, fp32[2*C], caller-zeroed): per-channel sum / sum of squares of dx -- the
 * sum is the bias gradient of the layer that produced x. */
int urir_conv2d_dgrad(const urir_conv_desc* d, const void* dy, const void* w_ck, const void* w_kc,
                      const float* bias, void* dx, float* stats, void* stream);
/* The same call when only the bias gradient is wanted from `stats`: stats[0..C) receives the channel sums of dx, stats[C..2C)
 * is left unspecified (the halo-tile kernels then skip the sums of squares: their epilogue is the bound on the 64-channel
 * input gradients of the net; the deep-layer kernel likewise; the remaining kernels still write both halves). Same buffer size, same zeroing rule. */
int urir_conv2d_dgrad_sums(const urir_conv_desc* d, const void* dy, const void* w_ck, const void* w_kc,
                           const float* bias, void* dx, float* stats, void* stream);
/* Conv2D / Conv2DTranspose weight gradient: dw[R,S,C,K] fp32 (HWIO) = sum x * dy; overwritten.
 * For a Conv2DTranspose layer pass x = gradient of its output, dy = its input. */
int urir_conv2d_wgrad(const urir_conv_desc* d, const void* x, const void* dy, float* dw,
                      void* stream);
/* The same operation for a 3x3 stride-2 layer on an even input (every Conv2DTranspose forward and every
 * stride-2 Conv2D input-gradient of the net) as ONE 2x2 stride-1 problem on the half-resolution grid with
 * GEMM-N = (row parity, column parity, channel): persistent halo-tile tcgen05 kernel, dy read once, the four
 * parity classes scattered by the epilogue. Weights in the w_up2 layout (urir_weight_prep_up2). Honours
 * d->accumulate (dx += ...). Returns URIR_ERR_UNSUP when urir_conv_path(d, 3) == 0. */
int urir_conv2d_dgrad_up2(const urir_conv_desc* d, const void* dy, const void* w_up2, const float* bias,
                          void* dx, void* stream);
/* fp32 HWIO [3][3][C][K] -> bf16 w_up2 [a*2+b][(ph, pw, c)][K] = w[2a+ph][2b+pw][c][k], zero where that tap
 * does not exist (16*C*K elements). */
int urir_weight_prep_up2(const float* w_hwio, void* w_up2, int C, int K, void* stream);
/* which kernel family URIR_IMPL_AUTO picks for this descriptor: op 0 = fprop, 1 = dgrad, 2 = wgrad;
 * returns 1 = tcgen05 implicit GEMM, 0 = CUDA-core direct convolution. op 3: 1 when urir_conv2d_dgrad_up2
 * supports the descriptor. Pure host query. */
int urir_conv_path(const urir_conv_desc* d, int op);
/* fp32 HWIO master weights -> the two bf16 operand layouts. */
int urir_weight_prep(const float* w_hwio, void* w_ck, void* w_kc, int taps, int C, int K,
                     void* stream);
/* Inference-mode BatchNormalization folded into the convolution that precedes it (one launch for all layers):
 * table_dev = n_entries x {w fp32 HWIO ptr, scale_shift fp32 [2K] ptr (urir_bn_finalize, inference mode), bias fp32 [K]
 * ptr, out w_kc bf16 [tap][K][C] ptr, out bias fp32 [K] ptr, taps, C, K} (int64, device memory):
 *   w_kc[t][k][c] = bf16(w[t][c][k] * scale[k]),  bias_out[k] = bias[k] * scale[k] + shift[k].
 * A fprop with these operands and act = URIR_ACT_RELU equals conv -> BN (moving statistics) -> ReLU. */
int urir_weight_fold_bn_batched(const int64_t* table_dev, int n_entries, void* stream);
/* the same for every kernel of a model in ONE launch: table_dev = n_entries x {w fp32 ptr, w_ck ptr,
 * w_kc ptr, taps, C, K} as int64 in device memory (refresh after each optimiser step). */
int urir_weight_prep_batched(const int64_t* table_dev, int n_entries, void* stream);
/* per-channel sum over pixels (bias gradients): out[c] = sum_p x[p*ld + coff + c]; overwritten. */
int urir_channel_sum(const void* x, int dtype, long long npix, int C, int ld, int coff,
                     float* out, void* stream);

/* ---- BatchNormalization + ReLU (u_net.py:367-369; Keras eps 1e-3, momentum .99) -------- */
/* training: stats = [sum, sumsq] over `count` pixels -> scale_shift[2C], mean_rstd[2C], and
 * moving stats update; inference (stats == NULL): scale/shift from the moving statistics. */
int urir_bn_finalize(const float* stats, double count, const float* gamma, const float* beta,
                     float* moving_mean, float* moving_var, float momentum, float eps,
                     int unbiased_moving_var, float* scale_shift, float* mean_rstd, int C,
                     void* stream);
/* y = relu(x*scale + shift); x, y bf16 (y may be a concat slice, or alias x). relu != 0. */
int urir_bn_relu_fwd(const void* x, int x_ld, int x_coff, const float* scale_shift,
                     void* y, int y_ld, int y_coff, long long npix, int C, int relu, void* stream);
/* training-mode urir_bn_finalize + urir_bn_relu_fwd in ONE launch: y = relu(BN(x)) with the batch statistics
 * taken from the conv epilogue's stats = [sum | sumsq]; also writes scale_shift / mean_rstd (for the backward
 * pass) and updates the moving statistics. */
int urir_bn_relu_fwd_train(const void* x, int x_ld, int x_coff, const float* stats, double count,
                           const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                           float momentum, float eps, int unbiased_moving_var, float* scale_shift,
                           float* mean_rstd, void* y, int y_ld, int y_coff, long long npix, int C,
                           void* stream);
/* sums[2C] = [sum g, sum g*xhat], g = dy * (x*scale+shift > 0). Overwritten; with prezeroed != 0 the caller
 * has zeroed `sums` (one memset for all layers instead of one per call). */
int urir_bn_relu_bwd_reduce(const void* dy, int dy_ld, int dy_coff, const void* x, int x_ld,
                            int x_coff, const float* scale_shift, const float* mean_rstd,
                            float* sums, long long npix, int C, int prezeroed, void* stream);
/* dx = gamma*rstd*(g - sum_g/n - xhat*sum_gx/n) (bf16); dgamma = sum_gx; dbeta = sum_g;
 * dbias (optional) = sum over pixels of dx = gradient of the preceding conv's bias (with prezeroed != 0 the
 * caller has zeroed it). */
int urir_bn_relu_bwd_apply(const void* dy, int dy_ld, int dy_coff, const void* x, int x_ld,
                           int x_coff, const float* scale_shift, const float* mean_rstd,
                           const float* gamma, const float* sums, void* dx, int dx_ld,
                           int dx_coff, float* dgamma, float* dbeta, float* dbias, long long npix,
                           int C, int prezeroed, void* stream);

/* ---- vector block (u_net.py:253-263) --------------------------------------------------- */
/* Embedding(2000,256) + Flatten: out bf16 [B, T*D] */
int urir_embedding_fwd(const int32_t* idx, const float* table, void* out, int B, int T, int D,
                       int vocab, void* stream);
/* dtable fp32 [vocab, D] (overwritten) += scatter of dx (fp32 or bf16, URIR_F32 | URIR_BF16) [B, T*D] */
int urir_embedding_bwd(const int32_t* idx, const void* dx, int dx_dtype, float* dtable, int B, int T,
                       int D, int vocab, void* stream);
/* Dense + Dropout on the tensor cores (a 1x1 convolution over B "pixels"):
 * out bf16 [B,N] = bf16(x bf16 [B,Kd] @ w + bias) * mask (mask fp32 [B,N] or NULL).
 * The kernel is passed in both bf16 layouts: w_kn [Kd][N] (Keras Dense kernel order) and w_nk [N][Kd]
 * (urir_weight_prep with taps = 1, C = Kd, K = N produces both). */
int urir_dense_fwd(const void* x, const void* w_kn, const void* w_nk, const float* bias,
                   const float* mask, void* out, int B, int Kd, int N, void* stream);
/* dy_eff = dy (bf16 [B,N]) * mask (written to the caller's bf16 [B,N] scratch `dy_eff`; unused when mask is NULL);
 * dw fp32 [Kd,N] = x^T dy_eff;  db fp32 [N];  dx bf16 [B,Kd] = dy_eff @ w^T.  dw / db / dx may each be NULL;
 * outputs are overwritten. */
int urir_dense_bwd(const void* x, const void* w_kn, const void* w_nk, const void* dy, const float* mask,
                   void* dy_eff, float* dw, float* db, void* dx, int B, int Kd, int N, void* stream);
/* inverted-dropout mask {0, 1/(1-rate)} from a counter-based generator; the stream position is
 * (seed, *step_dev) so CUDA-graph replays draw fresh masks. */
int urir_dropout_mask(float* mask, long long n, float rate, uint64_t seed,
                      const int32_t* step_dev, void* stream);

/* ---- loss (amp_phase_trainer.py:143-168 and main_training.py:203-235) ------------------ */
/* y_true, y_pred fp32 [npix, 2] (amp, phase). losses[4] (overwritten) =
 *   [w_amp*SSE + w_ph*sum(1-cos), mean(1-cos), mean sq err, 0];
 * grad (optional, fp32 [npix,2]) = dL/dy_pred, times y(1-y) when sigmoid_bwd != 0 (the head's
 * "sigmoid_layer", u_net.py:249). grad_bf16 (optional): the same gradient as bf16 with
 * `grad_bf16_ld` elements per pixel (only the first 2 are written) -- the TMA-loadable operand of the
 * head's tensor-core dgrad / wgrad. */
int urir_ampphase_loss(const float* y_true, const float* y_pred, long long npix, float w_amp,
                       float w_ph, int sigmoid_bwd, float* losses, float* grad, void* grad_bf16,
                       int grad_bf16_ld, void* stream);

/* The generic trainer's loss (trainer.py:146-156): squared error over BOTH channels. y_true, y_pred fp32 [npix, 2].
 * losses[4] (overwritten) = [w * SSE(amp and phase), mean(1 - cos) of the phase channel, mean sq err of the amp channel,
 * mean sq err over both]; grad (optional) = 2 w (y_pred - y_true), times y(1-y) when sigmoid_bwd != 0. */
int urir_mse2_loss(const float* y_true, const float* y_pred, long long npix, float w, int sigmoid_bwd, float* losses,
                   float* grad, void* stream);

/* ---- optimiser (amp_phase_trainer.py:30-35,139 ; Keras conventions) ------ */
/* flat Adam over n contiguous fp32 elements; lr and step (0-based count of completed steps)
 * are read from device memory so a captured graph follows the LR schedule. */
int urir_adam(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev,
              const int32_t* step_dev, float beta1, float beta2, float eps, void* stream);
int urir_sgd(float* p, const float* g, long long n, const float* lr_dev, void* stream);
/* tf.keras.optimizers.Nadam ("nadam" in the optimiser name, amp_phase_trainer.py:30-31; Keras momentum schedule
 * u_t = b1 (1 - 0.5 * 0.96^(0.004 t))). coef_dev: fp32[4] device state owned by the caller, [0] = running product of
 * u_i (the kernel treats *step_dev == 0 as a fresh state). Two launches; graph capturable. */
int urir_nadam(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev, const int32_t* step_dev,
               float* coef_dev, float beta1, float beta2, float eps, void* stream);
/* tensorflow_addons LAMB (the generic trainer's "lamb" option, trainer.py:37-38): Adam direction with a per-variable
 * trust ratio ||w|| / ||update||. table_dev = n_vars x {element offset, element count} (int64, device) of the variables
 * inside the flat buffers; upd = fp32 scratch of the flat size; norms = fp32[2 * n_vars] scratch. */
int urir_lamb(float* p, const float* g, float* m, float* v, float* upd, const int64_t* table_dev, int n_vars, float* norms,
              const float* lr_dev, const int32_t* step_dev, float beta1, float beta2, float eps, float weight_decay,
              void* stream);
int urir_step_increment(int32_t* step_dev, void* stream);
/* y += a*x (fp32) -- L2 kernel-regulariser gradient of the DP loss (main_training.py:232-233) */
int urir_axpy(float* y, const float* x, float a, long long n, void* stream);
/* out[0] (+)= scale * sum(x^2) */
int urir_sumsq(const float* x, long long n, float scale, float* out, int accumulate, void* stream);
/* the L2 kernel regulariser of ALL regularised tensors in one launch: table_dev = n_entries x {param fp32 ptr,
 * grad fp32 ptr, element count} (int64, device memory); out[0] = coef * sum ||W||^2 (overwritten), grad += 2*coef*W. */
int urir_l2_reg_batched(const int64_t* table_dev, int n_entries, float coef, float* out, void* stream);
/* elementwise helpers for the alternate block modes (u_net.py:337,359) and casts */
int urir_add_bf16(const void* a, const void* b, void* out, long long n, void* stream);
/* the same for C-channel slices of NHWC buffers (residual_block_1 / residual_block_2 Adds, u_net.py:337,359, and
 * their gradient fan-in); out may alias an input */
int urir_add_bf16_strided(const void* a, int a_ld, int a_coff, const void* b, int b_ld, int b_coff,
                          void* out, int out_ld, int out_coff, long long npix, int C, void* stream);
int urir_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream);
/* fp32 [npix][C] -> bf16 [npix][ld], ld >= C: the 16-byte-pitch bf16 copy of the 2-channel input
 * spectrogram that the stem's tensor-core weight-gradient reads through TMA (u_net.py:269-276). */
int urir_cast_pad_bf16(const float* x, void* y, long long npix, int C, int ld, void* stream);

/* ---- signal path (preprocess.py:13-113, postprocess.py:78-133) ------------------------- */
/* Loader mean removal + FeatureExtractor.extract + Normalizer.normalize + TensorPadder:
 * wav fp32 [B, n_samples] -> spec fp32 [B, H_pad, W_pad, 2] */
int urir_stft_ampphase(const float* wav, int B, const urir_stft_desc* d, float* spec,
                       void* stream);
/* PostProcess.post_process, algorithm 'ph': spec fp32 [B,H_pad,W_pad,2] -> wav fp32
 * [B, n_samples] */
int urir_istft_from_ampphase(const float* spec, int B, const urir_stft_desc* d, float* wav,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* URIR_H_ */

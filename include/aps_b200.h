/*
 * aps_b200.h — C ABI of libaps_b200.so: the B200 (sm_100a) kernels behind the APS hot path.
 *
 * The reference (funcwj/aps) has no FFI: its boundary is the Python nn.Module surface
 * (SURVEY.md §8b).  This header is the boundary a maintainer would bind from Python
 * (ctypes — see INTEGRATION.md) underneath those modules.  Every entry point names the
 * reference code it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`; all tensors fp32,
 *     row-major contiguous unless a leading dimension is passed;
 *   - the library never allocates or frees device memory and keeps no pointer after a call
 *     returns; scratch is caller-provided;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and return immediately;
 *   - return value 0 = ok; < 0 = error, text via aps_b200_last_error().  CUDA failures carry the
 *     CUDA error string (so "out of memory" stays greppable for the reference trainer's OOM
 *     guard, aps/trainer/ddp.py:146 + aps/const.py:23).
 */
#ifndef APS_B200_H_
#define APS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APS_B200_ABI_VERSION 6

/* library / device ------------------------------------------------------------------------ */
int aps_b200_abi_version(void);
/* Selects `device`, checks it is compute capability 10.x, caches the SM count. */
int aps_b200_init(int device);
/* Copies the calling thread's last error message (NUL terminated) into buf. */
int aps_b200_last_error(char* buf, size_t len);

/* STFT front-end --------------------------------------------------------------------------
 * One descriptor drives both the fused feature kernel (F1) and the complex STFT kernel (F2).
 * Replaces aps/transform/utils.py:227-290 (_forward_stft: reflect pad :257-260, per-frame
 * Kaldi pre-emphasis :263-272, window*DFT :262/:274, one-sided slice :281-284),
 * utils.py:363-415 (_pytorch_stft), asr.py:84 (RescaleTransform), asr.py:111-113
 * (PreEmphasisTransform).
 */
typedef struct aps_b200_stft_desc {
    int32_t nfft;          /* FFT size, power of two in [64, 1024]                               */
    int32_t frame_width;   /* samples covered by a frame: nfft (librosa/torch) or frame_len (kaldi) */
    int32_t hop;           /* frame shift                                                         */
    int32_t center_pad;    /* reflect padding on both sides in samples (0 = center False)         */
    int32_t rescale;       /* 1: x <- rint(x*32767) first (audio_norm=False)                      */
    float   utt_preemph;   /* utterance-level pre-emphasis ("emph" token), 0 = off                */
    float   frame_preemph; /* per-frame Kaldi pre-emphasis coefficient, 0 = off                   */
    float   frame_one_minus; /* float32(1 - frame_preemph) as the reference computes it            */
    float   scale;         /* 1, or 1/sqrt(nfft) when stft_normalized                             */
    const float* window;   /* [frame_width] analysis window (already centre-padded if needed)     */
    const float* twiddles; /* table from aps_b200_fft_tables_host(nfft, inverse=0), on device     */
} aps_b200_stft_desc;

/* Number of floats in the twiddle table for `nfft` (0 if nfft is unsupported). */
int64_t aps_b200_fft_table_floats(int nfft);
/* Fills a HOST buffer of aps_b200_fft_table_floats(nfft) floats; the caller uploads it once. */
int aps_b200_fft_tables_host(int nfft, int inverse, float* out_host);

/* Number of frames for `num_samples` (bit-exact integer rule of utils.py:653-662). */
int64_t aps_b200_num_frames(int64_t num_samples, int frame_width, int hop, int center_pad);

/* Epilogue of the fused feature kernel.
 * Replaces asr.py:296-303 (MagnitudeTransform), :216-223 (TFTranspose), :350-357 (Power),
 * :416-428 (MelTransform), :453-464 (LogTransform), :576-618 (CmvnTransform, per-frame and
 * global variants; "all band" runs as aps_b200_cmvn_allband afterwards).
 */
typedef struct aps_b200_feat_desc {
    int32_t power;         /* 1: |X|, 2: |X|^2                                                    */
    int32_t num_mels;      /* 0: linear spectrogram output (nfft/2+1 dims)                        */
    const int32_t* mel_start; /* [num_mels] first bin of each band                                */
    const int32_t* mel_len;   /* [num_mels] number of bins of each band                           */
    const float* mel_weight;  /* [num_mels, mel_stride] band weights, zero padded                 */
    int32_t mel_stride;
    int32_t log_mode;      /* 0 none, 1 log(clamp(x, min=log_eps)), 2 log(log_lower_bound + x)    */
    float   log_eps;
    float   log_lower_bound;
    int32_t cmvn_mode;     /* 0 none, 1 per frame over the feature axis, 2 global mean/std        */
    int32_t norm_mean;
    int32_t norm_var;
    float   cmvn_eps;
    const float* gmean;    /* [dims] (cmvn_mode 2) */
    const float* gstd;     /* [dims] (cmvn_mode 2) */
    int32_t* nan_count;    /* optional: += number of NaN feature values written (asr.py:41-45)   */
    const float* aug_mask; /* optional SpecAugment 0/1 mask [rows, T, dims] multiplied into the features AFTER cmvn
                              (asr.py:656-684 with mask_zero=True; the mask is drawn by the host RNG, augment.py:13-53) */
} aps_b200_feat_desc;

/* F1: wav [rows, num_samples] (row stride ld_wav floats) -> feats [rows, T, dims],
 * T = aps_b200_num_frames(num_samples, ...), dims = num_mels or nfft/2+1.               */
int aps_b200_feats_fwd(const float* wav, int64_t rows, int64_t num_samples, int64_t ld_wav,
                       const aps_b200_stft_desc* stft, const aps_b200_feat_desc* feat,
                       float* out, void* stream);

/* F2: wav [rows, num_samples] -> packed STFT [rows, nfft/2+1, T, 2] (real, imag) or, with
 * polar != 0, (sqrt(re^2+im^2+polar_eps), atan2(im, re)) — utils.py:285-288.             */
int aps_b200_stft_fwd(const float* wav, int64_t rows, int64_t num_samples, int64_t ld_wav,
                      const aps_b200_stft_desc* stft, int polar, float polar_eps,
                      float* out, void* stream);

/* F3: packed STFT [rows, nfft/2+1, T, 2] -> wav [rows, aps_b200_istft_num_samples(T, ...)].
 * `stft->twiddles` must be the INVERSE table (aps_b200_fft_tables_host(nfft, 1, ...)),
 * `stft->scale` = 1/nfft (or 1/sqrt(nfft) when normalized), `stft->window` the synthesis window
 * of `frame_width` samples; hop / center_pad as in the forward transform.  `polar` != 0 reads
 * (magnitude, phase).  Replaces aps/transform/utils.py:293-360 (_inverse_stft: Hermitian mirror
 * :327-332, iDFT conv_transpose1d :336, window^2 overlap-add :345-349, centre trim :354-357,
 * divide :358) and :418-469 (_pytorch_istft).                                              */
int64_t aps_b200_istft_num_samples(int64_t num_frames, int frame_width, int hop, int center_pad);
int aps_b200_istft_fwd(const float* spec, int64_t rows, int64_t num_frames,
                       const aps_b200_stft_desc* stft, int polar, float eps, float* out,
                       void* stream);

/* F1b: features from a packed STFT [rows, channels, num_bins, T, 2]: reference channel ->
 * sqrt(re^2+im^2+mag_eps) -> power -> [mel] -> [log] -> [cmvn] -> out[rows, T, ld_out] (first
 * `dims` columns).  Replaces aps/transform/enh.py:39-49 (RefChannelTransform) + the layer chain
 * asr.py:296-303, :216-223, :350-357, :416-428, :453-464, :576-618 as composed by
 * enh.py:518-529 and run by enh.py:595-613.                                               */
int aps_b200_spec_feats_fwd(const float* spec, int64_t rows, int64_t channels, int64_t ref_channel,
                            int64_t num_bins, int64_t num_frames, float mag_eps,
                            const aps_b200_feat_desc* feat, float* out, int64_t ld_out,
                            void* stream);

/* IPD: out[n, t, col0 + m*F + k] = cos(angle(x[n, l_m, k, t]) - angle(x[n, r_m, k, t])), and with
 * with_sin != 0 the sines in columns col0 + (P+m)*F + k.  index_l / index_r: device int32[P].
 * Replaces aps/transform/enh.py:67-76 (PhaseTransform) + :112-143 (IpdTransform).          */
int aps_b200_ipd_fwd(const float* spec, int64_t batch, int64_t channels, int64_t num_bins,
                     int64_t num_frames, const int32_t* index_l, const int32_t* index_r,
                     int64_t num_pairs, int with_sin, float* out, int64_t ld_out, int64_t col0,
                     void* stream);

/* Utterance-level "all band" CMVN over (T, dims) in place — asr.py:587-596. x: [rows, T, dims] */
int aps_b200_cmvn_allband(float* x, int64_t rows, int64_t T, int64_t dims, int norm_mean,
                          int norm_var, float eps, void* stream);

/* Multi-channel front-end (MVDR) ------------------------------------------------------------
 * Complex spectrograms are passed as separate real / imaginary base pointers plus the element
 * strides {n, c, f, t} of the [N, C, F, T] view (aps.cplx.ComplexTensor keeps two real tensors,
 * aps/cplx.py:18-33; they are normally the two halves of a packed STFT, stride_t = 2).
 * Masks are [N, T, F] or [N, F, T] views given by their strides {n, t, f}. 2 <= C <= 6.
 */

/* out[n, f] = max_t |mask[n, t, f]| over t < lens[n] (lens may be NULL).
 * Replaces the th.norm(mask, inf, dim=1) of aps/asr/filter/mvdr.py:112-113 (+ padding :109-111). */
int aps_b200_mask_colmax(const float* mask, int64_t stride_n, int64_t stride_t, int64_t stride_f,
                         int64_t batch, int64_t num_frames, int64_t num_bins, const int64_t* lens,
                         float* out, void* stream);

/* F4: Rs[n,f] = sum_t ms X X^H / max(sum_t ms, den_eps), Rn likewise with the noise mask, where
 * ms = (t < lens[n] ? mask_s : 0) / (max_s[n,f] + norm_eps) (max_s NULL: no normalisation) and the
 * noise mask is mask_n processed the same way or, when mask_n is NULL, 1 - ms.  Rs / Rn are
 * [N, F, C, C, 2] (Rn may be NULL).  Replaces mvdr.py:103-116 (_process_mask) + :42-61
 * (estimate_covar, four real batched GEMMs per covariance through aps/cplx.py:242-252).        */
int aps_b200_covar_fwd(const float* x_real, const float* x_imag, const int64_t* x_strides,
                       int64_t batch, int64_t channels, int64_t num_bins, int64_t num_frames,
                       const float* mask_s, const int64_t* mask_s_strides, const float* max_s,
                       const float* mask_n, const int64_t* mask_n_strides, const float* max_n,
                       const int64_t* lens, float norm_eps, float den_eps, float* Rs, float* Rn,
                       void* stream);

/* logits[n, c] of the reference-channel attention: gvec . tanh(proj . |offdiag-mean(Rs)[n, c, :]| + b).
 * Replaces mvdr.py:158-173 (ChannelAttention.forward up to the softmax).                      */
int aps_b200_mvdr_ref_logits(const float* Rs, int64_t batch, int64_t num_bins, int64_t channels,
                             const float* proj_weight, const float* proj_bias,
                             const float* gvec_weight, const float* gvec_bias, int64_t att_dim,
                             float* logits, void* stream);

/* F5: weight[n, f, :, 2] = (Rn + eps I)^-1 Rs u / (tr((Rn + eps I)^-1 Rs) + eps), u = softmax(logits[n]).
 * Replaces mvdr.py:174 (softmax), :75-101 (_derive_weight), aps/cplx.py:268-278 (inverse through
 * the real 2C x 2C matrix), mvdr.py:19-26 (trace), aps/cplx.py:221-226 (complex division).      */
int aps_b200_mvdr_weights(const float* Rs, const float* Rn, const float* logits, int64_t batch,
                          int64_t num_bins, int64_t channels, float eps, float* weight,
                          void* stream);

/* Y[n, f, t] = sum_c conj(weight[n, f, c]) X[n, c, f, t]; y_real / y_imag contiguous [N, F, T].
 * Replaces mvdr.py:29-39 (beamform).                                                          */
int aps_b200_beamform_fwd(const float* x_real, const float* x_imag, const int64_t* x_strides,
                          int64_t batch, int64_t channels, int64_t num_bins, int64_t num_frames,
                          const float* weight, float* y_real, float* y_imag, void* stream);

/* Dense layers ---------------------------------------------------------------------------------
 * out[m, n] = alpha * post(act(sum_k x[m, k] * weight[n, k] + bias[n])) + beta * residual[m, n],
 * post(v) = v * post_scale[n] + post_shift[n] when given (an eval-mode BatchNorm behind the activation)
 * (exact fp32 accumulate).  act: 0 none, 1 relu, 2 swish, 3 tanh, 4 sigmoid, 5 prelu, 6 glu (column
 * pairs (2j, 2j+1) -> out[:, j] = v0 * sigmoid(v1)), 7 leaky relu, 8 gelu (erf).
 */
typedef struct aps_b200_epilogue {
    const float* bias;         /* [N] or NULL */
    int32_t act;
    float   alpha;
    const float* prelu_slope;  /* [1] or [N] (prelu_per_channel) */
    int32_t prelu_per_channel;
    float   leaky_slope;
    const float* residual;     /* [M, ld_residual] or NULL */
    int64_t ld_residual;
    float   beta;
    const float* post_scale;   /* [N] or NULL (with post_shift) */
    const float* post_shift;
} aps_b200_epilogue;

/* x [rows, in_features] (row stride ld_x), weight [out_features, in_features] (torch Linear layout,
 * row stride ld_w).  Replaces F.linear / 1x1 Conv1d of aps/asr/transformer/impl.py:388-393, :454-475,
 * :62-83, aps/asr/base/encoder.py:415-441 (outp), aps/sse/bss/tcn.py:112-159.               */
int aps_b200_linear_fwd(const float* x, int64_t rows, int64_t in_features, int64_t ld_x,
                        const float* weight, int64_t ld_w, int64_t out_features,
                        const aps_b200_epilogue* epi, float* out, int64_t ld_out, void* stream);

/* Tensor-core engine (tcgen05.mma kind::tf32, double-buffered accumulators in TMEM, persistent 128 x BN
 * tiles) with a 3xTF32 split: hi = rn_tf32(v), lo = rn_tf32(v - hi) and hi*hi + hi*lo + lo*hi accumulated in
 * fp32.  WEIGHTS are passed pre-split (aps_b200_tf32_split, once per module; TMA loads them); ACTIVATIONS are
 * passed as plain fp32 and are gathered + split inside the kernel by producer warps, so neither a hi/lo copy
 * nor an im2col matrix of an activation is written to memory.  Same epilogue contract as aps_b200_linear_fwd.
 * Needs in_features % 4 == 0 and 16-byte aligned rows (linear), Cin % 32 == 0 (convolutions).           */
int aps_b200_tf32_split(const float* x, int64_t rows, int64_t cols, int64_t ld_x, float* hi, float* lo,
                        int64_t ld_out, void* stream);
/* F.linear: aps/asr/transformer/impl.py:62-83, :388-393, :454-475; 1x1 convs of aps/sse/bss/tcn.py:112-159 */
int aps_b200_linear_tc_fwd(const float* x, int64_t rows, int64_t in_features, int64_t ld_x,
                           const float* weight_hi, const float* weight_lo, int64_t ld_w,
                           int64_t out_features, const aps_b200_epilogue* epi, float* out,
                           int64_t ld_out, void* stream);
/* Encoder-stack variant of aps_b200_linear_tc_fwd (same reference call sites).  The tensor core reads an fp32 operand
 * as TF32 by dropping the low 13 mantissa bits, so an activation x can be fed RAW as its own "hi" part when its
 * companion  x_lo = rn_tf32(x - trunc_tf32(x))  exists: with x_lo != NULL both operand sides are loaded by TMA and the
 * kernel runs without gather / split warps.  out_lo != NULL makes the epilogue write the companion of the result for
 * the next layer (same leading dimension as out).  ksplit > 1 (needs x_lo) cuts K into slices whose RAW partial sums go
 * to out + s * split_stride (epilogue must be empty); aps_b200_layernorm2_fwd reduces them.                         */
int aps_b200_linear_tc2_fwd(const float* x, const float* x_lo, int64_t rows, int64_t in_features, int64_t ld_x,
                            const float* weight_hi, const float* weight_lo, int64_t ld_w,
                            int64_t out_features, const aps_b200_epilogue* epi, float* out, float* out_lo,
                            int64_t ld_out, int32_t ksplit, int64_t split_stride, void* stream);
/* aps_b200_conv2d_nhwc_tc_fwd that also writes the lo companion of its output */
int aps_b200_conv2d_nhwc_tc2_fwd(const float* x, int64_t batch, int64_t height, int64_t width,
                                 int64_t in_channels, const float* weight_hi, const float* weight_lo,
                                 int64_t out_channels, int kernel_h, int kernel_w, int stride_h, int stride_w,
                                 int pad_h, int pad_w, int dil_h, int dil_w, const aps_b200_epilogue* epi,
                                 float* out, float* out_lo, void* stream);
/* Conv2d as an implicit GEMM (same geometry / layouts as aps_b200_conv2d_nhwc_fwd; weight_hi / weight_lo are
 * the split [Cout, KH*KW*Cin] filter): aps/asr/base/component.py:251-307, aps/sse/enh/dcunet.py:24-45      */
int aps_b200_conv2d_nhwc_tc_fwd(const float* x, int64_t batch, int64_t height, int64_t width,
                                int64_t in_channels, const float* weight_hi, const float* weight_lo,
                                int64_t out_channels, int kernel_h, int kernel_w, int stride_h, int stride_w,
                                int pad_h, int pad_w, int dil_h, int dil_w, const aps_b200_epilogue* epi,
                                float* out, void* stream);
/* ConvTranspose2d as an implicit gather GEMM (as aps_b200_conv_transpose2d_nhwc_fwd; stride_w == 1): dcunet.py:48-70.
 * x_skip != NULL: the input is the "cat" skip connection of the DCCRN decoder on stacked complex channels,
 * [re(x) | re(x_skip) | im(x) | im(x_skip)] (aps/sse/enh/dcunet.py:258-262, dccrn.py:285), with x and x_skip of
 * in_channels / 2 channels each, read in place — the concatenated tensor is never materialised.                   */
int aps_b200_conv_transpose2d_nhwc_tc_fwd(const float* x, const float* x_skip, int64_t batch, int64_t height,
                                          int64_t width, int64_t in_channels, const float* weight_hi,
                                          const float* weight_lo, int64_t out_channels, int kernel_h,
                                          int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                                          int out_pad_h, int out_pad_w, const aps_b200_epilogue* epi,
                                          float* out, void* stream);

/* Implicit-GEMM convolution on channels-last data: x [B, H, W, Cin], weight [Cout, KH, KW, Cin],
 * out [B, OH, OW, Cout] with OH = (H + 2 pad - dil (K-1) - 1) / stride + 1.  Replaces the
 * Conv2d(+BatchNorm eval, folded by the caller)+ReLU block of aps/asr/base/component.py:251-307 and
 * the complex convolutions of aps/sse/enh/dcunet.py:24-45 (as one real convolution on stacked
 * real/imaginary channels).                                                                   */
int aps_b200_conv2d_nhwc_fwd(const float* x, int64_t batch, int64_t height, int64_t width,
                             int64_t in_channels, const float* weight, int64_t out_channels,
                             int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h,
                             int pad_w, int dil_h, int dil_w, const aps_b200_epilogue* epi,
                             float* out, void* stream);

/* Transposed convolution, channels-last: x [B, H, W, Cin], weight [Cout, KH, KW, Cin] (i.e. the torch
 * ConvTranspose2d weight [Cin, Cout, KH, KW] permuted), out [B, OH, OW, Cout] with
 * OH = (H-1)*stride - 2*pad + K + out_pad.  Replaces the nn.ConvTranspose2d pairs of
 * aps/sse/enh/dcunet.py:48-69 (ComplexConvTranspose2d, as one real transposed convolution on stacked
 * real/imaginary channels).                                                                   */
int aps_b200_conv_transpose2d_nhwc_fwd(const float* x, int64_t batch, int64_t height, int64_t width,
                                       int64_t in_channels, const float* weight, int64_t out_channels,
                                       int kernel_h, int kernel_w, int stride_h, int stride_w,
                                       int pad_h, int pad_w, int out_pad_h, int out_pad_w,
                                       const aps_b200_epilogue* epi, float* out, void* stream);

/* Complex ratio mask: per position p, (mr, mi) = mask[p*ld_mask + col_real|col_imag];
 * a = sqrt(mr^2 + mi^2 + eps); g = act(a) (0 none, 1 relu, 3 tanh, 4 sigmoid); m' = g*m/a;
 * out[p] = m' (apply = 0) or stft[p] * m' (complex product, apply = 1); out / stft are [positions, 2].
 * Replaces aps/sse/bss/dccrn.py:223-231 (_sep, cplx=True).                                    */
int aps_b200_cmask_fwd(const float* mask, int64_t ld_mask, int64_t col_real, int64_t col_imag,
                       const float* stft, int64_t positions, int act, float eps, int apply, float* out,
                       void* stream);

/* Remaining feature-chain tokens (rows a10, f2, f3) ----------------------------------------------------
 * SpecAugment apply on x [batch, channels, frames, dims] with mask [batch, frames, dims] (0/1, host RNG):
 * mask_zero != 0: out = x * mask; else out = mask == 0 ? mean(x) : x (global mean, asr.py:680-683).  workspace: >=
 * aps_b200_specaug_workspace_bytes(numel) bytes (fp64 partial sums, deterministic order); may be NULL for mask_zero. */
int64_t aps_b200_specaug_workspace_bytes(int64_t numel);
int aps_b200_specaug_apply(const float* x, int64_t batch, int64_t channels, int64_t frames, int64_t dims,
                           const float* mask, int32_t mask_zero, void* workspace, int64_t workspace_bytes,
                           float* out, void* stream);
/* Context splicing with edge clamping + frame subsampling: x [rows, frames, dims] -> out [rows, frames / subsampling,
 * (lctx + rctx + 1) * dims] (asr.py:687-728, utils.py:193-224).                                              */
int aps_b200_splice_fwd(const float* x, int64_t rows, int64_t frames, int64_t dims, int32_t lctx, int32_t rctx,
                        int32_t subsampling, float* out, void* stream);
/* One delta order: out[r, t, f] = sum_c scale[c] * in[r, clamp(t + c - ctx), f]; element (r, t, f) of in / out lives at
 * r * row_stride + t * frame_stride + f, so the slots of the concatenated / stacked result are written in place
 * (asr.py:731-781).                                                                                            */
int aps_b200_delta_fwd(const float* in, int64_t in_row_stride, int64_t in_frame_stride, int64_t rows, int64_t frames,
                       int64_t dims, int32_t ctx, const float* scale, float* out, int64_t out_row_stride,
                       int64_t out_frame_stride, void* stream);
/* Per-utterance speed perturbation (polyphase resampling): utterance n uses filter choice[n] (weights[i] is
 * [dst_sr[i], src_sr[i], taps[i]], utils.py:159-190) or is copied when choice[n] == num_filters; out [batch, ld_out] is
 * zero padded (asr.py:168-195, augment.py:85-109).                                                            */
int aps_b200_speed_perturb_fwd(const float* wav, int64_t batch, int64_t num_samples, int64_t ld_wav,
                               const int32_t* choice, int32_t num_filters, const float* const* weights,
                               const int32_t* dst_sr, const int32_t* src_sr, const int32_t* taps, float* out,
                               int64_t ld_out, void* stream);

/* Encoder (non-GEMM) kernels ---------------------------------------------------------------------
 * Activations are token-major rows; row(n, t) = n*stride_n + t*stride_t.
 */

/* out = LayerNorm(alpha * x + residual) * gamma + beta over the last `dim` values of every row
 * (residual / gamma / beta may be NULL).  Replaces nn.LayerNorm + the residual / macaron-factor
 * arithmetic of aps/asr/transformer/impl.py:424-428 and :508-540.                              */
int aps_b200_layernorm_fwd(const float* x, int64_t ld_x, const float* residual, int64_t ld_residual,
                           float alpha, const float* gamma, const float* beta, float eps,
                           int64_t rows, int64_t dim, float* out, int64_t ld_out, void* stream);

/* LayerNorm that also finishes a split-K tensor-core GEMM (aps_b200_linear_tc2_fwd, ksplit = num_parts):
 *   v = alpha * (sum_p x[p * part_stride + ...] + bias) + residual;  out = normalize ? LN(v) * gamma + beta : v
 * and, when out_lo != NULL, the TF32 lo companion of out.  dim % 128 == 0, dim <= 1024, 16-byte aligned operands.
 * Same reference lines as aps_b200_layernorm_fwd plus the second Linear of the feed-forward modules
 * (aps/asr/transformer/impl.py:388-393, :499-541).                                                       */
int aps_b200_layernorm2_fwd(const float* x, int64_t ld_x, int32_t num_parts, int64_t part_stride,
                            const float* bias, const float* residual, int64_t ld_residual, float alpha,
                            const float* gamma, const float* beta, float eps, int32_t normalize,
                            int64_t rows, int64_t dim, float* out, float* out_lo, int64_t ld_out, void* stream);

/* Depthwise 1-D convolution over time: out[n, t, d] = epi(bias[d] + sum_k w[k, d] * x[n, t - left_pad
 * + k*dilation, d]) (zeros outside [0, T)); `weight_kd` is [kernel, channels] (tap-major).
 * Replaces the grouped nn.Conv1d (+ eval BatchNorm1d folded by the caller + activation) of
 * aps/asr/transformer/impl.py:456-465 and aps/sse/bss/tcn.py:141-151.                         */
int aps_b200_dwconv1d_fwd(const float* x, int64_t ld_x, int64_t batch, int64_t num_frames,
                          int64_t channels, int64_t stride_n, int64_t stride_t,
                          const float* weight_kd, const float* bias, int kernel, int dilation,
                          int left_pad, const aps_b200_epilogue* epi, float* out, int64_t ld_out,
                          void* stream);
/* ... with the TF32 lo companion of the output (out_lo, optional: channels % 4 == 0, 16-byte aligned rows) and optional
 * per-utterance lengths (device int64 [batch]): input frames t >= lens[n] read as zero, which is what an utterance sees
 * alone (ragged batched decoding, aps/asr/ctc.py:58-84).                                                          */
int aps_b200_dwconv1d2_fwd(const float* x, int64_t ld_x, int64_t batch, int64_t num_frames,
                           int64_t channels, int64_t stride_n, int64_t stride_t,
                           const float* weight_kd, const float* bias, int kernel, int dilation,
                           int left_pad, const aps_b200_epilogue* epi, float* out, float* out_lo,
                           int64_t ld_out, const int64_t* lens, void* stream);

/* Multi-head self-attention, softmax((q.k + pos_term) * scale [masked]) . v per (batch, head).
 * mode 0: no position term (aps/asr/transformer/impl.py:120-131 / torch MHA);
 * mode 1: "rel"  pos [2L-1, head_dim], term[l, s] = qpos[l] . pos[s - l + L - 1] (impl.py:240-261 with
 *         the digit_shift skew of aps/asr/transformer/utils.py:14-39 done by indexing);
 * mode 2: "xl"   pos [2L-1, heads*head_dim] (already projected), content query = qpos + rel_u,
 *         position query = qpos + rel_v (impl.py:324-344; the reference passes VALUE as qpos,
 *         impl.py:369).
 * key_padding_mask [batch, L] (1 = masked -> logit := padding_fill), attn_mask [L, L] additive.
 * Softmax, masking and context follow impl.py:95-118.  head_dim in {32, 64}.                   */
typedef struct aps_b200_attn_desc {
    const float* q; const float* k; const float* v; const float* qpos;
    int64_t ld_q, ld_k, ld_v, ld_qpos;
    int64_t stride_n, stride_t;
    int64_t batch, length, heads, head_dim;
    int32_t mode;
    const float* pos; int64_t ld_pos;
    const float* rel_u; const float* rel_v;
    const uint8_t* key_padding_mask; float padding_fill;
    const float* attn_mask;
    float scale;
} aps_b200_attn_desc;
int aps_b200_mhsa_fwd(const aps_b200_attn_desc* desc, float* out, int64_t ld_out, void* stream);
/* ... and the TF32 lo companion of the context rows */
int aps_b200_mhsa2_fwd(const aps_b200_attn_desc* desc, float* out, float* out_lo, int64_t ld_out, void* stream);

/* Per-utterance normalisation over time of token rows, row(n, t) = n*stride_n + t*stride_t:
 * per_channel = 0: statistics over (channels, frames) of each utterance — nn.GroupNorm(1, C), i.e. "cLN"
 * (aps/sse/bss/tcn.py:81-82), Normalize1d("LN") (aps/asr/base/component.py:95-96, LinearProj default,
 * aps/asr/transformer/proj.py:42) and GlobalChannelLayerNorm "gLN" (tcn.py:33-72);
 * per_channel = 1: statistics over frames of each (utterance, channel) — nn.GroupNorm(C, C), "IN" (tcn.py:83-84).
 * out = (x - mean) / sqrt(var + eps) * gamma[c] + beta[c] (biased variance; gamma/beta may be NULL), then
 * ReLU when relu != 0 (proj.py:54).  In place (out == x) is allowed.                                    */
int64_t aps_b200_utt_norm_workspace_bytes(int64_t batch, int64_t num_frames, int64_t channels);
int aps_b200_utt_norm_fwd(const float* x, int64_t ld_x, int64_t batch, int64_t num_frames, int64_t channels,
                          int64_t stride_n, int64_t stride_t, int per_channel, const float* gamma,
                          const float* beta, float eps, int relu, void* workspace, int64_t workspace_bytes,
                          float* out, int64_t ld_out, void* stream);

/* Transposed convolution with at most 8 output channels (the last DCCRN decoder layer, dccrn.py:133-147:
 * 2 * num_spks channels) on a dedicated kernel: 8 lanes per output pixel, weights in shared memory, exact fp32.
 * x_skip (same shape as x, may be NULL): the input is cat_complex(x, x_skip) read in place; in_channels counts
 * both.  Channels per tensor % 4 == 0 (% 8 with x_skip); kernel_h * kernel_w * in_channels * 8 floats <= 48 KB. */
int aps_b200_conv_transpose2d_nhwc_narrow_fwd(const float* x, const float* x_skip, int64_t batch, int64_t height,
                                              int64_t width, int64_t in_channels, const float* weight,
                                              int64_t out_channels, int kernel_h, int kernel_w, int stride_h,
                                              int stride_w, int pad_h, int pad_w, int out_pad_h, int out_pad_w,
                                              const aps_b200_epilogue* epi, float* out, void* stream);

/* LSTM recurrence of one layer and direction ---------------------------------------------------
 * torch.nn.LSTM semantics (gate order i, f, g, o; h_0 = c_0 = 0), as used by the DCCRN bottleneck
 * (aps/sse/bss/dccrn.py:20-50: LSTMP -> nn.LSTM, batch_first).  xg [rows, num_frames, ld_xg >= 4H]
 * holds x_t W_ih^T + b_ih + b_hh of every frame (one aps_b200_linear_*_fwd call); this entry adds
 * h_{t-1} W_hh^T (w_hh [4H, H] row-major), applies the cell update and writes h_t to
 * y [rows, num_frames, ld_y >= H] (a bidirectional layer passes ld_y = 2H and y + H for the
 * reverse direction, reverse != 0 walks the frames backwards).  cell: [rows, H] scratch.
 * One fused launch per frame, exact fp32 arithmetic.  hidden % 4 == 0, ld_y % 4 == 0.             */
int aps_b200_lstm_fwd(const float* xg, int64_t ld_xg, int64_t rows, int64_t num_frames, int64_t hidden,
                      const float* w_hh, int reverse, float* cell, float* y, int64_t ld_y, void* stream);
/* `groups` (<= APS_B200_LSTM_MAX_GROUPS) independent recurrences of the same shape advanced together, one launch per
 * frame for all of them: the real and imaginary LSTMs of DCCRN's complex bottleneck (dccrn.py:97-110), or the two
 * directions of a bidirectional layer (bit g of reverse_mask: group g walks the frames backwards).  Pointer arrays
 * are host arrays of device pointers; everything else as above.                                                    */
#define APS_B200_LSTM_MAX_GROUPS 4
int aps_b200_lstm_group_fwd(const float* const* xg, int64_t ld_xg, int64_t rows, int64_t num_frames, int64_t hidden,
                            const float* const* w_hh, int reverse_mask, float* const* cell, float* const* y,
                            int64_t ld_y, int groups, void* stream);
/* The same recurrence (forward direction) with the per-frame product h_{t-1} W_hh^T on the tcgen05 engine (3xTF32
 * split, fp32-level accuracy): per frame ONE grouped GEMM launch — the groups' hidden states are stacked along M, a row
 * block uses its own group's W_hh, the epilogue adds the frame's input projections — and one cell kernel (precise expf /
 * tanhf).  hidden % 32 == 0; rows_pad = rows rounded up to a multiple of 128 (rows of a group beyond `rows` are never
 * read back).  xg: [groups, rows_pad, num_frames, 4 hidden] input projections incl. both biases; w_hi / w_lo:
 * aps_b200_tf32_split of the stacked W_hh [groups * 4 hidden, hidden]; y: host array of `groups` device pointers, each
 * [rows, num_frames, ld_y >= hidden]; work: groups * rows_pad * hidden * 21 floats of scratch (cell state, two hidden
 * (x, x_lo) pairs, up to four K slices of gate pre-activations: the small per-frame GEMM is cut along K while its work
 * items fit one wave, the cell kernel adds the slices and the input projections).  Same reference lines as above.  */
int aps_b200_lstm_group_tc_fwd(const float* xg, int64_t rows, int64_t rows_pad, int64_t num_frames, int64_t hidden,
                               const float* w_hi, const float* w_lo, void* const* y, int64_t ld_y, int32_t groups,
                               float* work, void* stream);

/* Time-domain separation objectives ----------------------------------------------------------
 * Si-SNR / SNR between every estimate and every reference of an utterance in ONE pass over the
 * waveforms (fp64 sums of x, s, x^2, s^2, x.s; closed-form objective).  out[n, e, r] is what
 * sisnr_objf(est[e], ref[r]) / snr_objf(est[e], ref[r]) return for utterance n — the K x K
 * matrix permu_invarint_objf needs — so PIT is a min over K! sums of K entries afterwards.
 * Replaces aps/task/objf.py:133-163 (sisnr_objf), :166-198 (snr_objf) and the K!*K calls of
 * :289-336 (permu_invarint_objf) made by aps/task/sse.py:83-139 (TimeDomainTask / SisnrTask).  */
#define APS_B200_MAX_SIGNALS 4
typedef struct aps_b200_signal_list {
    const float* ptr[APS_B200_MAX_SIGNALS];  /* each [batch, num_samples], row stride ld[k] floats */
    int64_t ld[APS_B200_MAX_SIGNALS];
    int32_t count;
} aps_b200_signal_list;
typedef struct aps_b200_objf_desc {
    int32_t kind;          /* 0: Si-SNR (objf.py:133), 1: SNR (objf.py:166)                        */
    int32_t zero_mean;     /* Si-SNR: remove the per-utterance mean first (objf.py:151-153)        */
    int32_t non_negative;  /* 10*log10(1 + snr^2) instead of 20*log10(eps + snr)                   */
    float   eps;           /* EPSILON of the reference (float32 machine epsilon)                   */
    float   snr_max;       /* SNR only: > 0 enables the thresholded form (objf.py:183-190)         */
} aps_b200_objf_desc;
/* Scratch for the per-CTA partial sums (bytes); 0 for an unsupported shape. */
int64_t aps_b200_pair_objf_workspace_bytes(int64_t batch, int64_t num_samples, int num_signals);
/* est->count == ref->count == K in [1, 4]; out: [batch, K, K] fp32. */
int aps_b200_pair_objf_fwd(const aps_b200_signal_list* est, const aps_b200_signal_list* ref, int64_t batch,
                           int64_t num_samples, const aps_b200_objf_desc* desc, void* workspace,
                           int64_t workspace_bytes, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* APS_B200_H_ */

// gemm.cu — entry points of the fp32 GEMM core: dense layers and NHWC implicit-GEMM convolutions.
//
// Replaces (forward, inference) torch.nn.functional.linear / conv1d(k=1) / conv2d as used by
// /root/reference/aps/asr/transformer/impl.py:388-393,454-475 (feed-forward and point-wise conv layers),
// impl.py:62-83 (attention in/out projections), aps/asr/base/component.py:251-307 (Conv2d block, with the
// eval-mode BatchNorm folded into weight/bias by the caller), aps/asr/base/encoder.py:415-441
// (Conv2dEncoder.outp), aps/sse/bss/tcn.py:112-159 (1x1 convolutions of Conv1dBlock).
#include "../../include/aps_b200.h"
#include "common.cuh"
#include "gemm.cuh"

namespace apsb {

template <typename ALoader>
int launch_gemm(const ALoader& a, const float* Wt, long long ldw, int M, int N, int K, const Epilogue& e,
                cudaStream_t st) {
    const int wvec = (((uintptr_t)Wt & 15) == 0 && (ldw & 3) == 0 && (K & 3) == 0) ? 1 : 0;
    // big tiles when they still fill the machine, else 64x64
    const long long big = (long long)((M + 127) / 128) * ((N + 127) / 128);
    if (big >= 2LL * num_sms()) {
        dim3 grid((N + 127) / 128, (M + 127) / 128);
        gemm_kernel<128, 128, ALoader><<<grid, kGemmThreads, 0, st>>>(a, Wt, ldw, wvec, M, N, K, e);
    } else {
        dim3 grid((N + 63) / 64, (M + 63) / 64);
        gemm_kernel<64, 64, ALoader><<<grid, kGemmThreads, 0, st>>>(a, Wt, ldw, wvec, M, N, K, e);
    }
    APSB_LAUNCH_CHECK();
    return 0;
}

// Direct convolution for very narrow inputs (first layer of the conv2d front: Cin = 1, K = 9; first DCCRN encoder
// layer: stacked re/im, Cin = 2, K = 18): a GEMM would waste most of a 16-deep k-tile; this is a pure bandwidth kernel
// (one output row of Cout floats per position).  Weights are staged k-major in shared memory so channel reads are
// coalesced.
struct NarrowConvParams {
    const float* x;
    const float* w;     // [Cout, KH*KW]
    int H, W, KH, KW, sh, sw, ph, pw, dh, dw, OH, OW, Cout, Cin;
    long long M;
    Epilogue e;
};

// thread = (position, group of 4 output channels), positions walked with 32-bit arithmetic (the first version spent
// its time in emulated 64-bit divisions: 1.07 ms for a layer whose HBM bound is 0.09 ms)
// KH_, KW_, CIN_ > 0: compile-time kernel size (3x3 with 1 or 2 input channels: the two layers that exist in the
// reference models) — the tap loops unroll and the kernel drops from ~380 to ~130 instructions per 4 outputs (it is
// instruction-issue bound, not bandwidth bound); 0: run-time sizes.
template <int KH_, int KW_, int CIN_>
__global__ void __launch_bounds__(256) conv2d_narrow_kernel(const __grid_constant__ NarrowConvParams p) {
    const int KH = KH_ > 0 ? KH_ : p.KH, KW = KW_ > 0 ? KW_ : p.KW, CIN = CIN_ > 0 ? CIN_ : p.Cin;
    extern __shared__ __align__(16) float sw_[];      // [K][Cout4], k = (kh*KW + kw)*Cin + c, rows padded to 4 channels
    const int K = KH * KW * CIN;
    const int cgroups = (p.Cout + 3) / 4, cpad = cgroups * 4;
    for (int i = threadIdx.x; i < K * cpad; i += blockDim.x) {
        const int k = i / cpad, co = i - k * cpad;
        sw_[i] = co < p.Cout ? __ldg(p.w + (long long)co * K + k) : 0.f;
    }
    __syncthreads();
    const unsigned gpb = cgroups < 256 ? (unsigned)cgroups : 256u;        // channel groups handled side by side
    const unsigned ppb = 256u / gpb;                                      // positions per block pass
    const unsigned g0 = threadIdx.x % gpb, pp = threadIdx.x / gpb;
    if (pp >= ppb) return;
    const bool vec = (p.Cout & 3) == 0 && (p.e.ldo & 3) == 0 && ((uintptr_t)p.e.out & 15) == 0;
    for (unsigned m = blockIdx.x * ppb + pp; m < (unsigned)p.M; m += gridDim.x * ppb) {
        const unsigned t = m / (unsigned)p.OW, ow = m - t * (unsigned)p.OW;
        const unsigned nb = t / (unsigned)p.OH, oh = t - nb * (unsigned)p.OH;
        for (unsigned g = g0; g < (unsigned)cgroups; g += gpb) {
            const int co = (int)g * 4;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const float* img = p.x + (long long)nb * p.H * p.W * CIN;
#pragma unroll
            for (int kh = 0; kh < KH; ++kh) {
                const int ih = (int)oh * p.sh - p.ph + kh * p.dh;
                if (ih < 0 || ih >= p.H) continue;
#pragma unroll
                for (int kw = 0; kw < KW; ++kw) {
                    const int iw = (int)ow * p.sw - p.pw + kw * p.dw;
                    if (iw < 0 || iw >= p.W) continue;
                    const float* xp = img + (ih * p.W + iw) * CIN;
#pragma unroll
                    for (int c = 0; c < CIN; ++c) {
                        const float xv = __ldg(xp + c);
                        const float4 wv = *reinterpret_cast<const float4*>(sw_ + ((kh * KW + kw) * CIN + c) * cpad + co);
                        acc[0] = fmaf(xv, wv.x, acc[0]); acc[1] = fmaf(xv, wv.y, acc[1]);
                        acc[2] = fmaf(xv, wv.z, acc[2]); acc[3] = fmaf(xv, wv.w, acc[3]);
                    }
                }
            }
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = co + j;
                float v = acc[j] + ((p.e.bias && n < p.Cout) ? __ldg(p.e.bias + n) : 0.f);
                v = apply_act(v, p.e.act, p.e, n < p.Cout ? n : 0);
                if (p.e.post_scale && n < p.Cout) v = fmaf(v, __ldg(p.e.post_scale + n), __ldg(p.e.post_shift + n));
                o[j] = v * p.e.alpha;
            }
            float* op = p.e.out + (long long)m * p.e.ldo + co;
            if (vec) {
                *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (co + j < p.Cout) op[j] = o[j];
            }
        }
    }
}

static int fill_epilogue(Epilogue& e, const aps_b200_epilogue* d, int N, float* out, int64_t ldo) {
    APSB_CHECK_ARG(d && out, "null pointer argument");
    APSB_CHECK_ARG(d->act >= ACT_NONE && d->act <= ACT_GELU, "unknown activation %d", d->act);
    APSB_CHECK_ARG(d->act != ACT_GLU || (N % 2 == 0), "GLU needs an even number of columns");
    APSB_CHECK_ARG(d->act != ACT_PRELU || d->prelu_slope, "PReLU slope missing");
    e.bias = d->bias; e.act = d->act; e.alpha = d->alpha;
    e.slope = d->prelu_slope; e.slope_stride = d->prelu_per_channel ? 1 : 0; e.leak = d->leaky_slope;
    e.res = d->residual; e.ldres = d->ld_residual; e.beta = d->beta;
    e.post_scale = d->post_scale; e.post_shift = d->post_shift;
    APSB_CHECK_ARG(!d->post_scale == !d->post_shift, "post_scale and post_shift come together");
    APSB_CHECK_ARG(!(d->post_scale && d->act == ACT_GLU), "post affine is not available with GLU");
    e.out = out; e.ldo = ldo;
    const int ncols = d->act == ACT_GLU ? N / 2 : N;
    APSB_CHECK_ARG(ldo >= ncols, "ld_out %lld smaller than %d columns", (long long)ldo, ncols);
    APSB_CHECK_ARG(!d->residual || d->ld_residual >= ncols, "ld_residual too small");
    return 0;
}

}  // namespace apsb

using namespace apsb;

extern "C" int aps_b200_linear_fwd(const float* x, int64_t rows, int64_t in_features, int64_t ld_x,
                                   const float* weight, int64_t ld_w, int64_t out_features,
                                   const aps_b200_epilogue* epi, float* out, int64_t ld_out, void* stream) {
    APSB_CHECK_ARG(x && weight, "null pointer argument");
    APSB_CHECK_ARG(rows > 0 && in_features > 0 && out_features > 0 && ld_x >= in_features && ld_w >= in_features,
                   "bad shape");
    APSB_CHECK_ARG(rows < (1LL << 31) && out_features < (1LL << 31), "shape too large");
    Epilogue e{};
    if (int rc = fill_epilogue(e, epi, (int)out_features, out, ld_out)) return rc;
    PlainA a{x, ld_x, (int)rows, (int)in_features,
             (((uintptr_t)x & 15) == 0 && (ld_x & 3) == 0 && (in_features & 3) == 0) ? 1 : 0};
    return launch_gemm(a, weight, ld_w, (int)rows, (int)out_features, (int)in_features, e, (cudaStream_t)stream);
}

extern "C" int aps_b200_conv2d_nhwc_fwd(const float* x, int64_t batch, int64_t height, int64_t width,
                                        int64_t in_channels, const float* weight, int64_t out_channels,
                                        int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                                        int dil_h, int dil_w, const aps_b200_epilogue* epi, float* out, void* stream) {
    APSB_CHECK_ARG(x && weight, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && height > 0 && width > 0 && in_channels > 0 && out_channels > 0, "bad shape");
    APSB_CHECK_ARG(kernel_h > 0 && kernel_w > 0 && stride_h > 0 && stride_w > 0 && dil_h > 0 && dil_w > 0 &&
                       pad_h >= 0 && pad_w >= 0, "bad convolution geometry");
    const int64_t OH = (height + 2 * pad_h - dil_h * (kernel_h - 1) - 1) / stride_h + 1;
    const int64_t OW = (width + 2 * pad_w - dil_w * (kernel_w - 1) - 1) / stride_w + 1;
    APSB_CHECK_ARG(OH > 0 && OW > 0, "convolution output is empty");
    const int64_t M = batch * OH * OW, K = (int64_t)kernel_h * kernel_w * in_channels;
    APSB_CHECK_ARG(M < (1LL << 31) && K < (1LL << 31), "shape too large");
    Epilogue e{};
    const int ncols = (int)out_channels;
    if (int rc = fill_epilogue(e, epi, ncols, out, epi && epi->act == ACT_GLU ? ncols / 2 : ncols)) return rc;
    if (in_channels <= 4 && K <= 64 && epi->act != ACT_GLU && !epi->residual &&
        (size_t)K * (out_channels + 3) * 4 <= 48 * 1024) {
        NarrowConvParams c{};
        c.x = x; c.w = weight; c.H = (int)height; c.W = (int)width; c.KH = kernel_h; c.KW = kernel_w;
        c.sh = stride_h; c.sw = stride_w; c.ph = pad_h; c.pw = pad_w; c.dh = dil_h; c.dw = dil_w;
        c.OH = (int)OH; c.OW = (int)OW; c.Cout = ncols; c.Cin = (int)in_channels; c.M = M; c.e = e;
        const long long cg = (ncols + 3) / 4, ppb = 256 / (cg < 256 ? cg : 256);
        const long long blocks = (M + ppb - 1) / ppb;
        const unsigned grid = (unsigned)(blocks < (long long)num_sms() * 16 ? blocks : (long long)num_sms() * 16);
        const size_t smem = (size_t)K * cg * 16;
        cudaStream_t st = (cudaStream_t)stream;
        if (kernel_h == 3 && kernel_w == 3 && in_channels == 1) conv2d_narrow_kernel<3, 3, 1><<<grid, 256, smem, st>>>(c);
        else if (kernel_h == 3 && kernel_w == 3 && in_channels == 2) conv2d_narrow_kernel<3, 3, 2><<<grid, 256, smem, st>>>(c);
        else conv2d_narrow_kernel<0, 0, 0><<<grid, 256, smem, st>>>(c);
        APSB_LAUNCH_CHECK();
        return 0;
    }
    ConvA a{};
    a.x = x; a.Nb = (int)batch; a.H = (int)height; a.W = (int)width; a.Cin = (int)in_channels;
    a.KH = kernel_h; a.KW = kernel_w; a.sh = stride_h; a.sw = stride_w; a.ph = pad_h; a.pw = pad_w;
    a.dh = dil_h; a.dw = dil_w; a.OH = (int)OH; a.OW = (int)OW; a.M = (int)M; a.K = (int)K;
    a.vec = (((uintptr_t)x & 15) == 0 && (in_channels & 3) == 0) ? 1 : 0;
    return launch_gemm(a, weight, K, (int)M, ncols, (int)K, e, (cudaStream_t)stream);
}

extern "C" int aps_b200_conv_transpose2d_nhwc_fwd(const float* x, int64_t batch, int64_t height, int64_t width,
                                                  int64_t in_channels, const float* weight, int64_t out_channels,
                                                  int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h,
                                                  int pad_w, int out_pad_h, int out_pad_w,
                                                  const aps_b200_epilogue* epi, float* out, void* stream) {
    APSB_CHECK_ARG(x && weight, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && height > 0 && width > 0 && in_channels > 0 && out_channels > 0, "bad shape");
    APSB_CHECK_ARG(kernel_h > 0 && kernel_w > 0 && stride_h > 0 && stride_w > 0 && pad_h >= 0 && pad_w >= 0 &&
                       out_pad_h >= 0 && out_pad_w >= 0, "bad convolution geometry");
    const int64_t OH = (height - 1) * stride_h - 2 * pad_h + kernel_h + out_pad_h;
    const int64_t OW = (width - 1) * stride_w - 2 * pad_w + kernel_w + out_pad_w;
    APSB_CHECK_ARG(OH > 0 && OW > 0, "transposed convolution output is empty");
    const int64_t M = batch * OH * OW, K = (int64_t)kernel_h * kernel_w * in_channels;
    APSB_CHECK_ARG(M < (1LL << 31) && K < (1LL << 31), "shape too large");
    Epilogue e{};
    const int ncols = (int)out_channels;
    if (int rc = fill_epilogue(e, epi, ncols, out, ncols)) return rc;
    APSB_CHECK_ARG(epi->act != ACT_GLU, "GLU is not available here");
    TConvA a{};
    a.x = x; a.Nb = (int)batch; a.H = (int)height; a.W = (int)width; a.Cin = (int)in_channels;
    a.KH = kernel_h; a.KW = kernel_w; a.sh = stride_h; a.sw = stride_w; a.ph = pad_h; a.pw = pad_w;
    a.OH = (int)OH; a.OW = (int)OW; a.M = (int)M; a.K = (int)K;
    a.vec = (((uintptr_t)x & 15) == 0 && (in_channels & 3) == 0) ? 1 : 0;
    return launch_gemm(a, weight, K, (int)M, ncols, (int)K, e, (cudaStream_t)stream);
}

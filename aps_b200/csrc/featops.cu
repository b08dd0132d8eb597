// featops.cu — the remaining feature-chain tokens of AsrTransform as kernels:
//   specaug_apply   : SpecAugment mask application, multiply or fill with the global mean
//                     (/root/reference/aps/transform/asr.py:656-684; the masks themselves come from the host RNG in the
//                     reference's call order, aps/transform/augment.py:13-82)
//   splice          : context splicing with edge clamping + frame subsampling (asr.py:687-728, utils.py:193-224)
//   delta           : delta / delta-delta features (asr.py:731-781)
//   speed_perturb   : per-utterance polyphase resampling (asr.py:116-195, augment.py:85-109, utils.py:159-190)
// All are HBM-bound element-wise / gather kernels: coalesced over the feature (or sample) axis, 32-bit index arithmetic
// inside a row, one pass over the data.  The fused fbank kernel applies a zero-fill SpecAugment mask in its own epilogue
// (frontend.cu); these entry points serve every other chain.
#include "../../include/aps_b200.h"
#include "common.cuh"

namespace apsb {

// ---------------------------------------------------------------------------------------------- SpecAugment
// deterministic two-stage sum (fp64 partials, fixed order) for the global mean of mask_zero = False
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ x, long long n,
                                                           double* __restrict__ partials) {
    __shared__ double red[8];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) s += (double)x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        partials[blockIdx.x] = t;
    }
}

// x [N, C, T*F], mask [N, T*F] (0 / 1): out = x * mask, or x where mask != 0 else mean(x)
__global__ void __launch_bounds__(256) specaug_apply_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                                            long long total, unsigned tf, unsigned chans, int mask_zero,
                                                            const double* __restrict__ partials, int nparts,
                                                            float* __restrict__ out) {
    float fill = 0.f;
    if (!mask_zero) {
        double t = 0.0;
        for (int i = 0; i < nparts; ++i) t += partials[i];       // every thread: same order, same value
        fill = (float)(t / (double)total);
    }
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long nc = i / tf;                               // (n, c) plane
        const unsigned r = (unsigned)(i - nc * tf);
        const long long n = nc / chans;
        const float m = __ldg(mask + n * tf + r);
        const float v = x[i];
        out[i] = mask_zero ? v * m : (m == 0.f ? fill : v);
    }
}

// ---------------------------------------------------------------------------------------------- splice / delta
// out[row, to, j*F + f] = x[row, clamp(to*sub + j - lctx, 0, T-1), f],  j = 0 .. lctx + rctx
__global__ void __launch_bounds__(256) splice_kernel(const float* __restrict__ x, long long rows, unsigned T, unsigned F,
                                                     int lctx, int nctx, unsigned sub, unsigned To,
                                                     float* __restrict__ out) {
    const unsigned width = (unsigned)nctx * F;
    const long long total = rows * To * (long long)width;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long rt = i / width;
        const unsigned col = (unsigned)(i - rt * width);
        const long long row = rt / To;
        const unsigned to = (unsigned)(rt - row * To);
        const unsigned j = col / F, f = col - j * F;
        int t = (int)(to * sub) + (int)j - lctx;
        t = t < 0 ? 0 : (t > (int)T - 1 ? (int)T - 1 : t);
        out[i] = __ldg(x + (row * T + t) * F + f);
    }
}

// one delta order: out[row, t, f] = sum_c scale[c] * in[row, clamp(t + c - ctx), f]; in / out addressed with
// (row stride, frame stride) so the slots of a concatenated / stacked result are written in place
__global__ void __launch_bounds__(256) delta_kernel(const float* __restrict__ in, long long in_rs, long long in_ts,
                                                    long long rows, unsigned T, unsigned F, int ctx,
                                                    const float* __restrict__ scale, float* __restrict__ out,
                                                    long long out_rs, long long out_ts) {
    const long long total = rows * T * (long long)F;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long rt = i / F;
        const unsigned f = (unsigned)(i - rt * F);
        const long long row = rt / T;
        const int t = (int)(rt - row * T);
        float acc = 0.f;
        for (int c = -ctx; c <= ctx; ++c) {
            int tt = t + c;
            tt = tt < 0 ? 0 : (tt > (int)T - 1 ? (int)T - 1 : tt);
            // the reference multiplies then sums over the context axis in this order (th.sum(splice * scale, -1))
            acc += __ldg(in + row * in_rs + tt * in_ts + f) * __ldg(scale + c + ctx);
        }
        out[row * out_rs + t * out_ts + f] = acc;
    }
}

// ---------------------------------------------------------------------------------------------- speed perturb
struct PerturbParams {
    const float* wav;        // [N, S]
    long long ld;
    int N;
    long long S;
    const int* choice;       // [N]: filter index, or num_filters = keep the utterance
    int num_filters;
    const float* weight[4];  // [dst, src, K] each
    int dst[4], src[4], K[4];
    float* out;              // [N, ld_out], zero padded
    long long ld_out;
};

// out[n, b*dst + p] = sum_{q < src} sum_{k < K} w[p, q, k] * x[n, (b + k - (K-1)/2)*src + q]   (blocks outside [0, B): 0)
// — tf.conv1d over blocks of src samples with padding (K-1)/2, augment.py:100-109.  grid (chunks, N).
__global__ void __launch_bounds__(256) speed_perturb_kernel(const __grid_constant__ PerturbParams p) {
    extern __shared__ float sw[];                  // the filter of this utterance's choice
    const int n = blockIdx.y;
    const int c = p.choice[n];
    const float* x = p.wav + (long long)n * p.ld;
    float* o = p.out + (long long)n * p.ld_out;
    if (c >= p.num_filters) {                      // factor 1.0: copy, zero the padding
        for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.ld_out; i += (long long)gridDim.x * 256)
            o[i] = i < p.S ? x[i] : 0.f;
        return;
    }
    const int dst = p.dst[c], src = p.src[c], K = p.K[c], pad = (K - 1) / 2;
    for (int i = threadIdx.x; i < dst * src * K; i += 256) sw[i] = __ldg(p.weight[c] + i);
    __syncthreads();
    const long long blocks = p.S / src;
    const long long len = blocks * dst;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.ld_out; i += (long long)gridDim.x * 256) {
        float acc = 0.f;
        if (i < len) {
            const long long b = i / dst;
            const int ph = (int)(i - b * dst);
            const float* w = sw + ph * src * K;
            for (int k = 0; k < K; ++k) {
                const long long bb = b + k - pad;
                if (bb < 0 || bb >= blocks) continue;
                const float* xb = x + bb * src;
                for (int q = 0; q < src; ++q) acc = fmaf(w[q * K + k], __ldg(xb + q), acc);
            }
        }
        o[i] = acc;
    }
}

static unsigned grid_for(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace apsb

using namespace apsb;

extern "C" int64_t aps_b200_specaug_workspace_bytes(int64_t numel) {
    return numel > 0 ? (int64_t)sizeof(double) * 1024 : 0;
}

extern "C" int aps_b200_specaug_apply(const float* x, int64_t batch, int64_t channels, int64_t frames, int64_t dims,
                                      const float* mask, int32_t mask_zero, void* workspace, int64_t workspace_bytes,
                                      float* out, void* stream) {
    APSB_CHECK_ARG(x && mask && out, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && channels > 0 && frames > 0 && dims > 0, "bad shape");
    APSB_CHECK_ARG(frames * dims < (1LL << 31), "feature plane too large");
    const long long total = batch * channels * frames * dims;
    cudaStream_t st = (cudaStream_t)stream;
    int nparts = 0;
    double* parts = static_cast<double*>(workspace);
    if (!mask_zero) {
        APSB_CHECK_ARG(parts && workspace_bytes >= aps_b200_specaug_workspace_bytes(total), "specaug: workspace too small");
        nparts = (int)grid_for(total);
        if (nparts > 1024) nparts = 1024;
        sum_partials_kernel<<<nparts, 256, 0, st>>>(x, total, parts);
        APSB_LAUNCH_CHECK();
    }
    specaug_apply_kernel<<<grid_for(total), 256, 0, st>>>(x, mask, total, (unsigned)(frames * dims), (unsigned)channels,
                                                          mask_zero, parts, nparts, out);
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_splice_fwd(const float* x, int64_t rows, int64_t frames, int64_t dims, int32_t lctx, int32_t rctx,
                                   int32_t subsampling, float* out, void* stream) {
    APSB_CHECK_ARG(x && out && rows > 0 && frames > 0 && dims > 0, "bad arguments");
    APSB_CHECK_ARG(lctx >= 0 && rctx >= 0 && subsampling >= 1, "bad context / subsampling");
    APSB_CHECK_ARG(frames < (1LL << 30) && dims * (lctx + rctx + 1) < (1LL << 30), "shape too large");
    const long long To = subsampling == 1 ? frames : frames / subsampling;
    if (To == 0) return 0;
    const long long total = rows * To * dims * (lctx + rctx + 1);
    splice_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, rows, (unsigned)frames, (unsigned)dims, lctx,
                                                                     lctx + rctx + 1, (unsigned)subsampling, (unsigned)To, out);
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_delta_fwd(const float* in, int64_t in_row_stride, int64_t in_frame_stride, int64_t rows,
                                  int64_t frames, int64_t dims, int32_t ctx, const float* scale, float* out,
                                  int64_t out_row_stride, int64_t out_frame_stride, void* stream) {
    APSB_CHECK_ARG(in && out && scale && rows > 0 && frames > 0 && dims > 0 && ctx >= 0, "bad arguments");
    APSB_CHECK_ARG(frames < (1LL << 30) && dims < (1LL << 30), "shape too large");
    const long long total = rows * frames * dims;
    delta_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(in, in_row_stride, in_frame_stride, rows, (unsigned)frames,
                                                                    (unsigned)dims, ctx, scale, out, out_row_stride,
                                                                    out_frame_stride);
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_speed_perturb_fwd(const float* wav, int64_t batch, int64_t num_samples, int64_t ld_wav,
                                          const int32_t* choice, int32_t num_filters, const float* const* weights,
                                          const int32_t* dst_sr, const int32_t* src_sr, const int32_t* taps, float* out,
                                          int64_t ld_out, void* stream) {
    APSB_CHECK_ARG(wav && choice && out && batch > 0 && num_samples > 0 && ld_out > 0, "bad arguments");
    APSB_CHECK_ARG(num_filters >= 0 && num_filters <= 4, "speed perturb: at most 4 filters (got %d)", num_filters);
    APSB_CHECK_ARG(batch <= 65535, "speed perturb: batch too large");
    PerturbParams p{};
    p.wav = wav; p.ld = ld_wav; p.N = (int)batch; p.S = num_samples; p.choice = choice; p.num_filters = num_filters;
    p.out = out; p.ld_out = ld_out;
    size_t smem = 0;
    for (int i = 0; i < num_filters; ++i) {
        APSB_CHECK_ARG(weights && weights[i] && dst_sr[i] > 0 && src_sr[i] > 0 && taps[i] > 0, "speed perturb: bad filter %d", i);
        p.weight[i] = weights[i]; p.dst[i] = dst_sr[i]; p.src[i] = src_sr[i]; p.K[i] = taps[i];
        const size_t b = (size_t)dst_sr[i] * src_sr[i] * taps[i] * 4;
        smem = b > smem ? b : smem;
    }
    APSB_CHECK_ARG(smem <= 200 * 1024, "speed perturb: filter of %zu bytes does not fit shared memory", smem);
    static LaunchCache slots[64];
    LaunchCache& lc = launch_cache(slots);
    if ((int)smem > lc.smem_set) {
        APSB_CUDA(cudaFuncSetAttribute(speed_perturb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lc.smem_set = (int)smem;
    }
    long long chunks = (ld_out + 256 * 8 - 1) / (256 * 8);
    if (chunks > 1024) chunks = 1024;
    dim3 grid((unsigned)chunks, (unsigned)batch);
    speed_perturb_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    APSB_LAUNCH_CHECK();
    return 0;
}

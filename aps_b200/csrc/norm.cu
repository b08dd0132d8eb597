// norm.cu — per-utterance normalisation over time for batch-major token rows [N*T, C].
//
//   group mode       : statistics over (C, T) of every utterance  -> nn.GroupNorm(1, C) ("cLN" of
//                      aps/sse/bss/tcn.py:81-82, "LN" of aps/asr/base/component.py:95-96) and
//                      GlobalChannelLayerNorm (tcn.py:33-72, "gLN")
//   per-channel mode : statistics over T of every (utterance, channel) -> nn.GroupNorm(C, C) ("IN", tcn.py:83-84)
// y = (x - mean) / sqrt(var + eps) * gamma[c] + beta[c]  (biased variance), optional ReLU (LinearProj, proj.py:54).
// Two launches: per-slab fp64 partial sums (deterministic, no atomics), then the apply pass, which re-reads
// x (L2 resident at these sizes).  HBM bound: 2 reads + 1 write of the activation.
#include "../../include/aps_b200.h"
#include "common.cuh"

namespace apsb {

constexpr int kNormThreads = 256;

struct UttNormParams {
    const float* x;
    long long ldx, sn, st;
    int N, T, C, slabs, frames_per_slab, per_channel, relu;
    const float* gamma;
    const float* beta;
    float eps;
    double* ws;          // [N, slabs, C, 2]
    float* out;
    long long ldo;
};

// grid (slabs, N): thread owns channels tid, tid + 256, ... and walks the slab's frames (coalesced over channels)
__global__ void __launch_bounds__(kNormThreads) utt_norm_stats_kernel(const __grid_constant__ UttNormParams p) {
    const int n = blockIdx.y, slab = blockIdx.x;
    const int t0 = slab * p.frames_per_slab, t1 = min(p.T, t0 + p.frames_per_slab);
    for (int c = threadIdx.x; c < p.C; c += kNormThreads) {
        double s = 0.0, q = 0.0;
        for (int t = t0; t < t1; ++t) {
            const double v = __ldg(p.x + ((long long)n * p.sn + (long long)t * p.st) * p.ldx + c);
            s += v;
            q = fma(v, v, q);
        }
        double* w = p.ws + (((long long)n * p.slabs + slab) * p.C + c) * 2;
        w[0] = s;
        w[1] = q;
    }
}

__global__ void __launch_bounds__(kNormThreads) utt_norm_apply_kernel(const __grid_constant__ UttNormParams p) {
    const int n = blockIdx.y, slab = blockIdx.x;
    const int t0 = slab * p.frames_per_slab, t1 = min(p.T, t0 + p.frames_per_slab);
    __shared__ double red[2][kNormThreads / 32];
    __shared__ float g_mean, g_rstd;
    const double* base = p.ws + (long long)n * p.slabs * p.C * 2;
    if (!p.per_channel) {
        double s = 0.0, q = 0.0;
        for (int i = threadIdx.x; i < p.slabs * p.C; i += kNormThreads) {
            s += base[2 * i];
            q += base[2 * i + 1];
        }
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if ((threadIdx.x & 31) == 0) {
            red[0][threadIdx.x >> 5] = s;
            red[1][threadIdx.x >> 5] = q;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double ts = 0.0, tq = 0.0;
            for (int w = 0; w < kNormThreads / 32; ++w) {
                ts += red[0][w];
                tq += red[1][w];
            }
            const double cnt = (double)p.T * (double)p.C;
            const double mean = ts / cnt;
            const double var = fmax(tq / cnt - mean * mean, 0.0);
            g_mean = (float)mean;
            g_rstd = (float)(1.0 / sqrt(var + (double)p.eps));
        }
        __syncthreads();
    }
    for (int c = threadIdx.x; c < p.C; c += kNormThreads) {
        float mean, rstd;
        if (p.per_channel) {
            double s = 0.0, q = 0.0;
            for (int sl = 0; sl < p.slabs; ++sl) {
                s += base[((long long)sl * p.C + c) * 2];
                q += base[((long long)sl * p.C + c) * 2 + 1];
            }
            const double m = s / (double)p.T;
            const double var = fmax(q / (double)p.T - m * m, 0.0);
            mean = (float)m;
            rstd = (float)(1.0 / sqrt(var + (double)p.eps));
        } else {
            mean = g_mean;
            rstd = g_rstd;
        }
        const float g = p.gamma ? __ldg(p.gamma + c) : 1.f;
        const float b = p.beta ? __ldg(p.beta + c) : 0.f;
        const float a = rstd * g, sh = fmaf(-mean, a, b);
        for (int t = t0; t < t1; ++t) {
            const long long r = (long long)n * p.sn + (long long)t * p.st;
            float y = fmaf(__ldg(p.x + r * p.ldx + c), a, sh);
            if (p.relu) y = fmaxf(y, 0.f);
            p.out[r * p.ldo + c] = y;
        }
    }
}

static int norm_slabs(int64_t batch, int64_t frames) {
    int64_t want = (2LL * num_sms() + batch - 1) / batch;
    if (want > frames) want = frames;
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    return (int)want;
}

}  // namespace apsb

using namespace apsb;

extern "C" int64_t aps_b200_utt_norm_workspace_bytes(int64_t batch, int64_t num_frames, int64_t channels) {
    if (batch <= 0 || num_frames <= 0 || channels <= 0) return 0;
    return batch * norm_slabs(batch, num_frames) * channels * 2 * (int64_t)sizeof(double);
}

extern "C" int aps_b200_utt_norm_fwd(const float* x, int64_t ld_x, int64_t batch, int64_t num_frames, int64_t channels,
                                     int64_t stride_n, int64_t stride_t, int per_channel, const float* gamma,
                                     const float* beta, float eps, int relu, void* workspace, int64_t workspace_bytes,
                                     float* out, int64_t ld_out, void* stream) {
    APSB_CHECK_ARG(x && out && workspace, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && batch <= 65535 && num_frames > 0 && num_frames < (1LL << 31) && channels > 0 &&
                       channels < (1LL << 31) && ld_x >= channels && ld_out >= channels, "bad shape");
    APSB_CHECK_ARG(workspace_bytes >= aps_b200_utt_norm_workspace_bytes(batch, num_frames, channels),
                   "utt_norm: workspace too small (%lld bytes)", (long long)workspace_bytes);
    UttNormParams p{};
    p.x = x; p.ldx = ld_x; p.sn = stride_n; p.st = stride_t;
    p.N = (int)batch; p.T = (int)num_frames; p.C = (int)channels;
    p.slabs = norm_slabs(batch, num_frames);
    p.frames_per_slab = (p.T + p.slabs - 1) / p.slabs;
    p.per_channel = per_channel; p.relu = relu; p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.ws = static_cast<double*>(workspace); p.out = out; p.ldo = ld_out;
    dim3 grid(p.slabs, p.N);
    cudaStream_t st = (cudaStream_t)stream;
    utt_norm_stats_kernel<<<grid, kNormThreads, 0, st>>>(p);
    APSB_LAUNCH_CHECK();
    utt_norm_apply_kernel<<<grid, kNormThreads, 0, st>>>(p);
    APSB_LAUNCH_CHECK();
    return 0;
}

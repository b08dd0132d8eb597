// objf.cu — time-domain separation objectives (Si-SNR / SNR) as ONE pass over the waveforms.
//
// Replaces aps/task/objf.py:133-163 (sisnr_objf), :166-198 (snr_objf) and the K x K pair matrix that
// permu_invarint_objf (:289-336) builds by calling them K! * K times: every estimate and reference
// is read exactly once; the K + K + K + K + K*K sums (sum x, sum s, sum x^2, sum s^2, sum x.s) are
// accumulated in fp64, and the objective follows in closed form (fp64, so the cancellations in
// ||x - t||^2 = ||x||^2 - 2a<x,s> + a^2||s||^2 stay far below the fp32 noise of the reference).
// HBM bound: (E + R) * S * 4 bytes per utterance.
#include "../../include/aps_b200.h"
#include "common.cuh"

namespace apsb {

constexpr int kObjfThreads = 256;

template <int K> struct PairSums {
    double sx[K], ss[K], sxx[K], sss[K], sxs[K * K];
    static constexpr int kCount = 4 * K + K * K;
};

template <int K> __device__ __forceinline__ void accumulate(PairSums<K>& a, const float (&x)[K], const float (&s)[K]) {
#pragma unroll
    for (int e = 0; e < K; ++e) {
        const double xe = x[e];
        a.sx[e] += xe;
        a.sxx[e] = fma(xe, xe, a.sxx[e]);
#pragma unroll
        for (int r = 0; r < K; ++r) a.sxs[e * K + r] = fma(xe, (double)s[r], a.sxs[e * K + r]);
    }
#pragma unroll
    for (int r = 0; r < K; ++r) {
        const double sr = s[r];
        a.ss[r] += sr;
        a.sss[r] = fma(sr, sr, a.sss[r]);
    }
}

// grid (chunks * batch): CTA (c, n) = (blockIdx.x % chunks, blockIdx.x / chunks) reduces samples [c*span, (c+1)*span) of utterance n and writes its
// PairSums to partials[n][c][:] (fixed order -> deterministic finalisation, no atomics).
template <int K, bool VEC>
__global__ void __launch_bounds__(kObjfThreads) pair_sums_kernel(aps_b200_signal_list est, aps_b200_signal_list ref,
                                                               int64_t num_samples, int64_t span, int chunks,
                                                               double* __restrict__ partials) {
    const int64_t n = blockIdx.x / (unsigned)chunks;          // batch folded into gridDim.x: no 65535 limit on N
    const int64_t chunk = blockIdx.x - (unsigned)n * (unsigned)chunks;
    const int64_t begin = chunk * span;
    const int64_t end = min(begin + span, num_samples);
    const float* xp[K];
    const float* sp[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        xp[k] = est.ptr[k] + n * est.ld[k];
        sp[k] = ref.ptr[k] + n * ref.ld[k];
    }
    PairSums<K> acc;
    double* flat = reinterpret_cast<double*>(&acc);
#pragma unroll
    for (int i = 0; i < PairSums<K>::kCount; ++i) flat[i] = 0.0;

    if (VEC) {   // span, pointers and leading dimensions are multiples of 4 floats
        for (int64_t i = begin + 4 * threadIdx.x; i < end; i += 4 * kObjfThreads) {
            float4 xv[K], sv[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                xv[k] = __ldg(reinterpret_cast<const float4*>(xp[k] + i));
                sv[k] = __ldg(reinterpret_cast<const float4*>(sp[k] + i));
            }
            float x[K], s[K];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (i + j < end) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        x[k] = reinterpret_cast<const float*>(&xv[k])[j];
                        s[k] = reinterpret_cast<const float*>(&sv[k])[j];
                    }
                    accumulate<K>(acc, x, s);
                }
            }
        }
    } else {
        for (int64_t i = begin + threadIdx.x; i < end; i += kObjfThreads) {
            float x[K], s[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                x[k] = __ldg(xp[k] + i);
                s[k] = __ldg(sp[k] + i);
            }
            accumulate<K>(acc, x, s);
        }
    }

    __shared__ double red[kObjfThreads / 32][PairSums<K>::kCount];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < PairSums<K>::kCount; ++i) {
        double v = flat[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < PairSums<K>::kCount) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kObjfThreads / 32; ++w) v += red[w][threadIdx.x];
        partials[(n * chunks + chunk) * PairSums<K>::kCount + threadIdx.x] = v;
    }
}

// one thread per (utterance, estimate, reference) pair
template <int K>
__global__ void pair_objf_kernel(const double* __restrict__ partials, int64_t batch, int chunks, int64_t num_samples,
                                 aps_b200_objf_desc d, float* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= batch * K * K) return;
    const int64_t n = idx / (K * K);
    const int e = (int)(idx % (K * K)) / K, r = (int)(idx % (K * K)) % K;
    constexpr int C = PairSums<K>::kCount;
    double sx = 0, ss = 0, sxx = 0, sss = 0, sxs = 0;
    for (int c = 0; c < chunks; ++c) {
        const double* p = partials + (n * chunks + c) * C;
        sx += p[e];
        ss += p[K + r];
        sxx += p[2 * K + e];
        sss += p[3 * K + r];
        sxs += p[4 * K + e * K + r];
    }
    const double eps = d.eps, len = (double)num_samples;
    double res;
    if (d.kind == 0) {   // Si-SNR, objf.py:151-163
        if (d.zero_mean) {
            sxs -= sx * ss / len;
            sxx -= sx * sx / len;
            sss -= ss * ss / len;
        }
        sxx = fmax(sxx, 0.0);
        sss = fmax(sss, 0.0);
        const double a = sxs / (sss + eps);                       // t = a * s
        const double t_norm = fabs(a) * sqrt(sss);
        const double n_norm = sqrt(fmax(sxx - 2.0 * a * sxs + a * a * sss, 0.0));
        const double snr = t_norm / (n_norm + eps);
        res = d.non_negative ? 10.0 * log10(1.0 + snr * snr) : 20.0 * log10(eps + snr);
    } else {             // SNR, objf.py:183-198
        const double diff = fmax(sxx - 2.0 * sxs + sss, 0.0);     // ||x - s||^2
        if (d.snr_max > 0.f) {
            const double threshold = pow(10.0, -(double)d.snr_max / 10.0);
            res = 10.0 * log10(sss + eps) - 10.0 * log10(threshold * sss + diff + eps);
        } else {
            const double snr = sqrt(sss) / (sqrt(diff) + eps);
            res = d.non_negative ? 10.0 * log10(1.0 + snr * snr) : 20.0 * log10(eps + snr);
        }
    }
    out[idx] = (float)res;
}

static int objf_chunks(int64_t batch, int64_t num_samples) {
    int64_t want = (4LL * num_sms() + batch - 1) / batch;          // ~4 CTAs per SM in total
    int64_t most = (num_samples + 4 * kObjfThreads - 1) / (4 * kObjfThreads);
    if (want > most) want = most;
    if (want < 1) want = 1;
    if (want > 1024) want = 1024;
    return (int)want;
}

template <int K>
static int launch_objf(const aps_b200_signal_list& est, const aps_b200_signal_list& ref, int64_t batch,
                       int64_t num_samples, const aps_b200_objf_desc& d, double* ws, float* out, cudaStream_t st) {
    const int chunks = objf_chunks(batch, num_samples);
    int64_t span = (num_samples + chunks - 1) / chunks;
    span = (span + 3) / 4 * 4;
    bool vec = true;
    for (int k = 0; k < K; ++k) {
        vec = vec && (reinterpret_cast<uintptr_t>(est.ptr[k]) % 16 == 0) && est.ld[k] % 4 == 0;
        vec = vec && (reinterpret_cast<uintptr_t>(ref.ptr[k]) % 16 == 0) && ref.ld[k] % 4 == 0;
    }
    vec = vec && num_samples % 4 == 0;
    const unsigned grid = (unsigned)(batch * chunks);
    if (vec)
        pair_sums_kernel<K, true><<<grid, kObjfThreads, 0, st>>>(est, ref, num_samples, span, chunks, ws);
    else
        pair_sums_kernel<K, false><<<grid, kObjfThreads, 0, st>>>(est, ref, num_samples, span, chunks, ws);
    APSB_LAUNCH_CHECK();
    const int64_t total = batch * K * K;
    pair_objf_kernel<K><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(ws, batch, chunks, num_samples, d, out);
    APSB_LAUNCH_CHECK();
    return 0;
}

}  // namespace apsb

extern "C" int64_t aps_b200_pair_objf_workspace_bytes(int64_t batch, int64_t num_samples, int num_signals) {
    if (batch <= 0 || num_samples <= 0 || num_signals < 1 || num_signals > APS_B200_MAX_SIGNALS) return 0;
    const int64_t count = 4 * num_signals + num_signals * num_signals;
    return batch * apsb::objf_chunks(batch, num_samples) * count * (int64_t)sizeof(double);
}

extern "C" int aps_b200_pair_objf_fwd(const aps_b200_signal_list* est, const aps_b200_signal_list* ref, int64_t batch,
                                      int64_t num_samples, const aps_b200_objf_desc* desc, void* workspace,
                                      int64_t workspace_bytes, float* out, void* stream) {
    APSB_CHECK_ARG(est && ref && desc && out && workspace, "pair_objf: null argument");
    APSB_CHECK_ARG(est->count == ref->count, "pair_objf: %d estimates vs %d references", est->count, ref->count);
    APSB_CHECK_ARG(est->count >= 1 && est->count <= APS_B200_MAX_SIGNALS, "pair_objf: 1..%d signals supported, got %d",
                   APS_B200_MAX_SIGNALS, est->count);
    APSB_CHECK_ARG(batch > 0 && batch < (1LL << 20) && num_samples > 0, "pair_objf: bad shape %lld x %lld", (long long)batch,
                   (long long)num_samples);
    APSB_CHECK_ARG(desc->kind == 0 || desc->kind == 1, "pair_objf: unknown objective %d", desc->kind);
    APSB_CHECK_ARG(workspace_bytes >= aps_b200_pair_objf_workspace_bytes(batch, num_samples, est->count),
                   "pair_objf: workspace too small (%lld bytes)", (long long)workspace_bytes);
    for (int k = 0; k < est->count; ++k)
        APSB_CHECK_ARG(est->ptr[k] && ref->ptr[k], "pair_objf: null signal %d", k);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* ws = static_cast<double*>(workspace);
    switch (est->count) {
        case 1: return apsb::launch_objf<1>(*est, *ref, batch, num_samples, *desc, ws, out, st);
        case 2: return apsb::launch_objf<2>(*est, *ref, batch, num_samples, *desc, ws, out, st);
        case 3: return apsb::launch_objf<3>(*est, *ref, batch, num_samples, *desc, ws, out, st);
        default: return apsb::launch_objf<4>(*est, *ref, batch, num_samples, *desc, ws, out, st);
    }
}

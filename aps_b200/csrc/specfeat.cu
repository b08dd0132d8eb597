// specfeat.cu — features from an already computed packed STFT (the EnhTransform.forward path):
//   F1b: reference channel -> |X| -> (^2) -> [mel] -> [log] -> [cmvn], with the F/T transpose fused
//   IPD: cos / sin inter-channel phase differences written next to them in the same output row.
//
// Replaces /root/reference/aps/transform/enh.py:39-49 (RefChannelTransform), asr.py:296-303
// (Magnitude), :216-223 (TFTranspose), :350-357 (Power), :416-428 (Mel), :453-464 (Log),
// :576-618 (Cmvn) as chained by enh.py:518-529 / :595-613, and enh.py:67-76 + :112-143
// (PhaseTransform + IpdTransform).
#include "../../include/aps_b200.h"
#include "common.cuh"
#include "feat_epilogue.cuh"

namespace apsb {

constexpr int kSThreads = 256;
constexpr int kSG = 16;                      // lanes per frame
constexpr int kSTC = kSThreads / kSG;        // frames per chunk (one per group)

struct SpecFeatParams {
    const float* spec;      // [rows, C, F, T, 2] (C = 1 for single channel input)
    long long row_stride;   // floats between rows (C*F*T*2)
    long long ch_offset;    // floats to the reference channel (ref*F*T*2)
    long long rows;
    int F, T;
    float mag_eps;
    FeatParams ft;
    float* out;             // [rows, T, ld_out]
    int ld_out;
    int chunks_per_row;
    long long total_chunks;
};

template <int FI>
__global__ void __launch_bounds__(kSThreads) specfeat_kernel(const __grid_constant__ SpecFeatParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int FS = p.F | 1;  // odd row stride: conflict-free transposed stores
    float* sm_mag = reinterpret_cast<float*>(smem);
    int* sm_mel_i = reinterpret_cast<int*>(sm_mag + kSTC * FS);
    float* sm_mel_w = reinterpret_cast<float*>(sm_mel_i + 2 * p.ft.M + 40);
    const int tid = threadIdx.x, lane = tid & 31, l = tid & (kSG - 1), group = tid / kSG;
    const unsigned mask = 0xffffu << (lane & 16);

    if (p.ft.M > 0) {
        for (int i = tid; i < p.ft.M; i += kSThreads) {
            sm_mel_i[i] = __ldg(p.ft.mel_start + i);
            sm_mel_i[p.ft.M + i] = __ldg(p.ft.mel_len + i);
        }
        if (p.ft.mel_in_smem)
            for (int i = tid; i < p.ft.M * p.ft.mel_stride; i += kSThreads) sm_mel_w[i] = __ldg(p.ft.mel_w + i);
        __syncthreads();
        if (tid < FI) {
            int mx = 0;
            for (int d = kSG * tid; d < min(p.ft.M, kSG * tid + kSG); ++d) mx = max(mx, sm_mel_i[p.ft.M + d]);
            sm_mel_i[2 * p.ft.M + tid] = mx;
        }
    }

    for (long long chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x) {
        const long long row = chunk / p.chunks_per_row;
        const int t0 = (int)(chunk - row * p.chunks_per_row) * kSTC;
        const int nf = min(kSTC, p.T - t0);
        const float2* sp = reinterpret_cast<const float2*>(p.spec + row * p.row_stride + p.ch_offset);
        __syncthreads();
        for (int idx = tid; idx < p.F * kSTC; idx += kSThreads) {
            const int k = idx / kSTC, f = idx - k * kSTC;
            float v = 0.f;
            if (f < nf) {
                const float2 X = __ldg(sp + (long long)k * p.T + t0 + f);
                const float pw = fmaf(X.x, X.x, fmaf(X.y, X.y, p.mag_eps));
                v = (p.ft.power == 2) ? pw : sqrtf(pw);   // power 2: (sqrt(pw))^2 up to 1 ulp
            }
            sm_mag[f * FS + k] = v;
        }
        __syncthreads();
        const int f = group;  // one frame per group; invalid frames run on zeros and skip the store
        float* o = (f < nf) ? p.out + ((long long)row * p.T + (t0 + f)) * p.ld_out : nullptr;
        feature_epilogue<kSG, FI>(p.ft, sm_mag + f * FS, sm_mel_i, sm_mel_w, l, mask, o);
    }
}

// IPD: out[n, t, col0 + m*F + k] = cos|sin(angle(x[n, l_m, k, t]) - angle(x[n, r_m, k, t]))
struct IpdParams {
    const float* spec;  // [N, C, F, T, 2]
    int N, C, F, T, P, with_sin;
    const int* idx_l;
    const int* idx_r;
    float* out;
    int ld_out, col0;
};

__global__ void __launch_bounds__(256) ipd_kernel(const __grid_constant__ IpdParams p) {
    __shared__ float tc[32][33];
    __shared__ float tsn[32][33];
    const int n = blockIdx.z / p.P, m = blockIdx.z - n * p.P;
    const int t0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int cl = __ldg(p.idx_l + m), cr = __ldg(p.idx_r + m);
    const long long FT = (long long)p.F * p.T;
    const float2* xl = reinterpret_cast<const float2*>(p.spec) + ((long long)n * p.C + cl) * FT;
    const float2* xr = reinterpret_cast<const float2*>(p.spec) + ((long long)n * p.C + cr) * FT;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int kk = ty; kk < 32; kk += 8) {
        const int k = k0 + kk, t = t0 + tx;
        float c = 0.f, s = 0.f;
        if (k < p.F && t < p.T) {
            const float2 a = __ldg(xl + (long long)k * p.T + t), b = __ldg(xr + (long long)k * p.T + t);
            const float d = atan2f(a.y, a.x) - atan2f(b.y, b.x);
            c = cosf(d);
            s = sinf(d);
        }
        tc[kk][tx] = c;
        tsn[kk][tx] = s;
    }
    __syncthreads();
    for (int tt = ty; tt < 32; tt += 8) {
        const int t = t0 + tt, k = k0 + tx;
        if (t < p.T && k < p.F) {
            float* o = p.out + ((long long)n * p.T + t) * p.ld_out + p.col0;
            o[(long long)m * p.F + k] = tc[tx][tt];
            if (p.with_sin) o[(long long)(p.P + m) * p.F + k] = tsn[tx][tt];
        }
    }
}

// complex ratio mask (dccrn.py:217-232): m = (mr, mi) -> |m| -> g(|m|) m/|m|, optionally applied to the STFT
__global__ void __launch_bounds__(256) cmask_kernel(const float* __restrict__ m, long long ldm, int col_r, int col_i,
                                                    const float* __restrict__ stft, long long total, int act,
                                                    float eps, int apply, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float mr = __ldg(m + i * ldm + col_r), mi = __ldg(m + i * ldm + col_i);
    const float a = sqrtf(mr * mr + mi * mi + eps);
    float g = a;
    if (act == 4) g = 1.f / (1.f + expf(-a));
    else if (act == 3) g = tanhf(a);
    else if (act == 1) g = fmaxf(a, 0.f);
    mr = g * mr / a;
    mi = g * mi / a;
    float2 o = make_float2(mr, mi);
    if (apply) {
        const float2 s = __ldg(reinterpret_cast<const float2*>(stft) + i);
        o = make_float2(s.x * mr - s.y * mi, s.x * mi + s.y * mr);
    }
    reinterpret_cast<float2*>(out)[i] = o;
}

template <int FI>
static int launch_specfeat(SpecFeatParams& p, cudaStream_t st) {
    auto kern = specfeat_kernel<FI>;
    p.ft.mel_in_smem = (p.ft.M * p.ft.mel_stride * 4 <= 48 * 1024) ? 1 : 0;
    const int smem = (kSTC * (p.F | 1) + 2 * p.ft.M + 40 + (p.ft.mel_in_smem ? p.ft.M * p.ft.mel_stride : 0)) * 4 + 16;
    APSB_CHECK_ARG(smem <= 227 * 1024, "specfeat: %d bins need too much shared memory", p.F);
    static LaunchCache slots[64];                             // per instantiation and device
    LaunchCache& lc = launch_cache(slots);
    if (smem > lc.smem_set) {
        APSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        lc.smem_set = smem;
    }
    p.chunks_per_row = (p.T + kSTC - 1) / kSTC;
    p.total_chunks = p.rows * p.chunks_per_row;
    long long grid = (long long)num_sms() * 6;
    if (grid > p.total_chunks) grid = p.total_chunks;
    kern<<<(unsigned)grid, kSThreads, smem, st>>>(p);
    APSB_LAUNCH_CHECK();
    return 0;
}

}  // namespace apsb

using namespace apsb;

extern "C" int aps_b200_spec_feats_fwd(const float* spec, int64_t rows, int64_t channels, int64_t ref_channel,
                                       int64_t num_bins, int64_t num_frames, float mag_eps,
                                       const aps_b200_feat_desc* feat, float* out, int64_t ld_out, void* stream) {
    APSB_CHECK_ARG(spec && feat && out, "null pointer argument");
    APSB_CHECK_ARG(rows > 0 && channels > 0 && num_bins > 1 && num_frames > 0, "bad shape");
    APSB_CHECK_ARG(ref_channel >= 0 && ref_channel < channels, "reference channel %lld out of range",
                   (long long)ref_channel);
    SpecFeatParams p{};
    if (int rc = fill_feat_params(p.ft, feat, (int)num_bins)) return rc;
    APSB_CHECK_ARG(ld_out >= p.ft.D, "ld_out %lld smaller than the feature size %d", (long long)ld_out, p.ft.D);
    p.spec = spec; p.rows = rows; p.F = (int)num_bins; p.T = (int)num_frames;
    p.row_stride = channels * num_bins * num_frames * 2;
    p.ch_offset = ref_channel * num_bins * num_frames * 2;
    p.mag_eps = mag_eps; p.out = out; p.ld_out = (int)ld_out;
    p.ft.out_base = out;
    APSB_CHECK_ARG(!p.ft.aug_mask || ld_out == p.ft.D, "a fused SpecAugment mask needs ld_out == feature size");
    const int fi = (p.ft.D + kSG - 1) / kSG;
    cudaStream_t st = (cudaStream_t)stream;
    if (fi <= 8) return launch_specfeat<8>(p, st);
    if (fi <= 17) return launch_specfeat<17>(p, st);
    if (fi <= 33) return launch_specfeat<33>(p, st);
    return set_error(-1, "feature size %d too large (max 528)", p.ft.D);
}

extern "C" int aps_b200_ipd_fwd(const float* spec, int64_t batch, int64_t channels, int64_t num_bins,
                                int64_t num_frames, const int32_t* index_l, const int32_t* index_r,
                                int64_t num_pairs, int with_sin, float* out, int64_t ld_out, int64_t col0,
                                void* stream) {
    APSB_CHECK_ARG(spec && index_l && index_r && out, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && channels > 1 && num_bins > 0 && num_frames > 0 && num_pairs > 0, "bad shape");
    APSB_CHECK_ARG(batch * num_pairs <= 65535, "batch*pairs %lld exceeds the grid limit",
                   (long long)(batch * num_pairs));
    IpdParams p{};
    p.spec = spec; p.N = (int)batch; p.C = (int)channels; p.F = (int)num_bins; p.T = (int)num_frames;
    p.P = (int)num_pairs; p.with_sin = with_sin; p.idx_l = index_l; p.idx_r = index_r; p.out = out;
    p.ld_out = (int)ld_out; p.col0 = (int)col0;
    dim3 grid((p.T + 31) / 32, (p.F + 31) / 32, (unsigned)(batch * num_pairs));
    ipd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_cmask_fwd(const float* mask, int64_t ld_mask, int64_t col_real, int64_t col_imag,
                                  const float* stft, int64_t positions, int act, float eps, int apply, float* out,
                                  void* stream) {
    APSB_CHECK_ARG(mask && out && positions > 0 && (!apply || stft), "bad arguments");
    APSB_CHECK_ARG(act == 0 || act == 1 || act == 3 || act == 4, "mask non-linearity must be none/relu/tanh/sigmoid");
    cmask_kernel<<<(unsigned)((positions + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        mask, ld_mask, (int)col_real, (int)col_imag, stft, positions, act, eps, apply, out);
    APSB_LAUNCH_CHECK();
    return 0;
}

// mvdr.cu — multi-channel front-end kernels (SURVEY.md rows a14–a18, kernels F4 / F5):
//   mask_colmax   : per (utterance, bin) max over time of the (length-masked) TF mask
//   covar         : masked spatial covariances Rs and Rn for every (utterance, bin), ONE read of X
//   ref_logits    : channel-attention logits from Rs (two tiny GEMVs + tanh)
//   mvdr_weights  : softmax over channels, (Rn + eps I)^-1 Rs, trace, Souden MVDR weights
//   beamform      : Y = sum_c conj(w_c) X_c, second and last read of X
//
// Replaces /root/reference/aps/asr/filter/mvdr.py:103-116 (_process_mask), :42-61 (estimate_covar: four
// real batched GEMMs of 4x249 . 249x4 per covariance, aps/cplx.py:242-252), :158-174
// (ChannelAttention), :75-101 (_derive_weight; the inverse goes through a real 2C x 2C LU,
// aps/cplx.py:268-278), :19-26 (trace), :29-39 (beamform).
//
// Layout: X is given as separate real / imaginary pointers with element strides (they are usually the
// two interleaved halves of a packed STFT [N, C, F, T, 2], i.e. stride_t = 2); a warp owns one
// (utterance, bin) and its lanes stride over time, so reads of X are coalesced along T.
// Algorithmic HBM traffic per utterance (C = 4, T = 249, F = 257): X twice (2 x 4.10 MB), masks 0.26 MB
// (x2 with a noise mask), Y 0.51 MB = 8.96 MB.
#include "../../include/aps_b200.h"
#include "common.cuh"

namespace apsb {

constexpr int kMaxCh = 6;

struct CplxView {
    const float* re;
    const float* im;
    long long sn, sc, sf, st;  // element strides of the [N, C, F, T] view
};

// ------------------------------------------------------------------------------------------------
// max_t |mask[n, t, f]| with frames t >= len[n] treated as 0 (mvdr.py:109-114)
__global__ void __launch_bounds__(256) mask_colmax_kernel(const float* __restrict__ mask, long long sn, long long st,
                                                          long long sf, int N, int T, int F,
                                                          const long long* __restrict__ lens, float* __restrict__ out) {
    const int n = blockIdx.y;
    const int f = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = threadIdx.x >> 5;
    __shared__ float red[8][33];
    const int tmax = lens ? (int)min((long long)T, lens[n]) : T;
    float m = 0.f;
    if (f < F)
        for (int t = ty; t < tmax; t += 8) m = fmaxf(m, fabsf(__ldg(mask + n * sn + t * st + f * sf)));
    red[ty][threadIdx.x & 31] = m;
    __syncthreads();
    if (ty == 0 && f < F) {
        for (int j = 1; j < 8; ++j) m = fmaxf(m, red[j][threadIdx.x]);
        out[(long long)n * F + f] = m;
    }
}

// ------------------------------------------------------------------------------------------------
struct CovarParams {
    CplxView x;
    int N, C, F, T;
    const float* mask_s;
    const float* mask_n;       // nullptr: noise mask = 1 - processed speech mask
    long long msn, mst, msf;   // speech mask strides
    long long mnn, mnt, mnf;   // noise mask strides
    const float* max_s;        // [N, F] column max (nullptr: no normalisation)
    const float* max_n;
    const long long* lens;     // nullptr: no padding mask
    float norm_eps;            // EPSILON added to the column max (mvdr.py:113)
    float den_eps;             // clamp of the mask sum (mvdr.py:59)
    float* Rs;                 // [N, F, C, C, 2]
    float* Rn;                 // nullptr: skip
};

template <int C>
__global__ void __launch_bounds__(256) covar_kernel(const __grid_constant__ CovarParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * 8 + warp;  // (n, f)
    if (item >= (long long)p.N * p.F) return;
    const int n = (int)(item / p.F), f = (int)(item - (long long)n * p.F);
    const float* xr = p.x.re + n * p.x.sn + f * p.x.sf;
    const float* xi = p.x.im + n * p.x.sn + f * p.x.sf;
    const int tvalid = p.lens ? (int)min((long long)p.T, p.lens[n]) : p.T;
    // interleaved (re, im) pairs are read as float2: needs an 8-byte aligned base and even outer strides as well
    const bool packed = (p.x.im == p.x.re + 1) && (p.x.st == 2) && ((reinterpret_cast<uintptr_t>(p.x.re) & 7) == 0) &&
                        (((p.x.sn | p.x.sc | p.x.sf) & 1) == 0);
    const float inv_s = p.max_s ? 1.0f : 0.f;  // flag only
    const float ds = p.max_s ? (p.max_s[item] + p.norm_eps) : 1.f;
    const float dn = (p.mask_n && p.max_n) ? (p.max_n[item] + p.norm_eps) : 1.f;
    (void)inv_s;
    // Hermitian accumulators: diag (real) + strict upper triangle (complex)
    constexpr int NP = C * (C - 1) / 2;
    float sd[C], su_r[NP], su_i[NP], nd[C], nu_r[NP], nu_i[NP];
#pragma unroll
    for (int i = 0; i < C; ++i) sd[i] = nd[i] = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) su_r[i] = su_i[i] = nu_r[i] = nu_i[i] = 0.f;
    float sum_s = 0.f, sum_n = 0.f;

    for (int t = lane; t < p.T; t += 32) {
        float ms = (t < tvalid) ? __ldg(p.mask_s + n * p.msn + t * p.mst + f * p.msf) : 0.f;
        if (p.max_s) ms = ms / ds;
        float mn;
        if (p.mask_n) {
            mn = (t < tvalid) ? __ldg(p.mask_n + n * p.mnn + t * p.mnt + f * p.mnf) : 0.f;
            if (p.max_n) mn = mn / dn;
        } else {
            mn = 1.0f - ms;
        }
        float ar[C], ai[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (packed) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(xr + c * p.x.sc + 2LL * t));
                ar[c] = v.x;
                ai[c] = v.y;
            } else {
                ar[c] = __ldg(xr + c * p.x.sc + t * p.x.st);
                ai[c] = __ldg(xi + c * p.x.sc + t * p.x.st);
            }
        }
        sum_s += ms;
        sum_n += mn;
        int k = 0;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const float pw = fmaf(ar[i], ar[i], ai[i] * ai[i]);
            sd[i] = fmaf(ms, pw, sd[i]);
            nd[i] = fmaf(mn, pw, nd[i]);
#pragma unroll
            for (int j = i + 1; j < C; ++j, ++k) {
                // x_i conj(x_j)
                const float pr = fmaf(ar[i], ar[j], ai[i] * ai[j]);
                const float pi = fmaf(ai[i], ar[j], -ar[i] * ai[j]);
                su_r[k] = fmaf(ms, pr, su_r[k]);
                su_i[k] = fmaf(ms, pi, su_i[k]);
                nu_r[k] = fmaf(mn, pr, nu_r[k]);
                nu_i[k] = fmaf(mn, pi, nu_i[k]);
            }
        }
    }
    auto wsum = [](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    sum_s = wsum(sum_s);
    sum_n = wsum(sum_n);
#pragma unroll
    for (int i = 0; i < C; ++i) {
        sd[i] = wsum(sd[i]);
        nd[i] = wsum(nd[i]);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        su_r[i] = wsum(su_r[i]);
        su_i[i] = wsum(su_i[i]);
        nu_r[i] = wsum(nu_r[i]);
        nu_i[i] = wsum(nu_i[i]);
    }
    if (lane == 0) {
        const float den_s = fmaxf(sum_s, p.den_eps), den_n = fmaxf(sum_n, p.den_eps);
        float2* Rs = reinterpret_cast<float2*>(p.Rs) + item * C * C;
        float2* Rn = p.Rn ? reinterpret_cast<float2*>(p.Rn) + item * C * C : nullptr;
        int k = 0;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            Rs[i * C + i] = make_float2(sd[i] / den_s, 0.f);
            if (Rn) Rn[i * C + i] = make_float2(nd[i] / den_n, 0.f);
#pragma unroll
            for (int j = i + 1; j < C; ++j, ++k) {
                Rs[i * C + j] = make_float2(su_r[k] / den_s, su_i[k] / den_s);
                Rs[j * C + i] = make_float2(su_r[k] / den_s, -su_i[k] / den_s);
                if (Rn) {
                    Rn[i * C + j] = make_float2(nu_r[k] / den_n, nu_i[k] / den_n);
                    Rn[j * C + i] = make_float2(nu_r[k] / den_n, -nu_i[k] / den_n);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// logits[n, c] = gvec . tanh(proj . a[n, c, :] + b) + gb,  a[n, c, f] = | sum_{j != c} Rs[n, f, c, j] | / (C - 1)
__global__ void __launch_bounds__(256) ref_logits_kernel(const float* __restrict__ Rs, int N, int F, int C,
                                                         const float* __restrict__ pw, const float* __restrict__ pb,
                                                         const float* __restrict__ gw, const float* __restrict__ gb,
                                                         int A, float* __restrict__ logits) {
    extern __shared__ float sm[];
    float* a = sm;            // [F]
    float* part = sm + F;     // [8]
    const int n = blockIdx.x / C, c = blockIdx.x - n * C;
    const float2* R = reinterpret_cast<const float2*>(Rs) + (long long)n * F * C * C;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float sr = 0.f, si = 0.f;
        for (int j = 0; j < C; ++j)
            if (j != c) {
                const float2 v = __ldg(R + ((long long)f * C + c) * C + j);
                sr += v.x;
                si += v.y;
            }
        sr /= (float)(C - 1);
        si /= (float)(C - 1);
        a[f] = sqrtf(sr * sr + si * si);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int r = warp; r < A; r += 8) {
        const float* w = pw + (long long)r * F;
        float d = 0.f;
        for (int f = lane; f < F; f += 32) d = fmaf(__ldg(w + f), a[f], d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0) acc = fmaf(__ldg(gw + r), tanhf(d + __ldg(pb + r)), acc);
    }
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += part[i];
        logits[blockIdx.x] = s + __ldg(gb);
    }
}

// ------------------------------------------------------------------------------------------------
// one thread per (n, f): u = softmax(logits[n, :]); M = (Rn + eps I)^-1 Rs in double precision
// (Gauss-Jordan with partial pivoting on the complex matrix); w = M u / (tr M + eps)
template <int C>
__global__ void __launch_bounds__(128) mvdr_weights_kernel(const float* __restrict__ Rs, const float* __restrict__ Rn,
                                                           const float* __restrict__ logits, long long items, int F,
                                                           float eps, float* __restrict__ w) {
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= items) return;
    const int n = (int)(item / F);
    float u[C];
    {
        float mx = -INFINITY, s = 0.f;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, logits[n * C + c]);
        for (int c = 0; c < C; ++c) {
            u[c] = expf(logits[n * C + c] - mx);
            s += u[c];
        }
        for (int c = 0; c < C; ++c) u[c] /= s;
    }
    // augmented system [A | B] with A = Rn + eps I, B = Rs; reduce A to identity
    double ar[C][C], ai[C][C], br[C][C], bi[C][C];
    const float2* pa = reinterpret_cast<const float2*>(Rn) + item * C * C;
    const float2* pb = reinterpret_cast<const float2*>(Rs) + item * C * C;
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const float2 va = pa[i * C + j], vb = pb[i * C + j];
            ar[i][j] = (double)(i == j ? va.x + eps : va.x);   // fp32 add like the reference (mvdr.py:90)
            ai[i][j] = va.y;
            br[i][j] = vb.x;
            bi[i][j] = vb.y;
        }
#pragma unroll
    for (int col = 0; col < C; ++col) {
        // partial pivoting
        int piv = col;
        double best = ar[col][col] * ar[col][col] + ai[col][col] * ai[col][col];
#pragma unroll
        for (int r = col + 1; r < C; ++r) {
            const double m = ar[r][col] * ar[r][col] + ai[r][col] * ai[r][col];
            if (m > best) {
                best = m;
                piv = r;
            }
        }
#pragma unroll
        for (int r = col + 1; r < C; ++r)
            if (r == piv) {
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    double t;
                    t = ar[col][j]; ar[col][j] = ar[r][j]; ar[r][j] = t;
                    t = ai[col][j]; ai[col][j] = ai[r][j]; ai[r][j] = t;
                    t = br[col][j]; br[col][j] = br[r][j]; br[r][j] = t;
                    t = bi[col][j]; bi[col][j] = bi[r][j]; bi[r][j] = t;
                }
            }
        // scale the pivot row by 1 / a[col][col]
        const double pr = ar[col][col], pi = ai[col][col], pd = 1.0 / (pr * pr + pi * pi);
        const double ir = pr * pd, ii = -pi * pd;
#pragma unroll
        for (int j = 0; j < C; ++j) {
            double xr = ar[col][j], xi = ai[col][j];
            ar[col][j] = xr * ir - xi * ii;
            ai[col][j] = xr * ii + xi * ir;
            xr = br[col][j];
            xi = bi[col][j];
            br[col][j] = xr * ir - xi * ii;
            bi[col][j] = xr * ii + xi * ir;
        }
        // eliminate the column from every other row
#pragma unroll
        for (int r = 0; r < C; ++r) {
            if (r == col) continue;
            const double fr = ar[r][col], fi = ai[r][col];
#pragma unroll
            for (int j = 0; j < C; ++j) {
                ar[r][j] -= fr * ar[col][j] - fi * ai[col][j];
                ai[r][j] -= fr * ai[col][j] + fi * ar[col][j];
                br[r][j] -= fr * br[col][j] - fi * bi[col][j];
                bi[r][j] -= fr * bi[col][j] + fi * br[col][j];
            }
        }
    }
    // b now holds M = A^-1 B
    double tr = (double)eps, ti = 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        tr += br[i][i];
        ti += bi[i][i];
    }
    const double sc = 1.0 / (tr * tr + ti * ti);
    float2* po = reinterpret_cast<float2*>(w) + item * C;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        double nr = 0.0, ni = 0.0;
#pragma unroll
        for (int j = 0; j < C; ++j) {
            nr += br[i][j] * (double)u[j];
            ni += bi[i][j] * (double)u[j];
        }
        po[i] = make_float2((float)((nr * tr + ni * ti) * sc), (float)((ni * tr - nr * ti) * sc));
    }
}

// ------------------------------------------------------------------------------------------------
// Y[n, f, t] = sum_c conj(w[n, f, c]) X[n, c, f, t]; yr / yi are contiguous [N, F, T]
struct BeamParams {
    CplxView x;
    int N, C, F, T;
    const float* w;        // [N, F, C, 2]
    float* yr;
    float* yi;
};

template <int C>
__global__ void __launch_bounds__(256) beamform_kernel(const __grid_constant__ BeamParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * 8 + warp;
    if (item >= (long long)p.N * p.F) return;
    const int n = (int)(item / p.F), f = (int)(item - (long long)n * p.F);
    const float* xr = p.x.re + n * p.x.sn + f * p.x.sf;
    const float* xi = p.x.im + n * p.x.sn + f * p.x.sf;
    // interleaved (re, im) pairs are read as float2: needs an 8-byte aligned base and even outer strides as well
    const bool packed = (p.x.im == p.x.re + 1) && (p.x.st == 2) && ((reinterpret_cast<uintptr_t>(p.x.re) & 7) == 0) &&
                        (((p.x.sn | p.x.sc | p.x.sf) & 1) == 0);
    float wr[C], wi[C];
    const float2* pw = reinterpret_cast<const float2*>(p.w) + item * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float2 v = __ldg(pw + c);
        wr[c] = v.x;
        wi[c] = v.y;
    }
    float* yr = p.yr + item * p.T;
    float* yi = p.yi + item * p.T;
    for (int t = lane; t < p.T; t += 32) {
        float sr = 0.f, si = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float a, b;
            if (packed) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(xr + c * p.x.sc + 2LL * t));
                a = v.x;
                b = v.y;
            } else {
                a = __ldg(xr + c * p.x.sc + t * p.x.st);
                b = __ldg(xi + c * p.x.sc + t * p.x.st);
            }
            // conj(w) x = (wr a + wi b) + i (wr b - wi a)
            sr = fmaf(wr[c], a, fmaf(wi[c], b, sr));
            si = fmaf(wr[c], b, fmaf(-wi[c], a, si));
        }
        yr[t] = sr;
        yi[t] = si;
    }
}

static int check_view(const float* re, const float* im, int64_t C) {
    APSB_CHECK_ARG(re && im, "null spectrogram pointer");
    APSB_CHECK_ARG(C >= 2 && C <= kMaxCh, "channel count %lld not in [2, %d]", (long long)C, kMaxCh);
    return 0;
}

}  // namespace apsb

using namespace apsb;

#define APSB_DISPATCH_C(C_, CALL)                 \
    switch (C_) {                                 \
        case 2: { constexpr int CC = 2; CALL; } break; \
        case 3: { constexpr int CC = 3; CALL; } break; \
        case 4: { constexpr int CC = 4; CALL; } break; \
        case 5: { constexpr int CC = 5; CALL; } break; \
        case 6: { constexpr int CC = 6; CALL; } break; \
        default: return set_error(-1, "unsupported channel count %d", (int)(C_)); \
    }

extern "C" int aps_b200_mask_colmax(const float* mask, int64_t stride_n, int64_t stride_t, int64_t stride_f,
                                    int64_t batch, int64_t num_frames, int64_t num_bins, const int64_t* lens,
                                    float* out, void* stream) {
    APSB_CHECK_ARG(mask && out && batch > 0 && num_frames > 0 && num_bins > 0, "bad arguments");
    APSB_CHECK_ARG(batch <= 65535, "batch too large");
    dim3 grid((unsigned)((num_bins + 31) / 32), (unsigned)batch);
    mask_colmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mask, stride_n, stride_t, stride_f, (int)batch,
                                                               (int)num_frames, (int)num_bins,
                                                               reinterpret_cast<const long long*>(lens), out);
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_covar_fwd(const float* x_real, const float* x_imag, const int64_t* x_strides, int64_t batch,
                                  int64_t channels, int64_t num_bins, int64_t num_frames, const float* mask_s,
                                  const int64_t* mask_s_strides, const float* max_s, const float* mask_n,
                                  const int64_t* mask_n_strides, const float* max_n, const int64_t* lens,
                                  float norm_eps, float den_eps, float* Rs, float* Rn, void* stream) {
    if (int rc = check_view(x_real, x_imag, channels)) return rc;
    APSB_CHECK_ARG(x_strides && mask_s && mask_s_strides && Rs, "null pointer argument");
    APSB_CHECK_ARG(!mask_n || mask_n_strides, "noise mask strides missing");
    APSB_CHECK_ARG(batch > 0 && num_bins > 0 && num_frames > 0, "bad shape");
    CovarParams p{};
    p.x = {x_real, x_imag, x_strides[0], x_strides[1], x_strides[2], x_strides[3]};
    p.N = (int)batch; p.C = (int)channels; p.F = (int)num_bins; p.T = (int)num_frames;
    p.mask_s = mask_s; p.msn = mask_s_strides[0]; p.mst = mask_s_strides[1]; p.msf = mask_s_strides[2];
    p.mask_n = mask_n;
    if (mask_n) { p.mnn = mask_n_strides[0]; p.mnt = mask_n_strides[1]; p.mnf = mask_n_strides[2]; }
    p.max_s = max_s; p.max_n = max_n;
    p.lens = reinterpret_cast<const long long*>(lens);
    p.norm_eps = norm_eps; p.den_eps = den_eps; p.Rs = Rs; p.Rn = Rn;
    const long long items = (long long)batch * num_bins;
    const unsigned grid = (unsigned)((items + 7) / 8);
    APSB_DISPATCH_C(channels, (covar_kernel<CC><<<grid, 256, 0, (cudaStream_t)stream>>>(p)));
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_mvdr_ref_logits(const float* Rs, int64_t batch, int64_t num_bins, int64_t channels,
                                        const float* proj_weight, const float* proj_bias, const float* gvec_weight,
                                        const float* gvec_bias, int64_t att_dim, float* logits, void* stream) {
    APSB_CHECK_ARG(Rs && proj_weight && proj_bias && gvec_weight && gvec_bias && logits, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && num_bins > 0 && channels >= 2 && att_dim > 0, "bad shape");
    const size_t smem = (size_t)(num_bins + 8) * sizeof(float);
    APSB_CHECK_ARG(smem <= 48 * 1024, "too many bins (%lld) for the attention kernel", (long long)num_bins);
    ref_logits_kernel<<<(unsigned)(batch * channels), 256, smem, (cudaStream_t)stream>>>(
        Rs, (int)batch, (int)num_bins, (int)channels, proj_weight, proj_bias, gvec_weight, gvec_bias, (int)att_dim,
        logits);
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_mvdr_weights(const float* Rs, const float* Rn, const float* logits, int64_t batch,
                                     int64_t num_bins, int64_t channels, float eps, float* weight, void* stream) {
    APSB_CHECK_ARG(Rs && Rn && logits && weight, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && num_bins > 0, "bad shape");
    const long long items = (long long)batch * num_bins;
    const unsigned grid = (unsigned)((items + 127) / 128);
    APSB_DISPATCH_C(channels, (mvdr_weights_kernel<CC><<<grid, 128, 0, (cudaStream_t)stream>>>(
                                  Rs, Rn, logits, items, (int)num_bins, eps, weight)));
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_beamform_fwd(const float* x_real, const float* x_imag, const int64_t* x_strides, int64_t batch,
                                     int64_t channels, int64_t num_bins, int64_t num_frames, const float* weight,
                                     float* y_real, float* y_imag, void* stream) {
    if (int rc = check_view(x_real, x_imag, channels)) return rc;
    APSB_CHECK_ARG(x_strides && weight && y_real && y_imag, "null pointer argument");
    BeamParams p{};
    p.x = {x_real, x_imag, x_strides[0], x_strides[1], x_strides[2], x_strides[3]};
    p.N = (int)batch; p.C = (int)channels; p.F = (int)num_bins; p.T = (int)num_frames;
    p.w = weight; p.yr = y_real; p.yi = y_imag;
    const long long items = (long long)batch * num_bins;
    const unsigned grid = (unsigned)((items + 7) / 8);
    APSB_DISPATCH_C(channels, (beamform_kernel<CC><<<grid, 256, 0, (cudaStream_t)stream>>>(p)));
    APSB_LAUNCH_CHECK();
    return 0;
}

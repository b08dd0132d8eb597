// gemm.cuh — fp32 SIMT GEMM core with pluggable A-operand loaders and fused epilogues.
//
//   C[m, n] = epilogue( sum_k A(m, k) * W[n, k] )        (W row-major [N, K]: a torch Linear / conv weight)
//
// This is the EXACT-fp32 contraction used by every dense layer of the encoder / TCN / DCCRN paths in
// round 1: the 1e-4 parity budget against an fp32 reference rules out single-pass TF32/BF16 tensor-core
// inputs; the tcgen05 3xTF32 path is the planned replacement behind the same entry
// points (DESIGN.md "GEMM precision").  Tiles BM x BN x 16, 256 threads, each thread a (TM x TN) micro
// tile split into 4-wide groups BM/2 (BN/2) apart so that every shared-memory fragment read is one
// conflict-free LDS.128; global operands are fetched one k-tile ahead into registers.
//
// A-loaders: PlainA (row-major activations) and ConvA (implicit im2col of an NHWC tensor: replaces
// F.unfold/cuDNN for aps/asr/base/component.py:306 Conv2d and the conv1d layers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace apsb {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SWISH = 2, ACT_TANH = 3, ACT_SIGMOID = 4, ACT_PRELU = 5, ACT_GLU = 6,
                 ACT_LEAKY = 7, ACT_GELU = 8 };

struct Epilogue {
    const float* bias;     // [N] or nullptr
    int act;               // Act; ACT_GLU pairs columns (2j, 2j+1) -> output column j = v0 * sigmoid(v1)
    float alpha;           // v = alpha * act(acc + bias)
    const float* slope;    // PReLU slopes: [1] or [N] (slope_stride 0 / 1); LeakyReLU: negative slope in `leak`
    int slope_stride;
    float leak;
    const float* post_scale;  // optional per-column affine applied right after the activation:
    const float* post_shift;  //   v = act(.) * post_scale[n] + post_shift[n]   (eval BatchNorm behind a PReLU)
    const float* res;      // optional residual, added AFTER activation and alpha: out = v + beta * res[m, n]
    long long ldres;
    float beta;
    float* out;
    long long ldo;
    float* out_lo;         // tensor-core engine only: optional TF32 "lo" companion of `out` (same leading dimension)
    int dbg_nobias;        // tuning builds only: skip the bias loads of the tensor-core epilogue
};

__device__ __forceinline__ float apply_act(float v, int act, const Epilogue& e, int n) {
    switch (act) {
        case ACT_RELU: return fmaxf(v, 0.f);
        case ACT_SWISH: return v / (1.f + __expf(-v));
        case ACT_TANH: return tanhf(v);
        case ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
        case ACT_PRELU: return v >= 0.f ? v : v * __ldg(e.slope + (long long)n * e.slope_stride);
        case ACT_LEAKY: return v >= 0.f ? v : v * e.leak;
        case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        default: return v;
    }
}

// ---- A loaders: fetch 4 consecutive k of row m (k % 4 == 0); zero outside the matrix ------------------
struct PlainA {
    const float* A;
    long long lda;
    int M, K;
    int vec;  // rows 16-byte aligned and K % 4 == 0
    struct Row { const float* p; bool ok; };
    __device__ __forceinline__ Row row(int m) const { return {A + (long long)m * lda, m < M}; }
    __device__ __forceinline__ float4 load4(const Row& r, int k) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!r.ok) return v;
        if (vec) {
            if (k < K) v = __ldg(reinterpret_cast<const float4*>(r.p + k));
        } else {
            if (k < K) v.x = __ldg(r.p + k);
            if (k + 1 < K) v.y = __ldg(r.p + k + 1);
            if (k + 2 < K) v.z = __ldg(r.p + k + 2);
            if (k + 3 < K) v.w = __ldg(r.p + k + 3);
        }
        return v;
    }
};

// implicit im2col of x[Nb, H, W, Cin] (NHWC): m -> (nb, oh, ow), k -> (kh, kw, c)
struct ConvA {
    const float* x;
    int Nb, H, W, Cin, KH, KW, sh, sw, ph, pw, dh, dw, OH, OW;
    int M, K;
    int vec;  // Cin % 4 == 0 and x 16-byte aligned
    struct Row { int nb, ih0, iw0; bool ok; };
    __device__ __forceinline__ Row row(int m) const {
        Row r;
        r.ok = m < M;
        const int mm = r.ok ? m : 0;
        const int ow = mm % OW, t = mm / OW;
        const int oh = t % OH;
        r.nb = t / OH;
        r.ih0 = oh * sh - ph;
        r.iw0 = ow * sw - pw;
        return r;
    }
    __device__ __forceinline__ float at(const Row& r, int k) const {
        if (k >= K) return 0.f;
        const int c = k % Cin, t = k / Cin;
        const int kw = t % KW, kh = t / KW;
        const int ih = r.ih0 + kh * dh, iw = r.iw0 + kw * dw;
        if (ih < 0 || ih >= H || iw < 0 || iw >= W) return 0.f;
        return __ldg(x + (((long long)r.nb * H + ih) * W + iw) * Cin + c);
    }
    __device__ __forceinline__ float4 load4(const Row& r, int k) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!r.ok) return v;
        if (vec) {
            if (k >= K) return v;
            const int c = k % Cin, t = k / Cin;
            const int kw = t % KW, kh = t / KW;
            const int ih = r.ih0 + kh * dh, iw = r.iw0 + kw * dw;
            if (ih < 0 || ih >= H || iw < 0 || iw >= W) return v;
            return __ldg(reinterpret_cast<const float4*>(x + (((long long)r.nb * H + ih) * W + iw) * Cin + c));
        }
        v.x = at(r, k); v.y = at(r, k + 1); v.z = at(r, k + 2); v.w = at(r, k + 3);
        return v;
    }
};

// implicit gather of a TRANSPOSED convolution: out[oh, ow] = sum x[ih, iw] w[kh, kw] over oh = ih*sh - ph + kh
struct TConvA {
    const float* x;
    int Nb, H, W, Cin, KH, KW, sh, sw, ph, pw, OH, OW;
    int M, K;
    int vec;
    struct Row { int nb, oh, ow; bool ok; };
    __device__ __forceinline__ Row row(int m) const {
        Row r;
        r.ok = m < M;
        const int mm = r.ok ? m : 0;
        r.ow = mm % OW;
        const int t = mm / OW;
        r.oh = t % OH;
        r.nb = t / OH;
        return r;
    }
    __device__ __forceinline__ const float* src(const Row& r, int k) const {
        const int c = k % Cin, t = k / Cin;
        const int kw = t % KW, kh = t / KW;
        const int nh = r.oh + ph - kh, nw = r.ow + pw - kw;
        if (nh < 0 || nw < 0 || nh % sh || nw % sw) return nullptr;
        const int ih = nh / sh, iw = nw / sw;
        if (ih >= H || iw >= W) return nullptr;
        return x + (((long long)r.nb * H + ih) * W + iw) * Cin + c;
    }
    __device__ __forceinline__ float4 load4(const Row& r, int k) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!r.ok) return v;
        if (vec) {
            if (k >= K) return v;
            const float* p = src(r, k);
            return p ? __ldg(reinterpret_cast<const float4*>(p)) : v;
        }
        float* o = &v.x;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (k + i < K) {
                const float* p = src(r, k + i);
                o[i] = p ? __ldg(p) : 0.f;
            }
        return v;
    }
};

constexpr int kGemmThreads = 256;
constexpr int kBK = 16;

template <int BM, int BN, typename ALoader>
__global__ void __launch_bounds__(kGemmThreads) gemm_kernel(const ALoader a, const float* __restrict__ Wt, long long ldw,
                                                            int wvec, int M, int N, int K, const Epilogue e) {
    constexpr int TM = BM / 16, TN = BN / 16;      // micro tile (8x8 for 128x128, 4x4 for 64x64)
    constexpr int HM = TM / 2, HN = TN / 2;        // two groups per dimension, BM/2 (BN/2) apart
    static_assert(HM == 4 || HM == 2, "tile");
    constexpr int LDA = BM + 4, LDB = BN + 4;
    __shared__ __align__(16) float As[2][kBK][LDA];
    __shared__ __align__(16) float Bs[2][kBK][LDB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // global fetch assignment: each thread fetches AR rows x one float4 of k for A, BR for B
    constexpr int AR = BM * kBK / 4 / kGemmThreads, BR = BN * kBK / 4 / kGemmThreads;
    const int kq = (tid & 3) * 4, r0 = tid >> 2;   // k offset inside the tile, first row
    typename ALoader::Row arow[AR];
#pragma unroll
    for (int i = 0; i < AR; ++i) arow[i] = a.row(m0 + r0 + i * 64);
    float4 pa[AR], pb[BR];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < AR; ++i) pa[i] = a.load4(arow[i], k0 + kq);
#pragma unroll
        for (int i = 0; i < BR; ++i) {
            const int n = n0 + r0 + i * 64, k = k0 + kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < N) {
                const float* w = Wt + (long long)n * ldw;
                if (wvec) {
                    if (k < K) v = __ldg(reinterpret_cast<const float4*>(w + k));
                } else {
                    if (k < K) v.x = __ldg(w + k);
                    if (k + 1 < K) v.y = __ldg(w + k + 1);
                    if (k + 2 < K) v.z = __ldg(w + k + 2);
                    if (k + 3 < K) v.w = __ldg(w + k + 3);
                }
            }
            pb[i] = v;
        }
    };
    auto stash = [&](int b) {
#pragma unroll
        for (int i = 0; i < AR; ++i) {
            const int r = r0 + i * 64;
            As[b][kq][r] = pa[i].x; As[b][kq + 1][r] = pa[i].y; As[b][kq + 2][r] = pa[i].z; As[b][kq + 3][r] = pa[i].w;
        }
#pragma unroll
        for (int i = 0; i < BR; ++i) {
            const int r = r0 + i * 64;
            Bs[b][kq][r] = pb[i].x; Bs[b][kq + 1][r] = pb[i].y; Bs[b][kq + 2][r] = pb[i].z; Bs[b][kq + 3][r] = pb[i].w;
        }
    };
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    fetch(0);
    stash(0);
    __syncthreads();
    const int nk = (K + kBK - 1) / kBK;
    for (int kt = 0; kt < nk; ++kt) {
        const int b = kt & 1;
        if (kt + 1 < nk) fetch((kt + 1) * kBK);
#pragma unroll
        for (int k = 0; k < kBK; ++k) {
            float af[TM], bf[TN];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (HM == 4) {
                    const float4 v = *reinterpret_cast<const float4*>(&As[b][k][g * (BM / 2) + ty * 4]);
                    af[g * 4] = v.x; af[g * 4 + 1] = v.y; af[g * 4 + 2] = v.z; af[g * 4 + 3] = v.w;
                } else {
                    const float2 v = *reinterpret_cast<const float2*>(&As[b][k][g * (BM / 2) + ty * 2]);
                    af[g * 2] = v.x; af[g * 2 + 1] = v.y;
                }
                if (HN == 4) {
                    const float4 v = *reinterpret_cast<const float4*>(&Bs[b][k][g * (BN / 2) + tx * 4]);
                    bf[g * 4] = v.x; bf[g * 4 + 1] = v.y; bf[g * 4 + 2] = v.z; bf[g * 4 + 3] = v.w;
                } else {
                    const float2 v = *reinterpret_cast<const float2*>(&Bs[b][k][g * (BN / 2) + tx * 2]);
                    bf[g * 2] = v.x; bf[g * 2 + 1] = v.y;
                }
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            stash(b ^ 1);
            __syncthreads();
        }
    }
    // ---- epilogue ----------------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / HM) * (BM / 2) + ty * HM + (i % HM);
        if (m >= M) continue;
        if (e.act == ACT_GLU) {
#pragma unroll
            for (int j = 0; j < TN; j += 2) {
                const int n = n0 + (j / HN) * (BN / 2) + tx * HN + (j % HN);
                if (n + 1 < N) {
                    const float v0 = acc[i][j] + (e.bias ? __ldg(e.bias + n) : 0.f);
                    const float v1 = acc[i][j + 1] + (e.bias ? __ldg(e.bias + n + 1) : 0.f);
                    float v = e.alpha * (v0 * (1.f / (1.f + __expf(-v1))));
                    const int no = n >> 1;
                    if (e.res) v = fmaf(e.beta, __ldg(e.res + (long long)m * e.ldres + no), v);
                    e.out[(long long)m * e.ldo + no] = v;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = n0 + (j / HN) * (BN / 2) + tx * HN + (j % HN);
                if (n >= N) continue;
                float v = acc[i][j] + (e.bias ? __ldg(e.bias + n) : 0.f);
                v = apply_act(v, e.act, e, n);
                if (e.post_scale) v = fmaf(v, __ldg(e.post_scale + n), __ldg(e.post_shift + n));
                v *= e.alpha;
                if (e.res) v = fmaf(e.beta, __ldg(e.res + (long long)m * e.ldres + n), v);
                e.out[(long long)m * e.ldo + n] = v;
            }
        }
    }
}

// host-side launcher shared by the entry points (defined in gemm.cu)
template <typename ALoader>
int launch_gemm(const ALoader& a, const float* Wt, long long ldw, int M, int N, int K, const Epilogue& e,
                cudaStream_t st);

}  // namespace apsb

// gemm.cu — entry points of the fp32 GEMM core: dense layers and NHWC implicit-GEMM convolutions.
//
// Replaces (forward, inference) torch.nn.functional.linear / conv1d(k=1) / conv2d as used by
// /root/reference/aps/asr/transformer/impl.py:388-393,454-475 (feed-forward and point-wise conv layers),
// impl.py:62-83 (attention in/out projections), aps/asr/base/component.py:251-307 (Conv2d block, with the
// eval-mode BatchNorm folded into weight/bias by the caller), aps/asr/base/encoder.py:415-441
// (Conv2dEncoder.outp), aps/sse/bss/tcn.py:112-159 (1x1 convolutions of Conv1dBlock).
#include "../../include/aps_b200.h"
#include "common.cuh"
#include <stdlib.h>
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
    int rw;             // conv2d_thin3x3_kernel: floats per staged input row
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

// 3x3 convolution of a 1- or 2-channel image (conv2d front of the encoder: [N, T, 80] -> 256 channels; first DCCRN
// encoder layer: stacked re/im).  The layer is a pure WRITE-bandwidth problem (506 MB of output for 8 MB of input at
// B = 64; a plain fill of that tensor takes 68 us), so everything is arranged around issuing as few instructions per
// stored float4 as possible: a thread owns 4 output channels for its whole life and keeps their 9 * CIN * 4 weights and
// the bias in REGISTERS (as f32x2 pairs: FFMA2 halves the FMA issue slots); a block walks output rows (nb, oh).
// Second version (round 2, visit O: ncu of the first showed 99 instructions per stored float4, 9.7 long-scoreboard
// stall cycles per issue): the three input rows of an output row are staged in shared memory WITH their zero padding
// by cp.async one row ahead (double buffer), so the position loop has no bounds test, no predicated global load and
// no wait on L2 — 6 to 9 shared-memory loads, 18 * CIN FFMA2, the activation and one 16-byte store per position.
__device__ __forceinline__ unsigned long long thin_pack(float lo, float hi) {
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}

template <int CIN, int ACT>
__global__ void __launch_bounds__(256) conv2d_thin3x3_kernel(const __grid_constant__ NarrowConvParams p) {
    constexpr int K = 9 * CIN;
    extern __shared__ __align__(16) float thin_rows[];            // [2][3][rw]
    const unsigned groups = (unsigned)p.Cout >> 2;                // power of two <= 256 (host check)
    const unsigned ppb = 256u / groups;                           // position lanes of the block
    const unsigned g = threadIdx.x & (groups - 1), pl = threadIdx.x / groups;
    unsigned long long w01[K], w23[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float* wp = p.w + (long long)(4 * g) * K + k;
        w01[k] = thin_pack(__ldg(wp), __ldg(wp + K));
        w23[k] = thin_pack(__ldg(wp + 2 * K), __ldg(wp + 3 * K));
    }
    const float4 b = p.e.bias ? __ldg(reinterpret_cast<const float4*>(p.e.bias) + g) : make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned long long b01 = thin_pack(b.x, b.y), b23 = thin_pack(b.z, b.w);
    const unsigned rows = (unsigned)(p.M / p.OW);                 // Nb * OH
    const float alpha = p.e.alpha, leak = p.e.leak;
    const int rw = p.rw, wc = p.W * CIN, lead = p.pw * CIN;       // staged row: lead zeros, W * CIN samples, zeros up to rw
    // stage the 3 input rows of output row `row` into buffer `buf` (asynchronously; padding written as zeros)
    auto stage = [&](unsigned row, int buf) {
        const unsigned nb = row / (unsigned)p.OH, oh = row - nb * (unsigned)p.OH;
        const int ih0 = (int)oh * p.sh - p.ph;
        const float* img = p.x + (long long)nb * p.H * wc;
        float* dst = thin_rows + buf * 3 * rw;
        for (int i = threadIdx.x; i < 3 * rw; i += 256) {
            const int kh = i / rw, col = i - kh * rw, ih = ih0 + kh, ic = col - lead;
            if (ih >= 0 && ih < p.H && ic >= 0 && ic < wc) {
                const unsigned d = (unsigned)__cvta_generic_to_shared(dst + i);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(img + (long long)ih * wc + ic) : "memory");
            } else {
                dst[i] = 0.f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int buf = 0;
    if (blockIdx.x < rows) stage(blockIdx.x, 0);
    const int step = (int)ppb * p.sw * CIN;                       // floats between this lane's consecutive positions
    for (unsigned row = blockIdx.x; row < rows; row += gridDim.x, buf ^= 1) {
        const unsigned next = row + gridDim.x;
        if (next < rows) {
            stage(next, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                          // this row's samples (and zeros) are visible
        const float* r0 = thin_rows + buf * 3 * rw + (int)pl * p.sw * CIN;
        float* op = p.e.out + ((long long)row * p.OW + pl) * p.e.ldo + 4 * g;
        const long long ostep = (long long)ppb * p.e.ldo;
        for (unsigned ow = pl; ow < (unsigned)p.OW; ow += ppb, r0 += step, op += ostep) {
            unsigned long long a01 = b01, a23 = b23;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int t = 0; t < 3 * CIN; ++t) {
                    const float xv = r0[kh * rw + t];
                    const unsigned long long x2 = thin_pack(xv, xv);
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a01) : "l"(x2), "l"(w01[kh * 3 * CIN + t]));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a23) : "l"(x2), "l"(w23[kh * 3 * CIN + t]));
                }
            }
            float4 a = make_float4(__uint_as_float((unsigned)a01), __uint_as_float((unsigned)(a01 >> 32)),
                                   __uint_as_float((unsigned)a23), __uint_as_float((unsigned)(a23 >> 32)));
            if (ACT == ACT_RELU) {
                a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
            } else if (ACT == ACT_LEAKY) {
                a.x = a.x >= 0.f ? a.x : a.x * leak; a.y = a.y >= 0.f ? a.y : a.y * leak;
                a.z = a.z >= 0.f ? a.z : a.z * leak; a.w = a.w >= 0.f ? a.w : a.w * leak;
            }
            if (alpha != 1.f) { a.x *= alpha; a.y *= alpha; a.z *= alpha; a.w *= alpha; }
            *reinterpret_cast<float4*>(op) = a;
        }
        __syncthreads();                                          // the buffer is free for the row after next
    }
}

// Transposed convolution with very few OUTPUT channels (<= 8): the last DCCRN decoder layer maps 64 channels to
// 2 * num_spks (dccrn.py:133-147).  On the GEMM engine such a layer pays for a 64-column tile and is bound by the
// operand gather (6.9 ms at B = 128 x 4 s for 1.2 GB of traffic); here 8 lanes share one output pixel, each reads
// one float4 of the pixel's channels per tap (a warp request covers whole 128-byte pixel rows of 4 pixels), the
// weights sit in shared memory as [tap][channel][CO] and the 8 partial sums meet in three shuffles.  `x2` is the
// skip tensor of a "cat" connection read in place (channel order of cat_complex: [x_re, skip_re, x_im, skip_im]).
struct NarrowTConvParams {
    const float* x;
    const float* x2;
    int H, W, Cx, cat_c, Cin, KH, KW, sh, sw, ph, pw, OH, OW, M, Cout;
    const float* w;       // [Cout, KH, KW, Cin]
    Epilogue e;
};

// input index i with i * stride == n (n >= 0), or -1; strides 1 and 2 (every reference model) without a division
__device__ __forceinline__ int tconv_src(int n, int stride) {
    if (stride == 1) return n;
    if (stride == 2) return (n & 1) ? -1 : (n >> 1);
    const int i = n / stride;
    return i * stride == n ? i : -1;
}

// channel quads per tap in the shared-memory weight layout (padded so that the swizzle stays in range) and the swizzle
__host__ __device__ __forceinline__ int tconv_slots(int cin) { return ((cin / 4 + 15) / 16) * 16; }
__device__ __forceinline__ int tconv_swz(int quad) { return quad ^ (((quad >> 3) & 1) << 2); }

template <int CO, int KW_>   // KW_ > 0: compile-time kernel width (3 in every reference model), 0: run-time
__global__ void __launch_bounds__(256) tconv_narrow_kernel(const __grid_constant__ NarrowTConvParams p) {
    // weights as [tap][c % 4][swz(c / 4)][CO]: the 8 lanes of a pixel hold consecutive channel quads, so their float4
    // reads of one (tap, c % 4) are 128 contiguous bytes; the XOR moves the second half of a cat segment pair
    // (quads 8..11 next to 0..3 when each tensor has 32 channels) onto the other 16 banks.  The first version,
    // [tap][c][CO], had 64-byte lane strides: 4-way conflicts, shared-memory pipe 92 % busy (ncu).
    extern __shared__ __align__(16) float sw_[];
    constexpr int KWM = KW_ > 0 ? KW_ : 1;            // taps gathered together (all loads of a kernel row in flight)
    const int KW = KW_ > 0 ? KW_ : p.KW;
    const int K = p.KH * KW * p.Cin;
    const int SL = tconv_slots(p.Cin);
    for (int i = threadIdx.x; i < K * CO; i += blockDim.x) {
        const int k = i / CO, co = i - k * CO;
        const int tap = k / p.Cin, c = k - tap * p.Cin;
        sw_[((tap * 4 + (c & 3)) * SL + tconv_swz(c >> 2)) * CO + co] = co < p.Cout ? __ldg(p.w + (long long)co * K + k) : 0.f;
    }
    __syncthreads();
    const int j = threadIdx.x & 7, slot = threadIdx.x >> 3;
    const bool two = p.x2 != nullptr;
    for (unsigned base = blockIdx.x * 32u; base < (unsigned)p.M; base += gridDim.x * 32u) {
        const unsigned m = base + slot;
        const bool valid = m < (unsigned)p.M;
        const unsigned t = m / (unsigned)p.OW, ow = m - t * (unsigned)p.OW;
        const unsigned nb = t / (unsigned)p.OH, oh = t - nb * (unsigned)p.OH;
        float acc[CO];
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[c] = 0.f;
        if (valid) {
            for (int kh = 0; kh < p.KH; ++kh) {
                const int nh = (int)oh + p.ph - kh;
                if (nh < 0) continue;
                const int ih = tconv_src(nh, p.sh);                 // -1: no input row feeds this tap
                if (ih < 0 || ih >= p.H) continue;
                const long long rowp = ((long long)nb * p.H + ih) * p.W;
                for (int kw0 = 0; kw0 < KW; kw0 += KWM) {
                    for (int q = j * 4; q < p.Cx; q += 32) {
                        // phase 1: every load of this kernel row (taps x both tensors) is issued before any is used
                        float4 v[KWM][2];
                        bool ok[KWM];
#pragma unroll
                        for (int i = 0; i < KWM; ++i) {
                            const int nw = (int)ow + p.pw - (kw0 + i);
                            const int iw = nw < 0 ? -1 : tconv_src(nw, p.sw);
                            ok[i] = iw >= 0 && iw < p.W;
                            const long long off = (rowp + iw) * p.Cx + q;
                            v[i][0] = ok[i] ? __ldg(reinterpret_cast<const float4*>(p.x + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            v[i][1] = (ok[i] && two) ? __ldg(reinterpret_cast<const float4*>(p.x2 + off))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        // phase 2
#pragma unroll
                        for (int i = 0; i < KWM; ++i) {
                            if (!ok[i]) continue;
                            const float* wt = sw_ + (kh * KW + kw0 + i) * 4 * SL * CO;
#pragma unroll
                            for (int tn = 0; tn < 2; ++tn) {
                                if (tn && !two) break;
                                int c = q;
                                if (p.cat_c) {
                                    const int half = q >= p.cat_c;
                                    c = (2 * half + tn) * p.cat_c + q - half * p.cat_c;
                                }
                                const float* wp = wt + tconv_swz(c >> 2) * CO;
                                const float4 vv = v[i][tn];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float xv = e == 0 ? vv.x : e == 1 ? vv.y : e == 2 ? vv.z : vv.w;
#pragma unroll
                                    for (int g = 0; g < CO / 4; ++g) {
                                        const float4 wv = *reinterpret_cast<const float4*>(wp + e * SL * CO + 4 * g);
                                        acc[4 * g + 0] = fmaf(xv, wv.x, acc[4 * g + 0]);
                                        acc[4 * g + 1] = fmaf(xv, wv.y, acc[4 * g + 1]);
                                        acc[4 * g + 2] = fmaf(xv, wv.z, acc[4 * g + 2]);
                                        acc[4 * g + 3] = fmaf(xv, wv.w, acc[4 * g + 3]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        // the 8 lanes of a pixel hold partial sums over their channels; afterwards lane j owns output channel j
        float mine = 0.f;
#pragma unroll
        for (int c = 0; c < CO; ++c) {
            float v = acc[c];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            if (c == j) mine = v;
        }
        if (valid && j < p.Cout) {
            float v = mine + (p.e.bias ? __ldg(p.e.bias + j) : 0.f);
            v = apply_act(v, p.e.act, p.e, j);
            if (p.e.post_scale) v = fmaf(v, __ldg(p.e.post_scale + j), __ldg(p.e.post_shift + j));
            v *= p.e.alpha;
            if (p.e.res) v = fmaf(p.e.beta, __ldg(p.e.res + (long long)m * p.e.ldres + j), v);
            p.e.out[(long long)m * p.e.ldo + j] = v;
        }
    }
}

// Same operation for the common geometry (kernel width 3, stride 1 along the width axis): the 8 lanes of a group own
// FOUR neighbouring output pixels of a row.  Their taps overlap — 6 input columns feed 4 pixels x 3 taps — so a kernel
// row costs 12 float4 loads instead of 24, and every weight float4 read from shared memory is used for 16 FMAs instead
// of 4 (the one-pixel kernel above is bound by exactly those shared-memory reads).
template <int CO>
__global__ void __launch_bounds__(256) tconv_narrow_tiled_kernel(const __grid_constant__ NarrowTConvParams p) {
    constexpr int P = 4, KW = 3, NC = P + KW - 1;
    extern __shared__ __align__(16) float sw_[];
    const int K = p.KH * KW * p.Cin;
    const int SL = tconv_slots(p.Cin);
    for (int i = threadIdx.x; i < K * CO; i += blockDim.x) {
        const int k = i / CO, co = i - k * CO;
        const int tap = k / p.Cin, c = k - tap * p.Cin;
        sw_[((tap * 4 + (c & 3)) * SL + tconv_swz(c >> 2)) * CO + co] = co < p.Cout ? __ldg(p.w + (long long)co * K + k) : 0.f;
    }
    __syncthreads();
    const int j = threadIdx.x & 7, slot = threadIdx.x >> 3;
    const bool two = p.x2 != nullptr;
    const unsigned gpr = ((unsigned)p.OW + P - 1) / P;                 // pixel groups per output row
    const unsigned G = ((unsigned)p.M / (unsigned)p.OW) * gpr;
    for (unsigned base = blockIdx.x * 32u; base < G; base += gridDim.x * 32u) {
        const unsigned g = base + slot;
        const bool valid = g < G;
        const unsigned t = g / gpr, og = g - t * gpr;                  // t = nb * OH + oh
        const unsigned nb = t / (unsigned)p.OH, oh = t - nb * (unsigned)p.OH;
        const int ow0 = (int)og * P;
        float acc[P][CO];
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int c = 0; c < CO; ++c) acc[i][c] = 0.f;
        if (valid) {
            const int iw0 = ow0 + p.pw - (KW - 1);                     // input column of (pixel 0, tap KW - 1)
            for (int kh = 0; kh < p.KH; ++kh) {
                const int nh = (int)oh + p.ph - kh;
                if (nh < 0) continue;
                const int ih = tconv_src(nh, p.sh);
                if (ih < 0 || ih >= p.H) continue;
                const long long rowp = ((long long)nb * p.H + ih) * p.W;
                for (int q = j * 4; q < p.Cx; q += 32) {
                    float4 v[NC][2];
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const int iw = iw0 + c;
                        const bool ok = iw >= 0 && iw < p.W;
                        const long long off = (rowp + iw) * p.Cx + q;
                        v[c][0] = ok ? __ldg(reinterpret_cast<const float4*>(p.x + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        v[c][1] = (ok && two) ? __ldg(reinterpret_cast<const float4*>(p.x2 + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int kw = 0; kw < KW; ++kw) {
                        const float* wt = sw_ + (kh * KW + kw) * 4 * SL * CO;
#pragma unroll
                        for (int tn = 0; tn < 2; ++tn) {
                            if (tn && !two) break;
                            int c = q;
                            if (p.cat_c) {
                                const int half = q >= p.cat_c;
                                c = (2 * half + tn) * p.cat_c + q - half * p.cat_c;
                            }
                            const float* wp = wt + tconv_swz(c >> 2) * CO;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
#pragma unroll
                                for (int g4 = 0; g4 < CO / 4; ++g4) {
                                    const float4 wv = *reinterpret_cast<const float4*>(wp + e * SL * CO + 4 * g4);
#pragma unroll
                                    for (int i = 0; i < P; ++i) {
                                        const float4 vv = v[i + KW - 1 - kw][tn];       // pixel i, tap kw -> column i + 2 - kw
                                        const float xv = e == 0 ? vv.x : e == 1 ? vv.y : e == 2 ? vv.z : vv.w;
                                        acc[i][4 * g4 + 0] = fmaf(xv, wv.x, acc[i][4 * g4 + 0]);
                                        acc[i][4 * g4 + 1] = fmaf(xv, wv.y, acc[i][4 * g4 + 1]);
                                        acc[i][4 * g4 + 2] = fmaf(xv, wv.z, acc[i][4 * g4 + 2]);
                                        acc[i][4 * g4 + 3] = fmaf(xv, wv.w, acc[i][4 * g4 + 3]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        // sum the 8 channel slices; afterwards lane j owns pixel j / 2, channels (j % 2) * CO / 2 ...
        float mine[CO / 2];
#pragma unroll
        for (int c = 0; c < CO / 2; ++c) mine[c] = 0.f;
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int c = 0; c < CO; ++c) {
                float v = acc[i][c];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (i == (j >> 1) && c / (CO / 2) == (j & 1)) mine[c % (CO / 2)] = v;
            }
        const int pi = j >> 1;
        if (valid && ow0 + pi < p.OW) {
            const long long m = (long long)t * p.OW + ow0 + pi;
#pragma unroll
            for (int c = 0; c < CO / 2; ++c) {
                const int n = (j & 1) * (CO / 2) + c;
                if (n >= p.Cout) continue;
                float v = mine[c] + (p.e.bias ? __ldg(p.e.bias + n) : 0.f);
                v = apply_act(v, p.e.act, p.e, n);
                if (p.e.post_scale) v = fmaf(v, __ldg(p.e.post_scale + n), __ldg(p.e.post_shift + n));
                v *= p.e.alpha;
                if (p.e.res) v = fmaf(p.e.beta, __ldg(p.e.res + m * p.e.ldres + n), v);
                p.e.out[m * p.e.ldo + n] = v;
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
        // 3x3, 1 or 2 input channels, Cout / 4 a power of two, plain bias + {none, relu, leaky}: weights-in-registers kernel
        const bool thin = kernel_h == 3 && kernel_w == 3 && dil_h == 1 && dil_w == 1 && (in_channels == 1 || in_channels == 2) &&
                          (ncols & 3) == 0 && cg <= 256 && (cg & (cg - 1)) == 0 && !epi->post_scale &&
                          (e.act == ACT_NONE || e.act == ACT_RELU || e.act == ACT_LEAKY) && (e.ldo & 3) == 0 &&
                          ((uintptr_t)out & 15) == 0 && (!e.bias || ((uintptr_t)e.bias & 15) == 0) && !getenv("APS_B200_NO_THIN_CONV") &&
                          (size_t)2 * 3 * (((width + 2 * pad_w) * in_channels + 3) & ~3LL) * sizeof(float) <= 48 * 1024;   // staged rows
        if (thin) {
            const long long rows = batch * OH;
            const unsigned tg = (unsigned)(rows < (long long)num_sms() * 3 ? rows : (long long)num_sms() * 3);
            c.rw = (int)(((width + 2 * pad_w) * in_channels + 3) & ~3LL);
            const size_t tsm = (size_t)2 * 3 * c.rw * sizeof(float);
#define APSB_THIN(CI, A) conv2d_thin3x3_kernel<CI, A><<<tg, 256, tsm, st>>>(c)
            if (in_channels == 1) {
                if (e.act == ACT_RELU) APSB_THIN(1, ACT_RELU); else if (e.act == ACT_LEAKY) APSB_THIN(1, ACT_LEAKY); else APSB_THIN(1, ACT_NONE);
            } else {
                if (e.act == ACT_RELU) APSB_THIN(2, ACT_RELU); else if (e.act == ACT_LEAKY) APSB_THIN(2, ACT_LEAKY); else APSB_THIN(2, ACT_NONE);
            }
#undef APSB_THIN
            APSB_LAUNCH_CHECK();
            return 0;
        }
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

extern "C" int aps_b200_conv_transpose2d_nhwc_narrow_fwd(const float* x, const float* x_skip, int64_t batch,
                                                         int64_t height, int64_t width, int64_t in_channels,
                                                         const float* weight, int64_t out_channels, int kernel_h,
                                                         int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                                                         int out_pad_h, int out_pad_w, const aps_b200_epilogue* epi,
                                                         float* out, void* stream) {
    APSB_CHECK_ARG(x && weight, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && height > 0 && width > 0 && in_channels > 0 && out_channels > 0, "bad shape");
    APSB_CHECK_ARG(kernel_h > 0 && kernel_w > 0 && stride_h > 0 && stride_w > 0 && pad_h >= 0 && pad_w >= 0 &&
                       out_pad_h >= 0 && out_pad_w >= 0, "bad convolution geometry");
    APSB_CHECK_ARG(out_channels <= 8, "narrow transposed convolution: at most 8 output channels (got %lld)",
                   (long long)out_channels);
    const int64_t cx = x_skip ? in_channels / 2 : in_channels;      // channels of each tensor that is read
    APSB_CHECK_ARG(cx % (x_skip ? 8 : 4) == 0 && (!x_skip || in_channels % 2 == 0) && ((uintptr_t)x & 15) == 0 &&
                       ((uintptr_t)x_skip & 15) == 0,
                   "narrow transposed convolution: %lld channels per tensor need 16-byte aligned float4 groups",
                   (long long)cx);
    const int64_t OH = (height - 1) * stride_h - 2 * pad_h + kernel_h + out_pad_h;
    const int64_t OW = (width - 1) * stride_w - 2 * pad_w + kernel_w + out_pad_w;
    APSB_CHECK_ARG(OH > 0 && OW > 0, "transposed convolution output is empty");
    const int64_t M = batch * OH * OW, K = (int64_t)kernel_h * kernel_w * in_channels;
    const int co = out_channels <= 4 ? 4 : 8;
    const size_t smem = (size_t)kernel_h * kernel_w * 4 * tconv_slots((int)in_channels) * co * 4;
    APSB_CHECK_ARG(M < (1LL << 31) - 64 && K < (1LL << 24) && smem <= 48 * 1024, "shape too large for the narrow kernel");
    NarrowTConvParams c{};
    const int ncols = (int)out_channels;
    if (int rc = fill_epilogue(c.e, epi, ncols, out, ncols)) return rc;
    APSB_CHECK_ARG(epi->act != ACT_GLU, "GLU is not available here");
    c.x = x; c.x2 = x_skip; c.H = (int)height; c.W = (int)width; c.Cx = (int)cx; c.cat_c = x_skip ? (int)(cx / 2) : 0;
    c.Cin = (int)in_channels; c.KH = kernel_h; c.KW = kernel_w; c.sh = stride_h; c.sw = stride_w; c.ph = pad_h;
    c.pw = pad_w; c.OH = (int)OH; c.OW = (int)OW; c.M = (int)M; c.Cout = ncols; c.w = weight;
    const long long blocks = (M + 31) / 32;
    const unsigned grid = (unsigned)(blocks < (long long)num_sms() * 16 ? blocks : (long long)num_sms() * 16);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool no_tiled = getenv("APS_B200_TCONV_NARROW_1PX") != nullptr;     // A/B switch
    if (kernel_w == 3 && stride_w == 1 && !no_tiled) {
        const long long groups = (M / OW) * ((OW + 3) / 4), gblocks = (groups + 31) / 32;
        const unsigned tgrid = (unsigned)(gblocks < (long long)num_sms() * 16 ? gblocks : (long long)num_sms() * 16);
        if (co == 4) tconv_narrow_tiled_kernel<4><<<tgrid, 256, smem, st>>>(c);
        else tconv_narrow_tiled_kernel<8><<<tgrid, 256, smem, st>>>(c);
    } else if (kernel_w == 3) {
        if (co == 4) tconv_narrow_kernel<4, 3><<<grid, 256, smem, st>>>(c);
        else tconv_narrow_kernel<8, 3><<<grid, 256, smem, st>>>(c);
    } else {
        if (co == 4) tconv_narrow_kernel<4, 0><<<grid, 256, smem, st>>>(c);
        else tconv_narrow_kernel<8, 0><<<grid, 256, smem, st>>>(c);
    }
    APSB_LAUNCH_CHECK();
    return 0;
}

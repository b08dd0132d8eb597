// encoder.cu — the non-GEMM kernels of the transformer / conformer encoder forward (rows a21–a23):
//   layernorm       : y = LN(alpha * x + res) * gamma + beta  (residual add + macaron scale fused)
//   mhsa            : multi-head self-attention with absolute / relative ("rel") / Transformer-XL ("xl")
//                     position terms, key-padding and additive masks, streaming (online) softmax
//   dwconv1d        : depthwise convolution over time (+ folded BatchNorm + activation) on token-major rows
//
// Replaces /root/reference/aps/asr/transformer/impl.py:95-118 (context_weight), :120-131 / :240-261 /
// :324-344 (dot_att variants; the pad/transpose/view skew of aps/asr/transformer/utils.py:14-39
// `digit_shift` becomes an index: shifted[l, s] = term[l, s - l + L - 1]), impl.py:454-465 (conformer
// depthwise conv + BatchNorm1d + Swish), impl.py:392-393 / :476-480 (LayerNorm with residual), and the
// dilated depthwise convolutions of aps/sse/bss/tcn.py:141-151.
//
// Activations are token-major rows [N*T, D] with row(n, t) = n*stride_n + t*stride_t (the encoder keeps
// batch-major order so no N<->T transposes are materialised).
#include "../../include/aps_b200.h"
#include "common.cuh"
#include "gemm.cuh"

#include <math.h>
#include <stdlib.h>

namespace apsb {

// ------------------------------------------------------------------------------------------------
// one warp per row
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long ldx,
                                                        const float* __restrict__ res, long long ldr, float alpha,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, long long M, int D,
                                                        float* __restrict__ out, long long ldo) {
    const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    const float* xr = x + m * ldx;
    const float* rr = res ? res + m * ldr : nullptr;
    auto val = [&](int d) { return rr ? fmaf(alpha, xr[d], rr[d]) : alpha * xr[d]; };
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += val(d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float q = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float c = val(d) - mean;
        q = fmaf(c, c, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float inv = rsqrtf(q / (float)D + eps);
    float* orow = out + m * ldo;
    for (int d = lane; d < D; d += 32) {
        float y = (val(d) - mean) * inv;
        if (gamma) y = fmaf(y, __ldg(gamma + d), beta ? __ldg(beta + d) : 0.f);
        orow[d] = y;
    }
}

// LayerNorm that also finishes a split-K GEMM: v = alpha * (sum_p x[p] + bias) + res, y = LN(v) * gamma + beta, and
// writes the TF32 "lo" companion of y for a TMA-fed tensor-core consumer (tc_gemm.cu MODE 3).  One warp per row,
// 16-byte vector traffic, the row stays in registers between the statistics and the normalisation (D <= 1024,
// D % 128 == 0: the encoder widths); `gamma == nullptr` with `do_norm == 0` makes it a plain reduce + bias + residual.
__device__ __forceinline__ float ln_tf32_lo(float x) {
    const float l = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    return __uint_as_float((__float_as_uint(l) + 0x1000u) & 0xffffe000u);
}

template <int VPL>   // float4 vectors per lane: D = 128 * VPL
__global__ void __launch_bounds__(256) layernorm2_kernel(const float* __restrict__ x, long long ldx, int nparts,
                                                         long long part_stride, const float* __restrict__ bias,
                                                         const float* __restrict__ res, long long ldr, float alpha,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float eps, int do_norm, long long M, float* __restrict__ out,
                                                         float* __restrict__ out_lo, long long ldo) {
    pdl_trigger();
    pdl_wait();
    const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    constexpr int D = 128 * VPL;
    float4 v[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const int d = 4 * lane + 128 * j;
        float4 a = __ldg(reinterpret_cast<const float4*>(x + m * ldx + d));
        if (nparts > 1) {
            // split-K slices: all loads first (up to 7 in flight instead of one L2 round trip per slice), then the adds in
            // slice order — deterministic, and x + 0 leaves the sum of a shorter split untouched
            float4 b[7];
#pragma unroll
            for (int p = 1; p < 8; ++p)
                b[p - 1] = p < nparts ? __ldg(reinterpret_cast<const float4*>(x + p * part_stride + m * ldx + d))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int p = 0; p < 7; ++p) { a.x += b[p].x; a.y += b[p].y; a.z += b[p].z; a.w += b[p].w; }
            for (int p = 8; p < nparts; ++p) {
                const float4 c = __ldg(reinterpret_cast<const float4*>(x + p * part_stride + m * ldx + d));
                a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
            }
        }
        if (bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + d));
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        if (res) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(res + m * ldr + d));
            a.x = fmaf(alpha, a.x, r.x); a.y = fmaf(alpha, a.y, r.y); a.z = fmaf(alpha, a.z, r.z); a.w = fmaf(alpha, a.w, r.w);
        } else {
            a.x *= alpha; a.y *= alpha; a.z *= alpha; a.w *= alpha;
        }
        v[j] = a;
    }
    if (do_norm) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)D;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
            q = fmaf(a, a, q); q = fmaf(b, b, q); q = fmaf(c, c, q); q = fmaf(d, d, q);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float inv = rsqrtf(q / (float)D + eps);
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int d = 4 * lane + 128 * j;
            float4 y = make_float4((v[j].x - mean) * inv, (v[j].y - mean) * inv, (v[j].z - mean) * inv, (v[j].w - mean) * inv);
            if (gamma) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + d));
                const float4 b = beta ? __ldg(reinterpret_cast<const float4*>(beta + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
                y.x = fmaf(y.x, g.x, b.x); y.y = fmaf(y.y, g.y, b.y); y.z = fmaf(y.z, g.z, b.z); y.w = fmaf(y.w, g.w, b.w);
            }
            v[j] = y;
        }
    }
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const int d = 4 * lane + 128 * j;
        *reinterpret_cast<float4*>(out + m * ldo + d) = v[j];
        if (out_lo)
            *reinterpret_cast<float4*>(out_lo + m * ldo + d) =
                make_float4(ln_tf32_lo(v[j].x), ln_tf32_lo(v[j].y), ln_tf32_lo(v[j].z), ln_tf32_lo(v[j].w));
    }
}

// ------------------------------------------------------------------------------------------------
struct DwParams {
    const float* x;
    long long ldx;
    const float* w;      // [Kw, D] (tap-major: coalesced over channels)
    const float* bias;   // [D] or nullptr
    int N, T, D, Kw, dil, lpad;   // out[t] = sum_k w[k] * x[t - lpad + k*dil]
    long long sn, st;    // row(n, t) = n*sn + t*st
    Epilogue e;          // e.out_lo: optional TF32 "lo" companion of the output (V == 4 path)
    const long long* lens;   // optional [N]: frames t >= lens[n] read as ZERO (ragged batches: an utterance must see the
                             // zero padding it would see alone, not its neighbours' padded activations)
};

// thread = (row, group of V channels); rows are walked with 32-bit arithmetic (64-bit divisions per element made the
// first version 6x slower than its HBM bound) and, when V == 4, 16-byte vector loads / stores.
template <int V>
__global__ void __launch_bounds__(256) dwconv1d_kernel(const __grid_constant__ DwParams p) {
    pdl_trigger();
    pdl_wait();
    const unsigned groups = (unsigned)(p.D / V);                 // channel groups per row (D % V == 0)
    const unsigned rows_per_block = 256u / groups > 0 ? 256u / groups : 1u;
    const unsigned g = threadIdx.x % groups, rb = threadIdx.x / groups;
    if (groups <= 256 && rb >= rows_per_block) return;
    const unsigned total_rows = (unsigned)p.N * (unsigned)p.T;
    for (unsigned gi = g; gi < groups; gi += 256) {              // groups > 256: one row per block, several passes
        const unsigned r = blockIdx.x * rows_per_block + rb;
        if (r >= total_rows) return;
        const unsigned n = r / (unsigned)p.T, t = r - n * (unsigned)p.T;
        const int d = (int)gi * V;
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = p.bias ? __ldg(p.bias + d + j) : 0.f;
        const int tlim = p.lens ? (int)min((long long)p.T, p.lens[n]) : p.T;
        for (int k = 0; k < p.Kw; ++k) {
            const int tt = (int)t - p.lpad + k * p.dil;
            if (tt >= 0 && tt < tlim) {
                const float* xp = p.x + ((long long)n * p.sn + (long long)tt * p.st) * p.ldx + d;
                const float* wp = p.w + (long long)k * p.D + d;
                if (V == 4) {
                    const float4 xv = __ldg(reinterpret_cast<const float4*>(xp)), wv = __ldg(reinterpret_cast<const float4*>(wp));
                    acc[0] = fmaf(wv.x, xv.x, acc[0]); acc[1 % V] = fmaf(wv.y, xv.y, acc[1 % V]);
                    acc[2 % V] = fmaf(wv.z, xv.z, acc[2 % V]); acc[3 % V] = fmaf(wv.w, xv.w, acc[3 % V]);
                } else {
                    acc[0] = fmaf(__ldg(wp), __ldg(xp), acc[0]);
                }
            }
        }
        const long long m = (long long)n * p.sn + (long long)t * p.st;
        float o[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float v = apply_act(acc[j], p.e.act, p.e, d + j);
            if (p.e.post_scale) v = fmaf(v, __ldg(p.e.post_scale + d + j), __ldg(p.e.post_shift + d + j));
            v *= p.e.alpha;
            if (p.e.res) v = fmaf(p.e.beta, __ldg(p.e.res + m * p.e.ldres + d + j), v);
            o[j] = v;
        }
        if (V == 4) {
            *reinterpret_cast<float4*>(p.e.out + m * p.e.ldo + d) = make_float4(o[0], o[1 % V], o[2 % V], o[3 % V]);
            if (p.e.out_lo)
                *reinterpret_cast<float4*>(p.e.out_lo + m * p.e.ldo + d) =
                    make_float4(ln_tf32_lo(o[0]), ln_tf32_lo(o[1 % V]), ln_tf32_lo(o[2 % V]), ln_tf32_lo(o[3 % V]));
        } else {
            p.e.out[m * p.e.ldo + d] = o[0];
        }
    }
}

// ------------------------------------------------------------------------------------------------
struct AttnParams {
    const float* q;      // rows (n, t), head h at columns h*dh .. ; row stride ldq
    const float* k;
    const float* v;
    const float* qpos;   // rows used for the position term (= q, or v for the reference's xl path, Q14)
    long long ldq, ldk, ldv, ldqp;
    long long sn, st;    // row(n, t) = n*sn + t*st (rows of q/k/v/out)
    int N, L, H, dh;
    int mode;            // 0 abs, 1 rel (pos [2L-1, dh] shared by heads), 2 xl (pos [2L-1, H*dh])
    const float* pos;
    long long ldpos;
    const float* rel_u;  // xl: [H, dh] added to the content query
    const float* rel_v;  // xl: [H, dh] added to the position query
    const unsigned char* kpm;  // [N, L] key padding mask (1 = masked) or nullptr
    float kpm_fill;      // value written on masked keys (MIN_F32 for the aps path, -inf for torch's)
    const float* amask;  // [L, L] additive mask or nullptr
    float scale;         // 1/sqrt(dh)
    float* out;          // rows (n, t), columns h*dh + d
    float* out_lo;       // optional TF32 "lo" companion of out
    long long ldo;
};

constexpr int kAttnWarps = 16;   // queries per CTA: the K / V / position tiles are re-read once per query tile, so fewer, larger tiles
constexpr int kKeyTile = 32;

// CTA = (query tile, head, batch); one warp per query; lane = key inside the tile; dh <= 128
template <int DH>
__global__ void __launch_bounds__(kAttnWarps * 32) mhsa_kernel(const __grid_constant__ AttnParams p) {
    constexpr int LD = DH + 1;
    __shared__ float sK[kKeyTile][LD];
    __shared__ float sV[kKeyTile][LD];
    __shared__ float sR[kKeyTile + kAttnWarps][LD];
    __shared__ float sQ[kAttnWarps][DH];
    __shared__ float sQp[kAttnWarps][DH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kAttnWarps;
    const int i = i0 + warp;
    const bool qok = i < p.L;
    const int L = p.L;
    // queries of this CTA
    for (int idx = threadIdx.x; idx < kAttnWarps * DH; idx += blockDim.x) {
        const int w = idx / DH, d = idx - w * DH;
        const int ii = min(i0 + w, L - 1);
        const long long row = n * p.sn + ii * p.st;
        float qc = __ldg(p.q + row * p.ldq + h * DH + d);
        float qp = p.mode ? __ldg(p.qpos + row * p.ldqp + h * DH + d) : 0.f;
        if (p.mode == 2) {
            // reference xl path: content term uses (X + u), position term (X + v), X = the tensor passed as
            // "query" to dot_att — which is VALUE in the reference (impl.py:369)
            qc = __ldg(p.qpos + row * p.ldqp + h * DH + d) + __ldg(p.rel_u + h * DH + d);
            qp = qp + __ldg(p.rel_v + h * DH + d);
        }
        sQ[w][d] = qc;
        sQp[w][d] = qp;
    }
    float mrun = -INFINITY, lrun = 0.f;
    float acc[(DH + 31) / 32];
#pragma unroll
    for (int c = 0; c < (DH + 31) / 32; ++c) acc[c] = 0.f;

    for (int j0 = 0; j0 < L; j0 += kKeyTile) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < kKeyTile * DH; idx += blockDim.x) {
            const int j = idx / DH, d = idx - j * DH;
            const int jj = j0 + j;
            float kv = 0.f, vv = 0.f;
            if (jj < L) {
                const long long row = n * p.sn + jj * p.st;
                kv = __ldg(p.k + row * p.ldk + h * DH + d);
                vv = __ldg(p.v + row * p.ldv + h * DH + d);
            }
            sK[j][d] = kv;
            sV[j][d] = vv;
        }
        if (p.mode) {
            // position rows x = j - i + L - 1 for i in [i0, i0+W), j in [j0, j0+32): x0 = j0 - (i0+W-1) + L - 1
            const int x0 = j0 - (i0 + kAttnWarps - 1) + L - 1;
            for (int idx = threadIdx.x; idx < (kKeyTile + kAttnWarps) * DH; idx += blockDim.x) {
                const int r = idx / DH, d = idx - r * DH;
                const int x = x0 + r;
                float pv = 0.f;
                if (x >= 0 && x < 2 * L - 1)
                    pv = __ldg(p.pos + (long long)x * p.ldpos + (p.mode == 2 ? h * DH : 0) + d);
                sR[r][d] = pv;
            }
        }
        __syncthreads();
        const int j = j0 + lane;
        float s = -INFINITY;
        if (qok && j < L) {
            float a = 0.f;
#pragma unroll 8
            for (int d = 0; d < DH; ++d) a = fmaf(sQ[warp][d], sK[lane][d], a);
            if (p.mode) {
                // row index inside sR: x - x0 = (j - i + L - 1) - x0 = lane + (kAttnWarps - 1 - warp)
                const float* rr = sR[lane + kAttnWarps - 1 - warp];
                float b = 0.f;
#pragma unroll 8
                for (int d = 0; d < DH; ++d) b = fmaf(sQp[warp][d], rr[d], b);
                a += b;
            }
            s = a * p.scale;
            if (p.kpm && p.kpm[(long long)n * L + j]) s = p.kpm_fill;
            if (p.amask) s += __ldg(p.amask + (long long)i * L + j);
        }
        // online softmax update
        float tmax = s;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        const float mnew = fmaxf(mrun, tmax);
        const float corr = (mrun == -INFINITY) ? 0.f : __expf(mrun - mnew);
        const float pj = (s == -INFINITY) ? 0.f : __expf(s - mnew);
        float psum = pj;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
        if (mnew != -INFINITY) {
            lrun = lrun * corr + psum;
#pragma unroll
            for (int c = 0; c < (DH + 31) / 32; ++c) acc[c] *= corr;
            for (int jj = 0; jj < kKeyTile; ++jj) {
                const float pb = __shfl_sync(0xffffffffu, pj, jj);
#pragma unroll
                for (int c = 0; c < (DH + 31) / 32; ++c) {
                    const int d = lane + 32 * c;
                    if (d < DH) acc[c] = fmaf(pb, sV[jj][d], acc[c]);
                }
            }
            mrun = mnew;
        }
    }
    if (qok) {
        // a fully masked row (all -inf) gives NaN in the reference's softmax too
        const float inv = 1.f / lrun;
        float* o = p.out + (n * p.sn + i * p.st) * p.ldo + h * DH;
#pragma unroll
        for (int c = 0; c < (DH + 31) / 32; ++c) {
            const int d = lane + 32 * c;
            if (d < DH) {
                const float y = (lrun > 0.f) ? acc[c] * inv : NAN;
                o[d] = y;
                if (p.out_lo) p.out_lo[(n * p.sn + i * p.st) * p.ldo + h * DH + d] = ln_tf32_lo(y);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Register-tiled attention (round 2).  ncu of mhsa_kernel at the BASELINE size (profiles/r02_ncu_enc_small.txt): LSU pipe
// 63 %, 8.6 M shared-memory wavefronts — one LDS per FMA operand.  Here a CTA owns 64 queries of one (batch, head), walks
// the keys in tiles of 64 with a streaming softmax, and every thread computes a 4 x 4 block of the 64 x 64 score tile from
// 16-byte shared-memory reads (8 FMA per LDS instead of 0.5): query rows ty + 16a, key rows tx + 16b — for a fixed a / b the
// rows read by the lanes of a quarter warp are consecutive, which is conflict free with 68-float rows.  The relative
// position term q . R[j - i + L - 1] (the reference's pad / transpose skew, utils.py:14-39) needs rows
// (tx - ty) + 16 (b - a) + 63 of the staged window of R: 7 distinct rows per thread.  P . V is tiled the same way
// (queries ty + 16a, 4 head dims per thread).  Same arithmetic order per row whatever the batch.
// NT = tile edge / 16: 4 in general, 3 (48 x 48) when the whole sequence fits one such tile — the BASELINE encoder has
// L = 48, where a 64 x 64 tile spends 44 % of its FMAs on padding.  One tile either way, so the order of every sum is the same.
template <int DH, int NT>
__global__ void __launch_bounds__(256, 2) mhsa_tiled_kernel(const __grid_constant__ AttnParams p) {
    constexpr int kTQ = 16 * NT, kTK = 16 * NT;
    constexpr int LD = DH + 4;              // floats per staged row: 16-byte aligned, conflict free for 8 consecutive rows
    constexpr int C4 = DH / 4;
    extern __shared__ float4 attn_smem4[];
    float* sQ = reinterpret_cast<float*>(attn_smem4);
    float* sK = sQ + kTQ * LD;
    float* sV = sK + kTK * LD;
    float* sR = sV + kTK * LD;              // kTQ + kTK rows (window of the position table), modes 1 / 2
    float* sQp = sR + (p.mode ? (kTQ + kTK) * LD : 0);   // position query, mode 2 only (mode 1: the content query)
    float* sS = sK;                         // the score / probability tile re-uses the key tile (rows of LD floats)
    __shared__ float sMax[kTQ], sSum[kTQ], sCorr[kTQ];
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kTQ;
    const int L = p.L;
    // ---- queries of this CTA ----
    for (int idx = tid; idx < kTQ * C4; idx += 256) {
        const int r = idx / C4, c = (idx - r * C4) * 4;
        const int ii = min(i0 + r, L - 1);
        const long long row = (long long)n * p.sn + (long long)ii * p.st;
        float4 qc = __ldg(reinterpret_cast<const float4*>(p.q + row * p.ldq + h * DH + c));
        if (p.mode == 2) {
            // reference xl path: content term uses (X + u), position term (X + v), X = the tensor passed as "query" to
            // dot_att — which is VALUE in the reference (impl.py:369)
            const float4 x = __ldg(reinterpret_cast<const float4*>(p.qpos + row * p.ldqp + h * DH + c));
            const float4 u = __ldg(reinterpret_cast<const float4*>(p.rel_u + h * DH + c));
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.rel_v + h * DH + c));
            qc = make_float4(x.x + u.x, x.y + u.y, x.z + u.z, x.w + u.w);
            *reinterpret_cast<float4*>(sQp + r * LD + c) = make_float4(x.x + v.x, x.y + v.y, x.z + v.z, x.w + v.w);
        }
        *reinterpret_cast<float4*>(sQ + r * LD + c) = qc;
    }
    if (tid < kTQ) { sMax[tid] = -INFINITY; sSum[tid] = 0.f; }
    const float* sQq = p.mode == 2 ? sQp : sQ;
    float acc[NT][4];
#pragma unroll
    for (int a = 0; a < NT; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;

    for (int j0 = 0; j0 < L; j0 += kTK) {
        __syncthreads();                    // the previous tile's P . V is done with sV and sS
        for (int idx = tid; idx < kTK * C4; idx += 256) {
            const int r = idx / C4, c = (idx - r * C4) * 4;
            const int jj = j0 + r;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (jj < L) {
                const long long row = (long long)n * p.sn + (long long)jj * p.st;
                kv = __ldg(reinterpret_cast<const float4*>(p.k + row * p.ldk + h * DH + c));
                vv = __ldg(reinterpret_cast<const float4*>(p.v + row * p.ldv + h * DH + c));
            }
            *reinterpret_cast<float4*>(sK + r * LD + c) = kv;
            *reinterpret_cast<float4*>(sV + r * LD + c) = vv;
        }
        if (p.mode) {
            // window row w <-> table row x = j - i + L - 1 with w = (j - j0) - (i - i0) + kTQ - 1
            const int x0 = j0 - i0 - (kTQ - 1) + L - 1;
            for (int idx = tid; idx < (kTQ + kTK - 1) * C4; idx += 256) {
                const int r = idx / C4, c = (idx - r * C4) * 4;
                const int x = x0 + r;
                float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (x >= 0 && x < 2 * L - 1)
                    pv = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)x * p.ldpos + (p.mode == 2 ? h * DH : 0) + c));
                *reinterpret_cast<float4*>(sR + r * LD + c) = pv;
            }
        }
        __syncthreads();
        // ---- scores: s[a][b] = q[ty + 16a] . k[tx + 16b] (+ position term) ----
        float s[NT][NT];
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) s[a][b] = 0.f;
        const int wbase = tx - ty + (kTQ - 1) - 16 * (NT - 1);      // window row of (a, b): wbase + 16 (b - a + NT - 1)
#pragma unroll 2
        for (int d = 0; d < DH; d += 4) {
            float4 q4[NT], k4[NT];
#pragma unroll
            for (int a = 0; a < NT; ++a) q4[a] = *reinterpret_cast<const float4*>(sQ + (ty + 16 * a) * LD + d);
#pragma unroll
            for (int b = 0; b < NT; ++b) k4[b] = *reinterpret_cast<const float4*>(sK + (tx + 16 * b) * LD + d);
#pragma unroll
            for (int a = 0; a < NT; ++a)
#pragma unroll
                for (int b = 0; b < NT; ++b)
                    s[a][b] = fmaf(q4[a].x, k4[b].x, fmaf(q4[a].y, k4[b].y, fmaf(q4[a].z, k4[b].z, fmaf(q4[a].w, k4[b].w, s[a][b]))));
            if (p.mode) {
                float4 r4[2 * NT - 1];
#pragma unroll
                for (int c = 0; c < 2 * NT - 1; ++c) {
                    const int w = wbase + 16 * c;
                    r4[c] = (w >= 0 && w < kTQ + kTK - 1) ? *reinterpret_cast<const float4*>(sR + w * LD + d)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (p.mode == 2) {
#pragma unroll
                    for (int a = 0; a < NT; ++a) q4[a] = *reinterpret_cast<const float4*>(sQq + (ty + 16 * a) * LD + d);
                }
#pragma unroll
                for (int a = 0; a < NT; ++a)
#pragma unroll
                    for (int b = 0; b < NT; ++b) {
                        const float4 r = r4[b - a + NT - 1];
                        s[a][b] = fmaf(q4[a].x, r.x, fmaf(q4[a].y, r.y, fmaf(q4[a].z, r.z, fmaf(q4[a].w, r.w, s[a][b]))));
                    }
            }
        }
        __syncthreads();                    // every thread is done with the key tile: it becomes the score tile
#pragma unroll
        for (int a = 0; a < NT; ++a) {
            const int i = i0 + ty + 16 * a;
#pragma unroll
            for (int b = 0; b < NT; ++b) {
                const int j = j0 + tx + 16 * b;
                float v = -INFINITY;
                if (i < L && j < L) {
                    v = s[a][b] * p.scale;
                    if (p.kpm && p.kpm[(long long)n * L + j]) v = p.kpm_fill;
                    if (p.amask) v += __ldg(p.amask + (long long)i * L + j);
                }
                sS[(ty + 16 * a) * LD + tx + 16 * b] = v;
            }
        }
        __syncthreads();
        // ---- streaming softmax: warp w owns rows 8w .. 8w+7, lane = keys lane, lane + 32 ----
#pragma unroll
        for (int rr = 0; rr < kTQ / 8; ++rr) {
            const int r = warp * (kTQ / 8) + rr;
            const float v0 = sS[r * LD + lane], v1 = (lane + 32 < kTK) ? sS[r * LD + lane + 32] : -INFINITY;
            float tmax = fmaxf(v0, v1);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
            const float mold = sMax[r];
            const float mnew = fmaxf(mold, tmax);
            const float p0 = (v0 == -INFINITY) ? 0.f : __expf(v0 - mnew);
            const float p1 = (v1 == -INFINITY) ? 0.f : __expf(v1 - mnew);
            float psum = p0 + p1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
            sS[r * LD + lane] = p0;
            if (lane + 32 < kTK) sS[r * LD + lane + 32] = p1;
            if (lane == 0) {
                const float corr = (mold == -INFINITY) ? 0.f : __expf(mold - mnew);
                sCorr[r] = (mnew == -INFINITY) ? 1.f : corr;
                if (mnew != -INFINITY) { sSum[r] = sSum[r] * corr + psum; sMax[r] = mnew; }
            }
        }
        __syncthreads();
        // ---- acc[a][c] += P[ty + 16a][:] . V[:][4tx + c] ----
#pragma unroll
        for (int a = 0; a < NT; ++a) {
            const float corr = sCorr[ty + 16 * a];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] *= corr;
        }
#pragma unroll 2
        for (int j = 0; j < kTK; j += 4) {
            float4 p4[NT], v4[4];
#pragma unroll
            for (int a = 0; a < NT; ++a) p4[a] = *reinterpret_cast<const float4*>(sS + (ty + 16 * a) * LD + j);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) v4[jj] = *reinterpret_cast<const float4*>(sV + (j + jj) * LD + 4 * tx);
#pragma unroll
            for (int a = 0; a < NT; ++a) {
                acc[a][0] = fmaf(p4[a].x, v4[0].x, fmaf(p4[a].y, v4[1].x, fmaf(p4[a].z, v4[2].x, fmaf(p4[a].w, v4[3].x, acc[a][0]))));
                acc[a][1] = fmaf(p4[a].x, v4[0].y, fmaf(p4[a].y, v4[1].y, fmaf(p4[a].z, v4[2].y, fmaf(p4[a].w, v4[3].y, acc[a][1]))));
                acc[a][2] = fmaf(p4[a].x, v4[0].z, fmaf(p4[a].y, v4[1].z, fmaf(p4[a].z, v4[2].z, fmaf(p4[a].w, v4[3].z, acc[a][2]))));
                acc[a][3] = fmaf(p4[a].x, v4[0].w, fmaf(p4[a].y, v4[1].w, fmaf(p4[a].z, v4[2].w, fmaf(p4[a].w, v4[3].w, acc[a][3]))));
            }
        }
    }
    // ---- normalise and store (4 consecutive head dims per thread: 16-byte stores) ----
    if (4 * tx < DH) {
#pragma unroll
        for (int a = 0; a < NT; ++a) {
            const int i = i0 + ty + 16 * a;
            if (i >= L) continue;
            const float l = sSum[ty + 16 * a];
            // a fully masked row (all -inf) gives NaN in the reference's softmax too
            const float inv = 1.f / l;
            float4 y = (l > 0.f) ? make_float4(acc[a][0] * inv, acc[a][1] * inv, acc[a][2] * inv, acc[a][3] * inv)
                                 : make_float4(NAN, NAN, NAN, NAN);
            const long long off = ((long long)n * p.sn + (long long)i * p.st) * p.ldo + h * DH + 4 * tx;
            *reinterpret_cast<float4*>(p.out + off) = y;
            if (p.out_lo)
                *reinterpret_cast<float4*>(p.out_lo + off) =
                    make_float4(ln_tf32_lo(y.x), ln_tf32_lo(y.y), ln_tf32_lo(y.z), ln_tf32_lo(y.w));
        }
    }
}

}  // namespace apsb

using namespace apsb;

extern "C" int aps_b200_layernorm_fwd(const float* x, int64_t ld_x, const float* residual, int64_t ld_residual,
                                      float alpha, const float* gamma, const float* beta, float eps, int64_t rows,
                                      int64_t dim, float* out, int64_t ld_out, void* stream) {
    APSB_CHECK_ARG(x && out && rows > 0 && dim > 0 && ld_x >= dim && ld_out >= dim, "bad arguments");
    layernorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        x, ld_x, residual, ld_residual, alpha, gamma, beta, eps, rows, (int)dim, out, ld_out);
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_layernorm2_fwd(const float* x, int64_t ld_x, int32_t num_parts, int64_t part_stride,
                                       const float* bias, const float* residual, int64_t ld_residual, float alpha,
                                       const float* gamma, const float* beta, float eps, int32_t normalize,
                                       int64_t rows, int64_t dim, float* out, float* out_lo, int64_t ld_out,
                                       void* stream) {
    APSB_CHECK_ARG(x && out && rows > 0 && dim > 0 && ld_x >= dim && ld_out >= dim && num_parts >= 1, "bad arguments");
    APSB_CHECK_ARG(dim % 128 == 0 && dim <= 1024, "layernorm2: width %lld is not a multiple of 128 up to 1024", (long long)dim);
    const uintptr_t al = (uintptr_t)x | (uintptr_t)out | (uintptr_t)out_lo | (uintptr_t)bias | (uintptr_t)residual |
                         (uintptr_t)gamma | (uintptr_t)beta;
    APSB_CHECK_ARG((al & 15) == 0 && (ld_x & 3) == 0 && (ld_out & 3) == 0 && (ld_residual & 3) == 0 && (part_stride & 3) == 0,
                   "layernorm2: operands must be 16-byte aligned");
    APSB_CHECK_ARG(num_parts == 1 || part_stride >= rows * ld_x, "layernorm2: part_stride too small");
    const unsigned grid = (unsigned)((rows + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
#define APSB_LN2(V)                                                                                                    \
    case V:                                                                                                            \
        APSB_CUDA(launch_pdl(layernorm2_kernel<V>, dim3(grid), dim3(256), 0, st, x, (long long)ld_x, (int)num_parts,    \
                             (long long)part_stride, bias, residual, (long long)ld_residual, alpha, gamma, beta, eps, \
                             (int)normalize, (long long)rows, out, out_lo, (long long)ld_out));                       \
        break;
    switch ((int)(dim / 128)) {
        APSB_LN2(1) APSB_LN2(2) APSB_LN2(3) APSB_LN2(4) APSB_LN2(5) APSB_LN2(6) APSB_LN2(7) APSB_LN2(8)
    }
#undef APSB_LN2
    APSB_LAUNCH_CHECK();
    return 0;
}

static int dwconv1d_impl(const float* x, int64_t ld_x, int64_t batch, int64_t num_frames, int64_t channels,
                         int64_t stride_n, int64_t stride_t, const float* weight_kd, const float* bias,
                         int kernel, int dilation, int left_pad, const aps_b200_epilogue* epi, float* out,
                         float* out_lo, int64_t ld_out, const int64_t* lens, void* stream) {
    APSB_CHECK_ARG(x && weight_kd && epi && out, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && num_frames > 0 && channels > 0 && kernel > 0 && dilation > 0 && left_pad >= 0,
                   "bad shape");
    APSB_CHECK_ARG(epi->act != ACT_GLU, "GLU is not available for the depthwise convolution");
    DwParams p{};
    p.x = x; p.ldx = ld_x; p.w = weight_kd; p.bias = bias;
    p.N = (int)batch; p.T = (int)num_frames; p.D = (int)channels; p.Kw = kernel; p.dil = dilation; p.lpad = left_pad;
    p.sn = stride_n; p.st = stride_t;
    p.e.bias = nullptr; p.e.act = epi->act; p.e.alpha = epi->alpha; p.e.slope = epi->prelu_slope;
    p.e.slope_stride = epi->prelu_per_channel ? 1 : 0; p.e.leak = epi->leaky_slope;
    p.e.res = epi->residual; p.e.ldres = epi->ld_residual; p.e.beta = epi->beta; p.e.out = out; p.e.ldo = ld_out;
    p.e.out_lo = out_lo;
    p.lens = reinterpret_cast<const long long*>(lens);
    p.e.post_scale = epi->post_scale; p.e.post_shift = epi->post_shift;
    APSB_CHECK_ARG(!epi->post_scale == !epi->post_shift, "post_scale and post_shift come together");
    APSB_CHECK_ARG(epi->act != ACT_PRELU || epi->prelu_slope, "PReLU slope missing");
    const long long rows = (long long)batch * num_frames;
    APSB_CHECK_ARG(rows < (1LL << 31), "too many rows (%lld)", rows);
    const bool vec = (channels & 3) == 0 && (ld_x & 3) == 0 && (ld_out & 3) == 0 && ((uintptr_t)x & 15) == 0 &&
                     ((uintptr_t)out & 15) == 0 && ((uintptr_t)weight_kd & 15) == 0;
    APSB_CHECK_ARG(!out_lo || (vec && ((uintptr_t)out_lo & 15) == 0), "dwconv: a lo companion needs the 16-byte aligned path");
    const long long groups = vec ? channels / 4 : channels;
    const long long rows_per_block = groups >= 256 ? 1 : 256 / groups;
    const unsigned grid = (unsigned)((rows + rows_per_block - 1) / rows_per_block);
    if (vec) APSB_CUDA(launch_pdl(dwconv1d_kernel<4>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p));
    else APSB_CUDA(launch_pdl(dwconv1d_kernel<1>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p));
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_dwconv1d_fwd(const float* x, int64_t ld_x, int64_t batch, int64_t num_frames, int64_t channels,
                                     int64_t stride_n, int64_t stride_t, const float* weight_kd, const float* bias,
                                     int kernel, int dilation, int left_pad, const aps_b200_epilogue* epi, float* out,
                                     int64_t ld_out, void* stream) {
    return dwconv1d_impl(x, ld_x, batch, num_frames, channels, stride_n, stride_t, weight_kd, bias, kernel, dilation,
                         left_pad, epi, out, nullptr, ld_out, nullptr, stream);
}

extern "C" int aps_b200_dwconv1d2_fwd(const float* x, int64_t ld_x, int64_t batch, int64_t num_frames, int64_t channels,
                                      int64_t stride_n, int64_t stride_t, const float* weight_kd, const float* bias,
                                      int kernel, int dilation, int left_pad, const aps_b200_epilogue* epi, float* out,
                                      float* out_lo, int64_t ld_out, const int64_t* lens, void* stream) {
    return dwconv1d_impl(x, ld_x, batch, num_frames, channels, stride_n, stride_t, weight_kd, bias, kernel, dilation,
                         left_pad, epi, out, out_lo, ld_out, lens, stream);
}

static int mhsa_impl(const aps_b200_attn_desc* d, float* out, float* out_lo, int64_t ld_out, void* stream) {
    APSB_CHECK_ARG(d && out && d->q && d->k && d->v, "null pointer argument");
    APSB_CHECK_ARG(d->batch > 0 && d->length > 0 && d->heads > 0, "bad shape");
    APSB_CHECK_ARG(d->mode >= 0 && d->mode <= 2, "unknown attention mode %d", d->mode);
    APSB_CHECK_ARG(d->mode == 0 || (d->pos && d->qpos), "relative attention needs position rows");
    APSB_CHECK_ARG(d->mode != 2 || (d->rel_u && d->rel_v), "xl attention needs rel_u / rel_v");
    APSB_CHECK_ARG(d->batch <= 65535 && d->heads <= 65535, "grid too large");
    AttnParams p{};
    p.q = d->q; p.k = d->k; p.v = d->v; p.qpos = d->qpos;
    p.ldq = d->ld_q; p.ldk = d->ld_k; p.ldv = d->ld_v; p.ldqp = d->ld_qpos;
    p.sn = d->stride_n; p.st = d->stride_t;
    p.N = (int)d->batch; p.L = (int)d->length; p.H = (int)d->heads; p.dh = (int)d->head_dim; p.mode = d->mode;
    p.pos = d->pos; p.ldpos = d->ld_pos; p.rel_u = d->rel_u; p.rel_v = d->rel_v;
    p.kpm = d->key_padding_mask; p.kpm_fill = d->padding_fill; p.amask = d->attn_mask;
    p.scale = d->scale; p.out = out; p.out_lo = out_lo; p.ldo = ld_out;
    cudaStream_t st = (cudaStream_t)stream;
    // register-tiled kernel: head dim 64, every operand row 16-byte aligned (APS_B200_MHSA=simple forces the other one)
    const uintptr_t al = (uintptr_t)p.q | (uintptr_t)p.k | (uintptr_t)p.v | (uintptr_t)p.qpos | (uintptr_t)p.pos |
                         (uintptr_t)p.rel_u | (uintptr_t)p.rel_v | (uintptr_t)out | (uintptr_t)out_lo;
    const long long lds = p.ldq | p.ldk | p.ldv | p.ldqp | p.ldpos | p.ldo;
    const char* force = getenv("APS_B200_MHSA");
    if (p.dh == 64 && (al & 15) == 0 && (lds & 3) == 0 && !(force && force[0] == 's')) {
        constexpr int LD = 64 + 4;
        const bool tile64 = getenv("APS_B200_MHSA_TILE64") != nullptr;   // A/B switch
        const int tile = (p.L <= 48 && !tile64) ? 48 : 64;
        const int smem = (3 * tile + (p.mode ? 2 * tile : 0) + (p.mode == 2 ? tile : 0)) * LD * 4;
        static LaunchCache slots[2][64];
        LaunchCache& lc = launch_cache(slots[tile == 48]);
        auto kern = tile == 48 ? mhsa_tiled_kernel<64, 3> : mhsa_tiled_kernel<64, 4>;
        if (smem > lc.smem_set) {
            APSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            lc.smem_set = smem;
        }
        dim3 tg((unsigned)((p.L + tile - 1) / tile), (unsigned)p.H, (unsigned)p.N);
        APSB_CUDA(launch_pdl(kern, tg, dim3(256), (size_t)smem, st, p));
        APSB_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid((unsigned)((p.L + kAttnWarps - 1) / kAttnWarps), (unsigned)p.H, (unsigned)p.N);
    switch (p.dh) {
        case 32: mhsa_kernel<32><<<grid, kAttnWarps * 32, 0, st>>>(p); break;
        case 64: mhsa_kernel<64><<<grid, kAttnWarps * 32, 0, st>>>(p); break;
        default: return set_error(-1, "unsupported head dimension %d (32 or 64)", p.dh);
    }
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_mhsa_fwd(const aps_b200_attn_desc* d, float* out, int64_t ld_out, void* stream) {
    return mhsa_impl(d, out, nullptr, ld_out, stream);
}

extern "C" int aps_b200_mhsa2_fwd(const aps_b200_attn_desc* d, float* out, float* out_lo, int64_t ld_out, void* stream) {
    return mhsa_impl(d, out, out_lo, ld_out, stream);
}

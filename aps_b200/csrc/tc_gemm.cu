// tc_gemm.cu — persistent tcgen05 / TMEM / TMA GEMM with a 3xTF32 split for fp32-level accuracy (sm_100a).
//
//   C[m, n] = epilogue( sum_k A(m, k) * W[n, k] ),   W [N, K] row-major (K-major), A(m, k) gathered on the fly:
//       linear rows         A(m, k) = x[m*ld + k]
//       conv2d (NHWC)       implicit im2col,  m -> (nb, oh, ow), k -> (kh, kw, c)
//       conv_transpose2d    implicit gather,  oh = ih*sh - ph + kh
//
// Precision: the tensor core reads fp32 shared-memory operands as TF32 (10-bit mantissa).  Every operand is split into
// hi = rn_tf32(x) and lo = rn_tf32(x - hi) and the product is rebuilt as hi*hi + hi*lo + lo*hi in the fp32 TMEM
// accumulator (three `tcgen05.mma kind::tf32` per k-step, relative error ~1e-6 instead of TF32's 1e-3), which is what
// the 1e-4 parity budget of a 12-layer post-norm conformer needs.  Weights are split once per module
// (aps_b200_tf32_split); ACTIVATIONS ARE SPLIT INSIDE THIS KERNEL by the A-producer warps, so no hi/lo or im2col copy
// of an activation is ever written to HBM.
//
// One persistent CTA per SM (320 threads) walks 128 x BN output tiles; roles:
//   warp 0    : TMA producer for W_hi / W_lo tiles (BN rows x 32 floats, 128-byte swizzle), mbarrier tx
//   warp 1    : allocates TMEM (2 x BN columns: double-buffered accumulator) and issues the UMMAs
//   warps 2-5 : epilogue — tcgen05.ld of the finished accumulator while the NEXT tile's MMAs run into the other
//               buffer; 32x32 transposes through shared memory, bias / activation / GLU / affine / residual, coalesced
//               128-byte row stores
//   warps 6-9 : A producers — coalesced float4 gathers of the fp32 activation (rows, im2col patches or transposed-conv
//               taps), hi/lo split in registers, st.shared into the same 128-byte-swizzled K-major layout TMA would
//               write, fence.proxy.async, mbarrier arrive
// Replaces the cuBLAS / cuDNN calls behind F.linear, 1x1 convolutions and Conv2d / ConvTranspose2d of the reference
// (aps/asr/transformer/impl.py:62-83, :388-393, :454-475; aps/asr/base/component.py:251-307;
// aps/sse/bss/tcn.py:112-159; aps/sse/enh/dcunet.py:24-87).
#include "../../include/aps_b200.h"
#include "common.cuh"
#include "gemm.cuh"

#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

namespace apsb {

constexpr int TC_BM = 128, TC_BK = 32;
constexpr int TC_THREADS = 320;
constexpr int TC_PRODUCERS = 128;          // warps 6..9
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
// bounded spin: a protocol bug traps (CUDA error) instead of hanging the GPU
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(s_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tc_tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            s_u32(dst)),
        "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar))
                 : "memory");
}
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address  [0, 14)
    d |= (uint64_t)0 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset [32, 46)
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                           // layout type: SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// hi = rn_tf32(x), lo = rn_tf32(x - hi): x = hi + lo up to 2^-22 |x|
__device__ __forceinline__ float rn_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

// How the A operand is gathered (host-filled; see aps_b200_gemm_a_desc)
struct AGather {
    int mode;                 // 0 linear rows, 1 conv2d NHWC, 2 conv_transpose2d NHWC
    const float* x;
    long long ld;             // linear: floats between rows
    int H, W, Cin, KH, KW, sh, sw, ph, pw, dh, dw, OH, OW;
};

struct TcParams {
    int M, N, K;
    int tiles_n;
    long long tiles;
    AGather a;
    Epilogue e;
};

template <int BN> struct TcCfg {
    static constexpr int STAGES = BN == 256 ? 2 : (BN == 128 ? 3 : 4);
    static constexpr int B_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * TC_A_BYTES + 2 * B_BYTES;
    static constexpr int EPI_BYTES = 4 * 32 * 33 * 4;
    static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + 256 + 1024;   // + barriers + alignment slack
    static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                   const __grid_constant__ TcParams p) {
    using C = TcCfg<BN>;
    constexpr int S = C::STAGES;
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~(uintptr_t)1023);
    float* epi_tiles = reinterpret_cast<float*>(base + S * C::STAGE_BYTES);
    uint64_t* full_a = reinterpret_cast<uint64_t*>(base + S * C::STAGE_BYTES + C::EPI_BYTES);
    uint64_t* full_b = full_a + S;
    uint64_t* empty = full_b + S;
    uint64_t* tmem_full = empty + S;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + TC_BK - 1) / TC_BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < S; ++s) {
                tc_mbar_init(full_a + s, TC_PRODUCERS);
                tc_mbar_init(full_b + s, 1);
                tc_mbar_init(empty + s, 1);
            }
            for (int b = 0; b < 2; ++b) {
                tc_mbar_init(tmem_full + b, 1);
                tc_mbar_init(tmem_empty + b, 4);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)),
                     "r"(C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: weight tiles =================
        if (lane == 0) {
            uint32_t it = 0;
            for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                const int n_blk = (int)(tile % p.tiles_n);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % S;
                    const uint32_t ph = (it / S) & 1;
                    tc_mbar_wait(empty + s, ph ^ 1);
                    uint8_t* st = base + s * C::STAGE_BYTES + 2 * TC_A_BYTES;
                    tc_mbar_expect_tx(full_b + s, 2 * C::B_BYTES);
                    tc_tma_load_2d(&tmB, full_b + s, st, kb * TC_BK, n_blk * BN);
                    tc_tma_load_2d(&tmBlo, full_b + s, st + C::B_BYTES, kb * TC_BK, n_blk * BN);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B tf32, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(TC_BM >> 4) << 24);
            uint32_t it = 0, tcount = 0;
            for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tcount) {
                const uint32_t buf = tcount & 1;
                tc_mbar_wait(tmem_empty + buf, ((tcount >> 1) & 1) ^ 1);     // epilogue has drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % S;
                    const uint32_t ph = (it / S) & 1;
                    tc_mbar_wait(full_a + s, ph);
                    tc_mbar_wait(full_b + s, ph);
                    tc_fence_after();
                    const uint32_t sa = s_u32(base + s * C::STAGE_BYTES);
                    const uint32_t sal = sa + TC_A_BYTES, sb = sa + 2 * TC_A_BYTES, sbl = sb + C::B_BYTES;
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t da = tc_smem_desc(sa + k * 32), dal = tc_smem_desc(sal + k * 32);
                        const uint64_t db = tc_smem_desc(sb + k * 32), dbl = tc_smem_desc(sbl + k * 32);
                        tc_mma_tf32(d_tmem, da, db, idesc, (kb | k) != 0);
                        tc_mma_tf32(d_tmem, da, dbl, idesc, 1);
                        tc_mma_tf32(d_tmem, dal, db, idesc, 1);
                    }
                    tc_commit(empty + s);          // frees the stage once the MMAs above have read it
                }
                tc_commit(tmem_full + buf);        // accumulator complete
            }
        }
    } else if (warp < 6) {
        // ================= epilogue warps 2..5: TMEM lane quarter = warp % 4 =================
        // A 32x32 block comes out of TMEM with lane = row; it is transposed through a padded shared tile so that
        // lane = COLUMN afterwards: bias / residual loads and the output stores are full 128-byte rows.
        const int q = warp & 3;
        float* tile_s = epi_tiles + (warp - 2) * (32 * 33);
        const Epilogue& e = p.e;
        uint32_t tcount = 0;
        for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tcount) {
            const int n_blk = (int)(tile % p.tiles_n);
            const long long m_blk = tile / p.tiles_n;
            const uint32_t buf = tcount & 1;
            tc_mbar_wait(tmem_full + buf, (tcount >> 1) & 1);
            tc_fence_after();
            const long long m0 = m_blk * TC_BM + q * 32;
            const int ncols = min(BN, p.N - n_blk * BN);
            const int nchunks = (ncols + 31) >> 5;
            const long long rows_ll = (long long)p.M - m0;
            const int rows = rows_ll >= 32 ? 32 : (rows_ll > 0 ? (int)rows_ll : 0);
#pragma unroll 1
            for (int ch = 0; ch < nchunks; ++ch) {
                const int c0 = ch * 32;
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
                      "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
                      "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
                      "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ch == nchunks - 1) {           // last read of this accumulator buffer: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(tmem_empty + buf);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; ++j) tile_s[lane * 33 + j] = __uint_as_float(r[j]);
                __syncwarp();
                const int n = n_blk * BN + c0 + lane;             // this lane's column from here on
                const bool nok = n < p.N;
                const float bias = (e.bias && nok) ? __ldg(e.bias + n) : 0.f;
                if (e.act == ACT_GLU) {
                    const int no = n >> 1;
                    for (int rr = 0; rr < rows; ++rr) {
                        const float v = tile_s[rr * 33 + lane] + bias;
                        const float g = __shfl_down_sync(0xffffffffu, v, 1);
                        if (!(lane & 1) && n + 1 < p.N) {
                            const long long m = m0 + rr;
                            float o = e.alpha * (v * (1.f / (1.f + __expf(-g))));
                            if (e.res) o = fmaf(e.beta, __ldg(e.res + m * e.ldres + no), o);
                            e.out[m * e.ldo + no] = o;
                        }
                    }
                } else {
                    const float ps = (e.post_scale && nok) ? __ldg(e.post_scale + n) : 1.f;
                    const float pt = (e.post_scale && nok) ? __ldg(e.post_shift + n) : 0.f;
                    const float slope =
                        (e.act == ACT_PRELU && nok) ? __ldg(e.slope + (long long)n * e.slope_stride) : e.leak;
                    // rows in batches of 8: the eight residual loads are in flight together
                    for (int r0 = 0; r0 < rows; r0 += 8) {
                        float v8[8], res8[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int rr = r0 + u;
                            v8[u] = tile_s[min(rr, 31) * 33 + lane] + bias;
                            res8[u] = (e.res && nok && rr < rows) ? __ldg(e.res + (m0 + rr) * e.ldres + n) : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            float v = v8[u];
                            switch (e.act) {
                                case ACT_RELU: v = fmaxf(v, 0.f); break;
                                case ACT_SWISH: v = __fdividef(v, 1.f + __expf(-v)); break;
                                case ACT_TANH: v = tanhf(v); break;
                                case ACT_SIGMOID: v = __fdividef(1.f, 1.f + __expf(-v)); break;
                                case ACT_PRELU:
                                case ACT_LEAKY: v = v >= 0.f ? v : v * slope; break;
                                case ACT_GELU: v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); break;
                                default: break;
                            }
                            v = fmaf(v, ps, pt) * e.alpha;
                            v = fmaf(e.beta, res8[u], v);
                            if (nok && r0 + u < rows) e.out[(m0 + r0 + u) * e.ldo + n] = v;
                        }
                    }
                }
                __syncwarp();                      // tile_s is reused by the next chunk
            }
        }
    } else {
        // ================= A producers (warps 6..9) =================
        // thread -> 16-byte chunk c of rows rg, rg + 16, ..., rg + 112: a warp instruction reads 4 full 128-byte rows
        const int pt = threadIdx.x - 192;
        const int c = pt & 7, rg = pt >> 3;
        const AGather& a = p.a;
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const long long m_blk = tile / p.tiles_n;
            // per-row bases for this tile
            const float* rowp[8];      // linear: row pointer; conv: image base of the row's batch index
            int r_a[8], r_b[8];        // conv: ih0 / iw0;  tconv: oh + ph / ow + pw
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long m = m_blk * TC_BM + rg + 16 * i;
                if (m >= p.M) {
                    rowp[i] = nullptr; r_a[i] = 0; r_b[i] = 0;
                } else if (a.mode == 0) {
                    rowp[i] = a.x + m * a.ld; r_a[i] = 0; r_b[i] = 0;
                } else {
                    const int ow = (int)(m % a.OW);
                    const long long t = m / a.OW;
                    const int oh = (int)(t % a.OH);
                    const long long nb = t / a.OH;
                    rowp[i] = a.x + nb * a.H * a.W * a.Cin;
                    if (a.mode == 1) { r_a[i] = oh * a.sh - a.ph; r_b[i] = ow * a.sw - a.pw; }
                    else             { r_a[i] = oh + a.ph;        r_b[i] = ow + a.pw; }
                }
            }
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                const int k = kb * TC_BK + c * 4;
                float4 v[8];
                if (a.mode == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        v[i] = (rowp[i] && k < p.K) ? __ldg(reinterpret_cast<const float4*>(rowp[i] + k))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    // Cin % 32 == 0: the whole k-block lies inside one (kh, kw) tap
                    const int k0 = kb * TC_BK;
                    const int tap = k0 / a.Cin, cc = k0 - tap * a.Cin + c * 4;
                    const int kw = tap % a.KW, kh = tap / a.KW;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        int ih, iw;
                        bool ok = rowp[i] != nullptr && k < p.K;
                        if (a.mode == 1) {
                            ih = r_a[i] + kh * a.dh;
                            iw = r_b[i] + kw * a.dw;
                            ok = ok && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W;
                        } else {
                            const int nh = r_a[i] - kh, nw = r_b[i] - kw;
                            ok = ok && nh >= 0 && nw >= 0 && (nh % a.sh) == 0 && (nw % a.sw) == 0;
                            ih = nh / a.sh;
                            iw = nw / a.sw;
                            ok = ok && ih < a.H && iw < a.W;
                        }
                        v[i] = ok ? __ldg(reinterpret_cast<const float4*>(rowp[i] + ((long long)ih * a.W + iw) * a.Cin + cc))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                tc_mbar_wait(empty + s, ph ^ 1);   // loads above are already in flight
                uint8_t* st = base + s * C::STAGE_BYTES;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rg + 16 * i;
                    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
                    float4 h, l;
                    h.x = rn_tf32(v[i].x); l.x = rn_tf32(v[i].x - h.x);
                    h.y = rn_tf32(v[i].y); l.y = rn_tf32(v[i].y - h.y);
                    h.z = rn_tf32(v[i].z); l.z = rn_tf32(v[i].z - h.z);
                    h.w = rn_tf32(v[i].w); l.w = rn_tf32(v[i].w - h.w);
                    *reinterpret_cast<float4*>(st + off) = h;
                    *reinterpret_cast<float4*>(st + TC_A_BYTES + off) = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> UMMA reads
                tc_mbar_arrive(full_a + s);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS)
                     : "memory");
    }
}

__global__ void __launch_bounds__(256) tf32_split_kernel(const float* __restrict__ x, long long ldx,
                                                         float* __restrict__ hi, float* __restrict__ lo,
                                                         long long ldo, long long rows, int cols) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = rows * (cols >> 2);
    if (i >= total) return;
    const long long r = i / (cols >> 2);
    const int c = (int)(i - r * (cols >> 2)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
    float4 h, l;
    h.x = rn_tf32(v.x); l.x = rn_tf32(v.x - h.x);
    h.y = rn_tf32(v.y); l.y = rn_tf32(v.y - h.y);
    h.z = rn_tf32(v.z); l.z = rn_tf32(v.z - h.z);
    h.w = rn_tf32(v.w); l.w = rn_tf32(v.w - h.w);
    *reinterpret_cast<float4*>(hi + r * ldo + c) = h;
    *reinterpret_cast<float4*>(lo + r * ldo + c) = l;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Tensor maps depend only on (pointer, shape, box): weights are long-lived, so the encoded maps are cached.
struct MapKey {
    const void* ptr;
    long long rows, cols, ld;
    int box_rows;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h = h * 1000003u ^ std::hash<long long>()(k.rows * 131 + k.cols);
        h = h * 1000003u ^ std::hash<long long>()(k.ld * 7 + k.box_rows);
        return h;
    }
};

static int make_map(CUtensorMap* map, const float* ptr, long long rows, long long cols, long long ld, int box_rows) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, rows, cols, ld, box_rows};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *map = it->second;
            return 0;
        }
    }
    EncodeTiledFn fn = encode_fn();
    APSB_CHECK_ARG(fn, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    APSB_CHECK_ARG(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)rc);
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *map;
    return 0;
}

template <int BN>
static int launch_tc(const AGather& a, const float* W, const float* Wlo, long long ldw, int M, int N, int K,
                     const Epilogue& e, cudaStream_t st) {
    using C = TcCfg<BN>;
    CUtensorMap tB, tBl;
    if (int rc = make_map(&tB, W, N, K, ldw, BN)) return rc;
    if (int rc = make_map(&tBl, Wlo, N, K, ldw, BN)) return rc;
    static bool attr = false;
    if (!attr) {
        APSB_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr = true;
    }
    TcParams p{};
    p.M = M; p.N = N; p.K = K;
    p.tiles_n = (N + BN - 1) / BN;
    p.tiles = (long long)((M + TC_BM - 1) / TC_BM) * p.tiles_n;
    p.a = a; p.e = e;
    const long long grid = p.tiles < num_sms() ? p.tiles : num_sms();
    tc_gemm_kernel<BN><<<(unsigned)grid, TC_THREADS, C::SMEM, st>>>(tB, tBl, p);
    APSB_LAUNCH_CHECK();
    return 0;
}

// Tile width: 256 columns halve the A re-reads and the shared-memory traffic per MMA, but need enough tiles to fill
// the machine; narrow outputs take 128 / 64 so that more CTAs share the work.
static int run_tc(const AGather& a, const float* W, const float* Wlo, long long ldw, long long M, long long N,
                  long long K, const Epilogue& e, cudaStream_t st) {
    const long long tm = (M + TC_BM - 1) / TC_BM;
    const int sms = num_sms();
    const char* fe = getenv("APS_B200_TC_BN");                    // tuning / test aid: force the tile width
    const int force = fe ? atoi(fe) : 0;
    int bn;
    if (force == 64 || force == 128 || force == 256) bn = force;
    else if (N >= 256 && tm * ((N + 255) / 256) >= sms) bn = 256;
    else if (N >= 128 && tm * ((N + 127) / 128) >= sms) bn = 128;
    else if (N > 128 && tm * ((N + 127) / 128) * 2 >= sms) bn = 128;
    else bn = 64;
    if (bn == 256) return launch_tc<256>(a, W, Wlo, ldw, (int)M, (int)N, (int)K, e, st);
    if (bn == 128) return launch_tc<128>(a, W, Wlo, ldw, (int)M, (int)N, (int)K, e, st);
    return launch_tc<64>(a, W, Wlo, ldw, (int)M, (int)N, (int)K, e, st);
}

static int fill_tc_epilogue(Epilogue& e, const aps_b200_epilogue* epi, long long N, float* out, long long ld_out) {
    APSB_CHECK_ARG(epi && out, "null pointer argument");
    e.bias = epi->bias; e.act = epi->act; e.alpha = epi->alpha; e.slope = epi->prelu_slope;
    e.slope_stride = epi->prelu_per_channel ? 1 : 0; e.leak = epi->leaky_slope; e.res = epi->residual;
    e.ldres = epi->ld_residual; e.beta = epi->beta; e.post_scale = epi->post_scale; e.post_shift = epi->post_shift;
    e.out = out; e.ldo = ld_out;
    APSB_CHECK_ARG(e.act >= ACT_NONE && e.act <= ACT_GELU, "unknown activation %d", e.act);
    APSB_CHECK_ARG(e.act != ACT_GLU || (N % 2 == 0), "GLU needs an even number of columns");
    APSB_CHECK_ARG(e.act != ACT_PRELU || e.slope, "PReLU slope missing");
    APSB_CHECK_ARG(!epi->post_scale == !epi->post_shift, "post_scale and post_shift come together");
    APSB_CHECK_ARG(!(epi->post_scale && e.act == ACT_GLU), "post affine is not available with GLU");
    const long long ncols = e.act == ACT_GLU ? N / 2 : N;
    APSB_CHECK_ARG(ld_out >= ncols, "ld_out %lld smaller than %lld columns", ld_out, ncols);
    APSB_CHECK_ARG(!epi->residual || epi->ld_residual >= ncols, "ld_residual too small");
    return 0;
}

static int check_weights(const float* w_hi, const float* w_lo, long long ld_w, long long K) {
    APSB_CHECK_ARG(w_hi && w_lo, "null weight pointer");
    APSB_CHECK_ARG((K & 3) == 0 && (ld_w & 3) == 0 && ld_w >= K && ((uintptr_t)w_hi & 15) == 0 &&
                       ((uintptr_t)w_lo & 15) == 0,
                   "the tensor-core path needs 16-byte aligned weight rows (K %% 4 == 0)");
    return 0;
}

}  // namespace apsb

using namespace apsb;

extern "C" int aps_b200_tf32_split(const float* x, int64_t rows, int64_t cols, int64_t ld_x, float* hi, float* lo,
                                   int64_t ld_out, void* stream) {
    APSB_CHECK_ARG(x && hi && lo && rows > 0 && cols > 0, "bad arguments");
    APSB_CHECK_ARG((cols & 3) == 0 && (ld_x & 3) == 0 && (ld_out & 3) == 0 && ((uintptr_t)x & 15) == 0 &&
                       ((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0, "tf32 split needs 16-byte aligned rows");
    const long long total = rows * (cols >> 2);
    tf32_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, ld_x, hi, lo, ld_out, rows,
                                                                                       (int)cols);
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_linear_tc_fwd(const float* x, int64_t rows, int64_t in_features, int64_t ld_x,
                                      const float* weight_hi, const float* weight_lo, int64_t ld_w,
                                      int64_t out_features, const aps_b200_epilogue* epi, float* out, int64_t ld_out,
                                      void* stream) {
    APSB_CHECK_ARG(x, "null pointer argument");
    APSB_CHECK_ARG(rows > 0 && in_features > 0 && out_features > 0 && ld_x >= in_features, "bad shape");
    APSB_CHECK_ARG(rows < (1LL << 31) && out_features < (1LL << 31) && in_features < (1LL << 31), "shape too large");
    APSB_CHECK_ARG((ld_x & 3) == 0 && ((uintptr_t)x & 15) == 0,
                   "the tensor-core path needs 16-byte aligned activation rows");
    if (int rc = check_weights(weight_hi, weight_lo, ld_w, in_features)) return rc;
    Epilogue e{};
    if (int rc = fill_tc_epilogue(e, epi, out_features, out, ld_out)) return rc;
    AGather a{};
    a.mode = 0; a.x = x; a.ld = ld_x;
    return run_tc(a, weight_hi, weight_lo, ld_w, rows, out_features, in_features, e, (cudaStream_t)stream);
}

static int conv_geometry(AGather& a, const float* x, int64_t batch, int64_t height, int64_t width,
                         int64_t in_channels, int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h,
                         int pad_w) {
    APSB_CHECK_ARG(x, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && height > 0 && width > 0 && in_channels > 0, "bad shape");
    APSB_CHECK_ARG(kernel_h > 0 && kernel_w > 0 && stride_h > 0 && stride_w > 0 && pad_h >= 0 && pad_w >= 0,
                   "bad convolution geometry");
    APSB_CHECK_ARG(in_channels % TC_BK == 0 && ((uintptr_t)x & 15) == 0,
                   "the tensor-core convolution needs Cin %% 32 == 0 and a 16-byte aligned input (Cin = %lld)",
                   (long long)in_channels);
    APSB_CHECK_ARG(height < (1 << 20) && width < (1 << 20) && in_channels < (1 << 20), "shape too large");
    a.x = x; a.H = (int)height; a.W = (int)width; a.Cin = (int)in_channels; a.KH = kernel_h; a.KW = kernel_w;
    a.sh = stride_h; a.sw = stride_w; a.ph = pad_h; a.pw = pad_w;
    return 0;
}

extern "C" int aps_b200_conv2d_nhwc_tc_fwd(const float* x, int64_t batch, int64_t height, int64_t width,
                                           int64_t in_channels, const float* weight_hi, const float* weight_lo,
                                           int64_t out_channels, int kernel_h, int kernel_w, int stride_h,
                                           int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                           const aps_b200_epilogue* epi, float* out, void* stream) {
    AGather a{};
    if (int rc = conv_geometry(a, x, batch, height, width, in_channels, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                               pad_w))
        return rc;
    APSB_CHECK_ARG(dil_h > 0 && dil_w > 0 && out_channels > 0, "bad convolution geometry");
    a.mode = 1; a.dh = dil_h; a.dw = dil_w;
    const int64_t OH = (height + 2 * pad_h - dil_h * (kernel_h - 1) - 1) / stride_h + 1;
    const int64_t OW = (width + 2 * pad_w - dil_w * (kernel_w - 1) - 1) / stride_w + 1;
    APSB_CHECK_ARG(OH > 0 && OW > 0, "convolution output is empty");
    a.OH = (int)OH; a.OW = (int)OW;
    const int64_t M = batch * OH * OW, K = (int64_t)kernel_h * kernel_w * in_channels;
    APSB_CHECK_ARG(M < (1LL << 31) && K < (1LL << 31) && out_channels < (1LL << 31), "shape too large");
    if (int rc = check_weights(weight_hi, weight_lo, K, K)) return rc;
    Epilogue e{};
    const int64_t ncols = (epi && epi->act == ACT_GLU) ? out_channels / 2 : out_channels;
    if (int rc = fill_tc_epilogue(e, epi, out_channels, out, ncols)) return rc;
    return run_tc(a, weight_hi, weight_lo, K, M, out_channels, K, e, (cudaStream_t)stream);
}

extern "C" int aps_b200_conv_transpose2d_nhwc_tc_fwd(const float* x, int64_t batch, int64_t height, int64_t width,
                                                     int64_t in_channels, const float* weight_hi,
                                                     const float* weight_lo, int64_t out_channels, int kernel_h,
                                                     int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                                                     int out_pad_h, int out_pad_w, const aps_b200_epilogue* epi,
                                                     float* out, void* stream) {
    AGather a{};
    if (int rc = conv_geometry(a, x, batch, height, width, in_channels, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                               pad_w))
        return rc;
    APSB_CHECK_ARG(out_pad_h >= 0 && out_pad_w >= 0 && out_channels > 0, "bad convolution geometry");
    a.mode = 2; a.dh = 1; a.dw = 1;
    const int64_t OH = (height - 1) * stride_h - 2 * pad_h + kernel_h + out_pad_h;
    const int64_t OW = (width - 1) * stride_w - 2 * pad_w + kernel_w + out_pad_w;
    APSB_CHECK_ARG(OH > 0 && OW > 0, "transposed convolution output is empty");
    a.OH = (int)OH; a.OW = (int)OW;
    const int64_t M = batch * OH * OW, K = (int64_t)kernel_h * kernel_w * in_channels;
    APSB_CHECK_ARG(M < (1LL << 31) && K < (1LL << 31) && out_channels < (1LL << 31), "shape too large");
    if (int rc = check_weights(weight_hi, weight_lo, K, K)) return rc;
    APSB_CHECK_ARG(epi && epi->act != ACT_GLU, "GLU is not available here");
    Epilogue e{};
    if (int rc = fill_tc_epilogue(e, epi, out_channels, out, out_channels)) return rc;
    return run_tc(a, weight_hi, weight_lo, K, M, out_channels, K, e, (cudaStream_t)stream);
}

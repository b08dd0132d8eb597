// tc_gemm.cu — tcgen05 / TMEM / TMA GEMM with a 3xTF32 split for fp32-level accuracy (sm_100a).
//
//   C[m, n] = epilogue( sum_k A[m, k] * W[n, k] ),   A [M, K], W [N, K] row-major fp32 (both K-major)
//
// Precision: the tensor core reads fp32 shared-memory operands as TF32 (10-bit mantissa).  Every operand
// is therefore split beforehand (aps_b200_tf32_split) into hi = rn_tf32(x) and lo = rn_tf32(x - hi), both
// exactly representable in TF32, and the product is rebuilt as hi*hi + hi*lo + lo*hi in the fp32 TMEM
// accumulator: three `tcgen05.mma kind::tf32` per k-step, relative error ~1e-6 instead of TF32's 1e-3,
// which is what the 1e-4 parity budget of a 12-layer post-norm conformer needs.  The
// honest tensor-pipe ceiling is therefore one third of the TF32 peak.
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0  : TMA producer — four 2-D bulk tensor loads per k-block (A, A_lo, W, W_lo tiles of
//             rows x 32 floats = 128-byte swizzled rows) into a 3-stage shared-memory ring, mbarrier tx
//   warp 1  : allocates TMEM, then one elected lane issues the UMMAs (128 x BN x 8 per instruction,
//             descriptors advance 32 bytes per k-step inside the 128-byte swizzle atom) and commits
//             stage-free / accumulator-ready barriers with tcgen05.commit
//   warps 2-5: epilogue — tcgen05.ld (32 lanes x 32 columns per warp and step) -> bias / activation /
//             GLU / affine / residual -> global stores
// Replaces the cuBLAS sgemm calls behind F.linear / 1x1 convolutions of the reference
// (aps/asr/transformer/impl.py:62-83, :388-393, :454-475; aps/sse/bss/tcn.py:112-159).
#include "../../include/aps_b200.h"
#include "common.cuh"
#include "gemm.cuh"

#include <cuda.h>

namespace apsb {

constexpr int TC_BM = 128, TC_BK = 32, TC_STAGES = 3, TC_THREADS = 192;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
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

struct TcParams {
    int M, N, K;
    Epilogue e;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                   const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                   const TcParams p) {
    constexpr int A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BN * TC_BK * 4;
    constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + TC_STAGES * STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* acc_ready = empty + TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = blockIdx.y, n_blk = blockIdx.x;
    const int num_kb = (p.K + TC_BK - 1) / TC_BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAlo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < TC_STAGES; ++s) {
                tc_mbar_init(full + s, 1);
                tc_mbar_init(empty + s, 1);
            }
            tc_mbar_init(acc_ready, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1;
                tc_mbar_wait(empty + s, ph ^ 1);
                uint8_t* st = base + s * STAGE_BYTES;
                tc_mbar_expect_tx(full + s, STAGE_BYTES);
                tc_tma_load_2d(&tmA, full + s, st, kb * TC_BK, m_blk * TC_BM);
                tc_tma_load_2d(&tmAlo, full + s, st + A_BYTES, kb * TC_BK, m_blk * TC_BM);
                tc_tma_load_2d(&tmB, full + s, st + 2 * A_BYTES, kb * TC_BK, n_blk * BN);
                tc_tma_load_2d(&tmBlo, full + s, st + 2 * A_BYTES + B_BYTES, kb * TC_BK, n_blk * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B tf32, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(TC_BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1;
                tc_mbar_wait(full + s, ph);
                tc_fence_after();
                const uint32_t sa = s_u32(base + s * STAGE_BYTES);
                const uint32_t sal = sa + A_BYTES, sb = sa + 2 * A_BYTES, sbl = sb + B_BYTES;
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    const uint64_t da = tc_smem_desc(sa + k * 32), dal = tc_smem_desc(sal + k * 32);
                    const uint64_t db = tc_smem_desc(sb + k * 32), dbl = tc_smem_desc(sbl + k * 32);
                    tc_mma_tf32(tmem_base, da, db, idesc, (kb | k) != 0);
                    tc_mma_tf32(tmem_base, da, dbl, idesc, 1);
                    tc_mma_tf32(tmem_base, dal, db, idesc, 1);
                }
                tc_commit(empty + s);       // frees the stage once the MMAs above have read it
            }
            tc_commit(acc_ready);           // accumulator complete
        }
    } else {
        // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4 ----------------------------------------------
        // Each warp owns 32 accumulator rows.  A 32x32 block comes out of TMEM with lane = row; it is transposed
        // through a padded shared tile (the pipeline stages are free by now) so that lane = COLUMN afterwards:
        // bias / residual loads and the output stores are then full 128-byte rows instead of 32 scattered sectors.
        tc_mbar_wait(acc_ready, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int m0 = m_blk * TC_BM + q * 32;
        float* tile = reinterpret_cast<float*>(base) + (warp - 2) * (32 * 33);
        const Epilogue& e = p.e;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                  "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                  "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(r[j]);
            __syncwarp();
            const int n = n_blk * BN + c0 + lane;             // this lane's column from here on
            const bool nok = n < p.N;
            const float bias = (e.bias && nok) ? __ldg(e.bias + n) : 0.f;
            const int rows = min(32, p.M - m0);
            if (e.act == ACT_GLU) {
                const int no = n >> 1;
                for (int rr = 0; rr < rows; ++rr) {
                    const float v = tile[rr * 33 + lane] + bias;
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
                const float slope = (e.act == ACT_PRELU && nok) ? __ldg(e.slope + (long long)n * e.slope_stride) : e.leak;
                // rows in batches of 8: the eight residual loads are in flight together (memory-level parallelism)
                for (int r0 = 0; r0 < rows; r0 += 8) {
                    float v8[8], res8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int rr = r0 + u;
                        v8[u] = tile[min(rr, 31) * 33 + lane] + bias;
                        res8[u] = (e.res && nok && rr < rows) ? __ldg(e.res + (long long)(m0 + rr) * e.ldres + n) : 0.f;
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
                        if (nok && r0 + u < rows) e.out[(long long)(m0 + r0 + u) * e.ldo + n] = v;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
    }
}

// hi = rn_tf32(x), lo = rn_tf32(x - hi): x = hi + lo up to 2^-22 |x|
__device__ __forceinline__ float rn_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
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

// im2col of an NHWC tensor fused with the TF32 split: patch[m, (kh*KW + kw)*Cin + c] -> hi / lo [M, K]
struct Im2colParams {
    const float* x;
    int H, W, Cin, KH, KW, sh, sw, ph, pw, dh, dw, OH, OW;
    long long M;
    int K;
    float* hi;
    float* lo;
};

__global__ void __launch_bounds__(256) im2col_split_kernel(const __grid_constant__ Im2colParams p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int k4n = p.K >> 2;
    if (i >= p.M * k4n) return;
    const long long m = i / k4n;
    const int k = (int)(i - m * k4n) * 4;
    const int c = k % p.Cin, t = k / p.Cin;
    const int kw = t % p.KW, kh = t / p.KW;
    const int ow = (int)(m % p.OW);
    const long long t2 = m / p.OW;
    const int oh = (int)(t2 % p.OH);
    const long long nb = t2 / p.OH;
    const int ih = oh * p.sh - p.ph + kh * p.dh, iw = ow * p.sw - p.pw + kw * p.dw;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
        v = __ldg(reinterpret_cast<const float4*>(p.x + ((nb * p.H + ih) * p.W + iw) * p.Cin + c));
    float4 h, l;
    h.x = rn_tf32(v.x); l.x = rn_tf32(v.x - h.x);
    h.y = rn_tf32(v.y); l.y = rn_tf32(v.y - h.y);
    h.z = rn_tf32(v.z); l.z = rn_tf32(v.z - h.z);
    h.w = rn_tf32(v.w); l.w = rn_tf32(v.w - h.w);
    *reinterpret_cast<float4*>(p.hi + m * p.K + k) = h;
    *reinterpret_cast<float4*>(p.lo + m * p.K + k) = l;
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

static int make_map(CUtensorMap* map, const float* ptr, long long rows, long long cols, long long ld, int box_rows) {
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
    return 0;
}

template <int BN>
static int launch_tc(const float* A, const float* Alo, long long lda, const float* W, const float* Wlo, long long ldw,
                     int M, int N, int K, const Epilogue& e, cudaStream_t st) {
    CUtensorMap tA, tAl, tB, tBl;
    if (int rc = make_map(&tA, A, M, K, lda, TC_BM)) return rc;
    if (int rc = make_map(&tAl, Alo, M, K, lda, TC_BM)) return rc;
    if (int rc = make_map(&tB, W, N, K, ldw, BN)) return rc;
    if (int rc = make_map(&tBl, Wlo, N, K, ldw, BN)) return rc;
    constexpr int smem = TC_STAGES * (2 * TC_BM * TC_BK * 4 + 2 * BN * TC_BK * 4) + 1024 + 256;
    static bool attr = false;
    if (!attr) {
        APSB_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    TcParams p{M, N, K, e};
    dim3 grid((N + BN - 1) / BN, (M + TC_BM - 1) / TC_BM);
    tc_gemm_kernel<BN><<<grid, TC_THREADS, smem, st>>>(tA, tAl, tB, tBl, p);
    APSB_LAUNCH_CHECK();
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

extern "C" int aps_b200_im2col_tf32_split(const float* x, int64_t batch, int64_t height, int64_t width,
                                          int64_t in_channels, int kernel_h, int kernel_w, int stride_h, int stride_w,
                                          int pad_h, int pad_w, int dil_h, int dil_w, float* hi, float* lo,
                                          void* stream) {
    APSB_CHECK_ARG(x && hi && lo, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && height > 0 && width > 0 && in_channels > 0 && (in_channels & 3) == 0 &&
                       ((uintptr_t)x & 15) == 0 && ((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0,
                   "im2col split needs Cin %% 4 == 0 and 16-byte aligned buffers");
    Im2colParams p{};
    p.x = x; p.H = (int)height; p.W = (int)width; p.Cin = (int)in_channels; p.KH = kernel_h; p.KW = kernel_w;
    p.sh = stride_h; p.sw = stride_w; p.ph = pad_h; p.pw = pad_w; p.dh = dil_h; p.dw = dil_w;
    p.OH = (int)((height + 2 * pad_h - dil_h * (kernel_h - 1) - 1) / stride_h + 1);
    p.OW = (int)((width + 2 * pad_w - dil_w * (kernel_w - 1) - 1) / stride_w + 1);
    APSB_CHECK_ARG(p.OH > 0 && p.OW > 0, "convolution output is empty");
    p.M = batch * p.OH * p.OW;
    p.K = kernel_h * kernel_w * (int)in_channels;
    p.hi = hi; p.lo = lo;
    const long long total = p.M * (p.K >> 2);
    APSB_CHECK_ARG(total < (1LL << 31) * 256, "im2col too large");
    im2col_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    APSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int aps_b200_linear_tc_fwd(const float* x_hi, const float* x_lo, int64_t rows, int64_t in_features,
                                      int64_t ld_x, const float* weight_hi, const float* weight_lo, int64_t ld_w,
                                      int64_t out_features, const aps_b200_epilogue* epi, float* out, int64_t ld_out,
                                      void* stream) {
    const float *x = x_hi, *weight = weight_hi;
    APSB_CHECK_ARG(x && x_lo && weight && weight_lo && epi && out, "null pointer argument");
    APSB_CHECK_ARG(rows > 0 && in_features > 0 && out_features > 0, "bad shape");
    APSB_CHECK_ARG((in_features & 3) == 0 && (ld_x & 3) == 0 && (ld_w & 3) == 0 && ((uintptr_t)x & 15) == 0 &&
                       ((uintptr_t)x_lo & 15) == 0 && ((uintptr_t)weight & 15) == 0 && ((uintptr_t)weight_lo & 15) == 0,
                   "the tensor-core path needs 16-byte aligned rows (K %% 4 == 0)");
    APSB_CHECK_ARG(rows < (1LL << 31) && out_features < (1LL << 31) && in_features < (1LL << 31), "shape too large");
    Epilogue e{};
    e.bias = epi->bias; e.act = epi->act; e.alpha = epi->alpha; e.slope = epi->prelu_slope;
    e.slope_stride = epi->prelu_per_channel ? 1 : 0; e.leak = epi->leaky_slope; e.res = epi->residual;
    e.ldres = epi->ld_residual; e.beta = epi->beta; e.post_scale = epi->post_scale; e.post_shift = epi->post_shift;
    e.out = out; e.ldo = ld_out;
    APSB_CHECK_ARG(e.act >= ACT_NONE && e.act <= ACT_GELU, "unknown activation %d", e.act);
    APSB_CHECK_ARG(e.act != ACT_GLU || (out_features % 2 == 0), "GLU needs an even number of columns");
    cudaStream_t st = (cudaStream_t)stream;
    const long long tiles128 = ((rows + 127) / 128) * ((out_features + 127) / 128);
    if (tiles128 >= num_sms())
        return launch_tc<128>(x, x_lo, ld_x, weight, weight_lo, ld_w, (int)rows, (int)out_features, (int)in_features, e,
                              st);
    return launch_tc<64>(x, x_lo, ld_x, weight, weight_lo, ld_w, (int)rows, (int)out_features, (int)in_features, e, st);
}

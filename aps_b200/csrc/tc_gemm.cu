// tc_gemm.cu — persistent tcgen05 / TMEM / TMA GEMM with a 3xTF32 split for fp32-level accuracy (sm_100a).
//
//   C[m, n] = epilogue( sum_k A(m, k) * W[n, k] ),   W [N, K] row-major (K-major), A(m, k) gathered on the fly:
//       linear rows         A(m, k) = x[m*ld + k]
//       conv2d (NHWC)       implicit im2col,  m -> (nb, oh, ow), k -> (kh, kw, c)
//       conv_transpose2d    implicit gather,  oh = ih*sh - ph + kh
//       linear rows + lo    (MODE 3) the producer of x also wrote its companion xl = rn_tf32(x - trunc_tf32(x)): the tensor
//                           core reads an fp32 operand as TF32 by DROPPING the low 13 mantissa bits, so the raw x tile IS
//                           the "hi" operand and x = trunc(x) + xl holds to 2^-21; both tiles arrive by TMA and the kernel
//                           has no gather / split warps at all.  Optional split-K (partial tiles, reduced by the
//                           LayerNorm / reduce kernel) for the skinny [3200 x 2048] x [2048 x 256] shapes.
//
// Precision: the tensor core reads fp32 shared-memory operands as TF32 (10-bit mantissa).  Every operand is split into
// hi = rn_tf32(x) and lo = rn_tf32(x - hi) and the product is rebuilt as hi*hi + hi*lo + lo*hi in the fp32 TMEM
// accumulator (three `tcgen05.mma kind::tf32` per k-step, relative error ~1e-6 instead of TF32's 1e-3), which is what
// the 1e-4 parity budget of a 12-layer post-norm conformer needs.  Weights are split once per module
// (aps_b200_tf32_split); ACTIVATIONS ARE SPLIT INSIDE THIS KERNEL by the A-producer warps, so no hi/lo or im2col copy
// of an activation is ever written to HBM.
//
// One persistent CTA per SM (320 threads, 448 for convolution gathers on 64/128-wide tiles) walks 128 x BN output tiles; roles:
//   warp 0    : TMA producer for W_hi / W_lo tiles (BN rows x BK floats, swizzled), mbarrier tx
//   warp 1    : allocates TMEM (2 x BN columns: double-buffered accumulator) and issues the UMMAs
//   warps 2-5 : epilogue — tcgen05.ld of the finished accumulator while the NEXT tile's MMAs run into the other
//               buffer; 32x32 transposes through shared memory, bias / activation / GLU / affine / residual, coalesced
//               128-byte row stores
//   warps 6-9 (6-13): A producers — coalesced float4 gathers of the fp32 activation (rows, im2col patches or transposed-conv
//               taps), hi/lo split in registers (integer rounding: sm_100a emulates cvt.rna.tf32), st.shared into the same 128-byte-swizzled K-major layout TMA would
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
#include <vector>

namespace apsb {

constexpr int TC_BM = 128;
// A-producer warps: 4 for linear layers and the 256-wide tiles (MMA / shared-memory bound), 8 for the convolution
// gathers on 64- and 128-wide tiles, which are bound by the producers' own instruction latency (0.14 IPC per warp)
#ifndef APSB_TC_PW_CONV
#define APSB_TC_PW_CONV 8
#endif
#ifndef APSB_TC_PW_LINEAR
#define APSB_TC_PW_LINEAR 4
#endif
// default cluster size for the weight-tile multicast.  1 = off: MEASURED SLOWER on every shape of the encoder (B200,
// profiles/r02h_tc_cluster_sweep.txt: conv2 850 / 1065 / 1198 us and FFN-a 31.1 / 32.0 / 39.2 us at CL = 1 / 2 / 4) — the
// lock step couples the CTAs' pipelines (a stage is refilled only when the SLOWEST CTA has consumed it) and that costs
// more than the halved L2 -> SM weight traffic saves, i.e. the engine is not bound by that traffic.  The path stays
// (numerics verified for CL = 1, 2, 4) behind APS_B200_TC_CL for shapes where it might pay.
#ifndef APSB_TC_CLUSTER
#define APSB_TC_CLUSTER 1
#endif
// MODE 5 is MODE 3 for several independent problems of one shape side by side ("groups": the recurrent projections of the
// LSTMs that advance together): row block m of the stacked activations takes the weight rows of ITS group
#define TC_LIN3(M) ((M) == 3 || (M) == 5)
// MODE 3 (linear layer whose activation comes with its TF32 "lo" companion, see below) has NO producer warps: the TMA
// warp loads the A tiles as well.  MODE 4 (stride-2 convolution fed by 5-D TMA boxes) has four CONVERTER warps in their
// place: they derive the lo tile from the raw tile the TMA delivered, inside shared memory.
template <int BN, int MODE> struct TcRoles {
    static constexpr int PW = TC_LIN3(MODE) ? 0 : ((BN > 128 || MODE == 4) ? 4 : (MODE != 0 ? APSB_TC_PW_CONV : APSB_TC_PW_LINEAR));
    static constexpr int PRODUCERS = PW * 32;
    // epilogue warps: 4 (one per TMEM lane quarter), or 8 in MODE 3 (two per quarter taking alternate 32-column chunks):
    // a 32 x 32 block costs ~900 dependent-issue-bound instructions, and ONE warp per scheduler cannot hide their
    // latencies (trace r02k: 11-13 k cycles of epilogue per 128 x 128 tile against a 10 k main loop).  MODE 3 has no
    // producer warps, so the second set is free.
    static constexpr int EW = TC_LIN3(MODE) ? 8 : 4;
    static constexpr int THREADS = (2 + EW + PW) * 32;
};
constexpr int WARP_TMA = 0, WARP_MMA = 1, WARP_EPI0 = 2, WARP_PROD0 = 6;

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
// Wait for the phase with the given parity.  Fast path: one non-blocking test_wait — on B200 a try_wait costs ~300
// cycles even when the phase is already complete (scripts/ubench/umma_rate.cu), which starves the tensor pipe when the
// single MMA-issuing thread pays it twice per k-block.  Slow path: bounded try_wait spin; a protocol bug traps (CUDA
// error) instead of hanging the GPU.
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(s_u32(bar)), "r"(parity)
        : "memory");
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
// Wait used by the roles that are expected to wait LONG (TMA thread and producer pollers on a free stage, epilogue on a
// finished accumulator): try_wait with a suspend-time hint parks the thread in hardware instead of hot-spinning — an
// unhinted try_wait returns after a few cycles, and the resulting ~100 M polling instructions per launch were stealing
// the issue slots of the producer warps that share the scheduler (ncu: profiles/r01_tc_gemm_v2_spin.txt).
__device__ __forceinline__ void tc_mbar_wait_parked(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(s_u32(bar)), "r"(parity)
        : "memory");
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(s_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
        if (spin > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void tc_tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            s_u32(dst)),
        "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                               int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            s_u32(dst)),
        "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// the same box written into the shared memory of every CTA of the cluster named in `mask` (same CTA-relative offset), each
// destination's mbarrier (same offset) gets the complete_tx
__device__ __forceinline__ void tc_tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                                  uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
        "[%2], %5;" ::"r"(s_u32(dst)),
        "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tc_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar))
                 : "memory");
}
// commit that arrives on the mbarrier at the same offset in EVERY CTA of `mask`: a stage that was filled by a multicast
// may only be refilled when all CTAs of the cluster have finished reading it
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     s_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// shared-memory matrix descriptor: K-major operand whose rows are one swizzle span (SWZ = 128 or 64 bytes) wide,
// 8-row groups 8*SWZ bytes apart (canonical layouts Swizzle<3,4,3> / Swizzle<2,4,3>)
template <int SWZ>
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address  [0, 14)
    d |= (uint64_t)(SWZ == 128 ? 0 : 1) << 16;        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8 * SWZ) >> 4) << 32;            // stride byte offset [32, 46)
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)(SWZ == 128 ? 2 : 4) << 61;        // layout type: SWIZZLE_128B / SWIZZLE_64B
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

// "lo" companion of an activation the NEXT tensor-core GEMM will read raw (MODE 3): the tensor core truncates x to
// trunc(x) = bits & ~0x1fff; lo = rn_tf32(x - trunc(x)) (the subtraction is exact; integer rounding as in the producers)
__device__ __forceinline__ float tf32_lo(float x) {
    const float l = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    return __uint_as_float((__float_as_uint(l) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float4 tf32_lo4(const float4& v) {
    return make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
}

// Activation resolved at compile time (a run-time switch inside the unrolled element loops would replicate tanhf / erff
// dozens of times: the epilogue is instruction-issue bound, one warp per scheduler).
// (A 4-instruction sigmoid — ex2.approx.ftz + rcp.approx.ftz through inline PTX instead of __expf / __fdividef, which carry
// a range test and two scaling multiplies each — removed 500 of the 8 664 instructions of the 128-wide TMA-fed kernel and
// made the step SLOWER: 3.995 -> 4.13 ms, conv2 848 -> 973 us although its ReLU epilogue does not even use it.  Same
// layout sensitivity as KFORM below; measured in round 2, visit O, and left out.)
__device__ __forceinline__ float tc_glu_gate(float g) { return 1.f / (1.f + __expf(-g)); }

template <int ACT>
__device__ __forceinline__ float tc_act(float v, float slope) {
    if (ACT == ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == ACT_SWISH) return __fdividef(v, 1.f + __expf(-v));
    if (ACT == ACT_TANH) return tanhf(v);
    if (ACT == ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-v));
    if (ACT == ACT_PRELU || ACT == ACT_LEAKY) return v >= 0.f ? v : v * slope;
    if (ACT == ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    return v;
}

// 16-byte shared-memory accesses on a 32-bit shared-window address.  Through the generic pointer the compiler emitted
// LD.E.128 / ST.E.128 with 64-bit addresses AND kept every read-back behind the previous row's global store (it cannot
// prove that the output does not alias the tile): ncu source view of the FFN-a kernel, r02j — eight serialised
// ~200-cycle round trips per 32 x 32 block made the epilogue (14-16 k cycles per 128 x 128 tile) slower than the tile's
// main loop (10 k).
__device__ __forceinline__ void tc_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 tc_lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// Epilogue arithmetic on 4 consecutive columns of one output row (after the shared-memory transpose, see the epilogue
// warps): bias / activation / post affine / alpha / residual; the per-column vectors are the lane's own 4 columns.
template <int ACT>
__device__ __forceinline__ float4 tc_epilogue4(const float4& v, const float4& b4, const float4& ps4, const float4& pt4,
                                               const float4& sl4, const float4& res, float alpha, float beta) {
    float4 o;
    o.x = fmaf(beta, res.x, fmaf(tc_act<ACT>(v.x + b4.x, sl4.x), ps4.x, pt4.x) * alpha);
    o.y = fmaf(beta, res.y, fmaf(tc_act<ACT>(v.y + b4.y, sl4.y), ps4.y, pt4.y) * alpha);
    o.z = fmaf(beta, res.z, fmaf(tc_act<ACT>(v.z + b4.z, sl4.z), ps4.z, pt4.z) * alpha);
    o.w = fmaf(beta, res.w, fmaf(tc_act<ACT>(v.w + b4.w, sl4.w), ps4.w, pt4.w) * alpha);
    return o;
}

// How the A operand is gathered (host-filled; see aps_b200_gemm_a_desc)
struct AGather {
    int mode;                 // 0 linear rows, 1 conv2d NHWC, 2 conv_transpose2d NHWC
    const float* x;
    long long ld;             // linear: floats between rows
    int H, W, Cin, KH, KW, sh, sw, ph, pw, dh, dw, OH, OW;
    // Transposed convolution with stride sh > 1 along H: output row oh only sees the taps kh with (oh + ph - kh) % sh == 0.
    // The GEMM rows are therefore enumerated CLASS-MAJOR (all rows with oh % sh == 0 first, then oh % sh == 1, ...), so
    // that a 128-row tile lies in one class and every role skips the k-blocks of the taps that are all-zero for it:
    // the zero-insertion work of the reference's conv_transpose2d (half of it at stride 2) is never done.
    // Channel-concatenated input of the DCCRN decoder (dccrn.py "cat" connection on stacked complex channels):
    // logical channels [re_a | re_b | im_a | im_b] with a = x, b = x2, each source holding 2*cat_c channels.  The gather
    // reads the two tensors in place, so the concatenated tensor is never written (torch.cat was 9 % of the step).
    const float* x2;
    int cat_c;                // channels per part (0: single source x with Cin channels)
    int classes;              // sh (1 = plain row order)
    long long class_start[4]; // first GEMM row of each class
    int class_rows[4];        // output rows per image in each class
};

struct RowPos {
    int nb, oh, ow;
};

// GEMM row m -> (image, output row, output column) for the convolution modes.  M < 2^31 (host check), so all of this
// is 32-bit arithmetic: 64-bit integer divisions are emulated and would cost thousands of cycles per tile.
__device__ __forceinline__ RowPos tc_decode_row(const AGather& a, unsigned m) {
    RowPos r;
    if (a.mode == 2 && a.classes > 1) {
        int cls = 0;
#pragma unroll
        for (int c = 1; c < 4; ++c)
            if (c < a.classes && m >= (unsigned)a.class_start[c]) cls = c;
        const unsigned mm = m - (unsigned)a.class_start[cls];
        const unsigned t = mm / (unsigned)a.OW;
        r.ow = (int)(mm - t * (unsigned)a.OW);
        const unsigned nb = t / (unsigned)a.class_rows[cls];
        r.oh = (int)(t - nb * (unsigned)a.class_rows[cls]) * a.sh + cls;
        r.nb = (int)nb;
    } else {
        const unsigned t = m / (unsigned)a.OW;
        r.ow = (int)(m - t * (unsigned)a.OW);
        const unsigned nb = t / (unsigned)a.OH;
        r.oh = (int)(t - nb * (unsigned)a.OH);
        r.nb = (int)nb;
    }
    return r;
}
// class of a tile (rows m_first .. m_last), -1 if it straddles two classes or classes are off
__device__ __forceinline__ int tc_tile_class(const AGather& a, unsigned m_first, unsigned m_last) {
    if (a.mode != 2 || a.classes <= 1) return -1;
    int c0 = 0, c1 = 0;
#pragma unroll
    for (int c = 1; c < 4; ++c)
        if (c < a.classes) {
            if (m_first >= (unsigned)a.class_start[c]) c0 = c;
            if (m_last >= (unsigned)a.class_start[c]) c1 = c;
        }
    return c0 == c1 ? c0 : -1;
}
// does kernel row kh contribute to the output rows of class `cls`?
__device__ __forceinline__ bool tc_kh_valid(const AGather& a, int cls, int kh) {
    if (cls < 0) return true;
    if (a.sh == 2) return ((cls + a.ph - kh) & 1) == 0;          // every reference model; no emulated modulo
    return ((cls + a.ph - kh) % a.sh) == 0;
}
// output row index (in units of rows of the [.., Cout] output) of GEMM row m
__device__ __forceinline__ long long tc_out_row(const AGather& a, unsigned m) {
    if (a.mode == 2 && a.classes > 1) {
        const RowPos r = tc_decode_row(a, m);
        return ((long long)r.nb * a.OH + r.oh) * a.OW + r.ow;
    }
    return m;
}

struct TcParams {
    int M, N, K;
    int tiles_n;
    unsigned tiles;
    AGather a;
    Epilogue e;
    int epi_vec;                 // output / residual rows are 16-byte aligned and N % 4 == 0 (N % 8 for GLU)
    unsigned stiles;             // super tiles = ceil(row blocks / CL) x column blocks x K slices (CL = cluster size)
    int ksplit;                  // MODE 3: the K range is cut into `ksplit` slices, slice s -> out + s * split_stride
    long long split_stride;      //         (raw partial sums; bias / activation / residual happen in the reducing kernel)
    // MODE 4 (see tc_gemm_kernel): a row block is c4_rh whole output rows (oh) of ONE image — c4_rh * OW <= 128 GEMM rows —,
    // c4_tpi row blocks per image, c4_tiles_m in all; or, for rows longer than a tile (c4_cw < OW, c4_rh = 1), c4_cw
    // consecutive columns of one output row, c4_tpr such blocks per row
    int c4_rh, c4_tpi, c4_tiles_m, c4_cw, c4_tpr;
    // transposed convolution on the same kernel (c4_t = 1; stride_w = 1): a row block holds c4_rh output rows of ONE row
    // class (oh = cls + sh i: their input rows (oh + ph - kh) / sh are consecutive) or a column chunk of one row;
    // c4_ct[cls] row blocks per image and class, a.class_rows[cls] output rows per image and class
    int c4_t, c4_ct[4];
    int dbg;                     // debug builds: bit 0 skip the A stores, bit 1 skip the TMA loads, bit 2 skip the epilogue body
    unsigned long long* trace;   // debug builds (-DAPSB_TC_TRACE): [0] = event counter, then (event << 48 | clock) words
    int grp_rows, grp_n;         // MODE 5: activation rows per group (a multiple of 128) and weight rows per group (= N)
};

// Pipeline timeline of CTA 0 for tuning (compiled out unless APSB_TC_TRACE is defined).
#ifdef APSB_TC_TRACE
static unsigned long long* g_tc_trace = nullptr;
static int g_tc_trace_cap = 0;
#define TC_TRACE_WORDS 1024
#define TC_TR(ev)                                                                                        \
    do {                                                                                                 \
        if (p.trace && blockIdx.x == 0 && (!(p.dbg & 16) || (ev) == 3 || (ev) == 9 || (ev) == 4 || (ev) == 6)) {                   \
            const unsigned i_ = atomicAdd(reinterpret_cast<unsigned*>(tr_smem), 1u);                     \
            if (i_ + 1 < TC_TRACE_WORDS)                                                                 \
                tr_smem[1 + i_] = ((unsigned long long)(ev) << 48) | (clock64() & 0xFFFFFFFFFFFFull);    \
        }                                                                                                \
    } while (0)
#else
#define TC_TR(ev) do { } while (0)
#endif

// BN = 256 uses 16-float k-blocks (64-byte swizzle) so that FOUR 48 KB stages fit: with 32-float blocks only two
// 96 KB stages fit and neither the TMA weight loads nor the A gathers can run far enough ahead of the MMAs.
template <int BN, int MODE = 0> struct TcCfg {
    static constexpr int BK = BN == 256 ? 16 : 32;
    static constexpr int SWZ = BK * 4;                                          // bytes per operand row = swizzle span
    // BN = 64 keeps one stage less than would fit: the 48 KB it leaves become L1, which serves the kw reuse of the gathers
    static constexpr int STAGES = BN == 256 ? 4 : 3;
    static constexpr int A_BYTES = TC_BM * BK * 4;
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    // per epilogue warp a 32 x 32 block: rows padded to 36 floats (4 warps), or unpadded with an XOR swizzle of the
    // 16-byte columns (8 warps, MODE 3) — 8 x 4096 B: the padded form would not fit beside three 64 KB stages.  The
    // gather modes keep the small allocation: every KB of shared memory they do not claim stays L1 for the im2col
    // gathers (conv2 went 848 -> 1071 us when this was 32 KB for every mode).
    static constexpr int EPI_BYTES = TC_LIN3(MODE) ? 8 * 32 * 32 * 4 : 4 * 32 * 36 * 4;
#ifdef APSB_TC_TRACE
    static constexpr int TRACE_BYTES = 8 * 1024;
#else
    static constexpr int TRACE_BYTES = 0;
#endif
    static constexpr int EPI_ALLOC = EPI_BYTES;
    static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_ALLOC + 256 + 1024 + TRACE_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
};

// CL > 1: thread-block cluster of CL CTAs that work on CL consecutive ROW blocks of the same column block (and K slice)
// in lock step.  Every CTA fetches 1 / CL of each weight tile and MULTICASTS it to the whole cluster, so a weight tile
// crosses the L2 -> SM fabric once per CL row blocks.  (Round-2 finding: at these shapes the engine is bound by L2 -> SM
// bandwidth — FFN-a moves 210 MB in 28 us, conv2 re-reads its 4.7 MB filter for each of its 1000 row tiles — not by the
// tensor pipe.)  A stage is released to the TMA threads of all CL CTAs by multicast commits (empty barriers count CL).
template <int BN, int MODE, int CL>
__global__ void __launch_bounds__((TcRoles<BN, MODE>::THREADS), 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                   const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                   const __grid_constant__ TcParams p) {
    using C = TcCfg<BN, MODE>;
    constexpr int S = C::STAGES, BK = C::BK, SWZ = C::SWZ, A_BYTES = C::A_BYTES;
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~(uintptr_t)1023);
    uint8_t* after_stages = base + S * C::STAGE_BYTES;
    float* epi_tiles = reinterpret_cast<float*>(after_stages);
    uint64_t* full_a = reinterpret_cast<uint64_t*>(after_stages + C::EPI_ALLOC);
    uint64_t* full_b = full_a + S;
    uint64_t* empty = full_b + S;
    uint64_t* tmem_full = empty + S;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
#ifdef APSB_TC_TRACE
    unsigned long long* tr_smem = reinterpret_cast<unsigned long long*>(after_stages + C::EPI_ALLOC + 256);
    if (threadIdx.x == 0) tr_smem[0] = 0;
#endif

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    static_assert(CL == 1 || MODE != 2, "row classes of a transposed convolution skip k-blocks per tile: no lock step");
    static_assert(CL == 1 || MODE != 4, "the TMA-fed convolution is not clustered");
    const unsigned crank = CL > 1 ? tc_cluster_rank() : 0u;
    // cluster index / number of clusters.  Code generation of this 35 k-instruction function is touchy (measured, B200,
    // same-box A/B of library builds): with these two values in registers and the first epilogue body below the
    // 256-wide im2col kernel of the conformer front runs conv2 in 844 us, with the special registers re-read and the
    // second body in 1068 us (+26 %) — while every other gather-fed instantiation (DCCRN's convolutions and transposed
    // convolutions: 23.9 vs 26.6 ms for the 18 launches of a step) prefers exactly the opposite.  Each gets its form.
    constexpr bool KFORM = MODE == 1 && BN == 256;    // the conformer-front convolutions: see the epilogue
    const unsigned cid_v = blockIdx.x / CL, ncl_v = gridDim.x / CL;
#define cid (KFORM ? cid_v : blockIdx.x / CL)
#define ncl (KFORM ? ncl_v : gridDim.x / CL)
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
    // super tile -> flat tile index of THIS CTA (row block = CL * super row + rank; may lie beyond M: a dummy tile that
    // keeps the lock step, reads zeros and stores nothing)
    auto tile_of = [&](unsigned st) -> unsigned {
        if (CL == 1) return st;
        unsigned t2 = st, ks = 0;
        if (TC_LIN3(MODE) && p.ksplit > 1) {
            t2 = st / (unsigned)p.ksplit;
            ks = st - t2 * (unsigned)p.ksplit;
        }
        const unsigned msup = t2 / (unsigned)p.tiles_n, nb = t2 - msup * (unsigned)p.tiles_n;
        const unsigned flat = (msup * CL + crank) * (unsigned)p.tiles_n + nb;
        return (TC_LIN3(MODE) && p.ksplit > 1) ? flat * (unsigned)p.ksplit + ks : flat;
    };
    // k-blocks are walked tap by tap (convolutions: a k-block never crosses a (kh, kw) tap since Cin % 32 == 0) so that
    // a transposed-convolution tile can skip the taps that are zero for its row class; a linear layer is one "tap"
    // Order inside a tile: kernel row kh (skippable), then channel block cb, then kw INNERMOST: the three kw taps of one
    // (kh, cb) read the same input pixels shifted by one, so consecutive k-blocks hit the lines the previous one pulled
    // into L1 (the tex->L2 sector traffic of a 3x3 convolution drops towards a third).
    // MODE (0 linear, 1 conv2d, 2 conv_transpose2d) is a template parameter: each instantiation carries only its own
    // gather code — the producers are instruction-fetch bound when their per-k-block code does not fit the L0 i-cache.
    constexpr bool LIN = MODE == 0 || TC_LIN3(MODE);
    const int num_kh = LIN ? 1 : p.a.KH;
    const int num_kw = LIN ? 1 : p.a.KW;
    const int kb_per_tap = LIN ? (p.K + BK - 1) / BK : p.a.Cin / BK;
    // The channel blocks are walked in groups of 32 channels whatever BK is (SUB = 2 half-blocks for BK = 16), so the
    // order of the k-steps — and with it every rounding — does not depend on the tile width: a row's result is bit
    // identical for any batch size / sharding (tests: "batch-shard invariance").
    constexpr int SUB = LIN ? 1 : 32 / BK;
    const int cb32_per_tap = LIN ? kb_per_tap : p.a.Cin / 32;
    // MODE 3 work item: (row block, column block, K slice)
    struct TileIdx { int m_blk, n_blk, ks, kb0, kb1; };
    auto decode3 = [&](unsigned tile) {
        TileIdx t;
        unsigned t2 = tile;
        t.ks = 0;
        if (p.ksplit > 1) {
            t2 = tile / (unsigned)p.ksplit;
            t.ks = (int)(tile - t2 * (unsigned)p.ksplit);
        }
        t.m_blk = (int)(t2 / (unsigned)p.tiles_n);
        t.n_blk = (int)(t2 - (unsigned)t.m_blk * (unsigned)p.tiles_n);
        // slice boundaries in units of 32 k (whatever BK is): a row's rounding does not depend on the tile width
        const int groups = (p.K + 31) / 32;
        t.kb0 = (int)((long long)t.ks * groups / p.ksplit) * (32 / BK);
        t.kb1 = (int)((long long)(t.ks + 1) * groups / p.ksplit) * (32 / BK);
        if (t.kb1 > kb_per_tap) t.kb1 = kb_per_tap;
        return t;
    };
    auto tile_class = [&](unsigned tile) {
        const unsigned m_first = (tile / (unsigned)p.tiles_n) * TC_BM;
        const unsigned m_last = m_first + TC_BM - 1 < (unsigned)p.M ? m_first + TC_BM - 1 : (unsigned)p.M - 1;
        return tc_tile_class(p.a, m_first, m_last);
    };

    // MODE 4 row block -> image, row class (-1: plain convolution), first (class) row, first column, GEMM rows in use
    struct Tile4 { unsigned img; int cls, i0, ow0, valid; };
    auto decode4 = [&](unsigned m_blk) {
        Tile4 t;
        t.img = m_blk / (unsigned)p.c4_tpi;
        unsigned j = m_blk - t.img * (unsigned)p.c4_tpi;
        t.cls = -1;
        int rows_all = p.a.OH;
        if (p.c4_t) {
            int c = 0;
            while (c < 3 && j >= (unsigned)p.c4_ct[c]) { j -= (unsigned)p.c4_ct[c]; ++c; }
            t.cls = c;
            rows_all = p.a.class_rows[c];
        }
        if (p.c4_tpr > 1) {                // (row, column chunk)
            t.i0 = (int)(j / (unsigned)p.c4_tpr);
            t.ow0 = (int)(j - (unsigned)t.i0 * (unsigned)p.c4_tpr) * p.c4_cw;
            t.valid = min(p.c4_cw, p.a.OW - t.ow0);
        } else {
            t.i0 = (int)j * p.c4_rh;
            t.ow0 = 0;
            t.valid = min(p.c4_rh, rows_all - t.i0) * p.a.OW;
        }
        return t;
    };
    // Role -> warp assignment: the SM's schedulers prefer the HIGHEST warp id among eligible warps of a sub-partition
    // (warp % 4).  The A producers are the throughput-critical role, so they get the top ids (6-9); the single-thread
    // TMA / MMA roles (warps 0 / 1) mostly wait and must never out-prioritise them; warps 2-5 are the epilogue (TMEM
    // lane quarter = warp % 4).
    if (warp == WARP_TMA && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
        if (TC_LIN3(MODE)) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAlo) : "memory");
        }
        if (MODE == 4) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAlo) : "memory");
        }
    }
    if (warp == WARP_MMA) {
        if (lane == 0) {
            for (int s = 0; s < S; ++s) {
                tc_mbar_init(full_a + s, TC_LIN3(MODE) ? 1 : TcRoles<BN, MODE>::PW);
                tc_mbar_init(full_b + s, 1);
                tc_mbar_init(empty + s, CL);
            }
            for (int b = 0; b < 2; ++b) {
                tc_mbar_init(tmem_full + b, 1);
                tc_mbar_init(tmem_empty + b, TcRoles<BN, MODE>::EW);
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
    if (CL > 1) tc_cluster_sync();          // every CTA's barriers are initialised before a peer may signal them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TC_TR(9);
    // PDL: everything above (mbarriers, TMEM allocation, tensor-map prefetch) may overlap the tail of the previous kernel;
    // nothing below may start before that kernel has completed (activations, residuals, output buffers it still reads)
    pdl_wait();

    // W_hi / W_lo tile of one stage: the whole box (CL = 1), or this CTA's 1 / CL of its rows multicast to the cluster
    auto load_w = [&](uint64_t* bar, uint8_t* dst, int k0, int n0) {
        if (CL == 1) {
            tc_tma_load_2d(&tmB, bar, dst, k0, n0);
            tc_tma_load_2d(&tmBlo, bar, dst + C::B_BYTES, k0, n0);
        } else {
            constexpr int SL = BN / CL;                               // rows of the slice (a multiple of 8: swizzle atoms)
            const uint32_t off = crank * (uint32_t)(SL * SWZ);
            tc_tma_load_2d_mc(&tmB, bar, dst + off, k0, n0 + (int)crank * SL, CMASK);
            tc_tma_load_2d_mc(&tmBlo, bar, dst + C::B_BYTES + off, k0, n0 + (int)crank * SL, CMASK);
        }
    };
    if (warp == WARP_TMA) {
        // ================= TMA producer: weight tiles =================
        if (TC_LIN3(MODE)) {
            // operands of both sides by TMA: [A (raw x = hi) | A lo | W hi | W lo] per stage, one mbarrier transaction
            if (lane == 0) {
                uint32_t it = 0;
                for (unsigned st_ = cid; st_ < p.stiles; st_ += ncl) {
                    const unsigned tile = tile_of(st_);
                    const TileIdx t = decode3(tile);
                    for (int kb = t.kb0; kb < t.kb1; ++kb, ++it) {
                        const int s = it % S;
                        const uint32_t ph = (it / S) & 1;
                        tc_mbar_wait_parked(empty + s, ph ^ 1);
                        uint8_t* st = base + s * C::STAGE_BYTES;
                        tc_mbar_expect_tx(full_b + s, 2 * A_BYTES + 2 * C::B_BYTES);
                        tc_tma_load_2d(&tmA, full_b + s, st, kb * BK, t.m_blk * TC_BM);
                        tc_tma_load_2d(&tmAlo, full_b + s, st + A_BYTES, kb * BK, t.m_blk * TC_BM);
                        load_w(full_b + s, st + 2 * A_BYTES, kb * BK,
                               t.n_blk * BN + (MODE == 5 ? (t.m_blk * TC_BM / p.grp_rows) * p.grp_n : 0));
                        TC_TR(1);
                    }
                }
            }
        } else if constexpr (MODE == 4) {
            // MODE 4: the NHWC input is the 4-D tensor [image][h][w][c] of a tiled tensor map whose TRAVERSAL strides are the
            // convolution strides, so the A tile of tap (kh, kw) for c4_rh whole output rows (oh0.., all OW columns) is ONE box
            // {32 (16) channels, OW pixels from w = kw dw - pw in steps of sw, c4_rh rows from h = oh0 sh + kh dh - ph in steps
            // of sh, one image}.  Rows land in GEMM-row order (ow fastest) as 128-byte rows under the hardware swizzle; the
            // zero padding is the TMA's out-of-bounds fill (coordinates may be negative).  No gather instructions at all.
            if (lane == 0) {
                uint32_t it = 0;
                const uint32_t a_box = (uint32_t)(p.c4_rh * p.c4_cw) * (uint32_t)SWZ;
                for (unsigned st_ = cid; st_ < p.stiles; st_ += ncl) {
                    const unsigned m_blk = st_ / (unsigned)p.tiles_n;
                    const int n_blk = (int)(st_ - m_blk * (unsigned)p.tiles_n);
                    const Tile4 t = decode4(m_blk);
                    // convolution: rows oh0.. every sh, columns every sw (the map's traversal strides); transposed (sw = 1,
                    // plain box): class rows i0.. read input rows (cls + ph - kh) / sh + i0.., columns ow0 + pw - kw..
                    const int hc = t.cls < 0 ? t.i0 * p.a.sh - p.a.ph : t.i0;
                    const int wc = t.cls < 0 ? t.ow0 * p.a.sw - p.a.pw : t.ow0 + p.a.pw;
                    for (int kh = 0; kh < num_kh; ++kh) {
                        if (!tc_kh_valid(p.a, t.cls, kh)) continue;      // taps that are zero for this row class
                        const int h0 = t.cls < 0 ? hc + kh * p.a.dh : hc + (t.cls + p.a.ph - kh) / p.a.sh;
                        for (int cb32 = 0; cb32 < cb32_per_tap; ++cb32)
                        for (int kw = 0; kw < num_kw; ++kw)
                        for (int h = 0; h < SUB; ++h, ++it) {
                            const int kb = (kh * num_kw + kw) * kb_per_tap + cb32 * SUB + h;
                            const int s = it % S;
                            const uint32_t ph = (it / S) & 1;
                            tc_mbar_wait_parked(empty + s, ph ^ 1);
                            uint8_t* st = base + s * C::STAGE_BYTES;
                            tc_mbar_expect_tx(full_b + s, a_box + 2 * C::B_BYTES);
                            int c0 = cb32 * 32 + h * BK;
                            const CUtensorMap* am = &tmA;
                            if (p.a.cat_c) {               // [re_a | re_b | im_a | im_b]: which source, which of its channels
                                const int seg = (c0 >= p.a.cat_c) + (c0 >= 2 * p.a.cat_c) + (c0 >= 3 * p.a.cat_c);
                                c0 = (seg >> 1) * p.a.cat_c + (c0 - seg * p.a.cat_c);
                                if (seg & 1) am = &tmAlo;
                            }
                            tc_tma_load_4d(am, full_b + s, st, c0, t.cls < 0 ? wc + kw * p.a.dw : wc - kw, h0, (int)t.img);
                            load_w(full_b + s, st + 2 * A_BYTES, kb * BK, n_blk * BN);
                            TC_TR(1);
                        }
                    }
                }
            }
        } else if (lane == 0) {
            uint32_t it = 0;
            for (unsigned st_ = cid; st_ < p.stiles; st_ += ncl) {
                const unsigned tile = tile_of(st_);
                const int n_blk = (int)(tile % (unsigned)p.tiles_n);
                const int cls = tile_class(tile);
                for (int kh = 0; kh < num_kh; ++kh) {
                    if (!tc_kh_valid(p.a, cls, kh)) continue;
                    for (int cb32 = 0; cb32 < cb32_per_tap; ++cb32)
                    for (int kw = 0; kw < num_kw; ++kw)
                    for (int h = 0; h < SUB; ++h, ++it) {
                        const int kb = (kh * num_kw + kw) * kb_per_tap + cb32 * SUB + h;
                        const int s = it % S;
                        const uint32_t ph = (it / S) & 1;
                        tc_mbar_wait_parked(empty + s, ph ^ 1);
                        uint8_t* st = base + s * C::STAGE_BYTES + 2 * A_BYTES;
#ifdef APSB_TC_TRACE
                        if (p.dbg & 2) {
                            tc_mbar_arrive(full_b + s);
                            continue;
                        }
#endif
                        tc_mbar_expect_tx(full_b + s, 2 * C::B_BYTES);
                        load_w(full_b + s, st, kb * BK, n_blk * BN);
                        TC_TR(1);
                    }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B tf32, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(TC_BM >> 4) << 24);
            uint32_t it = 0, tcount = 0;
            for (unsigned st_ = cid; st_ < p.stiles; st_ += ncl, ++tcount) {
                const unsigned tile = tile_of(st_);
                const uint32_t buf = tcount & 1;
                tc_mbar_wait_parked(tmem_empty + buf, ((tcount >> 1) & 1) ^ 1);     // epilogue has drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                const int cls = MODE == 4 ? decode4(tile / (unsigned)p.tiles_n).cls : tile_class(tile);
                uint32_t first = 1;
                if (TC_LIN3(MODE)) {
                    const TileIdx t = decode3(tile);
                    for (int kb = t.kb0; kb < t.kb1; ++kb, ++it) {
                        const int s = it % S;
                        const uint32_t ph = (it / S) & 1;
                        tc_mbar_wait(full_b + s, ph);
                        tc_fence_after();
                        TC_TR(2);
                        const uint32_t sa = s_u32(base + s * C::STAGE_BYTES);
                        const uint32_t sal = sa + A_BYTES, sb = sa + 2 * A_BYTES, sbl = sb + C::B_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
                            const uint64_t da = tc_smem_desc<SWZ>(sa + k * 32), dal = tc_smem_desc<SWZ>(sal + k * 32);
                            const uint64_t db = tc_smem_desc<SWZ>(sb + k * 32), dbl = tc_smem_desc<SWZ>(sbl + k * 32);
                            tc_mma_tf32(d_tmem, da, db, idesc, (first && k == 0) ? 0u : 1u);
                            tc_mma_tf32(d_tmem, da, dbl, idesc, 1);
                            tc_mma_tf32(d_tmem, dal, db, idesc, 1);
                        }
                        first = 0;
                        if (CL == 1) tc_commit(empty + s); else tc_commit_mc(empty + s, CMASK);
                    }
                } else
                for (int kh = 0; kh < num_kh; ++kh) {
                  if (!tc_kh_valid(p.a, cls, kh)) continue;
                  for (int j = 0; j < kb_per_tap * num_kw; ++j, ++it) {
                    const int s = it % S;
                    const uint32_t ph = (it / S) & 1;
                    tc_mbar_wait(full_a + s, ph);
                    TC_TR(11);
                    tc_mbar_wait(full_b + s, ph);
                    tc_fence_after();
                    TC_TR(2);
                    const uint32_t sa = s_u32(base + s * C::STAGE_BYTES);
                    const uint32_t sal = sa + A_BYTES, sb = sa + 2 * A_BYTES, sbl = sb + C::B_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t da = tc_smem_desc<SWZ>(sa + k * 32), dal = tc_smem_desc<SWZ>(sal + k * 32);
                        const uint64_t db = tc_smem_desc<SWZ>(sb + k * 32), dbl = tc_smem_desc<SWZ>(sbl + k * 32);
                        tc_mma_tf32(d_tmem, da, db, idesc, (first && k == 0) ? 0u : 1u);
                        tc_mma_tf32(d_tmem, da, dbl, idesc, 1);
                        tc_mma_tf32(d_tmem, dal, db, idesc, 1);
                    }
                    first = 0;
                    // frees the stage once the MMAs above have read it (in every CTA of the cluster: multicast fills)
                    if (CL == 1) tc_commit(empty + s); else tc_commit_mc(empty + s, CMASK);
                  }
                }
                tc_commit(tmem_full + buf);        // accumulator complete
                TC_TR(3);
            }
        }
    } else if (warp < WARP_EPI0 + TcRoles<BN, MODE>::EW) {
        // ================= epilogue warps 2..5: TMEM lane quarter = warp % 4 =================
        // Fast path (p.epi_vec: 16-byte aligned output / residual rows, N % 4 == 0): lane = row straight out of TMEM,
        // vector loads / stores on the lane's own row.  Fallback (e.g. N = 257 mask rows): the 32x32 block is transposed
        // through a padded shared tile so that lane = column and the scalar accesses are still 128-byte rows.
        const int q = warp & 3;
        constexpr bool EW8 = TcRoles<BN, MODE>::EW == 8;
        const int ehalf = EW8 ? ((warp - WARP_EPI0) >> 2) : 0;     // which of the two warps of this lane quarter
        constexpr int ESTEP = EW8 ? 2 : 1;                         // chunks are dealt out alternately
        float* tile_s = epi_tiles + (warp - WARP_EPI0) * (EW8 ? 32 * 32 : 32 * 36);
        // byte offset of the 16-byte column c4 of row r in the warp's tile
        auto ts_off = [&](uint32_t r, uint32_t c4) -> uint32_t {
            return EW8 ? r * 128u + ((c4 ^ (r & 7u)) << 4) : r * 144u + (c4 << 4);
        };
        const Epilogue& e = p.e;
        const bool glu = e.act == ACT_GLU;
        uint32_t tcount = 0;
        for (unsigned st_ = cid; st_ < p.stiles; st_ += ncl, ++tcount) {
            const unsigned tile = tile_of(st_);
            int n_blk = (int)(tile % (unsigned)p.tiles_n);
            unsigned m_blk = tile / (unsigned)p.tiles_n;
            float* eout = e.out;                   // split-K: slice ks of the partial-sum workspace
            if (TC_LIN3(MODE)) {
                const TileIdx t = decode3(tile);
                n_blk = t.n_blk; m_blk = (unsigned)t.m_blk;
                eout += (long long)t.ks * p.split_stride;
            }
            const uint32_t buf = tcount & 1;
            // first GEMM row of this warp's lane quarter and the number of valid rows from there on (MODE 4: a row block is
            // c4_rh whole output rows of one image, packed from row 0 of the tile)
            long long m0_ = (long long)m_blk * TC_BM + q * 32, rows_ = 0;
            bool row_ok4 = false;
            long long mrow4 = 0;
            if constexpr (MODE == 4) {
                const Tile4 t = decode4(m_blk);
                // GEMM row r = i * OW + ow (whole rows) or ow (column chunk); output row of (i, ow): rows of a transposed
                // convolution's class lie sh apart
                const int r = q * 32 + lane;
                const int i = p.c4_tpr > 1 ? 0 : r / p.a.OW;
                const int orow = t.cls < 0 ? t.i0 + i : t.cls + p.a.sh * (t.i0 + i);
                row_ok4 = r < t.valid;
                mrow4 = ((long long)t.img * p.a.OH + orow) * p.a.OW + t.ow0 + (r - i * p.a.OW);
                m0_ = 0;
                rows_ = (long long)t.valid - q * 32;
            } else {
                rows_ = (long long)p.M - m0_;
            }
            const long long m0 = m0_;
            const int ncols = min(BN, p.N - n_blk * BN);
            const int nchunks = (ncols + 31) >> 5;
            const long long rows_ll = rows_;
            const int rows = rows_ll >= 32 ? 32 : (rows_ll > 0 ? (int)rows_ll : 0);
#ifdef APSB_TC_TRACE
            const bool row_ok = (MODE == 4 ? row_ok4 : lane < rows) && !(p.dbg & 64);      // ablation: no residual loads / output stores
#else
            const bool row_ok = MODE == 4 ? row_ok4 : lane < rows;
#endif
            // row of the output / residual matrices
            const long long mrow = MODE == 4 ? mrow4 : (row_ok ? tc_out_row(p.a, (unsigned)(m0 + lane)) : 0);
            // Vector path: TMEM delivers lane = row.  Storing that way makes every STG.128 of a warp touch 32 different
            // lines (32 LSU cycles per instruction — the epilogue, not the MMAs, bounded the K = 256 layers: r02b
            // microbenchmark, +8 us for a second output).  The 32 x 32 block is therefore transposed through this warp's
            // padded shared tile and lane (g = lane / 8, c = lane % 8) handles columns 4c..4c+3 of rows g, g+4, ..., g+28:
            // a warp instruction reads / writes four full 128-byte row segments, the residual comes in coalesced, and a
            // lane's bias / affine / slope columns are the same for all 8 rows (one 16-byte load per chunk).
            const int er = lane >> 3, ec = (lane & 7) * 4;
            int mr[8];                             // output row of (er + 4 i) (rows < 2^31), -1: beyond M
#pragma unroll
            for (int i = 0; i < 8; ++i)
                mr[i] = __shfl_sync(0xffffffffu, row_ok ? (int)mrow : -1, er + 4 * i);
            float4 res_nxt[8];
            auto fetch_res = [&](int ch, float4 (&dst)[8]) {
                const int n = n_blk * BN + ch * 32 + ec;
                const bool ok = p.epi_vec && e.res != nullptr && n < p.N;
                const long long col = glu ? (n >> 1) : n;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok && mr[i] >= 0) {
                        if (glu) {
                            const float2 t = __ldg(reinterpret_cast<const float2*>(e.res + (long long)mr[i] * e.ldres + col));
                            dst[i].x = t.x; dst[i].y = t.y;
                        } else {
                            dst[i] = __ldg(reinterpret_cast<const float4*>(e.res + (long long)mr[i] * e.ldres + col));
                        }
                    }
                }
            };
            // the first residual chunk is requested BEFORE the wait, i.e. while the tile's main loop is still running
            fetch_res(ehalf, res_nxt);
            if (lane == 0) tc_mbar_wait_parked(tmem_full + buf, (tcount >> 1) & 1);
            __syncwarp();
            tc_fence_after();
            if (threadIdx.x == 64) TC_TR(4);
            if (EW8 && ehalf >= nchunks) {         // nothing for this warp in a one-chunk tile: just hand the buffer back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(tmem_empty + buf);
            }
#pragma unroll 1
            for (int ch = ehalf; ch < nchunks; ch += ESTEP) {
                const int c0 = ch * 32;
                const int n0 = n_blk * BN + c0;
                uint32_t r[32];
                float4 res_cur[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) res_cur[j] = res_nxt[j];
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
                if (ch + ESTEP < nchunks) fetch_res(ch + ESTEP, res_nxt);
                // Two textually separate bodies on purpose (see KFORM above): the producer warps of the gather-fed
                // instantiations are instruction-fetch sensitive and the layout of this function decides their speed.
                if constexpr (!KFORM) {
                // per-column vectors of this lane's 4 columns: requested while the TMEM load is in flight
                const int n = n0 + ec;
                const bool nok = n < p.N;          // N % 4 == 0 (N % 8 for GLU) on the vector path
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 ps4 = make_float4(1.f, 1.f, 1.f, 1.f), pt4 = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 sl4 = make_float4(e.leak, e.leak, e.leak, e.leak);
                constexpr bool LEAN = TcRoles<BN, MODE>::THREADS > 400;   // 128-register variants: load them late instead
                auto load_cols = [&]() {
                    if ((TC_LIN3(MODE) || p.epi_vec) && nok) {
                        if (e.bias && !e.dbg_nobias) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
                        if (e.post_scale) {
                            ps4 = __ldg(reinterpret_cast<const float4*>(e.post_scale + n));
                            pt4 = __ldg(reinterpret_cast<const float4*>(e.post_shift + n));
                        }
                        if (e.act == ACT_PRELU) {
                            if (e.slope_stride) sl4 = __ldg(reinterpret_cast<const float4*>(e.slope + n));
                            else { const float s0 = __ldg(e.slope); sl4 = make_float4(s0, s0, s0, s0); }
                        }
                    }
                };
                if (!LEAN) load_cols();
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ch + ESTEP >= nchunks) {       // this warp's last read of the accumulator buffer: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(tmem_empty + buf);
                    if (threadIdx.x == 64) TC_TR(5);
                }
                if (TC_LIN3(MODE) || p.epi_vec) {      // MODE 3 is only launched with the vector epilogue (host check)
                    // ---- transpose: lane = row -> lane = (row group, 4 columns) ----
                    __syncwarp();                  // the previous chunk's reads of the tile are done
                    const uint32_t ts_b = s_u32(tile_s);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        tc_sts128(ts_b + ts_off((uint32_t)lane, (uint32_t)j4), r[4 * j4], r[4 * j4 + 1], r[4 * j4 + 2], r[4 * j4 + 3]);
                    __syncwarp();
                    // the rows of this lane come back GRP at a time BEFORE the arithmetic and global stores of the group
                    // (1 in the 448-thread variants, whose 128-register budget is taken by the producers' gather ring)
                    constexpr int GRP = LEAN ? 1 : 4;
                    if (LEAN) load_cols();
                    if (TC_LIN3(MODE) && p.ksplit > 1) {
                        // split-K slice: RAW partial sums (bias, activation, residual and the companion belong to the
                        // reducing kernel) — straight from the tile to memory.  Through the general path below this cost
                        // ~900 instructions per 32 x 32 block: 17 k cycles of a 39 k-cycle FFN-b work item (trace r02k).
#pragma unroll
                        for (int g = 0; g < 8; g += 4) {
                            float4 tv[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) tv[i] = tc_lds128(ts_b + ts_off((uint32_t)(er + 4 * (g + i)), (uint32_t)(ec >> 2)));
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (nok && mr[g + i] >= 0) *reinterpret_cast<float4*>(eout + (long long)mr[g + i] * e.ldo + n) = tv[i];
                        }
                    } else if (glu) {
                        // columns (2j, 2j+1) -> output column j: this lane's 4 columns give 2 outputs
#pragma unroll
                        for (int g = 0; g < 8; g += GRP) {
                            float4 tv[GRP];
#pragma unroll
                            for (int i = 0; i < GRP; ++i) tv[i] = tc_lds128(ts_b + ts_off((uint32_t)(er + 4 * (g + i)), (uint32_t)(ec >> 2)));
#pragma unroll
                            for (int i = 0; i < GRP; ++i) {
                                const float4 v = tv[i];
                                float2 o;
                                o.x = e.alpha * ((v.x + b4.x) * tc_glu_gate(v.y + b4.y));
                                o.y = e.alpha * ((v.z + b4.z) * tc_glu_gate(v.w + b4.w));
                                o.x = fmaf(e.beta, res_cur[g + i].x, o.x);
                                o.y = fmaf(e.beta, res_cur[g + i].y, o.y);
                                if (nok && mr[g + i] >= 0) *reinterpret_cast<float2*>(eout + (long long)mr[g + i] * e.ldo + (n >> 1)) = o;
                            }
                        }
                    } else {
#define TC_EPI_CASE(A)                                                                                                  \
    case A:                                                                                                             \
        _Pragma("unroll") for (int g = 0; g < 8; g += GRP) {                                                            \
            float4 tv[GRP];                                                                                             \
            _Pragma("unroll") for (int i = 0; i < GRP; ++i)                                                             \
                tv[i] = tc_lds128(ts_b + ts_off((uint32_t)(er + 4 * (g + i)), (uint32_t)(ec >> 2)));                    \
            _Pragma("unroll") for (int i = 0; i < GRP; ++i) {                                                           \
                const float4 o = tc_epilogue4<A>(tv[i], b4, ps4, pt4, sl4, res_cur[g + i], e.alpha, e.beta);            \
                if (nok && mr[g + i] >= 0) {                                                                            \
                    *reinterpret_cast<float4*>(eout + (long long)mr[g + i] * e.ldo + n) = o;                            \
                    if (e.out_lo) *reinterpret_cast<float4*>(e.out_lo + (long long)mr[g + i] * e.ldo + n) = tf32_lo4(o); \
                }                                                                                                       \
            }                                                                                                           \
        }                                                                                                               \
        break;
                        // bias + activation only (no post affine, alpha = 1, no residual: FFN-a, QKV): three dependent
                        // operations per element less than the general form
#define TC_EPI_PLAIN(A)                                                                                                 \
        _Pragma("unroll") for (int g = 0; g < 8; g += GRP) {                                                            \
            float4 tv[GRP];                                                                                             \
            _Pragma("unroll") for (int i = 0; i < GRP; ++i)                                                             \
                tv[i] = tc_lds128(ts_b + ts_off((uint32_t)(er + 4 * (g + i)), (uint32_t)(ec >> 2)));                    \
            _Pragma("unroll") for (int i = 0; i < GRP; ++i) {                                                           \
                const float4 o = make_float4(tc_act<A>(tv[i].x + b4.x, 0.f), tc_act<A>(tv[i].y + b4.y, 0.f),            \
                                             tc_act<A>(tv[i].z + b4.z, 0.f), tc_act<A>(tv[i].w + b4.w, 0.f));           \
                if (nok && mr[g + i] >= 0) {                                                                            \
                    *reinterpret_cast<float4*>(eout + (long long)mr[g + i] * e.ldo + n) = o;                            \
                    if (e.out_lo) *reinterpret_cast<float4*>(e.out_lo + (long long)mr[g + i] * e.ldo + n) = tf32_lo4(o); \
                }                                                                                                       \
            }                                                                                                           \
        }
                        const bool plain = TC_LIN3(MODE) && !e.res && !e.post_scale && e.alpha == 1.f;
                        if (plain && e.act == ACT_SWISH) {
                            TC_EPI_PLAIN(ACT_SWISH)
                        } else if (plain && e.act == ACT_NONE) {
                            TC_EPI_PLAIN(ACT_NONE)
                        } else
#undef TC_EPI_PLAIN
                        switch (e.act) {
                            TC_EPI_CASE(ACT_RELU)
                            TC_EPI_CASE(ACT_SWISH)
                            TC_EPI_CASE(ACT_TANH)
                            TC_EPI_CASE(ACT_SIGMOID)
                            TC_EPI_CASE(ACT_PRELU)
                            TC_EPI_CASE(ACT_LEAKY)
                            TC_EPI_CASE(ACT_GELU)
                            default:
                            TC_EPI_CASE(ACT_NONE)
                        }
#undef TC_EPI_CASE
                    }
                } else if constexpr (!TC_LIN3(MODE)) {
                    // ---- transposed scalar fallback (unaligned rows): same code as in the other body below ----
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; ++j) tile_s[lane * 33 + j] = __uint_as_float(r[j]);
                    __syncwarp();
                    const int n = n0 + lane;                          // this lane's column from here on (shadows the 4-column index)
                    const bool nok = n < p.N;
                    const float bias = (e.bias && nok) ? __ldg(e.bias + n) : 0.f;
                    if (glu) {
                        const int no = n >> 1;
                        for (int rr = 0; rr < rows; ++rr) {
                            const float v = tile_s[rr * 33 + lane] + bias;
                            const float g = __shfl_down_sync(0xffffffffu, v, 1);
                            if (!(lane & 1) && n + 1 < p.N) {
                                const long long m = tc_out_row(p.a, (unsigned)(m0 + rr));
                                float o = e.alpha * (v * tc_glu_gate(g));
                                if (e.res) o = fmaf(e.beta, __ldg(e.res + m * e.ldres + no), o);
                                eout[m * e.ldo + no] = o;
                            }
                        }
                    } else {
                        const float ps = (e.post_scale && nok) ? __ldg(e.post_scale + n) : 1.f;
                        const float pt = (e.post_scale && nok) ? __ldg(e.post_shift + n) : 0.f;
                        const float slope =
                            (e.act == ACT_PRELU && nok) ? __ldg(e.slope + (long long)n * e.slope_stride) : e.leak;
#pragma unroll 4
                        for (int rr = 0; rr < rows; ++rr) {
                            float v = tile_s[rr * 33 + lane] + bias;
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
                            if (nok) {
                                const long long m = tc_out_row(p.a, (unsigned)(m0 + rr));
                                if (e.res) v = fmaf(e.beta, __ldg(e.res + m * e.ldres + n), v);
                                eout[m * e.ldo + n] = v;
                            }
                        }
                    }
                    __syncwarp();                  // tile_s is reused by the next chunk
                }
                } else {
                // per-column vectors of this lane's 4 columns: requested while the TMEM load is in flight
                const int n = n0 + ec;
                const bool nok = n < p.N;          // N % 4 == 0 (N % 8 for GLU) on the vector path
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 ps4 = make_float4(1.f, 1.f, 1.f, 1.f), pt4 = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 sl4 = make_float4(e.leak, e.leak, e.leak, e.leak);
                // the 448-thread variants (8 producer warps, 128 registers) request these AFTER the transposition and read
                // the tile back one row at a time: the early / batched form spills there (DCCRN 43.6 -> 45.9 ms)
                constexpr bool LEAN = TcRoles<BN, MODE>::THREADS > 400;
                if (!LEAN && p.epi_vec && nok) {
                    if (e.bias && !e.dbg_nobias) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
                    if (e.post_scale) {
                        ps4 = __ldg(reinterpret_cast<const float4*>(e.post_scale + n));
                        pt4 = __ldg(reinterpret_cast<const float4*>(e.post_shift + n));
                    }
                    if (e.act == ACT_PRELU) {
                        if (e.slope_stride) sl4 = __ldg(reinterpret_cast<const float4*>(e.slope + n));
                        else { const float s0 = __ldg(e.slope); sl4 = make_float4(s0, s0, s0, s0); }
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ch == nchunks - 1) {           // last read of this accumulator buffer: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(tmem_empty + buf);
                    if (threadIdx.x == 64) TC_TR(5);
                }
                if (p.epi_vec) {
                    // ---- transpose: lane = row -> lane = (row group, 4 columns) ----
                    __syncwarp();                  // the previous chunk's reads of the tile are done
                    const uint32_t ts_w = s_u32(tile_s) + (uint32_t)lane * 144u;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        tc_sts128(ts_w + 16u * j4, r[4 * j4], r[4 * j4 + 1], r[4 * j4 + 2], r[4 * j4 + 3]);
                    __syncwarp();
                    // the rows of this lane come back four at a time BEFORE any arithmetic or global store of the group
                    const uint32_t ts_r = s_u32(tile_s) + (uint32_t)er * 144u + (uint32_t)ec * 4u;
                    float4 tv[8];
                    if constexpr (LEAN) {
                        if (nok) {
                            if (e.bias && !e.dbg_nobias) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
                            if (e.post_scale) {
                                ps4 = __ldg(reinterpret_cast<const float4*>(e.post_scale + n));
                                pt4 = __ldg(reinterpret_cast<const float4*>(e.post_shift + n));
                            }
                            if (e.act == ACT_PRELU) {
                                if (e.slope_stride) sl4 = __ldg(reinterpret_cast<const float4*>(e.slope + n));
                                else { const float s0 = __ldg(e.slope); sl4 = make_float4(s0, s0, s0, s0); }
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) tv[i] = tc_lds128(ts_r + (uint32_t)i * (4u * 144u));
                    }
                    if (glu) {
                        // columns (2j, 2j+1) -> output column j: this lane's 4 columns give 2 outputs
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if constexpr (LEAN) {
                                tv[i] = tc_lds128(ts_r + (uint32_t)i * (4u * 144u));
                            } else if (i == 4) {
#pragma unroll
                                for (int i2 = 4; i2 < 8; ++i2) tv[i2] = tc_lds128(ts_r + (uint32_t)i2 * (4u * 144u));
                            }
                            const float4 v = tv[i];
                            float2 o;
                            o.x = e.alpha * ((v.x + b4.x) * tc_glu_gate(v.y + b4.y));
                            o.y = e.alpha * ((v.z + b4.z) * tc_glu_gate(v.w + b4.w));
                            o.x = fmaf(e.beta, res_cur[i].x, o.x);
                            o.y = fmaf(e.beta, res_cur[i].y, o.y);
                            if (nok && mr[i] >= 0) *reinterpret_cast<float2*>(eout + (long long)mr[i] * e.ldo + (n >> 1)) = o;
                        }
                    } else {
#define TC_EPI_CASE(A)                                                                                                  \
    case A:                                                                                                             \
        _Pragma("unroll") for (int i = 0; i < 8; ++i) {                                                                 \
            if constexpr (LEAN) {                                                                                       \
                tv[i] = tc_lds128(ts_r + (uint32_t)i * (4u * 144u));                                                    \
            } else if (i == 4) {                                                                                        \
                _Pragma("unroll") for (int i2 = 4; i2 < 8; ++i2) tv[i2] = tc_lds128(ts_r + (uint32_t)i2 * (4u * 144u)); \
            }                                                                                                           \
            const float4 o = tc_epilogue4<A>(tv[i], b4, ps4, pt4, sl4, res_cur[i], e.alpha, e.beta);                    \
            if (nok && mr[i] >= 0) {                                                                                    \
                *reinterpret_cast<float4*>(eout + (long long)mr[i] * e.ldo + n) = o;                                               \
                if (e.out_lo) *reinterpret_cast<float4*>(e.out_lo + (long long)mr[i] * e.ldo + n) = tf32_lo4(o);                   \
            }                                                                                                           \
        }                                                                                                               \
        break;
                        switch (e.act) {
                            TC_EPI_CASE(ACT_RELU)
                            TC_EPI_CASE(ACT_SWISH)
                            TC_EPI_CASE(ACT_TANH)
                            TC_EPI_CASE(ACT_SIGMOID)
                            TC_EPI_CASE(ACT_PRELU)
                            TC_EPI_CASE(ACT_LEAKY)
                            TC_EPI_CASE(ACT_GELU)
                            default:
                            TC_EPI_CASE(ACT_NONE)
                        }
#undef TC_EPI_CASE
                    }
                } else {
                    // ---- transposed scalar fallback (unaligned rows) ----
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; ++j) tile_s[lane * 33 + j] = __uint_as_float(r[j]);
                    __syncwarp();
                    const int n = n0 + lane;                          // this lane's column from here on
                    const bool nok = n < p.N;
                    const float bias = (e.bias && nok) ? __ldg(e.bias + n) : 0.f;
                    if (glu) {
                        const int no = n >> 1;
                        for (int rr = 0; rr < rows; ++rr) {
                            const float v = tile_s[rr * 33 + lane] + bias;
                            const float g = __shfl_down_sync(0xffffffffu, v, 1);
                            if (!(lane & 1) && n + 1 < p.N) {
                                const long long m = tc_out_row(p.a, (unsigned)(m0 + rr));
                                float o = e.alpha * (v * tc_glu_gate(g));
                                if (e.res) o = fmaf(e.beta, __ldg(e.res + m * e.ldres + no), o);
                                eout[m * e.ldo + no] = o;
                            }
                        }
                    } else {
                        const float ps = (e.post_scale && nok) ? __ldg(e.post_scale + n) : 1.f;
                        const float pt = (e.post_scale && nok) ? __ldg(e.post_shift + n) : 0.f;
                        const float slope =
                            (e.act == ACT_PRELU && nok) ? __ldg(e.slope + (long long)n * e.slope_stride) : e.leak;
#pragma unroll 4
                        for (int rr = 0; rr < rows; ++rr) {
                            float v = tile_s[rr * 33 + lane] + bias;
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
                            if (nok) {
                                const long long m = tc_out_row(p.a, (unsigned)(m0 + rr));
                                if (e.res) v = fmaf(e.beta, __ldg(e.res + m * e.ldres + n), v);
                                eout[m * e.ldo + n] = v;
                            }
                        }
                    }
                    __syncwarp();                  // tile_s is reused by the next chunk
                }
                }   // !TC_LIN3(MODE)
            }
            if (threadIdx.x == 64) TC_TR(6);
        }
    } else if constexpr (MODE == 4) {
        // ================= lo converters (warps 6..9) =================
        // The TMA delivered the raw fp32 tile (the tensor core reads it as hi = trunc_tf32(x)); lo = rn_tf32(x - hi) goes to
        // the stage's second A buffer at the SAME byte offsets (the swizzle is a property of the address, and this is an
        // element-wise map), 16-byte chunks dealt out linearly: conflict free, 4 (BK 16) or 8 chunks per thread and k-block.
        const int pt = threadIdx.x - WARP_PROD0 * 32;
        uint32_t it = 0;
        for (unsigned st_ = cid; st_ < p.stiles; st_ += ncl) {
            uint32_t nkb = (uint32_t)(num_kh * num_kw * kb_per_tap);
            if (p.c4_t) {                   // only the taps that exist for this row class
                const int cls = decode4(st_ / (unsigned)p.tiles_n).cls;
                int nv = 0;
                for (int kh = 0; kh < num_kh; ++kh) nv += tc_kh_valid(p.a, cls, kh) ? 1 : 0;
                nkb = (uint32_t)(nv * num_kw * kb_per_tap);
            }
            for (uint32_t j = 0; j < nkb; ++j, ++it) {
                const int s = (int)(it % S);
                const uint32_t ph = (it / S) & 1;
                if (lane == 0) tc_mbar_wait_parked(full_b + s, ph);   // one poller per warp
                __syncwarp();
                const uint32_t sa = s_u32(base + s * C::STAGE_BYTES) + (uint32_t)pt * 16u;
#pragma unroll
                for (int i = 0; i < A_BYTES / (16 * 128); ++i) {
                    const float4 l = tf32_lo4(tc_lds128(sa + (uint32_t)i * 2048u));
                    tc_sts128(sa + (uint32_t)A_BYTES + (uint32_t)i * 2048u, __float_as_uint(l.x), __float_as_uint(l.y),
                              __float_as_uint(l.z), __float_as_uint(l.w));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> UMMA reads
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(full_a + s);      // one arrival per converter warp
            }
        }
    } else if constexpr (!TC_LIN3(MODE)) {
        // ================= A producers (warps 6..9) =================
        // thread -> 16-byte chunk c of rows rg, rg + RSTEP, ...: a warp instruction reads whole SWZ-byte row segments.
        // The gather for k-block i+1 is issued BEFORE block i is split and stored, so the L2 round trip is hidden.
        constexpr int CPR = BK / 4;                 // 16-byte chunks per operand row
        constexpr int RSTEP = TcRoles<BN, MODE>::PRODUCERS / CPR;   // rows covered by one pass of the producer threads
        constexpr int RPT = TC_BM / RSTEP;          // rows per thread (= CPR)
        const int pt = threadIdx.x - WARP_PROD0 * 32;
        const int c = pt % CPR, rg = pt / CPR;
        static_assert(RSTEP % 8 == 0, "rows of one thread must share the swizzle phase");
        const uint32_t prod_base = s_u32(base) + (uint32_t)rg * (uint32_t)SWZ +
                                   (uint32_t)((c ^ (SWZ == 128 ? (rg & 7) : ((rg >> 1) & 3))) << 4);
        const AGather& a = p.a;
        // Per-row geometry is resolved when the tile or the kernel row kh changes (rare); the per-k-block path is then a
        // bounds check on iw, one multiply-add and the load — about 50 instructions instead of 400, which matters
        // because this role runs straight-line code once per k-block and is instruction-fetch bound otherwise.
        const float* prow[RPT];     // linear: row pointer; conv: input row (nb, ih) of this kh; null: nothing to read
        int r_nb[RPT], r_a[RPT], r_b[RPT];   // image index (-1: row >= M); conv: oh*sh - ph / ow*sw - pw; tconv: oh + ph / ow + pw
        int lcls = -1;                          // row class of the current tile (-1: none / straddling)
        auto set_tile = [&](unsigned tile) {
            const unsigned m_blk = tile / (unsigned)p.tiles_n;
            const unsigned m0 = m_blk * TC_BM + rg;
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const unsigned m = m0 + RSTEP * i;
                    const bool in = m < (unsigned)p.M;
                    r_nb[i] = in ? 0 : -1; r_a[i] = 0; r_b[i] = 0;
                    prow[i] = in ? a.x + (long long)m * a.ld : nullptr;
                }
                return;
            }
            // Convolutions: the first row is decoded with divisions, the others follow by stepping RSTEP columns (the
            // divisions are emulated: decoding every row cost ~8 k cycles per tile and thread).  Tiles that straddle two
            // row classes of a strided transposed convolution (a handful per launch) decode every row.
            const bool cls_on = MODE == 2 && a.classes > 1;
            const bool stepping = !cls_on || lcls >= 0;
            const int rows_img = cls_on ? (lcls >= 0 ? a.class_rows[lcls] : 1) : a.OH;
            int nb = 0, j = 0, ow = 0;                      // image, row index inside the image (class), column
            if (stepping && m0 < (unsigned)p.M) {
                const unsigned mm = cls_on ? m0 - (unsigned)a.class_start[lcls] : m0;
                const unsigned t = mm / (unsigned)a.OW;
                ow = (int)(mm - t * (unsigned)a.OW);
                nb = (int)(t / (unsigned)rows_img);
                j = (int)(t - (unsigned)nb * (unsigned)rows_img);
            }
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const unsigned m = m0 + RSTEP * i;
                r_nb[i] = -1; r_a[i] = 0; r_b[i] = 0; prow[i] = nullptr;
                if (m < (unsigned)p.M) {
                    RowPos rp;
                    if (stepping) {
                        rp.nb = nb; rp.ow = ow; rp.oh = cls_on ? j * a.sh + lcls : j;
                        ow += RSTEP;
                        while (ow >= a.OW) {
                            ow -= a.OW;
                            if (++j == rows_img) { j = 0; ++nb; }
                        }
                    } else {
                        rp = tc_decode_row(a, m);
                    }
                    r_nb[i] = rp.nb;
                    if (MODE == 1) { r_a[i] = rp.oh * a.sh - a.ph; r_b[i] = rp.ow * a.sw - a.pw; }
                    else           { r_a[i] = rp.oh + a.ph;        r_b[i] = rp.ow + a.pw; }
                }
            }
        };
        auto set_kh = [&](int kh) {             // convolutions: input row of every tile row for kernel row kh
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                int ih;
                bool ok = r_nb[i] >= 0;
                if (MODE == 1) {
                    ih = r_a[i] + kh * a.dh;
                    ok = ok && ih >= 0 && ih < a.H;
                } else {
                    const int nh = r_a[i] - kh;
                    ok = ok && nh >= 0;
                    if (a.sh == 1) ih = nh;
                    else if (a.sh == 2) { ok = ok && !(nh & 1); ih = nh >> 1; }
                    else { ok = ok && (nh % a.sh) == 0; ih = nh / a.sh; }
                    ok = ok && ih < a.H;
                }
                const int csrc = a.cat_c ? 2 * a.cat_c : a.Cin;          // channels of the tensor(s) actually read
                // pointer to the pixel at iw = r_b[i] (the tap offset kw is added per k-block; may lie outside the row)
                prow[i] = ok ? a.x + (((long long)r_nb[i] * a.H + ih) * a.W + r_b[i]) * csrc : nullptr;
            }
        };
        // Gather of one k-block into registers.  Plain (L1-allocating) loads on purpose: L1 merges a warp's 16-byte lane
        // requests into 128-byte line requests and serves the kw-overlap of neighbouring taps; both L1-bypassing forms
        // that were tried (ld.global.nc.L1::no_allocate, cp.async.cg into a shared-memory ring) were 15-50 % slower.
        auto gather = [&](int cb, int kw, float4 (&v)[RPT]) {
            if (MODE == 0) {
                const int k = cb * BK + c * 4;
#pragma unroll
                for (int i = 0; i < RPT; ++i)
                    v[i] = (prow[i] && k < p.K) ? __ldg(reinterpret_cast<const float4*>(prow[i] + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                // Cin % 32 == 0: the whole k-block lies inside one (kh, kw) tap; stride_w == 1 for MODE 2 (host check)
                int coff = cb * BK + c * 4, csrc = a.Cin;
                long long delta = 0;
                if (a.cat_c) {                   // which of the four parts this (32-channel aligned) k-block lies in
                    const int c0 = cb * BK;                                  // c0 < 4 * cat_c: no integer division
                    const int seg = (c0 >= a.cat_c) + (c0 >= 2 * a.cat_c) + (c0 >= 3 * a.cat_c);
                    coff = (seg >> 1) * a.cat_c + (c0 - seg * a.cat_c) + c * 4;
                    csrc = 2 * a.cat_c;
                    delta = (seg & 1) ? (a.x2 - a.x) : 0;
                }
                const int dwk = MODE == 1 ? kw * a.dw : -kw;
                const long long common = (long long)dwk * csrc + coff + delta;     // same for every row of the k-block
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const bool ok = prow[i] != nullptr && (unsigned)(r_b[i] + dwk) < (unsigned)a.W;
                    v[i] = ok ? __ldg(reinterpret_cast<const float4*>(prow[i] + common)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        // k-blocks of gathers in flight (registers): as deep as the register budget of the role allows — the linear
        // mode carries no per-row convolution geometry, so it affords one more block
        constexpr int D = BK == 16 ? 4 : (MODE == 0 ? 3 : (TcRoles<BN, MODE>::PW == 8 ? 4 : 2));
        // iterator over the (tile, tap, k-block) sequence of this CTA, skipping taps that are zero for the tile's class
        unsigned lst = cid;                                 // super tile; ltile = this CTA's flat tile of it
        unsigned ltile = tile_of(lst);
        int lkh = 0, lcb = 0, lkw = 0, lh = 0;              // lcb counts 32-channel groups, lh the half inside (BK = 16)
        bool lfresh = true;                     // the tile's rows have not been decoded yet
        unsigned lmask = 0, lmask_tile = 0xffffffffu;
        int lkh_set = -1;                       // kernel row the prow[] pointers were computed for
        auto seek = [&]() {                     // move (ltile, lkh) to the next valid kernel row; false at the end
            while (lst < p.stiles) {
                if (lfresh && lmask_tile != ltile) {       // once per tile: class and the bit mask of its non-zero kernel rows
                    lcls = tile_class(ltile);
                    lmask = 0;
                    for (int kh = 0; kh < num_kh && kh < 32; ++kh) lmask |= tc_kh_valid(a, lcls, kh) ? (1u << kh) : 0u;
                    lmask_tile = ltile;
                }
                while (lkh < num_kh && !(lkh < 32 ? (bool)((lmask >> lkh) & 1u) : tc_kh_valid(a, lcls, lkh))) ++lkh;
                if (lkh < num_kh) return true;
                lst += ncl;
                ltile = tile_of(lst);
                lkh = 0;
                lcb = 0;
                lkw = 0;
                lh = 0;
                lfresh = true;
            }
            return false;
        };
        auto step = [&]() {                     // advance by one k-block: (half,) kw innermost, then channel group, then kh
            if (++lh == SUB) {
                lh = 0;
                if (++lkw == num_kw) {
                    lkw = 0;
                    if (++lcb == cb32_per_tap) {
                        lcb = 0;
                        ++lkh;
                    }
                }
            }
        };
        float4 ring[D][RPT];
#ifdef APSB_TC_TRACE
        long long acc_seek = 0, acc_tile = 0, acc_ld = 0;
#endif
        auto gather_next = [&](float4 (&dst)[RPT]) {
#ifdef APSB_TC_TRACE
            const long long g0_ = clock64();
#endif
            if (!seek()) return false;
#ifdef APSB_TC_TRACE
            const long long g1_ = clock64();
#endif
            if (lfresh) {
                set_tile(ltile);
                lfresh = false;
                lkh_set = -1;
            }
            if (MODE != 0 && lkh_set != lkh) {
                set_kh(lkh);
                lkh_set = lkh;
            }
#ifdef APSB_TC_TRACE
            const long long g2_ = clock64();
#endif
            gather(lcb * SUB + lh, lkw, dst);
            step();
#ifdef APSB_TC_TRACE
            acc_seek += g1_ - g0_; acc_tile += g2_ - g1_; acc_ld += clock64() - g2_;
#endif
            return true;
        };
        bool live[D];
#ifdef APSB_TC_TRACE
        long long acc_wait = 0, acc_store = 0, acc_gather = 0;
#endif
#pragma unroll
        for (int d = 0; d < D; ++d) live[d] = gather_next(ring[d]);
        for (long long it0 = 0;; it0 += D) {
            bool any = false;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const long long it = it0 + d;
                if (live[d]) {
                    any = true;
                    const int s = (int)(it % S);
                    const uint32_t ph = (uint32_t)(it / S) & 1;
#ifdef APSB_TC_TRACE
                    const long long c0_ = clock64();
#endif
                    if (lane == 0) tc_mbar_wait_parked(empty + s, ph ^ 1);   // one poller per warp
                    __syncwarp();
                    if (pt == 0) TC_TR(8);
#ifdef APSB_TC_TRACE
                    const long long c1_ = clock64();
#endif
                    // 32-bit shared-window addresses and st.shared: through the generic pointer the compiler emitted generic
                    // ST.E with 64-bit address arithmetic per store.  Rows rg + RSTEP * i share the swizzle term
                    // (RSTEP % 8 == 0), so the row offsets are immediates.
                    const uint32_t sa = prod_base + (uint32_t)s * (uint32_t)C::STAGE_BYTES;
#ifdef APSB_TC_TRACE
                    if (!(p.dbg & 1))
#endif
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        const float4 v = ring[d][i];
                        // hi = rn_tf32(v) as two integer ops: add half an ulp of the 10-bit mantissa, clear the low 13 bits
                        // (bit-identical to cvt.rna.tf32.f32 for finite inputs; sm_100a emulates that cvt with four
                        // instructions).  lo = v - hi is exact; the tensor core drops its low mantissa bits itself.
                        const uint32_t hx = (__float_as_uint(v.x) + 0x1000u) & 0xffffe000u;
                        const uint32_t hy = (__float_as_uint(v.y) + 0x1000u) & 0xffffe000u;
                        const uint32_t hz = (__float_as_uint(v.z) + 0x1000u) & 0xffffe000u;
                        const uint32_t hw = (__float_as_uint(v.w) + 0x1000u) & 0xffffe000u;
                        const float lx = v.x - __uint_as_float(hx), ly = v.y - __uint_as_float(hy);
                        const float lz = v.z - __uint_as_float(hz), lw = v.w - __uint_as_float(hw);
                        const uint32_t dst = sa + (uint32_t)(i * RSTEP * SWZ);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hx), "r"(hy), "r"(hz), "r"(hw)
                                     : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)A_BYTES), "f"(lx), "f"(ly),
                                     "f"(lz), "f"(lw)
                                     : "memory");
                    }
#ifdef APSB_TC_TRACE
                    if (!(p.dbg & 4))
#endif
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> UMMA reads
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(full_a + s);      // one arrival per producer warp
                    if (pt == 0) TC_TR(7);
#ifdef APSB_TC_TRACE
                    const long long c2_ = clock64();
                    acc_wait += c1_ - c0_;
                    acc_store += c2_ - c1_;
                    if (p.dbg & 8) { live[d] = seek(); if (live[d]) step(); } else
#endif
                    live[d] = gather_next(ring[d]);
#ifdef APSB_TC_TRACE
                    acc_gather += clock64() - c2_;
#endif
                }
            }
            if (!any) break;
        }
#ifdef APSB_TC_TRACE
        if (pt == 0 && p.trace && blockIdx.x == 0) {
            tr_smem[1020] = acc_wait; tr_smem[1021] = acc_store; tr_smem[1022] = acc_gather;
            tr_smem[1017] = acc_seek; tr_smem[1018] = acc_tile; tr_smem[1019] = acc_ld;
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) tc_cluster_sync();          // no CTA leaves while a peer may still multicast into it or signal its barriers
#ifdef APSB_TC_TRACE
    if (p.trace && blockIdx.x == 0)
        for (int i = threadIdx.x; i < TC_TRACE_WORDS; i += blockDim.x) p.trace[i] = tr_smem[i];
#endif
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS)
                     : "memory");
    }
}
#undef cid
#undef ncl

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
    int box_rows, box_cols;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
               box_cols == o.box_cols;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h = h * 1000003u ^ std::hash<long long>()(k.rows * 131 + k.cols);
        h = h * 1000003u ^ std::hash<long long>()(k.ld * 7 + k.box_rows + 1024 * k.box_cols);
        return h;
    }
};

static int make_map(CUtensorMap* map, const float* ptr, long long rows, long long cols, long long ld, int box_rows,
                    int box_cols) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, rows, cols, ld, box_rows, box_cols};
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
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    APSB_CHECK_ARG(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)rc);
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *map;
    return 0;
}

// MODE 4: geometry of a convolution whose A tiles are strided TMA boxes (see the kernel's TMA role)
struct Conv4 {
    int rh, tpi, tiles_m, cw, tpr;
    int t, ct[4];       // transposed convolution: row blocks per image and row class
};

// 4-D view [B][H][W][C] of the NHWC input; box = {box_c channels, OW pixels every sw, rh rows every sh, one image}
// (cuTensorMapEncodeTiled: "to load N elements along the i-th dimension, boxDim[i] must be N * elementStrides[i]")
static int make_map_conv(CUtensorMap* map, const float* ptr, long long B, long long H, long long W, long long Cc, int box_c,
                         int OW, int rh, int sh, int sw) {   // OW: output columns per box
    static std::mutex mu;
    struct Key4 { const void* ptr; long long B, H, W, C; int box_c, OW, rh, sh, sw; CUtensorMap map; };
    static std::vector<Key4> cache;
    {
        std::lock_guard<std::mutex> g(mu);
        for (const Key4& k : cache)
            if (k.ptr == ptr && k.B == B && k.H == H && k.W == W && k.C == Cc && k.box_c == box_c && k.OW == OW &&
                k.rh == rh && k.sh == sh && k.sw == sw) {
                *map = k.map;
                return 0;
            }
    }
    EncodeTiledFn fn = encode_fn();
    APSB_CHECK_ARG(fn, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[4] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Cc * 4, (cuuint64_t)(W * Cc) * 4, (cuuint64_t)(H * W * Cc) * 4};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(OW * sw), (cuuint32_t)(rh * sh), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    APSB_CHECK_ARG(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled (strided 4-D) failed with code %d", (int)rc);
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 256) cache.clear();
    cache.push_back(Key4{ptr, B, H, W, Cc, box_c, OW, rh, sh, sw, *map});
    return 0;
}

// launch with a cluster of CL CTAs (CL = 1: plain launch), optionally as a programmatic dependent launch.  The grid is
// `nclusters` x CL with nclusters bounded by what the device can keep resident at once (cudaOccupancyMaxActiveClusters:
// a cluster must sit inside one GPC, and 148 SMs are not a multiple of every GPC's width).
template <int BN, int MODE, int CL>
static int launch_tc_cl(const CUtensorMap& tB, const CUtensorMap& tBl, const CUtensorMap& tA, const CUtensorMap& tAl,
                        TcParams p, cudaStream_t st) {
    using C = TcCfg<BN, MODE>;
    static bool attr_done[64] = {false};       // function attributes are per device (one context per GPU)
    static int max_clusters[64] = {0};
    int dev = 0;
    APSB_CUDA(cudaGetDevice(&dev));
    auto kern = tc_gemm_kernel<BN, MODE, CL>;
    if (!attr_done[dev & 63]) {
        APSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        // smallest shared-memory carve-out that holds the CTA: what is left of the 228 KB array stays L1 for the gathers
        APSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (C::SMEM + 1024) * 100 / (228 * 1024) + 1));
        int mc = num_sms() / CL;
        if (CL > 1) {
            cudaLaunchConfig_t q{};
            q.gridDim = dim3((unsigned)(num_sms() / CL * CL));
            q.blockDim = dim3(TcRoles<BN, MODE>::THREADS);
            q.dynamicSmemBytes = C::SMEM;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.attrs = qa; q.numAttrs = 1;
            int n = 0;
            APSB_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &q));
            APSB_CHECK_ARG(n >= 1, "no cluster of %d CTAs of the tensor-core GEMM fits on this device", CL);
            mc = n < mc ? n : mc;
        }
        max_clusters[dev & 63] = mc;
        attr_done[dev & 63] = true;
    }
    const long long tiles_m = MODE == 4 ? (long long)p.c4_tiles_m : (p.M + TC_BM - 1) / TC_BM;
    const long long stiles = (tiles_m + CL - 1) / CL * p.tiles_n * p.ksplit;
    p.stiles = (unsigned)stiles;
    const long long ncl = stiles < max_clusters[dev & 63] ? stiles : max_clusters[dev & 63];
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CL > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(ncl * CL));
    cfg.blockDim = dim3(TcRoles<BN, MODE>::THREADS);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = na;
    APSB_CUDA(cudaLaunchKernelEx(&cfg, kern, tB, tBl, tA, tAl, p));
    APSB_LAUNCH_CHECK();
    return 0;
}

// cluster size of a launch: weight-tile multicast pays when several row blocks share a column block (row classes of a
// strided transposed convolution break the lock step: MODE 2 stays at 1).  APS_B200_TC_CL = 1 | 2 | 4 overrides.
template <int BN, int MODE>
static int launch_tc_mode(const CUtensorMap& tB, const CUtensorMap& tBl, const CUtensorMap& tA, const CUtensorMap& tAl,
                          const TcParams& p, int cl, cudaStream_t st) {
    if constexpr (MODE != 2) {
        if (cl == 4) return launch_tc_cl<BN, MODE, 4>(tB, tBl, tA, tAl, p, st);
        if (cl == 2) return launch_tc_cl<BN, MODE, 2>(tB, tBl, tA, tAl, p, st);
    }
    return launch_tc_cl<BN, MODE, 1>(tB, tBl, tA, tAl, p, st);
}

// `xlo` != nullptr selects MODE 3 (A tiles by TMA from x / xlo); `ksplit` > 1 cuts K into slices (MODE 3 only)
template <int BN>
static int launch_tc(const AGather& a, const float* W, const float* Wlo, long long ldw, int M, int N, int K,
                     const Epilogue& e, cudaStream_t st, const float* xlo = nullptr, int ksplit = 1,
                     long long split_stride = 0, const Conv4* c4 = nullptr, long long batch = 0, int grp_rows = 0,
                     int groups = 1) {
    using C = TcCfg<BN>;
    // cluster size (weight-tile multicast)
    const long long tiles_m_ = (M + TC_BM - 1) / TC_BM;
    int cl = (a.mode != 2 && tiles_m_ >= 8) ? APSB_TC_CLUSTER : 1;
    if (const char* ce = getenv("APS_B200_TC_CL")) {
        const int v = atoi(ce);
        if ((v == 1 || v == 2 || v == 4) && a.mode != 2) cl = v;
    }
    if (c4 || grp_rows) cl = 1;
    CUtensorMap tB, tBl, tA, tAl;
    const long long wrows = grp_rows ? (long long)N * groups : N;                 // MODE 5: the groups' weights stacked
    if (int rc = make_map(&tB, W, wrows, K, ldw, BN / cl, C::BK)) return rc;      // a CTA fetches BN / cl rows of the box
    if (int rc = make_map(&tBl, Wlo, wrows, K, ldw, BN / cl, C::BK)) return rc;
    tA = tB; tAl = tBl;
    if (xlo) {
        if (int rc = make_map(&tA, a.x, M, K, a.ld, TC_BM, C::BK)) return rc;
        if (int rc = make_map(&tAl, xlo, M, K, a.ld, TC_BM, C::BK)) return rc;
    }
    TcParams p{};
    p.M = M; p.N = N; p.K = K;
    p.tiles_n = (N + BN - 1) / BN;
    p.ksplit = xlo ? (ksplit < 1 ? 1 : ksplit) : 1;
    p.split_stride = split_stride;
    const long long tiles = (long long)((M + TC_BM - 1) / TC_BM) * p.tiles_n * p.ksplit;
    APSB_CHECK_ARG(tiles < (1LL << 31) - 1024, "too many output tiles (%lld)", tiles);
    p.tiles = (unsigned)tiles;
    p.a = a; p.e = e;
    {
        const bool glu = e.act == ACT_GLU;
        bool v = ((uintptr_t)e.out & 15) == 0 && (e.ldo & 3) == 0 && (N % (glu ? 8 : 4)) == 0;
        if (e.res) v = v && ((uintptr_t)e.res & 15) == 0 && (e.ldres & 3) == 0;
        if (e.bias) v = v && ((uintptr_t)e.bias & 15) == 0;
        if (e.post_scale) v = v && ((uintptr_t)e.post_scale & 15) == 0 && ((uintptr_t)e.post_shift & 15) == 0;
        if (e.act == ACT_PRELU && e.slope_stride) v = v && ((uintptr_t)e.slope & 15) == 0;
        p.epi_vec = v ? 1 : 0;
        APSB_CHECK_ARG(!e.out_lo || v, "a lo companion needs the vector epilogue (16-byte aligned bias / residual / output rows)");
    }
#ifdef APSB_TC_TRACE
    p.trace = g_tc_trace;
    p.dbg = getenv("APS_B200_TC_DBG") ? atoi(getenv("APS_B200_TC_DBG")) : 0;
    p.e.dbg_nobias = (p.dbg & 128) ? 1 : 0;
#endif
    APSB_CHECK_ARG(!(xlo && p.ksplit > 1 && !p.epi_vec), "split-K needs 16-byte aligned partial rows (N %% 4 == 0)");
    // the TMA-fed kernels have the vector epilogue only: unaligned outputs take the gather-fed kernel and its scalar path
    if (c4 && p.epi_vec) {
        if (c4->t) {        // transposed (stride_w = 1): plain boxes of rh consecutive input rows; the cat-skip tensor has its own map
            const long long csrc = a.cat_c ? 2LL * a.cat_c : a.Cin;
            if (int rc = make_map_conv(&tA, a.x, batch, a.H, a.W, csrc, C::BK, c4->cw, c4->rh, 1, 1)) return rc;
            tAl = tA;
            if (a.x2)
                if (int rc = make_map_conv(&tAl, a.x2, batch, a.H, a.W, csrc, C::BK, c4->cw, c4->rh, 1, 1)) return rc;
        } else {
            if (int rc = make_map_conv(&tA, a.x, batch, a.H, a.W, a.Cin, C::BK, c4->cw, c4->rh, a.sh, a.sw)) return rc;
        }
        p.c4_rh = c4->rh; p.c4_tpi = c4->tpi; p.c4_tiles_m = c4->tiles_m; p.c4_cw = c4->cw; p.c4_tpr = c4->tpr;
        p.c4_t = c4->t;
        for (int i = 0; i < 4; ++i) p.c4_ct[i] = c4->ct[i];
        p.tiles = (unsigned)((long long)c4->tiles_m * p.tiles_n);
        return launch_tc_cl<BN, 4, 1>(tB, tBl, tA, tAl, p, st);
    }
    if (grp_rows) {
        APSB_CHECK_ARG(xlo && p.epi_vec && grp_rows % TC_BM == 0, "grouped GEMM: needs the lo companion, aligned rows and 128-row groups");
        p.grp_rows = grp_rows; p.grp_n = N;
        return launch_tc_cl<BN, 5, 1>(tB, tBl, tA, tAl, p, st);
    }
    if (xlo && p.epi_vec) return launch_tc_mode<BN, 3>(tB, tBl, tA, tAl, p, cl, st);
    if (a.mode == 0) return launch_tc_mode<BN, 0>(tB, tBl, tA, tAl, p, cl, st);
    if (a.mode == 1) return launch_tc_mode<BN, 1>(tB, tBl, tA, tAl, p, cl, st);
    return launch_tc_mode<BN, 2>(tB, tBl, tA, tAl, p, 1, st);
}

// Tile width from a small cost model fitted to B200 measurements (profiles/r01_tc_gemm_v2_microbench.txt, cycles):
// a k-step of 1 costs ~87 (BN 256), ~64 (BN 128), ~53 (BN 64) cycles per tile in the main loop (+ ~5000 per tile for the
// pipeline refill), the epilogue ~4000 cycles per 32 columns and overlaps the next tile's main loop; tiles run in
// waves of one per SM.
static int run_tc(const AGather& a, const float* W, const float* Wlo, long long ldw, long long M, long long N,
                  long long K, const Epilogue& e, cudaStream_t st, const float* xlo = nullptr, int ksplit = 1,
                  long long split_stride = 0, const Conv4* c4 = nullptr, long long batch = 0, int grp_rows = 0,
                  int groups = 1) {
    if (c4) {       // TMA-fed convolution: MMA bound at every width, so the widest tile that is not mostly padding
        if (N > 128) return launch_tc<256>(a, W, Wlo, ldw, (int)M, (int)N, (int)K, e, st, nullptr, 1, 0, c4, batch);
        if (N > 64) return launch_tc<128>(a, W, Wlo, ldw, (int)M, (int)N, (int)K, e, st, nullptr, 1, 0, c4, batch);
        return launch_tc<64>(a, W, Wlo, ldw, (int)M, (int)N, (int)K, e, st, nullptr, 1, 0, c4, batch);
    }
    const long long tm = (M + TC_BM - 1) / TC_BM;
    const long long K0 = K;
    const int sms = num_sms();
    const char* fe = getenv("APS_B200_TC_BN");                    // tuning / test aid: force the tile width
    int bn = fe ? atoi(fe) : 0;
    if (bn != 64 && bn != 128 && bn != 256) {
        const int cand[3] = {256, 128, 64};
        // cycles per unit of k in the main loop: gather-fed tiles are bound by the producer warps (~the same time per
        // k-block whatever the width), TMA-fed tiles (MODE 3) by the MMAs and the shared-memory operand traffic
        // TMA-fed tiles (MODE 3): fitted to the round-2 trace / microbenchmark (profiles/r02l_tc_mode3_microbench.txt):
        // the main loop is bound by shared-memory operand traffic (~1170 / 1250 / 1700 cycles per 32 k at 64 / 128 / 256
        // columns), the pipeline runs on across tiles (one fill per launch), eight epilogue warps take ~1400 cycles per
        // 32 columns and overlap the next tile
        const double per_k_gather[3] = {87.0, 64.0, 53.0}, per_k_tma[3] = {53.0, 39.0, 37.0};
        if (ksplit > 1) K = (K + ksplit - 1) / ksplit;
        double best = 0.0;
        for (int i = 0; i < 3; ++i) {
            if (cand[i] > 64 && N <= cand[i] / 2) continue;       // more than half of the tile would be padding
            const long long tiles = tm * ((N + cand[i] - 1) / cand[i]) * (ksplit > 1 ? ksplit : 1);
            const double waves = (double)((tiles + sms - 1) / sms);
            double t;
            if (xlo) {
                const double kc = (double)K * per_k_tma[i], epi_c = 1400.0 * (cand[i] / 32);
                t = 4000.0 + (waves - 1.0) * (kc > epi_c ? kc : epi_c) + kc + epi_c;
            } else {
                const double main_c = (double)K * per_k_gather[i] + 5000.0, epi_c = 4000.0 * (cand[i] / 32);
                t = main_c + (waves - 1.0) * (main_c > epi_c ? main_c : epi_c) + epi_c;
            }
            if (bn == 0 || t < best) { best = t; bn = cand[i]; }
        }
    }
    if (bn == 256) return launch_tc<256>(a, W, Wlo, ldw, (int)M, (int)N, (int)K0, e, st, xlo, ksplit, split_stride, nullptr, 0, grp_rows, groups);
    if (bn == 128) return launch_tc<128>(a, W, Wlo, ldw, (int)M, (int)N, (int)K0, e, st, xlo, ksplit, split_stride, nullptr, 0, grp_rows, groups);
    return launch_tc<64>(a, W, Wlo, ldw, (int)M, (int)N, (int)K0, e, st, xlo, ksplit, split_stride, nullptr, 0, grp_rows, groups);
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

#ifdef APSB_TC_TRACE
extern "C" int aps_b200_tc_trace(unsigned long long* device_buffer, int capacity_words) {
    g_tc_trace = device_buffer;
    g_tc_trace_cap = capacity_words;
    return 0;
}
#endif

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

// Linear layer with the optional extras of the encoder stack: `x_lo` (companion of x: the kernel then loads both
// operand sides by TMA and runs without gather warps), `out_lo` (write the companion of the result for the next layer),
// `ksplit` > 1 (needs x_lo; slice s of K accumulates into out + s * split_stride as RAW partial sums — the epilogue must
// then be empty and aps_b200_layernorm2_fwd / the caller reduces the slices).
extern "C" int aps_b200_linear_tc2_fwd(const float* x, const float* x_lo, int64_t rows, int64_t in_features, int64_t ld_x,
                                       const float* weight_hi, const float* weight_lo, int64_t ld_w,
                                       int64_t out_features, const aps_b200_epilogue* epi, float* out, float* out_lo,
                                       int64_t ld_out, int32_t ksplit, int64_t split_stride, void* stream) {
    APSB_CHECK_ARG(x, "null pointer argument");
    APSB_CHECK_ARG(rows > 0 && in_features > 0 && out_features > 0 && ld_x >= in_features, "bad shape");
    APSB_CHECK_ARG(rows < (1LL << 31) && out_features < (1LL << 31) && in_features < (1LL << 31), "shape too large");
    APSB_CHECK_ARG((ld_x & 3) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)x_lo & 15) == 0,
                   "the tensor-core path needs 16-byte aligned activation rows");
    if (int rc = check_weights(weight_hi, weight_lo, ld_w, in_features)) return rc;
    Epilogue e{};
    if (int rc = fill_tc_epilogue(e, epi, out_features, out, ld_out)) return rc;
    e.out_lo = out_lo;
    if (out_lo)
        APSB_CHECK_ARG(((uintptr_t)out_lo & 15) == 0 && ((uintptr_t)out & 15) == 0 && (ld_out & 3) == 0 &&
                           (out_features & 3) == 0 && epi->act != ACT_GLU,
                       "a lo companion needs 16-byte aligned output rows and no GLU");
    if (ksplit > 1) {
        APSB_CHECK_ARG(x_lo, "split-K needs the lo companion of x (TMA-fed mode)");
        APSB_CHECK_ARG(!epi->bias && epi->act == ACT_NONE && !epi->residual && !epi->post_scale && epi->alpha == 1.f && !out_lo,
                       "split-K writes raw partial sums: the epilogue must be empty");
        APSB_CHECK_ARG(split_stride >= rows * ld_out, "split_stride too small");
        APSB_CHECK_ARG(ksplit <= (in_features + 31) / 32, "ksplit %d: every slice needs at least one 32-wide k group", ksplit);
    }
    AGather a{};
    a.mode = 0; a.x = x; a.ld = ld_x;
    return run_tc(a, weight_hi, weight_lo, ld_w, rows, out_features, in_features, e, (cudaStream_t)stream, x_lo,
                  ksplit < 1 ? 1 : ksplit, split_stride);
}

static int conv_geometry(AGather& a, const float* x, int64_t batch, int64_t height, int64_t width,
                         int64_t in_channels, int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h,
                         int pad_w) {
    APSB_CHECK_ARG(x, "null pointer argument");
    APSB_CHECK_ARG(batch > 0 && height > 0 && width > 0 && in_channels > 0, "bad shape");
    APSB_CHECK_ARG(kernel_h > 0 && kernel_w > 0 && stride_h > 0 && stride_w > 0 && pad_h >= 0 && pad_w >= 0,
                   "bad convolution geometry");
    APSB_CHECK_ARG(in_channels % 32 == 0 && ((uintptr_t)x & 15) == 0,
                   "the tensor-core convolution needs Cin %% 32 == 0 and a 16-byte aligned input (Cin = %lld)",
                   (long long)in_channels);
    APSB_CHECK_ARG(height < (1 << 20) && width < (1 << 20) && in_channels < (1 << 20), "shape too large");
    a.x = x; a.H = (int)height; a.W = (int)width; a.Cin = (int)in_channels; a.KH = kernel_h; a.KW = kernel_w;
    a.sh = stride_h; a.sw = stride_w; a.ph = pad_h; a.pw = pad_w;
    return 0;
}

static int conv2d_tc_impl(const float* x, int64_t batch, int64_t height, int64_t width,
                          int64_t in_channels, const float* weight_hi, const float* weight_lo,
                          int64_t out_channels, int kernel_h, int kernel_w, int stride_h,
                          int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                          const aps_b200_epilogue* epi, float* out, float* out_lo, void* stream) {
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
    e.out_lo = out_lo;
    if (out_lo)
        APSB_CHECK_ARG(((uintptr_t)out_lo & 15) == 0 && ((uintptr_t)out & 15) == 0 && (out_channels & 3) == 0 &&
                           epi->act != ACT_GLU, "a lo companion needs 16-byte aligned output rows and no GLU");
    // A tiles by TMA (MODE 4) when the strides are within the TMA's traversal-stride range: a 128-row tile is either some
    // whole output rows (short rows: the conformer front) or a chunk of one output row (long rows: DCCRN's time axis)
    Conv4 c4{};
    const bool tma_conv = stride_h <= 8 && stride_w <= 8 && (in_channels & 31) == 0 && ((uintptr_t)x & 15) == 0 &&
                          !getenv("APS_B200_NO_CONV_TMA");
    if (tma_conv) {
        const int64_t cw_max = 256 / stride_w < TC_BM ? 256 / stride_w : TC_BM;     // box extent <= 256 elements
        if (OW <= cw_max) {
            c4.cw = (int)OW; c4.tpr = 1;
            c4.rh = (int)(TC_BM / OW);
            if ((long long)c4.rh > OH) c4.rh = (int)OH;
            if (c4.rh * stride_h > 256) c4.rh = 256 / stride_h;
            c4.tpi = (int)((OH + c4.rh - 1) / c4.rh);
        } else {
            c4.tpr = (int)((OW + cw_max - 1) / cw_max);
            c4.cw = (int)((OW + c4.tpr - 1) / c4.tpr);        // even chunks
            c4.rh = 1;
            c4.tpi = (int)(OH * c4.tpr);
        }
        c4.tiles_m = (int)(batch * c4.tpi);
        // rows actually used over rows paid for (tile padding + the short last tile of every image / row)
        const double eff = (double)(OH * OW) / ((double)c4.tpi * TC_BM);
        if (eff >= 0.7 && batch * c4.tpi < (1LL << 30))
            return run_tc(a, weight_hi, weight_lo, K, M, out_channels, K, e, (cudaStream_t)stream, nullptr, 1, 0, &c4, batch);
    }
    return run_tc(a, weight_hi, weight_lo, K, M, out_channels, K, e, (cudaStream_t)stream);
}

extern "C" int aps_b200_conv2d_nhwc_tc_fwd(const float* x, int64_t batch, int64_t height, int64_t width,
                                           int64_t in_channels, const float* weight_hi, const float* weight_lo,
                                           int64_t out_channels, int kernel_h, int kernel_w, int stride_h,
                                           int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                           const aps_b200_epilogue* epi, float* out, void* stream) {
    return conv2d_tc_impl(x, batch, height, width, in_channels, weight_hi, weight_lo, out_channels, kernel_h, kernel_w,
                          stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, epi, out, nullptr, stream);
}

/* as above, and also writes the TF32 "lo" companion of the output (same shape) for a TMA-fed linear consumer */
extern "C" int aps_b200_conv2d_nhwc_tc2_fwd(const float* x, int64_t batch, int64_t height, int64_t width,
                                            int64_t in_channels, const float* weight_hi, const float* weight_lo,
                                            int64_t out_channels, int kernel_h, int kernel_w, int stride_h,
                                            int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                            const aps_b200_epilogue* epi, float* out, float* out_lo, void* stream) {
    return conv2d_tc_impl(x, batch, height, width, in_channels, weight_hi, weight_lo, out_channels, kernel_h, kernel_w,
                          stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, epi, out, out_lo, stream);
}

extern "C" int aps_b200_conv_transpose2d_nhwc_tc_fwd(const float* x, const float* x_skip, int64_t batch, int64_t height,
                                                     int64_t width, int64_t in_channels, const float* weight_hi,
                                                     const float* weight_lo, int64_t out_channels, int kernel_h,
                                                     int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                                                     int out_pad_h, int out_pad_w, const aps_b200_epilogue* epi,
                                                     float* out, void* stream) {
    AGather a{};
    if (int rc = conv_geometry(a, x, batch, height, width, in_channels, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                               pad_w))
        return rc;
    APSB_CHECK_ARG(out_pad_h >= 0 && out_pad_w >= 0 && out_channels > 0, "bad convolution geometry");
    APSB_CHECK_ARG(stride_w == 1, "the tensor-core transposed convolution needs stride_w == 1 (got %d)", stride_w);
    a.mode = 2; a.dh = 1; a.dw = 1;
    if (x_skip) {
        APSB_CHECK_ARG(in_channels % 128 == 0 && ((uintptr_t)x_skip & 15) == 0,
                       "a concatenated input needs four parts of a multiple of 32 channels (in_channels = %lld)",
                       (long long)in_channels);
        a.x2 = x_skip;
        a.cat_c = (int)(in_channels / 4);
    }
    const int64_t OH = (height - 1) * stride_h - 2 * pad_h + kernel_h + out_pad_h;
    const int64_t OW = (width - 1) * stride_w - 2 * pad_w + kernel_w + out_pad_w;
    APSB_CHECK_ARG(OH > 0 && OW > 0, "transposed convolution output is empty");
    a.OH = (int)OH; a.OW = (int)OW;
    // row classes (oh % stride_h): usable when every class owns at least one tap, else plain row order
    a.classes = 1;
    if (stride_h >= 2 && stride_h <= 4 && !getenv("APS_B200_TC_NO_CLASSES")) {
        bool ok = true;
        for (int r = 0; r < stride_h && ok; ++r) {
            bool any = false;
            for (int kh = 0; kh < kernel_h; ++kh) any = any || ((r + pad_h - kh) % stride_h) == 0;
            ok = any;
        }
        if (ok) {
            a.classes = stride_h;
            long long start = 0;
            for (int r = 0; r < stride_h; ++r) {
                a.class_start[r] = start;
                a.class_rows[r] = OH > r ? (int)((OH - r + stride_h - 1) / stride_h) : 0;
                start += batch * a.class_rows[r] * OW;
            }
            for (int r = stride_h; r < 4; ++r) { a.class_start[r] = start; a.class_rows[r] = 0; }
        }
    }
    const int64_t M = batch * OH * OW, K = (int64_t)kernel_h * kernel_w * in_channels;
    APSB_CHECK_ARG(M < (1LL << 31) && K < (1LL << 31) && out_channels < (1LL << 31), "shape too large");
    if (int rc = check_weights(weight_hi, weight_lo, K, K)) return rc;
    APSB_CHECK_ARG(epi && epi->act != ACT_GLU, "GLU is not available here");
    Epilogue e{};
    if (int rc = fill_tc_epilogue(e, epi, out_channels, out, out_channels)) return rc;
    // A tiles by TMA (MODE 4, transposed form): a 128-row tile is some output rows of ONE row class (their input rows
    // are consecutive) or a column chunk of one output row (DCCRN's time axis); every class needs a tap (a.classes)
    if ((stride_h == 1 || a.classes == stride_h) && (in_channels & 31) == 0 && ((uintptr_t)x & 15) == 0 &&
        !getenv("APS_B200_NO_CONV_TMA")) {
        Conv4 c4{};
        c4.t = 1;
        const int ncls = stride_h;
        if (stride_h == 1) { a.class_rows[0] = (int)OH; a.class_rows[1] = a.class_rows[2] = a.class_rows[3] = 0; }
        int max_rows = 0;
        for (int c = 0; c < ncls; ++c) max_rows = a.class_rows[c] > max_rows ? a.class_rows[c] : max_rows;
        if (OW <= TC_BM) {
            c4.cw = (int)OW; c4.tpr = 1;
            c4.rh = (int)(TC_BM / OW);
            if (c4.rh > max_rows) c4.rh = max_rows;
        } else {
            c4.tpr = (int)((OW + TC_BM - 1) / TC_BM);
            c4.cw = (int)((OW + c4.tpr - 1) / c4.tpr);
            c4.rh = 1;
        }
        c4.tpi = 0;
        for (int c = 0; c < 4; ++c) {
            c4.ct[c] = c < ncls ? ((a.class_rows[c] + c4.rh - 1) / c4.rh) * c4.tpr : 0;
            c4.tpi += c4.ct[c];
        }
        c4.tiles_m = (int)(batch * c4.tpi);
        const double eff = (double)(OH * OW) / ((double)c4.tpi * TC_BM);
        if (c4.rh >= 1 && eff >= 0.7 && batch * c4.tpi < (1LL << 30))
            return run_tc(a, weight_hi, weight_lo, K, M, out_channels, K, e, (cudaStream_t)stream, nullptr, 1, 0, &c4, batch);
    }
    return run_tc(a, weight_hi, weight_lo, K, M, out_channels, K, e, (cudaStream_t)stream);
}

// ---- LSTM recurrence on the tensor-core engine ---------------------------------------------------------------------
namespace apsb {
// gates -> cell / hidden state of one frame for all groups (rows of the stacked [groups * rows_pad, .] buffers)
struct LstmCellParams {
    const float* pre;     // [nparts][R, 4H] K-slice partial sums of h W_hh^T (i | f | g | o blocks of H), R = groups * rows_pad
    const float* xg;      // [R, frames, 4H]: this frame's input projections (pointer already at the frame)
    long long part_stride, ld_xg;
    int nparts;
    float* c;             // [R, H] cell state, in place
    float* h;             // [R, H] next hidden state (the next frame's GEMM operand) ...
    float* h_lo;          // ... and its TF32 lo companion
    float* y[APS_B200_LSTM_MAX_GROUPS];   // per group: output rows of this frame (row stride ld_y)
    long long ld_y;
    int rows, rows_pad, H;
    long long total;      // R * H
};

__global__ void __launch_bounds__(256) lstm_cell_kernel(const __grid_constant__ LstmCellParams p) {
    // ONE hidden unit per thread: the five precise transcendentals of a unit are a ~400-instruction dependent chain, and
    // the whole frame is less than one wave of threads, so the kernel's time IS that chain (four units per thread with
    // 16-byte accesses measured 8.7 us per frame under ncu — as long as the recurrent GEMM it follows)
    pdl_trigger();
    pdl_wait();
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= p.total) return;
    const int r = (int)(idx / p.H), u = (int)(idx - (long long)r * p.H);
    const int g = r / p.rows_pad, row = r - g * p.rows_pad;
    if (row >= p.rows) return;                        // padding rows of a group
    const float* xr = p.xg + (long long)r * p.ld_xg + u;
    float gi = __ldg(xr), gf = __ldg(xr + p.H), gg = __ldg(xr + 2 * p.H), go = __ldg(xr + 3 * p.H);
    for (int s = 0; s < p.nparts; ++s) {              // fixed order: deterministic
        const float* pr = p.pre + s * p.part_stride + (long long)r * 4 * p.H + u;
        gi += __ldg(pr); gf += __ldg(pr + p.H); gg += __ldg(pr + 2 * p.H); go += __ldg(pr + 3 * p.H);
    }
    auto sg = [](float x) { return 1.f / (1.f + expf(-x)); };     // precise: the recurrence amplifies rounding
    const long long o = (long long)r * p.H + u;
    const float c = sg(gf) * p.c[o] + sg(gi) * tanhf(gg);
    const float h = sg(go) * tanhf(c);
    p.c[o] = c;
    p.h[o] = h;
    p.h_lo[o] = tf32_lo(h);
    p.y[g][(long long)row * p.ld_y + u] = h;
}
}  // namespace apsb

/* LSTM recurrence (torch.nn.LSTM semantics, gate order i, f, g, o, zero initial state, forward direction) of `groups`
 * modules of one shape that advance together, with the per-frame product h_{t-1} W_hh^T on the tcgen05 engine (3xTF32):
 * per frame ONE grouped GEMM launch (the modules' hidden states stacked along M, each row block using its own
 * module's W_hh) whose epilogue adds the frame's input projections, and one cell kernel.  Replaces the fp32-FMA
 * recurrence of aps_b200_lstm_group_fwd (same reference lines) where hidden % 32 == 0.
 * xg: [groups, rows_pad, frames, 4 hidden] input projections incl. biases; w_hi / w_lo: [groups * 4 hidden, hidden];
 * y[g]: [rows, frames, >= hidden] with ld_y floats between frames; work: caller-provided scratch of
 * groups * rows_pad * hidden * 21 floats (cell, two (h, h_lo) pairs, up to four K slices of gate pre-activations). */
extern "C" int aps_b200_lstm_group_tc_fwd(const float* xg, int64_t rows, int64_t rows_pad, int64_t frames, int64_t hidden,
                                          const float* w_hi, const float* w_lo, void* const* y, int64_t ld_y,
                                          int32_t groups, float* work, void* stream) {
    APSB_CHECK_ARG(xg && w_hi && w_lo && y && work, "null pointer argument");
    APSB_CHECK_ARG(rows > 0 && frames > 0 && hidden > 0 && groups > 0 && groups <= APS_B200_LSTM_MAX_GROUPS, "bad shape");
    APSB_CHECK_ARG(rows_pad >= rows && rows_pad % TC_BM == 0, "rows_pad must be a multiple of %d", TC_BM);
    APSB_CHECK_ARG(hidden % 32 == 0, "the tensor-core recurrence needs hidden %% 32 == 0 (got %lld)", (long long)hidden);
    APSB_CHECK_ARG(((uintptr_t)xg & 15) == 0 && ((uintptr_t)work & 15) == 0 && (ld_y & 3) == 0, "16-byte alignment");
    APSB_CHECK_ARG(ld_y >= hidden, "ld_y %lld smaller than hidden %lld", (long long)ld_y, (long long)hidden);
    for (int g = 0; g < groups; ++g) APSB_CHECK_ARG(y[g] != nullptr, "null output pointer of group %d", g);
    cudaStream_t st = (cudaStream_t)stream;
    const long long R = (long long)groups * rows_pad, H = hidden;
    APSB_CHECK_ARG(R < (1LL << 31), "too many rows");
    float* cell = work;
    float* hbuf[2] = {cell + R * H, cell + 3 * R * H};      // each: h then h_lo
    float* pre = cell + 5 * R * H;
    APSB_CUDA(cudaMemsetAsync(work, 0, (size_t)(5 * R * H) * sizeof(float), st));     // c = h = 0 before the first frame
    // K slices: the per-frame GEMM is small (64 tiles of 128 x 128 at the DCCRN size) and strictly sequential over frames,
    // so it is cut along K while the work items still fit one wave; the cell kernel adds the slices
    int ks = 1;
    {
        const long long tiles128 = (R / TC_BM) * ((4 * H + 127) / 128);
        while (ks < 4 && tiles128 * ks * 2 <= num_sms() && H / (ks * 2) >= 128) ks *= 2;
    }
    LstmCellParams cp{};
    cp.pre = pre; cp.c = cell; cp.ld_y = frames * ld_y; cp.rows = (int)rows; cp.rows_pad = (int)rows_pad; cp.H = (int)H;
    cp.total = R * H; cp.nparts = ks; cp.part_stride = R * 4 * H; cp.ld_xg = frames * 4 * H;
    const unsigned cgrid = (unsigned)((cp.total + 255) / 256);
    for (int64_t t = 0; t < frames; ++t) {
        const float* hin = hbuf[t & 1];
        float* hout = hbuf[(t + 1) & 1];
        // pre = h_{t-1} W_hh^T + xg[:, :, t, :]
        Epilogue e{};
        e.act = ACT_NONE; e.alpha = 1.f; e.beta = 1.f; e.out = pre; e.ldo = 4 * H;
        AGather a{};
        a.mode = 0; a.x = hin; a.ld = H;
        // ks slices of RAW partial sums (ks = 1: the plain product); the input projections are added by the cell kernel
        if (int rc = run_tc(a, w_hi, w_lo, H, R, 4 * H, H, e, st, hin + R * H, ks, R * 4 * H, nullptr, 0, (int)rows_pad, groups)) return rc;
        cp.h = hout; cp.h_lo = hout + R * H; cp.xg = xg + t * 4 * H;
        for (int g = 0; g < groups; ++g) cp.y[g] = static_cast<float*>(y[g]) + t * ld_y;
        APSB_CUDA(launch_pdl(lstm_cell_kernel, dim3(cgrid), dim3(256), 0, st, cp));
    }
    APSB_LAUNCH_CHECK();
    return 0;
}

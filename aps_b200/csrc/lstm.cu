// lstm.cu — the recurrence of one LSTM layer / direction (torch.nn.LSTM semantics, gate order i, f, g, o),
// the bottleneck of DCCRN (aps/sse/bss/dccrn.py:20-50, LSTMP -> nn.LSTM) and of every other nn.LSTM user.
//
// The input projection x_t W_ih^T + b_ih + b_hh of ALL frames is one large GEMM done beforehand on the tensor-core
// engine (ops.linear); what is left per frame is gates = xg_t + h_{t-1} W_hh^T, the cell update and h_t — a
// [rows, H] x [H, 4H] product that is sequential over frames.  One launch per frame fuses the product with the cell
// update: a CTA owns 64 rows x 16 hidden units (all four gates of a unit sit in one thread, so the cell update needs
// no exchange), walks K = H in 32-wide chunks through a 4-stage cp.async ring and writes h_t straight into the
// layer output y[:, t, :], which is also where frame t + 1 reads h_{t-1} from.  Launches are chained with
// programmatic dependent launch: frame t + 1 prefetches W_hh and its input projections while frame t finishes.  Exact fp32 FMA arithmetic, precise
// expf / tanhf: the recurrence amplifies rounding, and parity is against the reference's fp32 CPU path.
//
// Bound by the fp32 FMA pipe: 2 * rows * 4H * H flop per frame (0.54 GFLOP at rows = 256, H = 512) against
// 148 SMs * 128 FMA/clk.  W_hh (4 MB at H = 512) and h_{t-1} stay L2 resident across frames.
#include "../../include/aps_b200.h"
#include "common.cuh"

namespace apsb {

constexpr int kLstmBM = 64;        // rows per CTA
constexpr int kLstmBU = 16;        // hidden units per CTA (x 4 gates = 64 gate columns)
constexpr int kLstmBK = 32;        // k chunk
constexpr int kLstmStages = 4;     // cp.async ring depth (3 chunks = 24 KB per CTA in flight: covers the L2 latency)
constexpr int kLstmThreads = 256;  // thread = 4 rows x 1 unit x 4 gates
constexpr int kLstmLd = kLstmBK + 4;                                   // padded row: conflict-free float4 reads
constexpr int kLstmTile = (kLstmBM + 4 * kLstmBU) * kLstmLd;           // floats per stage: h rows, then W_hh rows
constexpr int kLstmSmem = kLstmStages * kLstmTile * 4;

struct LstmStepParams {
    // up to APS_B200_LSTM_MAX_GROUPS independent recurrences of the same shape run side by side (grid.z): the two
    // LSTMs of DCCRN's complex bottleneck, or the two directions of a bidirectional layer.  One recurrence is
    // 128 CTAs at the DCCRN size — one per SM with a dependent-issue-bound inner loop; two of them co-resident
    // double the warps per scheduler
    const float* xg[APS_B200_LSTM_MAX_GROUPS];   // [rows, T, >= 4H] input projections (+ both biases)
    const float* w[APS_B200_LSTM_MAX_GROUPS];    // W_hh [4H, H]
    float* c[APS_B200_LSTM_MAX_GROUPS];          // [rows, H] cell state, updated in place
    float* y[APS_B200_LSTM_MAX_GROUPS];          // [rows, T, >= H] outputs; frame t - 1 (t + 1 reversed) is h_{t-1}
    long long ld_xg, ld_y;                       // floats between consecutive frames of a row
    int reverse_mask;                            // bit g: group g walks the frames backwards
    int rows, H, T, step;                        // step: 0 .. T - 1 (h = c = 0 before step 0)
};

__device__ __forceinline__ float sigmoid_precise(float x) { return 1.f / (1.f + expf(-x)); }

// 16-byte global -> shared copy, zero-filled when !valid
__device__ __forceinline__ void lstm_cp16(float* dst, const float* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kLstmThreads) lstm_step_kernel(const __grid_constant__ LstmStepParams p) {
    extern __shared__ __align__(16) float lstm_smem[];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int u0 = blockIdx.x * kLstmBU, r0 = blockIdx.y * kLstmBM, grp = blockIdx.z;
    const int H = p.H;
    const int unit = u0 + tx;
    const bool unit_ok = unit < H;
    const bool first = p.step == 0;
    const int nchunks = first ? 0 : (H + kLstmBK - 1) / kLstmBK;
    const bool rev = (p.reverse_mask >> grp) & 1;
    const int frame = rev ? p.T - 1 - p.step : p.step, frame_prev = rev ? frame + 1 : frame - 1;
    const long long ldx = (long long)p.T * p.ld_xg, ldy = (long long)p.T * p.ld_y;   // row strides
    const float* const g_xg = p.xg[grp] + (long long)frame * p.ld_xg;
    const float* const g_w = p.w[grp];
    float* const g_c = p.c[grp];
    float* const g_y = p.y[grp] + (long long)frame * p.ld_y;
    const float* const g_h = p.y[grp] + (long long)(first ? frame : frame_prev) * p.ld_y;

    // gate pre-activations from the input projection: independent of the previous frame, issued first, consumed last
    float xg[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = r0 + ty * 4 + r;
        const bool ok = unit_ok && row < p.rows;
#pragma unroll
        for (int g = 0; g < 4; ++g) xg[r][g] = ok ? __ldg(g_xg + (long long)row * ldx + (long long)g * H + unit) : 0.f;
    }

    // chunk -> ring stage: this thread copies two 16-byte pieces of the h tile and two of the W_hh tile (k-major rows,
    // exactly as they lie in global memory: no transposition, so the copies can be asynchronous)
    const int lrow = tid >> 3, lkq = tid & 7;                       // + 32 rows for the second piece
    auto issue_w = [&](int chunk) {
        float* st = lstm_smem + (chunk % kLstmStages) * kLstmTile + kLstmBM * kLstmLd;
        const int k = chunk * kLstmBK + 4 * lkq;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = lrow + 32 * j, g = col >> 4, u = u0 + (col & 15);
            const bool ok = u < H && k < H;
            lstm_cp16(st + col * kLstmLd + 4 * lkq, ok ? g_w + ((long long)g * H + u) * H + k : g_w, ok);
        }
    };
    auto issue_h = [&](int chunk) {
        float* st = lstm_smem + (chunk % kLstmStages) * kLstmTile;
        const int k = chunk * kLstmBK + 4 * lkq;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int row = lrow + 32 * j;
            const bool ok = r0 + row < p.rows && k < H;
            lstm_cp16(st + row * kLstmLd + 4 * lkq, ok ? g_h + (long long)(r0 + row) * ldy + k : g_w, ok);
        }
    };

    // W_hh does not depend on the previous frame: its first stages are requested before waiting for the previous
    // launch (programmatic dependent launch), which hides the launch latency and the first L2 round trip
#pragma unroll
    for (int st = 0; st < kLstmStages - 1; ++st)
        if (st < nchunks) issue_w(st);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#pragma unroll
    for (int st = 0; st < kLstmStages - 1; ++st) {
        if (st < nchunks) issue_h(st);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    float c_old[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = r0 + ty * 4 + r;
        c_old[r] = (unit_ok && row < p.rows && !first) ? g_c[(long long)row * H + unit] : 0.f;
    }

    // accumulators as packed pairs {sum over even k, sum over odd k}: the shared-memory float4s already hold
    // (k, k + 1) pairs in adjacent registers, so fma.rn.f32x2 (FFMA2) needs no packing and halves the issue slots of
    // the inner loop, which is what bounds it (512 FFMA + 64 LDS per chunk and warp otherwise)
    unsigned long long acc2[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int g = 0; g < 4; ++g) acc2[r][g] = 0ull;

    for (int chunk = 0; chunk < nchunks; ++chunk) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kLstmStages - 2) : "memory");
        __syncthreads();                      // chunk landed for everyone; everyone is done with chunk - 1's stage
        if (chunk + kLstmStages - 1 < nchunks) {
            issue_w(chunk + kLstmStages - 1);
            issue_h(chunk + kLstmStages - 1);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const float* As = lstm_smem + (chunk % kLstmStages) * kLstmTile + (ty * 4) * kLstmLd;
        const float* Ws = lstm_smem + (chunk % kLstmStages) * kLstmTile + (kLstmBM + tx) * kLstmLd;
#pragma unroll
        for (int kq = 0; kq < kLstmBK / 4; ++kq) {
            ulonglong2 a[4], w[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const ulonglong2*>(As + r * kLstmLd + 4 * kq);
#pragma unroll
            for (int g = 0; g < 4; ++g) w[g] = *reinterpret_cast<const ulonglong2*>(Ws + g * kLstmBU * kLstmLd + 4 * kq);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[r][g]) : "l"(a[r].x), "l"(w[g].x));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[r][g]) : "l"(a[r].y), "l"(w[g].y));
                }
        }
    }
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc2[r][g]));
            acc[r][g] = lo + hi;
        }

    if (!unit_ok) return;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = r0 + ty * 4 + r;
        if (row >= p.rows) continue;
        const float ig = sigmoid_precise(acc[r][0] + xg[r][0]);
        const float fg = sigmoid_precise(acc[r][1] + xg[r][1]);
        const float gg = tanhf(acc[r][2] + xg[r][2]);
        const float og = sigmoid_precise(acc[r][3] + xg[r][3]);
        const float c = fmaf(fg, c_old[r], ig * gg);
        g_c[(long long)row * H + unit] = c;
        g_y[(long long)row * ldy + unit] = og * tanhf(c);
    }
}

}  // namespace apsb

using namespace apsb;

extern "C" int aps_b200_lstm_group_fwd(const float* const* xg, int64_t ld_xg, int64_t rows, int64_t num_frames,
                                       int64_t hidden, const float* const* w_hh, int reverse_mask, float* const* cell,
                                       float* const* y, int64_t ld_y, int groups, void* stream) {
    APSB_CHECK_ARG(xg && w_hh && cell && y, "null pointer argument");
    APSB_CHECK_ARG(groups >= 1 && groups <= APS_B200_LSTM_MAX_GROUPS, "lstm: 1..%d groups (got %d)",
                   APS_B200_LSTM_MAX_GROUPS, groups);
    APSB_CHECK_ARG(rows > 0 && num_frames > 0 && hidden > 0 && rows < (1LL << 31) && num_frames < (1LL << 31) &&
                       hidden < (1LL << 29), "bad shape");
    APSB_CHECK_ARG(hidden % 4 == 0 && ld_y % 4 == 0 && ld_y >= hidden && ld_xg >= 4 * hidden,
                   "lstm: hidden (%lld) and ld_y (%lld) must be multiples of 4", (long long)hidden, (long long)ld_y);
    const int64_t grid_y = (rows + kLstmBM - 1) / kLstmBM;
    APSB_CHECK_ARG(grid_y <= 65535, "lstm: too many rows (%lld)", (long long)rows);
    LstmStepParams p{};
    for (int g = 0; g < groups; ++g) {
        APSB_CHECK_ARG(xg[g] && w_hh[g] && cell[g] && y[g], "null pointer argument (group %d)", g);
        APSB_CHECK_ARG(((uintptr_t)y[g] & 15) == 0 && ((uintptr_t)w_hh[g] & 15) == 0,
                       "lstm: y and w_hh must be 16-byte aligned (group %d)", g);
        p.xg[g] = xg[g]; p.w[g] = w_hh[g]; p.c[g] = cell[g]; p.y[g] = y[g];
    }
    p.ld_xg = ld_xg; p.ld_y = ld_y; p.reverse_mask = reverse_mask;
    p.rows = (int)rows; p.H = (int)hidden; p.T = (int)num_frames;
    static bool attr_set[64] = {};
    int dev = 0;
    APSB_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        APSB_CUDA(cudaFuncSetAttribute(lstm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLstmSmem));
        attr_set[dev] = true;
    }
    // every frame's launch may start (and prefetch its W_hh tiles and input projections) while the previous frame is
    // still running; the kernel waits for it (griddepcontrol.wait) before touching h_{t-1} or the cell state
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((hidden + kLstmBU - 1) / kLstmBU), (unsigned)grid_y, (unsigned)groups);
    cfg.blockDim = dim3(kLstmThreads);
    cfg.dynamicSmemBytes = kLstmSmem;
    cfg.stream = (cudaStream_t)stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (int64_t t = 0; t < num_frames; ++t) {
        p.step = (int)t;
        // The kernel prefetches its input projections BEFORE griddepcontrol.wait: safe against the previous STEP (which
        // does not write them), not against the GEMM that produces them — that kernel now triggers its dependents early
        // (pdl_trigger in tc_gemm.cu), so the first step is launched without the attribute and starts after the GEMM.
        cfg.numAttrs = t == 0 ? 0 : 1;
        APSB_CUDA(cudaLaunchKernelEx(&cfg, lstm_step_kernel, p));
    }
    return 0;
}

extern "C" int aps_b200_lstm_fwd(const float* xg, int64_t ld_xg, int64_t rows, int64_t num_frames, int64_t hidden,
                                 const float* w_hh, int reverse, float* cell, float* y, int64_t ld_y, void* stream) {
    return aps_b200_lstm_group_fwd(&xg, ld_xg, rows, num_frames, hidden, &w_hh, reverse ? 1 : 0, &cell, &y, ld_y, 1, stream);
}

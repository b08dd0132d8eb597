// lstm.cu — the recurrence of one LSTM layer / direction (torch.nn.LSTM semantics, gate order i, f, g, o),
// the bottleneck of DCCRN (aps/sse/bss/dccrn.py:20-50, LSTMP -> nn.LSTM) and of every other nn.LSTM user.
//
// The input projection x_t W_ih^T + b_ih + b_hh of ALL frames is one large GEMM done beforehand on the tensor-core
// engine (ops.linear); what is left per frame is gates = xg_t + h_{t-1} W_hh^T, the cell update and h_t — a
// [rows, H] x [H, 4H] product that is sequential over frames.  One launch per frame fuses the product with the cell
// update: a CTA owns 32 rows x 16 hidden units (all four gates of a unit sit in one thread, so the cell update needs
// no exchange), walks K = H in 32-wide chunks through double-buffered shared memory and writes h_t straight into the
// layer output y[:, t, :], which is also where frame t + 1 reads h_{t-1} from.  Exact fp32 FMA arithmetic, precise
// expf / tanhf: the recurrence amplifies rounding, and parity is against the reference's fp32 CPU path.
//
// Bound by the fp32 FMA pipe: 2 * rows * 4H * H flop per frame (0.54 GFLOP at rows = 256, H = 512) against
// 148 SMs * 128 FMA/clk.  W_hh (4 MB at H = 512) and h_{t-1} stay L2 resident across frames.
#include "../../include/aps_b200.h"
#include "common.cuh"

namespace apsb {

constexpr int kLstmBM = 32;        // rows per CTA
constexpr int kLstmBU = 16;        // hidden units per CTA (x 4 gates = 64 gate columns)
constexpr int kLstmBK = 32;        // k chunk
constexpr int kLstmThreads = 128;  // thread = 4 rows x 1 unit x 4 gates
constexpr int kLstmALd = kLstmBK + 4;

struct LstmStepParams {
    const float* xg;      // [rows][4H] of this frame, row stride ldx
    long long ldx;
    const float* h_prev;  // [rows][H] of the previous frame, row stride ldh; null on the first frame (h = 0)
    long long ldh;
    float* c;             // [rows, H] cell state, updated in place
    float* y;             // [rows][H] of this frame, row stride ldy
    long long ldy;
    const float* w;       // W_hh [4H, H]
    int rows, H, first;
};

__device__ __forceinline__ float sigmoid_precise(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(kLstmThreads) lstm_step_kernel(const __grid_constant__ LstmStepParams p) {
    __shared__ __align__(16) float As[2][kLstmBM][kLstmALd];
    __shared__ __align__(16) float Ws[2][kLstmBK][kLstmBU * 4];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int u0 = blockIdx.x * kLstmBU, r0 = blockIdx.y * kLstmBM;
    const int H = p.H;
    const int unit = u0 + tx;
    const bool unit_ok = unit < H;

    // gate pre-activations from the input projection and the old cell state: issued first, consumed last
    float xg[4][4], c_old[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = r0 + ty * 4 + r;
        const bool ok = unit_ok && row < p.rows;
#pragma unroll
        for (int g = 0; g < 4; ++g) xg[r][g] = ok ? __ldg(p.xg + (long long)row * p.ldx + (long long)g * H + unit) : 0.f;
        c_old[r] = (ok && !p.first) ? p.c[(long long)row * H + unit] : 0.f;
    }

    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[r][g] = 0.f;

    if (!p.first) {
        const int nchunks = (H + kLstmBK - 1) / kLstmBK;
        float4 ra[2], rw[4];
        auto load_global = [&](int chunk) {
            const int k0 = chunk * kLstmBK;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int idx = tid + kLstmThreads * j, kq = idx & 7, row = r0 + (idx >> 3), k = k0 + 4 * kq;
                ra[j] = (row < p.rows && k < H) ? __ldg(reinterpret_cast<const float4*>(p.h_prev + (long long)row * p.ldh + k))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = tid + kLstmThreads * j, kq = idx & 7, g = (idx >> 3) & 3, u = u0 + (idx >> 5), k = k0 + 4 * kq;
                rw[j] = (u < H && k < H) ? __ldg(reinterpret_cast<const float4*>(p.w + ((long long)g * H + u) * H + k))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto store_smem = [&](int buf) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int idx = tid + kLstmThreads * j, kq = idx & 7, row = idx >> 3;
                *reinterpret_cast<float4*>(&As[buf][row][4 * kq]) = ra[j];
            }
            // W transposed to [k][unit][gate]; the unit index is XOR-swizzled with k / 4 so that the 32 lanes of a
            // store (8 k-quads x 4 gates) hit 32 different banks, and the float4 reads below stay a permutation
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = tid + kLstmThreads * j, kq = idx & 7, g = (idx >> 3) & 3, u = idx >> 5;
                const int col = ((u ^ kq) << 2) + g;
                Ws[buf][4 * kq + 0][col] = rw[j].x;
                Ws[buf][4 * kq + 1][col] = rw[j].y;
                Ws[buf][4 * kq + 2][col] = rw[j].z;
                Ws[buf][4 * kq + 3][col] = rw[j].w;
            }
        };

        load_global(0);
        store_smem(0);
        __syncthreads();
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            const int buf = chunk & 1;
            if (chunk + 1 < nchunks) load_global(chunk + 1);
#pragma unroll
            for (int kq = 0; kq < kLstmBK / 4; ++kq) {
                float4 a[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(&As[buf][ty * 4 + r][4 * kq]);
                const int col = (tx ^ kq) << 2;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 w = *reinterpret_cast<const float4*>(&Ws[buf][4 * kq + e][col]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float av = e == 0 ? a[r].x : e == 1 ? a[r].y : e == 2 ? a[r].z : a[r].w;
                        acc[r][0] = fmaf(av, w.x, acc[r][0]);
                        acc[r][1] = fmaf(av, w.y, acc[r][1]);
                        acc[r][2] = fmaf(av, w.z, acc[r][2]);
                        acc[r][3] = fmaf(av, w.w, acc[r][3]);
                    }
                }
            }
            if (chunk + 1 < nchunks) store_smem(buf ^ 1);
            __syncthreads();
        }
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
        p.c[(long long)row * H + unit] = c;
        p.y[(long long)row * p.ldy + unit] = og * tanhf(c);
    }
}

}  // namespace apsb

using namespace apsb;

extern "C" int aps_b200_lstm_fwd(const float* xg, int64_t ld_xg, int64_t rows, int64_t num_frames, int64_t hidden,
                                 const float* w_hh, int reverse, float* cell, float* y, int64_t ld_y, void* stream) {
    APSB_CHECK_ARG(xg && w_hh && cell && y, "null pointer argument");
    APSB_CHECK_ARG(rows > 0 && num_frames > 0 && hidden > 0 && rows < (1LL << 31) && num_frames < (1LL << 31) &&
                       hidden < (1LL << 29), "bad shape");
    APSB_CHECK_ARG(hidden % 4 == 0 && ld_y % 4 == 0 && ld_y >= hidden && ld_xg >= 4 * hidden &&
                       ((uintptr_t)y & 15) == 0 && ((uintptr_t)w_hh & 15) == 0,
                   "lstm: hidden (%lld) and ld_y (%lld) must be multiples of 4 and y, w_hh 16-byte aligned",
                   (long long)hidden, (long long)ld_y);
    const int64_t grid_y = (rows + kLstmBM - 1) / kLstmBM;
    APSB_CHECK_ARG(grid_y <= 65535, "lstm: too many rows (%lld)", (long long)rows);
    LstmStepParams p{};
    p.ldx = num_frames * ld_xg;
    p.ldh = p.ldy = num_frames * ld_y;
    p.c = cell; p.w = w_hh; p.rows = (int)rows; p.H = (int)hidden;
    const dim3 grid((unsigned)((hidden + kLstmBU - 1) / kLstmBU), (unsigned)grid_y);
    for (int64_t t = 0; t < num_frames; ++t) {
        const int64_t f = reverse ? num_frames - 1 - t : t, fp = reverse ? f + 1 : f - 1;
        p.xg = xg + f * ld_xg;
        p.y = y + f * ld_y;
        p.first = t == 0;
        p.h_prev = t ? y + fp * ld_y : nullptr;
        lstm_step_kernel<<<grid, kLstmThreads, 0, (cudaStream_t)stream>>>(p);
    }
    APSB_LAUNCH_CHECK();
    return 0;
}

// istft.cu — F3: fused inverse rFFT + synthesis window + overlap-add + window^2 de-normalisation.
//
// Replaces /root/reference/aps/transform/utils.py:293-360 (_inverse_stft): the Hermitian
// re-mirror (:327-332), the dense iDFT as conv_transpose1d (:336), the SECOND conv_transpose1d
// that overlap-adds window^2 (:345-349, input independent), the centre trim (:354-357) and the
// divide (:358); and utils.py:418-469 (_pytorch_istft) for stft_mode="torch".
//
// A CTA owns TC hops of output samples of one row.  It loads the spectra of the TC + halo frames
// that touch those samples as a [bin][frame] shared tile (8*NF-byte runs from the [F, T, 2]
// layout), each lane group inverse-transforms one frame in registers (fft_core.cuh), the windowed
// frames land in a shared frame buffer, and a final pass gathers them in a FIXED order
// (deterministic, no atomics), divides by the on-the-fly sum of window^2 and writes the samples
// coalesced: 2056 B read + 1024 B written per frame at 512/256.
#include "../../include/aps_b200.h"
#include "common.cuh"
#include "fft_core.cuh"

#include <math.h>

namespace apsb {

constexpr int kIThreads = 256;

struct IstftParams {
    const float* spec;    // [rows, F, T, 2]
    float* out;           // [rows, S_out]
    long long rows;
    int T, nfft, width, hop, pad, polar;
    long long S_out;      // samples per row written: (T-1)*hop + width - 2*pad
    float scale;          // 1/nfft, or 1/sqrt(nfft) when normalized
    float eps;
    const float* window;  // [width]
    const float2* tables; // inverse tables
    int TC, halo, chunks_per_row, nf_log2;
    long long total_chunks;
};

struct ISmem {
    int win, tw, ptw, tile, fb, buf, total;
};

__host__ __device__ inline int ialign16(int x) { return (x + 15) & ~15; }

template <int NC>
__host__ __device__ inline ISmem make_ilayout(int nfft, int NF) {
    using P = FFTPlan<NC>;
    ISmem s;
    int off = 0;
    s.win = off;  off = ialign16(off + nfft * 4);
    s.tw = off;   off = ialign16(off + P::TW_TOTAL * 8);
    s.ptw = off;  off = ialign16(off + NC * 8);
    s.tile = off; off = ialign16(off + (NC + 1) * (NF + 1) * 8);
    s.fb = off;   off = ialign16(off + NF * (nfft + 2) * 4);
    s.buf = off;  off = ialign16(off + (kIThreads / P::G) * P::BUF * 8);
    s.total = off;
    return s;
}

template <int NC>
__global__ void __launch_bounds__(kIThreads) istft_kernel(const __grid_constant__ IstftParams p) {
    using P = FFTPlan<NC>;
    constexpr int G = P::G;
    constexpr int NGROUPS = kIThreads / G;
    extern __shared__ __align__(16) unsigned char smem[];
    const int NF = p.TC + p.halo;
    const ISmem L = make_ilayout<NC>(p.nfft, NF);
    float* sm_win = reinterpret_cast<float*>(smem + L.win);
    float2* sm_tw = reinterpret_cast<float2*>(smem + L.tw);
    float2* sm_ptw = reinterpret_cast<float2*>(smem + L.ptw);
    float2* sm_tile = reinterpret_cast<float2*>(smem + L.tile);
    float* sm_fb = reinterpret_cast<float*>(smem + L.fb);

    const int tid = threadIdx.x, lane = tid & 31, l = tid & (G - 1), group = tid / G;
    const unsigned mask = group_mask(G, lane);
    float2* buf = reinterpret_cast<float2*>(smem + L.buf) + group * P::BUF;
    const int ts = NF + 1;          // tile row stride (float2)
    const int fs = p.nfft + 2;      // frame buffer row stride (floats, even: float2 stores)

    for (int i = tid; i < p.nfft; i += kIThreads) sm_win[i] = (i < p.width) ? __ldg(p.window + i) : 0.f;
    for (int i = tid; i < P::TW_TOTAL; i += kIThreads) sm_tw[i] = __ldg(p.tables + i);
    for (int i = tid; i < NC; i += kIThreads) {
        float2 t = __ldg(p.tables + P::TW_TOTAL + i);
        sm_ptw[i] = make_float2(p.scale * t.x, p.scale * t.y);
    }

    // Round-2 structure (ncu r02a: 195 us, long-scoreboard 3.3 / issue, 25 % warps active): the frame count of a chunk
    // (NF = TC + halo) is a POWER OF TWO and a multiple of the lane groups, so (a) the tile index splits with shifts
    // instead of an emulated division per element and a bin's NF frames are one aligned-size run, (b) the inverse
    // FFTs take exactly NF / NGROUPS full passes (TC = 16 + 1 halo frame used to cost a second pass for one frame),
    // (c) eight independent 8-byte loads per thread are in flight before the first is consumed, and (d) the overlap-add
    // indexes samples relative to the chunk in 32-bit arithmetic (the 64-bit divisions per output sample are gone).
    const int lg = p.nf_log2;
    const unsigned total = (unsigned)(NC + 1) << lg;
    for (long long chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x) {
        const long long row = chunk / p.chunks_per_row;
        const int c = (int)(chunk - row * p.chunks_per_row);
        const bool last = (c == p.chunks_per_row - 1);
        const int tb = c * p.TC;                       // first hop owned by this CTA
        const int f0 = tb - p.halo;                    // first frame loaded (may be < 0)
        // samples (un-trimmed coordinates) owned: [tb*hop, s_end)
        const long long s_beg = (long long)tb * p.hop;
        const long long s_full = (long long)(p.T - 1) * p.hop + p.width;
        const long long s_end = last ? s_full : (long long)(tb + p.TC) * p.hop;
        const int f_hi = min(p.T, tb + p.TC);          // frames f0 .. f_hi-1 are needed
        const int nfr = f_hi - f0;

        __syncthreads();
        // ---- load the [bin][frame] tile -----------------------------------------------------------
        const float2* sp = reinterpret_cast<const float2*>(p.spec) + row * (long long)(NC + 1) * p.T;
        for (unsigned base = tid; base < total; base += kIThreads * 8) {
            float2 X[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned idx = base + u * kIThreads;
                const int k = (int)(idx >> lg), f = (int)(idx & ((1u << lg) - 1u));
                const int t = f0 + f;
                X[u] = (idx < total && t >= 0 && t < f_hi) ? __ldg(sp + (unsigned)k * (unsigned)p.T + (unsigned)t)
                                                           : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned idx = base + u * kIThreads;
                if (idx >= total) break;
                const int k = (int)(idx >> lg), f = (int)(idx & ((1u << lg) - 1u));
                float2 v = X[u];
                if (p.polar) {
                    float sn, cs;
                    sincosf(v.y, &sn, &cs);
                    v = make_float2(v.x * cs, v.x * sn);
                }
                if (k == 0 || k == NC) v.y = 0.f;      // a C2R transform ignores Im of DC / Nyquist
                sm_tile[k * ts + f] = v;
            }
        }
        __syncthreads();

        // ---- inverse transform of every frame (NF is a multiple of NGROUPS) ------------------------
        for (int f = group; f < NF; f += NGROUPS) {
            if (f - group >= nfr) break;               // whole pass beyond the chunk's frames (warp-uniform)
            const int fe = min(f, nfr - 1);
            float2 v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int k = l + G * q;
                const float2 X = sm_tile[k * ts + fe];
                const float2 Xp = sm_tile[(NC - k) * ts + fe];
                v[q] = irfft_merge(X, Xp, sm_ptw[k], p.scale);
            }
            group_fft<NC, true>(v, buf, sm_tw, l, mask);
            if (f < nfr) {
                float* fr = sm_fb + f * fs;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int n = l + G * q;
                    const float2 w = *reinterpret_cast<const float2*>(sm_win + 2 * n);
                    *reinterpret_cast<float2*>(fr + 2 * n) = make_float2(v[q].x * w.x, v[q].y * w.y);
                }
            }
            __syncwarp(mask);
        }
        __syncthreads();

        // ---- deterministic overlap-add + de-normalisation ------------------------------------------
        float* o = p.out + row * p.S_out;
        const int j_end = (int)(s_end - s_beg);
        const int t_last = p.T - 1 - tb;               // last existing frame, relative to tb
        for (int j = tid; j < j_end; j += kIThreads) {
            const long long so = s_beg + j - p.pad;
            if (so < 0 || so >= p.S_out) continue;
            // frames tr (relative to tb) with tr*hop <= j < tr*hop + width, clipped to the frames that exist
            const int tr_hi = min(t_last, (int)((unsigned)j / (unsigned)p.hop));
            const int lo = j - p.width + 1;
            int tr_lo = lo <= 0 ? -(int)((unsigned)(-lo) / (unsigned)p.hop) : (int)(((unsigned)lo + (unsigned)p.hop - 1u) / (unsigned)p.hop);
            tr_lo = max(tr_lo, -min(tb, p.halo));
            float acc = 0.f, den = 0.f;
            for (int tr = tr_lo; tr <= tr_hi; ++tr) {
                const int n = j - tr * p.hop;
                const float w = sm_win[n];
                acc += sm_fb[(tr + p.halo) * fs + n];
                den = fmaf(w, w, den);
            }
            o[so] = acc / (den + p.eps);
        }
    }
}

template <int NC>
static int launch_istft(IstftParams& p, cudaStream_t st) {
    auto kern = istft_kernel<NC>;
    constexpr int NGROUPS = kIThreads / FFTPlan<NC>::G;
    p.halo = (p.width - 1) / p.hop;
    // frames per chunk: a power of two, a multiple of the lane groups, large enough to own at least one hop
    int NF = NGROUPS < 16 ? 16 : NGROUPS;
    while (NF - p.halo < 1 || (NF - p.halo) * 4 < NF) NF *= 2;          // keep the halo re-computation below 3/4
    APSB_CHECK_ARG(NF <= 4096, "istft: hop %d is too small for a window of %d samples", p.hop, p.width);
    int TC = NF - p.halo;
    ISmem L = make_ilayout<NC>(p.nfft, NF);
    p.nf_log2 = 0;
    while ((1 << p.nf_log2) < NF) ++p.nf_log2;
    APSB_CHECK_ARG(L.total <= 227 * 1024, "istft: shared memory need %d B exceeds 227 KB (hop %d too small?)", L.total,
                   p.hop);
    p.TC = TC;
    p.chunks_per_row = (p.T + TC - 1) / TC;
    p.total_chunks = p.rows * p.chunks_per_row;
    static LaunchCache slots[64];                             // per instantiation and device
    LaunchCache& lc = launch_cache(slots);
    if (L.total > lc.smem_set) {
        APSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        lc.smem_set = L.total;
    }
    if (L.total != lc.occ_smem) {
        int o = 0;
        APSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kIThreads, L.total));
        lc.occ = o < 1 ? 1 : o;
        lc.occ_smem = L.total;
    }
    long long grid = (long long)num_sms() * lc.occ;
    if (grid > p.total_chunks) grid = p.total_chunks;
    kern<<<(unsigned)grid, kIThreads, L.total, st>>>(p);
    APSB_LAUNCH_CHECK();
    return 0;
}

}  // namespace apsb

using namespace apsb;

extern "C" int64_t aps_b200_istft_num_samples(int64_t num_frames, int frame_width, int hop, int center_pad) {
    return (num_frames - 1) * (int64_t)hop + frame_width - 2LL * center_pad;
}

extern "C" int aps_b200_istft_fwd(const float* spec, int64_t rows, int64_t num_frames,
                                  const aps_b200_stft_desc* d, int polar, float eps, float* out, void* stream) {
    APSB_CHECK_ARG(spec && out && d && d->window && d->twiddles, "null pointer argument");
    APSB_CHECK_ARG(rows > 0 && num_frames > 0, "bad shape rows=%lld frames=%lld", (long long)rows,
                   (long long)num_frames);
    APSB_CHECK_ARG(d->frame_width > 0 && d->frame_width <= d->nfft, "frame_width %d not in (0, nfft]", d->frame_width);
    APSB_CHECK_ARG(d->hop > 0 && d->center_pad >= 0, "bad hop / padding");
    IstftParams p{};
    p.spec = spec; p.out = out; p.rows = rows; p.T = (int)num_frames;
    p.nfft = d->nfft; p.width = d->frame_width; p.hop = d->hop; p.pad = d->center_pad; p.polar = polar;
    p.S_out = aps_b200_istft_num_samples(num_frames, d->frame_width, d->hop, d->center_pad);
    APSB_CHECK_ARG(p.S_out > 0, "center padding %d leaves no samples", d->center_pad);
    p.scale = d->scale; p.eps = eps; p.window = d->window;
    p.tables = reinterpret_cast<const float2*>(d->twiddles);
    cudaStream_t st = (cudaStream_t)stream;
    switch (d->nfft) {
        case 64: return launch_istft<32>(p, st);
        case 128: return launch_istft<64>(p, st);
        case 256: return launch_istft<128>(p, st);
        case 512: return launch_istft<256>(p, st);
        case 1024: return launch_istft<512>(p, st);
        default: return set_error(-1, "unsupported FFT size %d (need a power of two in [64, 1024])", d->nfft);
    }
}

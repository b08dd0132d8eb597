// feat_epilogue.cuh — shared tail of the feature kernels: banded mel filterbank, log, per-frame or
// global CMVN and the coalesced store of one frame's feature vector by a group of G lanes.
//
// Replaces /root/reference/aps/transform/asr.py:416-428 (MelTransform.forward, a dense GEMM on a
// 97.6 %-sparse matrix), :453-464 (LogTransform), :576-585 + :607-611 (CmvnTransform per-band/global).
#pragma once
#include <cuda_runtime.h>

namespace apsb {

struct FeatParams {
    int power, M, mel_stride, log_mode, cmvn_mode, norm_mean, norm_var, D;
    const int* mel_start;
    const int* mel_len;
    const float* mel_w;
    int mel_in_smem;
    float log_eps, log_lb, cmvn_eps;
    const float* gmean;
    const float* gstd;
    int* nan_count;   // optional device counter of NaN feature values (asr.py:41 check_valid)
    const float* aug_mask;   // optional SpecAugment 0/1 mask, same layout as the output [rows, T, D]
    const float* out_base;   // start of the output (the mask row of a frame sits at the same offset as its output row)
};

template <int G>
__device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

// mag      : this frame's |X| or |X|^2, D_in = F values (shared memory)
// sm_mel_i : [2*M + FI] band start / length, then per lane-slot i the max length over bands d = l + G*i (shared);  sm_mel_w: [M, mel_stride] weights in shared memory
//            (used when p.mel_in_smem, else p.mel_w in global memory)
// o        : output row of this frame (D floats) or nullptr to skip the store
// FULL: the caller guarantees D == G*FI, M == D and mel weights in shared memory (every guard folds away)
template <int G, int FI, bool FULL = false>
__device__ __forceinline__ void feature_epilogue(const FeatParams& p, const float* __restrict__ mag,
                                                 const int* __restrict__ sm_mel_i,
                                                 const float* __restrict__ sm_mel_w, int l, unsigned mask,
                                                 float* __restrict__ o) {
    float feat[FI];
    const int D = FULL ? G * FI : p.D;
#pragma unroll
    for (int i = 0; i < FI; ++i) {
        const int d = l + G * i;
        float acc = 0.f;
        if (FULL || d < D) {
            if (FULL || p.M > 0) {
                const int s0 = sm_mel_i[d], n = sm_mel_i[p.M + d];
                const float* mg = mag + s0;
                if (FULL || p.mel_in_smem) {   // keep the two address spaces apart: LDS, not generic LD
                    // uniform trip count for the whole group (weights are zero padded to mel_stride and the
                    // magnitude buffer has slack behind bin F-1), so there is no lane divergence
                    const float* wr = sm_mel_w + d * p.mel_stride;
                    const int nmax = sm_mel_i[2 * p.M + i];
#pragma unroll 4
                    for (int t = 0; t < nmax; ++t) acc = fmaf(wr[t], mg[t], acc);
                } else {
                    const float* wr = p.mel_w + d * p.mel_stride;
                    for (int t = 0; t < n; ++t) acc = fmaf(__ldg(wr + t), mg[t], acc);
                }
            } else {
                acc = mag[d];
            }
        }
        feat[i] = acc;
    }
    // clamp mode: __logf (MUFU.LG2, abs err < 4e-7) — far inside the 1e-4 budget after CMVN;
    // lower-bound mode keeps the fully accurate logf because log(lb + x) can sit next to 0
    if (p.log_mode == 1) {
        const float le = p.log_eps;
#pragma unroll
        for (int i = 0; i < FI; ++i) feat[i] = __logf(feat[i] < le ? le : feat[i]);   // NaN propagates like th.clamp
    } else if (p.log_mode == 2) {
        const float lb = p.log_lb;
#pragma unroll
        for (int i = 0; i < FI; ++i) feat[i] = logf(lb + feat[i]);
    }
    if (p.cmvn_mode == 1) {
        const float invD = 1.0f / (float)D;
        float mean = 0.f;
        if (p.norm_mean || p.norm_var) {
            // shifted mean: pivot + mean(x - pivot).  Exact for a constant frame (digital silence), where
            // 1/sqrt(var+eps) ~ 2900 would otherwise amplify one ulp of the mean into a visible feature.
            const float pivot = __shfl_sync(mask, feat[0], (threadIdx.x & 31) & ~(G - 1));
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < FI; ++i) s += (l + G * i < D) ? feat[i] - pivot : 0.f;
            mean = fmaf(group_sum<G>(s, mask), invD, pivot);
        }
        if (p.norm_mean) {
#pragma unroll
            for (int i = 0; i < FI; ++i) feat[i] -= mean;
        }
        if (p.norm_var) {
            const float ctr = p.norm_mean ? 0.f : mean;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < FI; ++i) {
                const float dlt = feat[i] - ctr;
                s += (l + G * i < D) ? dlt * dlt : 0.f;
            }
            const float var = group_sum<G>(s, mask) * invD;
            const float inv = 1.0f / sqrtf(var + p.cmvn_eps);
#pragma unroll
            for (int i = 0; i < FI; ++i) feat[i] *= inv;
        }
    } else if (p.cmvn_mode == 2) {
#pragma unroll
        for (int i = 0; i < FI; ++i) {
            const int d = l + G * i;
            if (d < D) {
                if (p.norm_mean) feat[i] -= __ldg(p.gmean + d);
                if (p.norm_var) feat[i] = feat[i] / __ldg(p.gstd + d);
            }
        }
    }
    if (o != nullptr && p.aug_mask != nullptr) {
        // SpecAugment with mask_zero (asr.py:678-679): x * mask, fused here so the augmented features are written once
        const float* mrow = p.aug_mask + (o - p.out_base);
#pragma unroll
        for (int i = 0; i < FI; ++i) {
            const int d = l + G * i;
            if (d < D) feat[i] *= __ldg(mrow + d);
        }
    }
    if (o != nullptr) {
        bool bad = false;   // any NaN among this lane's values (the host only tests the counter against 0)
#pragma unroll
        for (int i = 0; i < FI; ++i) {
            const int d = l + G * i;
            if (d < D) {
                o[d] = feat[i];
                bad |= (feat[i] != feat[i]);
            }
        }
        if (bad && p.nan_count != nullptr) atomicAdd(p.nan_count, 1);
    }
}

// host-side validation + copy of the public descriptor; returns 0 or an error code
int fill_feat_params(FeatParams& p, const struct aps_b200_feat_desc* feat, int num_bins);

}  // namespace apsb

// fft_core.cuh — register/shared-memory FFT building blocks for sm_100a.
//
// A "group" of G = NC/16 lanes (NC = complex FFT length = nfft/2, G in {2..32}, always
// inside one warp) owns one frame.  Every lane keeps 16 complex values in registers;
// lane l, register q holds element  l + G*q  of the current stage (both before the
// first and after the last pass).  The transform is a Stockham auto-sort FFT with
// radices {r0, 16[, 16]}: each pass does 16/R radix-R butterflies per lane fully in
// registers and exchanges through a padded shared-memory buffer private to the group
// (index i lives at i + i/16, which makes both the strided stores and the unit-stride
// loads bank-conflict free).  The real-input transform uses the N/2 packing
// z[n] = x[2n] + i x[2n+1]; the split into the one-sided spectrum pairs bin k with bin
// NC-k, which lives in lane (G-l)%G, register 15-q (or 16-q on lane 0) — fetched with
// warp shuffles, never through shared memory.
//
// This replaces the reference's dense two-sided DFT GEMM
// (/root/reference/aps/transform/utils.py:62-112 init_kernel, :262-274 matmul/conv1d).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace apsb {

// Complex arithmetic on Blackwell's packed fp32 instructions (add/sub/mul/fma.f32x2 -> FADD2 / FMUL2 / FFMA2 in SASS):
// one instruction per complex add and two per complex multiply instead of two and four.  The packed forms have the
// same lane throughput as the scalar ones (scripts/ubench/f32x2.cu) but halve the ISSUE slots, and ptxas folds the
// real/imaginary swaps and negations of the butterflies (multiplication by -i) into operand modifiers (.LO_HI, .NP).
// Results are IEEE-identical for add/sub; the multiply rounds a.y*b.y first instead of a.x*b.x (same error bound).
#ifndef APSB_F32X2
#define APSB_F32X2 1
#endif
#if APSB_F32X2
__device__ __forceinline__ unsigned long long f2_pack(float x, float y) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ float2 f2_unpack(unsigned long long r) {
    float2 c;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(r));
    return c;
}
__device__ __forceinline__ float2 operator+(float2 a, float2 b) {
    unsigned long long r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b.x, b.y)));
    return f2_unpack(r);
}
__device__ __forceinline__ float2 operator-(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b.x, b.y)));
    return f2_unpack(r);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    // (a.x b.x - a.y b.y, a.x b.y + a.y b.x) = a.x * (b.x, b.y) + (-a.y, a.y) * (b.y, b.x)
    unsigned long long t, r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(t) : "l"(f2_pack(-a.y, a.y)), "l"(f2_pack(b.y, b.x)));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(a.x, a.x)), "l"(f2_pack(b.x, b.y)), "l"(t));
    return f2_unpack(r);
}
// element-wise product (window * samples)
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b.x, b.y)));
    return f2_unpack(r);
}
#else
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#endif
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <bool INV>
__device__ __forceinline__ void bfly2(float2& a, float2& b) {
    float2 t = a - b;
    a = a + b;
    b = t;
}

// 4-point DFT, natural order in and out: (a,b,c,d) -> (X0,X1,X2,X3)
template <bool INV>
__device__ __forceinline__ void bfly4(float2& a, float2& b, float2& c, float2& d) {
    float2 s0 = a + c, s1 = a - c, s2 = b + d, s3 = mul_mi<INV>(b - d);
    a = s0 + s2;
    c = s0 - s2;
    b = s1 + s3;
    d = s1 - s3;
}

#define APSB_C8 0.70710678118654752440f   // cos(pi/4)
#define APSB_C16 0.92387953251128675613f  // cos(pi/8)
#define APSB_S16 0.38268343236508977173f  // sin(pi/8)

// W16^m (forward: exp(-2 pi i m/16)) as a compile-time-foldable constant
template <bool INV>
__device__ __forceinline__ float2 w16(int m) {
    const float c[10] = {1.f, APSB_C16, APSB_C8, APSB_S16, 0.f, -APSB_S16, -APSB_C8, -APSB_C16, -1.f, -APSB_C16};
    const float s[10] = {0.f, APSB_S16, APSB_C8, APSB_C16, 1.f, APSB_C16, APSB_C8, APSB_S16, 0.f, -APSB_S16};
    return make_float2(c[m], INV ? s[m] : -s[m]);
}

// Small DFTs on a register array u[R], natural order in and out.
template <int R, bool INV>
struct SmallFFT;

template <bool INV>
struct SmallFFT<2, INV> {
    __device__ __forceinline__ static void run(float2 (&u)[2]) { bfly2<INV>(u[0], u[1]); }
};
template <bool INV>
struct SmallFFT<4, INV> {
    __device__ __forceinline__ static void run(float2 (&u)[4]) { bfly4<INV>(u[0], u[1], u[2], u[3]); }
};
template <bool INV>
struct SmallFFT<8, INV> {
    // n = 2*n1 + n2, k = k1 + 4*k2
    __device__ __forceinline__ static void run(float2 (&u)[8]) {
        bfly4<INV>(u[0], u[2], u[4], u[6]);  // n2 = 0 : A[k1][0] at 2*k1
        bfly4<INV>(u[1], u[3], u[5], u[7]);  // n2 = 1 : A[k1][1] at 2*k1+1
        u[3] = cmul(u[3], w16<INV>(2));      // W8^1
        u[5] = mul_mi<INV>(u[5]);            // W8^2
        u[7] = cmul(u[7], w16<INV>(6));      // W8^3
        float2 o[8];
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            float2 a = u[2 * k1], b = u[2 * k1 + 1];
            o[k1] = a + b;
            o[k1 + 4] = a - b;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = o[i];
    }
};
template <bool INV>
struct SmallFFT<16, INV> {
    // n = 4*n1 + n2, k = k1 + 4*k2
    __device__ __forceinline__ static void run(float2 (&u)[16]) {
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) bfly4<INV>(u[n2], u[n2 + 4], u[n2 + 8], u[n2 + 12]);  // A[k1][n2] at n2+4*k1
#pragma unroll
        for (int k1 = 1; k1 < 4; ++k1) {
#pragma unroll
            for (int n2 = 1; n2 < 4; ++n2) {
                const int m = n2 * k1;
                if (m == 4)
                    u[n2 + 4 * k1] = mul_mi<INV>(u[n2 + 4 * k1]);
                else
                    u[n2 + 4 * k1] = cmul(u[n2 + 4 * k1], w16<INV>(m));
            }
        }
        float2 o[16];
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            float2 a = u[4 * k1], b = u[4 * k1 + 1], c = u[4 * k1 + 2], d = u[4 * k1 + 3];
            bfly4<INV>(a, b, c, d);
            o[k1] = a;
            o[k1 + 4] = b;
            o[k1 + 8] = c;
            o[k1 + 12] = d;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) u[i] = o[i];
    }
};

// exp(-i pi q / 16), q = 0..15: the split twiddle of bin k = l + G q is W^l * this constant (pi k / NC = pi l / NC + pi q / 16
// for every supported NC), so a lane keeps ONE table value in a register and derives its 16 split twiddles with two packed
// instructions each instead of 16 shared-memory loads per frame (the F1 kernel is shared-memory-wavefront bound).
__device__ __forceinline__ float2 split_step(int q) {
    const float c[16] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                         0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
                         0.f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                         -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f};
    const float s[16] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                         0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
                         1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                         0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
    return make_float2(c[q], -s[q]);
}

// ---- plan: radices for a complex length NC (32..512) ------------------------------------------
template <int NC>
struct FFTPlan {
    static_assert(NC == 32 || NC == 64 || NC == 128 || NC == 256 || NC == 512, "unsupported FFT size");
    static constexpr int G = NC / 16;                       // lanes per frame
    static constexpr int R0 = (NC == 256) ? 16 : (NC == 512 ? 2 : NC / 16);
    static constexpr int NPASS = (NC == 256) ? 2 : (NC == 512 ? 3 : 2);
    static constexpr int R1 = 16;
    static constexpr int R2 = 16;                           // only when NPASS == 3
    static constexpr int NS1 = R0;                          // Ns of pass 1
    static constexpr int NS2 = R0 * R1;                     // Ns of pass 2
    // twiddle table sizes ((R-1)*Ns entries per pass with Ns > 1)
    static constexpr int TW1 = (R1 - 1) * NS1;
    static constexpr int TW2 = (NPASS == 3) ? (R2 - 1) * NS2 : 0;
    static constexpr int TW_TOTAL = TW1 + TW2;
    // padded exchange buffer (float2); the +8 makes consecutive groups' buffers start 64 bytes apart
    // modulo 128, so the two half-warps of a warp (G = 16) use complementary shared-memory banks
    static constexpr int BUF = NC + NC / 16 + 8;
};

__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }

__device__ __forceinline__ unsigned group_mask(int G, int lane) {
    return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
}

// One Stockham pass over the 16 registers of every lane of the group.
//   v[q]    : element l + G*q (input);  if LAST, also the output in the same layout
//   buf     : group-private padded exchange buffer
//   tw      : (R-1) x Ns twiddles of this pass, tw[(r-1)*Ns + k] = exp(-/+ 2 pi i k r / (Ns R))
template <int NC, int R, int Ns, bool INV, bool LAST>
__device__ __forceinline__ void stockham_pass(float2 (&v)[16], float2* __restrict__ buf,
                                              const float2* __restrict__ tw, int l) {
    constexpr int G = NC / 16, NB = 16 / R;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int j = l + G * i;
        const int k = j & (Ns - 1);
        float2 u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            u[r] = v[i + r * NB];
            if (Ns > 1 && r > 0) u[r] = cmul(u[r], tw[(r - 1) * Ns + k]);
        }
        SmallFFT<R, INV>::run(u);
        if (LAST) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[i + r * NB] = u[r];
        } else {
            const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) buf[pad16(j0 + r * Ns)] = u[r];
        }
    }
}

template <int NC>
__device__ __forceinline__ void load_regs(float2 (&v)[16], const float2* __restrict__ buf, int l) {
    constexpr int G = NC / 16;
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = buf[pad16(l + G * q)];
}

// Full complex FFT of length NC on the group's registers (in: v[q] = z[l+G q], out: Z[l+G q]).
// `tw` points at the concatenated pass tables (FFTPlan<NC>::TW_TOTAL entries, direction baked in).
template <int NC, bool INV>
__device__ __forceinline__ void group_fft(float2 (&v)[16], float2* __restrict__ buf,
                                          const float2* __restrict__ tw, int l, unsigned mask) {
    using P = FFTPlan<NC>;
    stockham_pass<NC, P::R0, 1, INV, false>(v, buf, tw, l);
    __syncwarp(mask);
    load_regs<NC>(v, buf, l);
    if (P::NPASS == 2) {
        stockham_pass<NC, P::R1, P::NS1, INV, true>(v, buf, tw, l);
    } else {
        __syncwarp(mask);
        stockham_pass<NC, P::R1, P::NS1, INV, false>(v, buf, tw, l);
        __syncwarp(mask);
        load_regs<NC>(v, buf, l);
        stockham_pass<NC, P::R2, P::NS2, INV, true>(v, buf, tw + P::TW1, l);
    }
}

// Fetch the partner of bin k = l + G q, i.e. Z[(NC - k) mod NC], from lane (G-l)%G.
template <int NC>
__device__ __forceinline__ void fetch_partner(const float2 (&v)[16], float2 (&p)[16], int lane, int l,
                                              unsigned mask) {
    constexpr int G = NC / 16;
    const int src = (lane & ~(G - 1)) | ((G - l) & (G - 1));
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        // lanes l != 0 need register 15-q of the source, lane 0 needs its own register (16-q)&15
        float sx = __shfl_sync(mask, v[15 - q].x, src);
        float sy = __shfl_sync(mask, v[15 - q].y, src);
        const float2 own = v[(16 - q) & 15];
        p[q] = (l == 0) ? own : make_float2(sx, sy);
    }
}

// One partner value (see fetch_partner) for register q; all lanes of the group must call it together.
template <int NC>
__device__ __forceinline__ float2 partner_of(const float2 (&v)[16], int q_static_15_minus_q, int q_static_own,
                                             int src, int l, unsigned mask) {
    const float sx = __shfl_sync(mask, v[q_static_15_minus_q].x, src);
    const float sy = __shfl_sync(mask, v[q_static_15_minus_q].y, src);
    const float2 own = v[q_static_own];
    return (l == 0) ? own : make_float2(sx, sy);
}

// Split the packed transform Z (length NC) into the one-sided real-input spectrum.
//   X[k] = 0.5 (Z[k] + conj Zp) - 0.5 i W_{2NC}^k (Z[k] - conj Zp),   k = l + G q
// `ptw[k]` holds 0.5*scale*(cos(pi k/NC), -sin(pi k/NC)), `half` = 0.5*scale.
__device__ __forceinline__ float2 rfft_split(float2 z, float2 zp, float2 hw, float half) {
    const float er = z.x + zp.x, ei = z.y - zp.y, dr = z.x - zp.x, di = z.y + zp.y;
    float2 x;
    x.x = fmaf(hw.x, di, fmaf(hw.y, dr, half * er));
    x.y = fmaf(-hw.x, dr, fmaf(hw.y, di, half * ei));
    return x;
}

// Inverse of rfft_split: from one-sided X[k], X[NC-k] rebuild Z[k] of the packed inverse transform
//   Z[k] = (X[k] + conj Xp) + i conj(W_{2NC}^k) (X[k] - conj Xp)   (un-normalised, scale via hw/half)
// `hw` = s*(cos(pi k/NC), -sin(pi k/NC)) (same table convention as the forward split, any scale s),
// `half` = s.
__device__ __forceinline__ float2 irfft_merge(float2 x, float2 xp, float2 hw, float half) {
    const float er = x.x + xp.x, ei = x.y - xp.y, dr = x.x - xp.x, di = x.y + xp.y;
    // i * conj(W) * D,  conj(W) = (c, +s') with hw = (c, -s')  =>  conj(W) = (hw.x, -hw.y)
    // i*(a+ib)(dr+i di) = -(a di + b dr) + i (a dr - b di),  a = hw.x, b = -hw.y
    float2 z;
    z.x = fmaf(-hw.x, di, fmaf(hw.y, dr, half * er));
    z.y = fmaf(hw.x, dr, fmaf(hw.y, di, half * ei));
    return z;
}

}  // namespace apsb

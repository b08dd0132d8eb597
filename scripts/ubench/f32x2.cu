// microbenchmark: scalar FFMA/FADD vs packed fma.rn.f32x2 / add.f32x2 throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE> __global__ void k(float* out, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    if (MODE == 0) {
#pragma unroll 8
        for (int i = 0; i < ITERS; ++i) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    } else if (MODE == 1) {
        unsigned long long p0, p1, p2, p3, pa, pb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(x0), "f"(x1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(x2), "f"(x3));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(x4), "f"(x5));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(x6), "f"(x7));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
#pragma unroll 8
        for (int i = 0; i < ITERS; ++i) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
        }
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(p0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(p1));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x4), "=f"(x5) : "l"(p2));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x6), "=f"(x7) : "l"(p3));
    } else if (MODE == 2) {
#pragma unroll 8
        for (int i = 0; i < ITERS; ++i) {
            x0 = x0 + a; x1 = x1 + a; x2 = x2 + a; x3 = x3 + a; x4 = x4 + b; x5 = x5 + b; x6 = x6 + b; x7 = x7 + b;
        }
    } else if (MODE == 3) {
        unsigned long long p0, p1, p2, p3, pa;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(x0), "f"(x1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(x2), "f"(x3));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(x4), "f"(x5));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(x6), "f"(x7));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(b));
#pragma unroll 8
        for (int i = 0; i < ITERS; ++i) {
            asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pa));
            asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pa));
            asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pa));
            asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pa));
        }
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(p0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(p1));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x4), "=f"(x5) : "l"(p2));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x6), "=f"(x7) : "l"(p3));
    } else if (MODE == 4) {   // mixed: packed fma + scalar alu-pipe op (lop3) to see co-issue
        unsigned long long p0, p1, p2, p3, pa, pb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(x0), "f"(x1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(x2), "f"(x3));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(x4), "f"(x5));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(x6), "f"(x7));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
        unsigned u0 = threadIdx.x, u1 = u0 * 3, u2 = u0 * 5, u3 = u0 * 7;
#pragma unroll 8
        for (int i = 0; i < ITERS; ++i) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
            u0 = (u0 ^ u1) + 0x9e37u;
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
            u1 = (u1 ^ u2) + 0x9e37u;
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
            u2 = (u2 ^ u3) + 0x9e37u;
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
            u3 = (u3 ^ u0) + 0x9e37u;
        }
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(p0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(p1));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x4), "=f"(x5) : "l"(p2));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x6), "=f"(x7) : "l"(p3));
        x0 += __uint_as_float((u0 ^ u1 ^ u2 ^ u3) & 1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
template <int MODE> void run(const char* name, float* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = 148 * 4, threads = 512;
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double lane_ops = double(blocks) * threads * ITERS * 8;   // 8 scalar results per iteration
    printf("%-28s %8.3f ms  %7.2f T lane-ops/s (x2 flops for fma)  err=%s\n", name, ms, lane_ops / ms * 1e-9,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
    run<0>("scalar FFMA", out); run<1>("fma.rn.f32x2", out); run<2>("scalar FADD", out); run<3>("add.f32x2", out);
    run<4>("fma.f32x2 + alu mix", out);
    return 0;
}

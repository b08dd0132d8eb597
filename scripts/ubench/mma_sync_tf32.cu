// mma_sync_tf32.cu — throughput of the legacy tensor path (mma.sync.m16n8k8 tf32, HMMA in SASS) on sm_100a.
// Question: would a 3xTF32 recurrent GEMM inside the per-frame LSTM kernel (M = 64 rows per CTA, too small and too
// latency-sensitive for a tcgen05 pipeline) beat the fp32 FMA pipe (128 FMA/clk/SM)?  Needs > 3 x 128 MAC/clk/SM.
// Each warp keeps 8 independent accumulator tiles in flight; prints MAC/clk/SM for 4, 8 and 16 warps per SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void mma_loop(float* out, int iters, long long* clk) {
    float c[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[t][j] = 0.f;
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f000000u, 0x3f800000u};
    uint32_t b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 8; ++t)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    long long* clk;
    cudaMalloc(&out, sms * 512 * 4);
    cudaMalloc(&clk, 8);
    const int iters = 4096;
    for (int warps : {4, 8, 16}) {
        mma_loop<<<sms, warps * 32>>>(out, 64, clk);
        mma_loop<<<sms, warps * 32>>>(out, iters, clk);
        long long c = 0;
        cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        const double macs = (double)warps * iters * 8 * (16 * 8 * 8);       // per SM
        printf("%2d warps/SM: %lld clk, %.0f tf32 MAC/clk/SM (fp32 FMA pipe: 128; 3xTF32 break-even: 384) %s\n", warps, c,
               macs / (double)c, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

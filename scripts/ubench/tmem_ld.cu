// microbenchmark: latency of tcgen05.ld.32x32b.xN + tcgen05.wait::ld per warp (4 warps, one TMEM lane quarter each)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int NW>
__global__ void __launch_bounds__(128, 1) k(long long* out, float* sink, int iters) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5;
    float acc = 0.f;
    long long t0 = 0, t1 = 0;
    if (warp < NW) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint32_t r[32];
            const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)((i & 7) * 32);
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
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]);
        }
        t1 = clock64();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}
template <int NW> void run(long long* dout, float* sink) {
    const int iters = 1000;
    k<NW><<<148, 128>>>(dout, sink, iters);
    cudaDeviceSynchronize();
    k<NW><<<148, 128>>>(dout, sink, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
    printf("tcgen05.ld.32x32b.x32 + wait + 32 FADD, %d warp(s): %.1f clk per iteration (%s)\n", NW, (double)h / iters, cudaGetErrorString(e));
}
int main() {
    long long* dout; float* sink;
    cudaMalloc(&dout, 64); cudaMalloc(&sink, 148 * 128 * 4);
    run<1>(dout, sink); run<4>(dout, sink);
    return 0;
}

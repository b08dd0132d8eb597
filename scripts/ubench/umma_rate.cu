// microbenchmark: cycles per tcgen05.mma (kind::tf32 / kind::f16, M=128, N in {64,128,256}) issued back to back by one
// thread from fixed shared-memory operands; variants: one accumulator vs two alternating accumulators.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
template <int KIND>  // 0 tf32, 1 bf16
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND, int BN, int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar2)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar3)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        // idesc: D fp32 (1<<4); A/B format at bits 7 / 10: tf32 = 2, bf16 = 1; N>>3 at 17; M>>4 at 24
        const uint32_t fmt = KIND == 0 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sa = s_u32(base), sb = s_u32(base + 64 * 1024);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            // MODE 0: same operands, one accumulator; 1: two accumulators alternating; 2: operands walk through 4 k-steps and 2 stages
            const uint32_t d = (MODE == 1) ? tm + (i & 1) * BN : tm;
            const uint32_t ko = (MODE == 2) ? ((i & 3) * 32 + ((i >> 2) & 1) * 32768) : 0;
            if (MODE == 5 && (i % 6) == 5) {
                for (int w = 0; w < 2; ++w) {
                    uint32_t dn = 0;
                    while (!dn)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 1;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(dn) : "r"(s_u32(&bar3)) : "memory");
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            mma<KIND>(d, desc128(sa + ko), desc128(sb + ko), idesc, 1);
            if (MODE >= 3 && (i % 6) == 5) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(&bar2)) : "memory");
                if (MODE == 7) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (MODE == 8) {
                    for (int w = 0; w < 2; ++w) {
                        uint32_t dn = 0;
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], 1;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(dn) : "r"(s_u32(&bar3)) : "memory");
                        if (!dn) __trap();
                    }
                }
                if (MODE == 4 || MODE == 6) {
                    for (int w = 0; w < 2; ++w) {
                        uint32_t dn = 0;
                        while (!dn)
                            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 1;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                         : "=r"(dn) : "r"(s_u32(&bar3)) : "memory");
                    }
                    if (MODE == 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
            }
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(s_u32(&bar)) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}
template <int KIND, int BN, int MODE> void run(const char* name, long long* dout, int grid) {
    const int iters = 2000, smem = 170 * 1024;
    cudaFuncSetAttribute(k<KIND, BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<KIND, BN, MODE><<<grid, 128, smem>>>(dout, iters);
    cudaDeviceSynchronize();
    k<KIND, BN, MODE><<<grid, 128, smem>>>(dout, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
    const double flop = 2.0 * 128 * BN * (KIND == 0 ? 8 : 16);
    printf("%-34s grid %3d: issue %6.1f clk/mma, complete %6.1f clk/mma -> %6.0f flop/clk/SM  (%s)\n", name, grid,
           (double)h[0] / iters, (double)h[1] / iters, flop * iters / (double)h[1], cudaGetErrorString(e));
}
int main() {
    long long* dout; cudaMalloc(&dout, 64);
    run<0, 256, 3>("tf32 N=256 commit/6", dout, 148);
    run<0, 256, 4>("tf32 N=256 commit/6 + 2 waits after", dout, 148);
    run<0, 256, 6>("tf32 N=256 commit/6 + 2 try_waits, no fence", dout, 148);
    run<0, 256, 7>("tf32 N=256 commit/6 + fence only", dout, 148);
    run<0, 256, 8>("tf32 N=256 commit/6 + 2 test_waits", dout, 148);
    run<0, 64, 3>("tf32 N=64 commit/6", dout, 148);
    run<0, 64, 4>("tf32 N=64 commit/6 + 2 waits after", dout, 148);
    for (int grid : {148}) {
        run<0, 64, 0>("tf32 N=64  same acc", dout, grid);
        run<0, 128, 0>("tf32 N=128 same acc", dout, grid);
        run<0, 256, 0>("tf32 N=256 same acc", dout, grid);
        run<0, 256, 1>("tf32 N=256 two accs", dout, grid);
        run<0, 256, 2>("tf32 N=256 walking operands", dout, grid);
        run<0, 64, 2>("tf32 N=64  walking operands", dout, grid);
        run<1, 64, 0>("bf16 N=64  same acc", dout, grid);
        run<1, 128, 0>("bf16 N=128 same acc", dout, grid);
        run<1, 256, 0>("bf16 N=256 same acc", dout, grid);
        run<1, 256, 2>("bf16 N=256 walking operands", dout, grid);
    }
    return 0;
}

// common.cuh — error plumbing shared by all translation units of libaps_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

namespace apsb {

// last error text of the calling thread, returned by aps_b200_last_error()
extern thread_local char g_last_error[512];
int set_error(int code, const char* fmt, ...);
int num_sms();

#define APSB_CHECK_ARG(cond, ...)                                  \
    do {                                                           \
        if (!(cond)) return apsb::set_error(-1, __VA_ARGS__);      \
    } while (0)

#define APSB_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return apsb::set_error(-(int)e__ - 1000, "%s failed: %s (%s:%d)", #call,             \
                                   cudaGetErrorString(e__), __FILE__, __LINE__);                 \
    } while (0)

// after a kernel launch: surfaces launch errors (the text contains "out of memory" for OOM so the
// reference trainer's guard, aps/trainer/ddp.py:146, keeps working through the Python shell)
#define APSB_LAUNCH_CHECK() APSB_CUDA(cudaGetLastError())

// Function attributes (opt-in dynamic shared memory) and occupancy answers are PER DEVICE: launchers keep one slot per
// device index (the Python layer supports several GPUs per process).  Concurrent first calls from two host threads may
// both set the attribute — the calls are idempotent, so the race is benign.
struct LaunchCache {
    int smem_set = -1, occ_smem = -1, occ = 1;
};
inline LaunchCache& launch_cache(LaunchCache (&slots)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    return slots[dev & 63];
}

// Programmatic dependent launch (PDL).  Kernels of the encoder stack call pdl_trigger() first thing — the NEXT kernel of the
// stream (or graph) may then be scheduled as soon as SMs free up, instead of after this grid has drained — and pdl_wait()
// before they touch anything a predecessor wrote or still reads: the launch latency, CTA start-up and (tensor-core GEMM)
// mbarrier / TMEM / tensor-map set-up of kernel N+1 overlap the tail of kernel N.  Both instructions are no-ops for a
// kernel launched without the attribute.  APS_B200_PDL=0 turns the attribute off.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace apsb

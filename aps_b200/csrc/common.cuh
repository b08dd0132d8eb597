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

}  // namespace apsb

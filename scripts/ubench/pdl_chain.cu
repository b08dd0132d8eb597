// pdl_chain.cu — what does one link of a dependent launch chain cost?  (open question behind csrc/lstm.cu: the
// recurrence is one launch per frame; 17.7 us per frame measured against an 8.6 us FMA floor.)
// A chain of K tiny kernels, each reading the previous one's output: (a) plain stream order, (b) programmatic dependent
// launch with griddepcontrol.wait at the top, (c) the same with a prologue (W_hh-like prefetch) before the wait.
// Prints microseconds per link for grids of 128 and 256 CTAs x 256 threads.
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void link_plain(const float* in, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + 1.f;
}

template <int PROLOGUE>
__global__ void link_pdl(const float* in, float* out, const float* konst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float k = 0.f;
    if (PROLOGUE && i < n) {
#pragma unroll
        for (int j = 0; j < PROLOGUE; ++j) k += __ldg(konst + (i + j * 4096) % n);      // independent of the predecessor
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (i < n) out[i] = in[i] + 1.f + 0.f * k;
}

template <typename F>
static float time_chain(F launch, int links, cudaStream_t st) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 64; ++i) launch(i);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int i = 0; i < links; ++i) launch(i);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / links;
}

int main() {
    cudaStream_t st;
    cudaStreamCreate(&st);
    const int links = 2000;
    for (int ctas : {128, 256}) {
        const int n = ctas * 256;
        float *a, *b, *k;
        cudaMalloc(&a, n * 4);
        cudaMalloc(&b, n * 4);
        cudaMalloc(&k, n * 4);
        cudaMemset(a, 0, n * 4);
        cudaMemset(b, 0, n * 4);
        cudaMemset(k, 0, n * 4);
        float* buf[2] = {a, b};
        const float plain = time_chain([&](int i) { link_plain<<<ctas, 256, 0, st>>>(buf[i & 1], buf[(i + 1) & 1], n); }, links, st);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(ctas);
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const float pdl0 = time_chain([&](int i) { cudaLaunchKernelEx(&cfg, link_pdl<0>, (const float*)buf[i & 1], buf[(i + 1) & 1], (const float*)k, n); }, links, st);
        const float pdl8 = time_chain([&](int i) { cudaLaunchKernelEx(&cfg, link_pdl<8>, (const float*)buf[i & 1], buf[(i + 1) & 1], (const float*)k, n); }, links, st);
        printf("%d CTAs: plain %.2f us/link, PDL %.2f us/link, PDL + 8-load prologue %.2f us/link (%s)\n", ctas, plain, pdl0, pdl8,
               cudaGetErrorString(cudaGetLastError()));
        cudaFree(a);
        cudaFree(b);
        cudaFree(k);
    }
    return 0;
}

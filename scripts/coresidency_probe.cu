// co-residency probe (profiles/README.md, "K2 streamed"): does a small kernel on a second stream become resident while a persistent kernel of 3 CTAs x 256 threads x 80 registers holds every SM?  Registers are allocated per SM sub-partition: a 32-register warp fits beside it, a 48- or 64-register one does not.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cores scripts/coresidency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int REGS>
__global__ void __launch_bounds__(256, 3) big(unsigned long long *out, long long ns, float *sink)
{
    extern __shared__ char sm[];
    float acc[40];
    for (int i = 0; i < 40; ++i) acc[i] = threadIdx.x * i;
    const unsigned long long t0 = gt();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t0;
    while ((long long) (gt() - t0) < ns) { for (int i = 0; i < 40; ++i) acc[i] = acc[i] * 1.0001f + sm[(threadIdx.x + i) & 1023]; }
    float s = 0; for (int i = 0; i < 40; ++i) s += acc[i];
    if (s == 12345.f) *sink = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = gt();
}
template <int R>
__global__ void __launch_bounds__(32, 32) small(unsigned long long *out, float *sink, int n)
{
    float acc[R];
    for (int i = 0; i < R; ++i) acc[i] = threadIdx.x * i + n;
    for (int k = 0; k < n; ++k) for (int i = 0; i < R; ++i) acc[i] = acc[i] * 1.0001f + acc[(i + 1) % R];
    float s = 0; for (int i = 0; i < R; ++i) s += acc[i];
    if (s == 12345.f) *sink = s;
    if (threadIdx.x == 0) { const unsigned long long t = gt(); atomicMin(out + 2, t); atomicMax(out + 3, t); }
}
template <int R> void run_small(cudaStream_t b, unsigned long long *d, float *sink) { small<R><<<296, 32, 0, b>>>(d, sink, 3); }
int main()
{
    unsigned long long *d, h[4];
    float *sink;
    cudaMalloc(&d, 32); cudaMalloc(&sink, 4);
    cudaStream_t a, b;
    cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    const int smem = 42270;
    cudaFuncSetAttribute(big<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int variant = 0; variant < 4; ++variant) {
        int pct = -1;
        cudaFuncSetAttribute(big<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);

        int occ_big = 0, occ_small = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_big, big<0>, 256, smem);

        cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, big<0>);
        cudaFuncAttributes fb;
        if (variant == 0) cudaFuncGetAttributes(&fb, small<4>); else if (variant == 1) cudaFuncGetAttributes(&fb, small<28>); else if (variant == 2) cudaFuncGetAttributes(&fb, small<44>); else cudaFuncGetAttributes(&fb, small<56>);
        h[0] = h[1] = 0; h[2] = ~0ull; h[3] = 0;
        cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice);
        big<0><<<444, 256, smem, a>>>(d, 20000000ll, sink);
        if (variant == 0) run_small<4>(b, d, sink); else if (variant == 1) run_small<28>(b, d, sink); else if (variant == 2) run_small<44>(b, d, sink); else run_small<56>(b, d, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("variant %d (carveout %d): big regs %d occ %d, small regs %d occ %d; big ran %.2f ms; small first start +%.2f ms, last +%.2f ms  err=%s\n", variant, pct, fa.numRegs, occ_big,
               fb.numRegs, occ_small, (h[1] - h[0]) / 1e6, ((double) h[2] - (double) h[0]) / 1e6, ((double) h[3] - (double) h[0]) / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

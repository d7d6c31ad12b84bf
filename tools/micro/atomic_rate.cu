// Same-address atomicAdd throughput on the device this runs on: 296 CTAs x 8 warps, lane 0 of every warp adds to one of
// `naddr` counters (128 bytes apart) `iters` times, with the result consumed (ATOM) or dropped (RED).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atomic_rate atomic_rate.cu && ./atomic_rate
#include <cstdio>
#include <cuda_runtime.h>
template <bool RET> __global__ void k(unsigned *c, int naddr, int iters, unsigned *sink)
{
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned *a = c + 32 * (warp % naddr);
    unsigned acc = 0;
    if ((threadIdx.x & 31) == 0)
        for (int i = 0; i < iters; i++) {
            if (RET) acc += atomicAdd(a, 32u);
            else atomicAdd(a, 32u);
        }
    if (acc == 0xdeadbeef) *sink = acc;
}
int main()
{
    unsigned *c, *sink;
    cudaMalloc(&c, 1 << 20); cudaMalloc(&sink, 4); cudaMemset(c, 0, 1 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, ctas = 296, warps = ctas * 8;
    for (int ret = 1; ret >= 0; ret--)
        for (int naddr : {1, 2, 4, 8, 64, 2368}) {
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (ret) k<true><<<ctas, 256>>>(c, naddr, iters, sink); else k<false><<<ctas, 256>>>(c, naddr, iters, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%s naddr %4d: %.3f ms, %.2f ns per atomic (all), %.1f ns per atomic per address, latency seen by a warp %.0f ns\n", ret ? "ATOM" : "RED ", naddr, ms,
                   ms * 1e6 / ((double)warps * iters), ms * 1e6 / ((double)warps * iters) * naddr, ms * 1e6 / iters);
        }
    return 0;
}

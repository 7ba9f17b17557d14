// Microbenchmark: two (or three) 7-warp CTAs per SM where warp `idle` of every CTA issues no DMMA --
// how do the remaining 6 warps per CTA land on the four schedulers' tensor units?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(224) k(double *out, int iters, int idle, long long *cyc) {
    extern __shared__ double pad[];
    const int w = threadIdx.x >> 5;
    double c0 = 0, c1 = 0, e0 = 0, e1 = 0, a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    if (w != idle) {
        for (int it = 0; it < iters; ++it) { dmma(c0, c1, a, b); dmma(e0, e1, b, a); }
    }
    __syncthreads();
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + e0 + e1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 4 * 224 * 8); cudaMalloc(&cyc, 148 * 4 * 8);
    const int iters = 65536;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int per_sm = 1; per_sm <= 3; ++per_sm) {
        // dynamic shared memory sized so that exactly per_sm CTAs fit on an SM
        const size_t smem = per_sm == 1 ? 100 * 1024 : per_sm == 2 ? 80 * 1024 : 60 * 1024;
        for (int idle = -1; idle < 7; ++idle) {
            k<<<148 * per_sm, 224, smem>>>(out, iters, idle, cyc);
            cudaDeviceSynchronize();
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            k<<<148 * per_sm, 224, smem>>>(out, iters, idle, cyc);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const int active = idle < 0 ? 7 : 6;
            const double cycles = ms * 1e-3 * 1.965e9;
            printf("CTAs/SM %d idle warp %2d: %.3f DMMA per cycle per SM (%d DMMA warps per SM, kernel %.3f ms)\n", per_sm, idle,
                   2.0 * iters * active * per_sm / cycles, active * per_sm, ms);
        }
    }
    return 0;
}

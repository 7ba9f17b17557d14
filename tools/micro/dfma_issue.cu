// Microbenchmark: DFMA issue rate per warp and per SM on sm_100a as a function of the number of
// independent accumulator chains per thread and of resident warps per SM.  Each operand comes from
// a distinct register (like a register-resident matrix-vector product).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, int iters, long long *cyc) {
    double a[CH], p[CH];
    for (int i = 0; i < CH; ++i) { a[i] = threadIdx.x * 1e-3 + i; p[i] = 1.0 + 1e-9 * (i + threadIdx.x); }
    double v = 1.0 + 1e-12 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) a[i] = fma(p[i], v, a[i]);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CH; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void run(int warps_per_sm) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 2048 * 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    // one CTA per SM with warps_per_sm warps
    k<CH><<<148, 32 * warps_per_sm>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<CH><<<148, 32 * warps_per_sm>>>(out, iters, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_warp = (double)h / ((double)iters * CH);
    double sm_rate = (double)iters * CH * warps_per_sm / (double)h;
    printf("chains %2d warps/SM %2d: %.2f cycles per DFMA per warp, %.2f warp-DFMA per cycle per SM (%.1f lanes/clk), kernel %.3f ms\n",
           CH, warps_per_sm, per_warp, sm_rate, 32 * sm_rate, ms);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 2, 4, 8, 16, 32}) { run<1>(w); }
    for (int w : {1, 4, 8, 16}) { run<2>(w); }
    for (int w : {1, 4, 8, 16}) { run<4>(w); }
    for (int w : {1, 4, 8, 16, 32}) { run<8>(w); }
    for (int w : {1, 4, 8, 16}) { run<16>(w); }
    for (int w : {4, 8}) { run<64>(w); }
    return 0;
}

// Microbenchmark: FP64 tensor-core (mma.sync.m8n8k4.f64 -> DMMA) issue rate on sm_100a as a function
// of independent accumulator chains per warp and resident warps per SM (operands in registers).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k(double *out, int iters, long long *cyc) {
    double c0[CH], c1[CH], a[CH], b[CH];
    for (int i = 0; i < CH; ++i) { c0[i] = c1[i] = 0.0; a[i] = 1.0 + 1e-9 * (threadIdx.x + i); b[i] = 1.0 - 1e-9 * (threadIdx.x + 2 * i); }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma(c0[i], c1[i], a[i], b[i]);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void run(int warps) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 2048 * 8); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    k<CH><<<148, 32 * warps>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    k<CH><<<148, 32 * warps>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_warp = (double)h / ((double)iters * CH);
    double sm_rate = (double)iters * CH * warps / (double)h;
    printf("chains %2d warps/SM %2d: %.2f cycles per DMMA per warp, %.3f DMMA per cycle per SM = %.1f TFLOP/s at 1.965 GHz x 148 SMs\n",
           CH, warps, per_warp, sm_rate, sm_rate * 512 * 1.965e9 * 148 / 1e12);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 14, 16, 28, 32}) run<1>(w);
    for (int w : {1, 4, 7, 8, 14, 16, 28}) run<2>(w);
    for (int w : {1, 4, 8, 14, 16}) run<4>(w);
    for (int w : {1, 4, 8, 16}) run<8>(w);
    return 0;
}

// Shared helpers for the qspectra_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <vector>
#include "../../include/qspectra_b200.h"

typedef double2 cplx;

__host__ __device__ __forceinline__ cplx cmake(double re, double im) {
    cplx z; z.x = re; z.y = im; return z;
}
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cscale(double s, cplx a) { return cmake(s * a.x, s * a.y); }
// acc += a * b   (4 DFMA)
__device__ __forceinline__ void cfma(cplx &acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
// acc += s * b  with real s (2 DFMA)
__device__ __forceinline__ void rfma(cplx &acc, double s, cplx b) {
    acc.x = fma(s, b.x, acc.x);
    acc.y = fma(s, b.y, acc.y);
}
__device__ __forceinline__ double cabs1(cplx a) { return fabs(a.x) + fabs(a.y); }
__device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }

// -i * E(t) of a Gaussian pulse in the rotating frame (pulse.py:110-114, eom.py:87-94)
__device__ __forceinline__ cplx pulse_coefficient(const qsx_pulse &p, double t) {
    double dt = t - p.t_peak;
    double env = p.scale * exp(-dt * dt * p.inv_two_sigma_sq);
    double s, c;
    sincos(p.detuning * dt, &s, &c);
    double er = env * c, ei = env * s;
    if (p.conjugate) ei = -ei;
    return cmake(ei, -er);               // -i * (er + i ei)
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide reductions; `scratch` needs 32 doubles of shared memory.  All
// threads get the result.  Contains two barriers.
__device__ __forceinline__ double block_max(double v, double *scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = scratch[0];
    for (int i = 1; i < nw; ++i) r = fmax(r, scratch[i]);
    return r;
}
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = 0;
    for (int i = 0; i < nw; ++i) r += scratch[i];
    return r;
}

// atomic max on non-negative doubles through their (order-preserving) bit pattern
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long *>(addr),
              (unsigned long long)__double_as_longlong(v));
}

// ---- host side ----------------------------------------------------------
void qsx_set_error(const char *fmt, ...);
extern std::atomic<uint64_t> qsx_launch_counter;
// bytes this library has moved across PCIe since load (bench.py: e2e.h2d/d2h_bytes_per_step)
extern std::atomic<uint64_t> qsx_h2d_counter, qsx_d2h_counter;

#define QSX_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) {                                             \
            qsx_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,        \
                          cudaGetErrorString(e__));                           \
            return QSX_ERR_CUDA;                                              \
        }                                                                     \
    } while (0)

#define QSX_REQUIRE(cond, ...)                                                \
    do {                                                                      \
        if (!(cond)) {                                                        \
            qsx_set_error(__VA_ARGS__);                                       \
            return QSX_ERR_INVALID;                                           \
        }                                                                     \
    } while (0)

// dense_wide.cu: exp(L dt) for state dimensions above the single-CTA propagator kernel
int qsx_dense_expm_wide(const cplx *Lt, int M, int n_gen, const double *lnorm_dev, double dt, cplx *Pt,
                        unsigned long long *gemms, cudaStream_t stream);
int qsx_dense_map_gemm(const cplx *Lt, int M, int n_runs, int R, const int *run_gen_dev, const int *run_save_dev,
                       const cplx *y0, int nt, const cplx *S, int save_rows, long long S_stride, cplx *out,
                       cudaStream_t stream);

// dense_real.cu: generators in Hermitian coordinates (real form)
int qsx_real_form_launch(const cplx *Lt, int M, int n_gen, const int32_t *perm_host, double *Gt, double *gnorm,
                         double *defect, cudaStream_t stream);

int qsx_fused_expm_launch(const cplx *Lt, int M, int n_gen, const int32_t *perm_host, double dt, double *P,
                          double *defect, unsigned long long *gemm_count, cudaStream_t stream);

// Device scratch comes from a small caching pool (power-of-two size classes, blocks up
// to 64 MB are kept for reuse; larger ones go straight to cudaMalloc/cudaFree): the
// per-call metadata buffers of the propagate entry points must not cost a
// cudaMalloc + synchronising cudaFree each time.
cudaError_t qsx_pool_alloc(void **ptr, size_t bytes);
void qsx_pool_free(void *ptr);

// small RAII device buffer for handle-owned tables and per-call scratch
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return qsx_pool_alloc(reinterpret_cast<void **>(&p), count * sizeof(T));
    }
    cudaError_t upload(const T *src, size_t count, cudaStream_t s) {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        qsx_h2d_counter += count * sizeof(T);
        return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    cudaError_t upload(const std::vector<T> &v, cudaStream_t s) { return upload(v.data(), v.size(), s); }
    void release() {
        if (p) qsx_pool_free(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

// Dormand-Prince 5(4) tableau (Hairer, Norsett, Wanner, "Solving ODEs I", II.5)
#define DP_C2 (1.0 / 5.0)
#define DP_C3 (3.0 / 10.0)
#define DP_C4 (4.0 / 5.0)
#define DP_C5 (8.0 / 9.0)
#define DP_A21 (1.0 / 5.0)
#define DP_A31 (3.0 / 40.0)
#define DP_A32 (9.0 / 40.0)
#define DP_A41 (44.0 / 45.0)
#define DP_A42 (-56.0 / 15.0)
#define DP_A43 (32.0 / 9.0)
#define DP_A51 (19372.0 / 6561.0)
#define DP_A52 (-25360.0 / 2187.0)
#define DP_A53 (64448.0 / 6561.0)
#define DP_A54 (-212.0 / 729.0)
#define DP_A61 (9017.0 / 3168.0)
#define DP_A62 (-355.0 / 33.0)
#define DP_A63 (46732.0 / 5247.0)
#define DP_A64 (49.0 / 176.0)
#define DP_A65 (-5103.0 / 18656.0)
#define DP_A71 (35.0 / 384.0)
#define DP_A73 (500.0 / 1113.0)
#define DP_A74 (125.0 / 192.0)
#define DP_A75 (-2187.0 / 6784.0)
#define DP_A76 (11.0 / 84.0)
#define DP_E1 (71.0 / 57600.0)
#define DP_E3 (-71.0 / 16695.0)
#define DP_E4 (71.0 / 1920.0)
#define DP_E5 (-17253.0 / 339200.0)
#define DP_E6 (22.0 / 525.0)
#define DP_E7 (-1.0 / 40.0)

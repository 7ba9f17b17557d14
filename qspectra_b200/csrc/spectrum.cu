// K7: Fourier transform of response functions on the device.
//
// Replaces simulate/utils.py:154-219 (`fourier_transform`: zero-pad the signal
// symmetrically around t = 0, ifftshift -> fft -> fftshift, flip for sign = +1) for
// signals sampled on t = 0, dt, ..., (n-1) dt.  The padded signal has N = 2n - 1
// points of which only n are non-zero, so the whole pipeline collapses to
//
//     X[k] = dt * sum_{j<n} x[j] exp(sign * 2 pi i * j * (k - n + 1) / N),   k = 0..N-1,
//
// a (N x n) x (n x columns) complex matrix product whose left factor is generated from
// a table of the N-th roots of unity held in shared memory (exact integer phase
// arithmetic: no accumulated twiddle error, no padding traffic, no transposes for the
// middle-axis case).  Two-dimensional spectra call it twice (t1 with sign -1, t3 with
// sign +1; response.py:430-455) on the (n_t1, n_t2, n_t3) signal without leaving HBM.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int BK = 64, BC = 64, JC = 16, DFT_THREADS = 256;

struct DftArgs {
    const cplx *x;      // [outer][n][inner]
    cplx *out;          // [outer][N][inner]
    long long outer, inner, cols;
    int n, N;
    double dt;
    int sign;
};

// KFAST: lanes run over k (inner == 1: the transformed axis is the contiguous one);
// otherwise lanes run over columns (inner > 1: columns are contiguous).
template <bool KFAST>
__global__ void __launch_bounds__(DFT_THREADS) dft_sym_kernel(DftArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *tw = reinterpret_cast<cplx *>(smem_raw);          // [N] exp(sign 2 pi i r / N)
    cplx *xs = tw + a.N;                                    // [JC][BC]
    const int tid = threadIdx.x;
    for (int r = tid; r < a.N; r += DFT_THREADS) {
        double s, c;
        sincospi(2.0 * (double)r / (double)a.N, &s, &c);
        tw[r] = cmake(c, a.sign > 0 ? s : -s);
    }
    const int tk = KFAST ? tid % 16 : tid / 16, tc = KFAST ? tid / 16 : tid % 16;
    const long long n_ctile = (a.cols + BC - 1) / BC;
    const long long n_ktile = (a.N + BK - 1) / BK;
    for (long long t = blockIdx.x; t < n_ctile * n_ktile; t += gridDim.x) {
        const long long c0 = (t / n_ktile) * BC;
        const int k0 = (int)(t % n_ktile) * BK;
        cplx acc[4][4];
        int q[4], r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + tk + 16 * i;
            q[i] = ((k - a.n + 1) % a.N + a.N) % a.N;        // phase increment per sample
            r[i] = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[i][b] = cmake(0, 0);
        }
        for (int j0 = 0; j0 < a.n; j0 += JC) {
            __syncthreads();
            for (int e = tid; e < JC * BC; e += DFT_THREADS) {
                const int jj = KFAST ? e % JC : e / BC, cc = KFAST ? e / JC : e % BC;
                const long long c = c0 + cc;
                const int j = j0 + jj;
                cplx v = cmake(0, 0);
                if (c < a.cols && j < a.n) {
                    const long long o = c / a.inner, i = c % a.inner;
                    v = __ldg(&a.x[(o * a.n + j) * a.inner + i]);
                }
                xs[jj * BC + cc] = v;
            }
            __syncthreads();
#pragma unroll 4
            for (int jj = 0; jj < JC; ++jj) {
                cplx w[4], v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    w[i] = tw[r[i]];
                    r[i] += q[i];
                    if (r[i] >= a.N) r[i] -= a.N;
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) v[b] = xs[jj * BC + tc + 16 * b];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int b = 0; b < 4; ++b) cfma(acc[i][b], w[i], v[b]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + tk + 16 * i;
            if (k >= a.N) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const long long c = c0 + tc + 16 * b;
                if (c >= a.cols) continue;
                const long long o = c / a.inner, ii = c % a.inner;
                a.out[(o * a.N + k) * a.inner + ii] = cscale(a.dt, acc[i][b]);
            }
        }
    }
}

}  // namespace

extern "C" int qsx_fourier_transform(const void *x_dev, int64_t outer, int32_t n, int64_t inner,
                                     double dt, int32_t sign, void *out_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(x_dev && out_dev && outer > 0 && inner > 0 && n > 0, "qsx_fourier_transform: bad arguments");
    QSX_REQUIRE(sign == 1 || sign == -1, "qsx_fourier_transform: sign must be +1 or -1");
    QSX_REQUIRE(n <= 4096, "qsx_fourier_transform: at most 4096 samples along the transformed axis");
    DftArgs a;
    a.x = (const cplx *)x_dev; a.out = (cplx *)out_dev;
    a.outer = outer; a.inner = inner; a.cols = outer * inner;
    a.n = n; a.N = 2 * n - 1; a.dt = dt; a.sign = sign;
    const size_t smem = ((size_t)a.N + JC * BC) * sizeof(cplx);
    int dev = 0, sms = 0;
    QSX_CUDA(cudaGetDevice(&dev));
    QSX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long tiles = ((a.cols + BC - 1) / BC) * ((a.N + BK - 1) / BK);
    const int grid = (int)std::min<long long>(tiles, (long long)sms * 4);
    if (inner == 1) {
        QSX_CUDA(cudaFuncSetAttribute(dft_sym_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dft_sym_kernel<true><<<grid, DFT_THREADS, smem, stream>>>(a);
    } else {
        QSX_CUDA(cudaFuncSetAttribute(dft_sym_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dft_sym_kernel<false><<<grid, DFT_THREADS, smem, stream>>>(a);
    }
    qsx_launch_counter += 1;
    QSX_CUDA(cudaGetLastError());
    return QSX_OK;
}

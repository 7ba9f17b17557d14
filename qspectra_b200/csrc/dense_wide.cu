// Tiled complex FP64 GEMM on the FP64 tensor cores (mma.sync.m8n8k4.f64 -> DMMA) and the two
// users it has on the propagation path:
//
//   * qsx_dense_expm_wide -- exp(L dt) for state dimensions above the single-CTA kernel of
//     dense.cu (57 .. 1024; FMO 'fe' = 147, 'gg,ge,eg,ee' = 64): the same algorithm
//     (|A|_inf <= 1/2 scaling, degree-14 Taylor polynomial in Paterson-Stockmeyer form,
//     squarings), each matrix product one launch of the GEMM below over all generators.
//     Replaces the ZVODE loop for a constant generator on a uniform grid (reference
//     simulate/utils.py:45-49, liouville_space.py:316-341) where round 1 called
//     torch.linalg.matrix_exp.
//   * qsx_response_contract -- K6, the signal contraction of the third-order response,
//       S[ab][c] += sum_u w_u sum_i X_u[ab][i] Y_u[c][i]
//     (reference response.py:336, `np.einsum('ci,abi', V_Gt3, V_rho2)`, summed over ensemble
//     members and polarisation configurations u with their weights; decorators.py:55-61,
//     86-92), where round 1 called torch.einsum.  One CTA owns an output tile and walks all
//     units, so the sum over u is deterministic and needs no atomics.
//
// Tile: 32 x 32 complex outputs per CTA, four warps (8 rows x 32 columns each), K in chunks
// of 16 through planar (re / im) shared-memory tiles whose leading dimensions (4 mod 16
// doubles) make the A- and B-fragment loads of a half-warp bank-conflict free.  A complex
// product is three real DMMA products (Gauss), as in dense.cu.
#include "common.cuh"
#include <algorithm>

namespace {

__device__ __forceinline__ void dmma884w(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

constexpr int BM = 32, BN = 32, KC = 16;
constexpr int LDA = 20;      // doubles per row of the A tile [BM][KC]      (4 mod 16)
constexpr int LDB = 36;      // doubles per row of the B tile [KC][BN]      (4 mod 16)
constexpr int LDT = 20;      // doubles per row of the transposed B tile [BN][KC]

struct ZgemmArgs {
    // C_u = A_u B_u (+ epilogue), or with `reduce`: C += sum_u w_u A_u op(B_u)
    const cplx *A, *B;
    cplx *C;
    int M, N, K;                    // C is M x N, contraction length K
    int lda, ldb, ldc;              // row strides (elements)
    long long sA, sB, sC;           // strides between units
    const int *ib;                  // B operand of unit u is B + ib[u] * sB (null: u)
    // reduce with groups: unit u multiplies A_u = sum_{j < grp_count[u]} w[grp_first[u] + j] A[grp_first[u] + j]
    // (summed while the A tile is staged) with B_u -- units that share their B operand cost one product
    const int *grp_first, *grp_count;
    int n_units;
    int b_transposed;               // 1: B_u is stored [N][K] (C = A B^T)
    int reduce;                     // 1: one output, summed over the units with weights w
    int n_split;                    // reduce: the units are dealt out to n_split CTAs per output tile; CTA z writes
                                    // its partial sum to part[z] (deterministic second pass adds them to C)
    cplx *part;                     // [n_split][M][ldc] (reduce with n_split > 1)
    const cplx *w;                  // [n_units] or null (= 1)
    // epilogue of the unit-wise form: C = AB + c0 I + c1 X + c2 Y  (X, Y like C)
    const cplx *X, *Y;
    double c0, c1, c2;
};

__global__ void __launch_bounds__(128) zgemm_dmma_kernel(const ZgemmArgs a) {
    __shared__ __align__(16) double Ar[BM * LDA], Ai[BM * LDA];
    constexpr int BSZ = KC * LDB > BN * LDT ? KC * LDB : BN * LDT;
    __shared__ __align__(16) double Br[BSZ], Bi[BSZ];                // [KC][LDB] or [BN][LDT]
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
    // reduce: contiguous block of units per split
    const int per = a.reduce ? (a.n_units + a.n_split - 1) / a.n_split : 1;
    const int u_begin = a.reduce ? (int)blockIdx.z * per : (int)blockIdx.z;
    const int u_end = a.reduce ? min(a.n_units, u_begin + per) : (int)blockIdx.z + 1;
    const int nK = (a.K + KC - 1) / KC;
    const int n_chunks = max(0, u_end - u_begin) * nK;      // (unit, K chunk) pairs of this CTA, in order

    // The three real products of the complex one (Gauss); in the reducing form the unit weights are
    // applied while the A tile is staged, so the accumulators run over all units of the CTA.
    double p1[4][2], p2[4][2], p3[4][2];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) { p1[nb][0] = p1[nb][1] = p2[nb][0] = p2[nb][1] = p3[nb][0] = p3[nb][1] = 0.0; }

    // global -> registers for chunk c (four A and four B elements per thread); the loads of the
    // next chunk are in flight while the tensor cores work on the current one
    constexpr int EPT = BM * KC / 128;
    cplx ra[EPT], rb[EPT];
    auto fetch = [&](int c) {
        const int u = u_begin + c / nK, k0 = (c % nK) * KC;
        const cplx *Bu = a.B + (size_t)(a.ib ? __ldg(&a.ib[u]) : u) * a.sB;
        if (a.grp_first) {
            const int first = __ldg(&a.grp_first[u]), count = __ldg(&a.grp_count[u]);
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int i = threadIdx.x + 128 * e, r = i / KC, k = i % KC;
                cplx v = cmake(0, 0);
                if (row0 + r < a.M && k0 + k < a.K) {
                    const cplx *Aj = a.A + (size_t)first * a.sA + (size_t)(row0 + r) * a.lda + k0 + k;
                    for (int j = 0; j < count; ++j) cfma(v, __ldg(&a.w[first + j]), __ldg(&Aj[(size_t)j * a.sA]));
                }
                ra[e] = v;
            }
        } else {
            const cplx *Au = a.A + (size_t)u * a.sA;
            const cplx wu = (a.reduce && a.w) ? __ldg(&a.w[u]) : cmake(1.0, 0.0);
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int i = threadIdx.x + 128 * e, r = i / KC, k = i % KC;
                cplx v = cmake(0, 0);
                if (row0 + r < a.M && k0 + k < a.K) v = __ldg(&Au[(size_t)(row0 + r) * a.lda + k0 + k]);
                ra[e] = (a.reduce && a.w) ? cmul(wu, v) : v;
            }
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int i = threadIdx.x + 128 * e;
            cplx v = cmake(0, 0);
            if (a.b_transposed) {       // B stored [N][K]: tile [BN][KC], k fastest
                const int c2 = i / KC, k = i % KC;
                if (col0 + c2 < a.N && k0 + k < a.K) v = __ldg(&Bu[(size_t)(col0 + c2) * a.ldb + k0 + k]);
            } else {                    // B stored [K][N]: tile [KC][BN], column fastest
                const int k = i / BN, c2 = i % BN;
                if (k0 + k < a.K && col0 + c2 < a.N) v = __ldg(&Bu[(size_t)(k0 + k) * a.ldb + col0 + c2]);
            }
            rb[e] = v;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int i = threadIdx.x + 128 * e;
            const int r = i / KC, k = i % KC;
            Ar[r * LDA + k] = ra[e].x; Ai[r * LDA + k] = ra[e].y;
            if (a.b_transposed) {
                Br[r * LDT + k] = rb[e].x; Bi[r * LDT + k] = rb[e].y;
            } else {
                const int kk = i / BN, cc = i % BN;
                Br[kk * LDB + cc] = rb[e].x; Bi[kk * LDB + cc] = rb[e].y;
            }
        }
    };
    static_assert(BM == BN && BM * KC == BN * KC, "one element count for the A and B tiles");

    if (n_chunks > 0) fetch(0);
    for (int c = 0; c < n_chunks; ++c) {
        __syncthreads();                    // the previous chunk's fragments have been read
        stash();
        __syncthreads();
        if (c + 1 < n_chunks) fetch(c + 1);
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
            const double ar = Ar[(8 * wrp + g) * LDA + 4 * ks + t];
            const double ai = Ai[(8 * wrp + g) * LDA + 4 * ks + t];
            const double as = ar + ai;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const int o = a.b_transposed ? (8 * nb + g) * LDT + 4 * ks + t : (4 * ks + t) * LDB + 8 * nb + g;
                const double br = Br[o], bi = Bi[o];
                dmma884w(p1[nb][0], p1[nb][1], ar, br);
                dmma884w(p2[nb][0], p2[nb][1], ai, bi);
                dmma884w(p3[nb][0], p3[nb][1], as, br + bi);
            }
        }
    }
    const int r = row0 + 8 * wrp + g;
    if (r >= a.M) return;
    cplx *Cu = a.C + (a.reduce ? 0 : (size_t)blockIdx.z * a.sC);
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = col0 + 8 * nb + 2 * t + j;
            if (c >= a.N) continue;
            const size_t o = (size_t)r * a.ldc + c;
            cplx v = cmake(p1[nb][j] - p2[nb][j], p3[nb][j] - p1[nb][j] - p2[nb][j]);
            if (a.reduce && a.n_split > 1) {
                a.part[(size_t)blockIdx.z * a.M * a.ldc + o] = v;
                continue;
            }
            if (a.reduce) {
                const cplx old = Cu[o];
                v.x += old.x; v.y += old.y;
            } else {
                if (a.X) { const cplx x = a.X[(size_t)blockIdx.z * a.sC + o]; v.x += a.c1 * x.x; v.y += a.c1 * x.y; }
                if (a.Y) { const cplx y = a.Y[(size_t)blockIdx.z * a.sC + o]; v.x += a.c2 * y.x; v.y += a.c2 * y.y; }
                if (r == c) v.x += a.c0;
            }
            Cu[o] = v;
        }
}

// out[g] = s_g * in[g] with s_g = dt / 2^sq (uniform sq), and B4 = c12 I + c13 A + c14 A^2
__global__ void wide_scale_kernel(const cplx *__restrict__ in, cplx *__restrict__ out, size_t n, double s) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = cscale(s, in[i]);
}
__global__ void wide_poly2_kernel(const cplx *__restrict__ A, const cplx *__restrict__ A2, cplx *__restrict__ out,
                                  int M, size_t n, double c0, double c1, double c2) {
    const size_t mm = (size_t)M * M;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e = i % mm;
        cplx v = cmake(c1 * A[i].x + c2 * A2[i].x, c1 * A[i].y + c2 * A2[i].y);
        if (e / M == e % M) v.x += c0;
        out[i] = v;
    }
}

// C[i] += sum_z part[z][i] in split order
__global__ void zgemm_add_partials_kernel(const cplx *__restrict__ part, int n_split, size_t n, cplx *__restrict__ C) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx v = C[i];
        for (int z = 0; z < n_split; ++z) {
            const cplx p = part[(size_t)z * n + i];
            v.x += p.x; v.y += p.y;
        }
        C[i] = v;
    }
}

int launch_zgemm(const ZgemmArgs &a, cudaStream_t stream) {
    dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, a.reduce ? std::max(1, a.n_split) : a.n_units);
    zgemm_dmma_kernel<<<grid, 128, 0, stream>>>(a);
    qsx_launch_counter += 1;
    QSX_CUDA(cudaGetLastError());
    return QSX_OK;
}

}  // namespace

// exp(L_g dt) for every generator, transposed storage in and out (exp commutes with the
// transpose).  lnorm_host: inf-norms of the generators.  Returns the number of complex GEMMs
// per generator in *gemms.
int qsx_dense_expm_wide(const cplx *Lt, int M, int n_gen, const double *lnorm_dev, double dt, cplx *Pt,
                        unsigned long long *gemms, cudaStream_t stream) {
    std::vector<double> nrm(n_gen);
    qsx_d2h_counter += n_gen * sizeof(double);
    QSX_CUDA(cudaMemcpyAsync(nrm.data(), lnorm_dev, n_gen * sizeof(double), cudaMemcpyDeviceToHost, stream));
    QSX_CUDA(cudaStreamSynchronize(stream));
    double worst = 0;
    for (double x : nrm) {
        QSX_REQUIRE(x == x && x < 1e300, "qsx_dense_expm: non-finite generator");
        worst = std::max(worst, fabs(dt) * x);
    }
    int sq = 0;
    while (worst > 0.5 && sq < 60) { worst *= 0.5; ++sq; }
    const double scale = dt / (double)(1ULL << sq);
    const size_t mm = (size_t)M * M, n = mm * n_gen;
    DevBuf<cplx> A, A2, A3, U;
    QSX_CUDA(A.alloc(n)); QSX_CUDA(A2.alloc(n)); QSX_CUDA(A3.alloc(n)); QSX_CUDA(U.alloc(n));
    static const double inv_fact[15] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040,
                                        1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800,
                                        1.0 / 479001600, 1.0 / 6227020800.0, 1.0 / 87178291200.0};
    const int eb = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    wide_scale_kernel<<<eb, 256, 0, stream>>>(Lt, A.p, n, scale);
    qsx_launch_counter += 1;
    ZgemmArgs g;
    g.M = g.N = g.K = M; g.lda = g.ldb = g.ldc = M; g.sA = g.sB = g.sC = (long long)mm;
    g.n_units = n_gen; g.b_transposed = 0; g.reduce = 0; g.w = nullptr; g.n_split = 1; g.part = nullptr; g.ib = nullptr; g.grp_first = g.grp_count = nullptr;
    g.X = g.Y = nullptr; g.c0 = g.c1 = g.c2 = 0.0;
    int rc;
    // all factors are polynomials in A, so the order of the products is immaterial
    g.A = A.p; g.B = A.p; g.C = A2.p;
    if ((rc = launch_zgemm(g, stream))) return rc;                     // A^2
    g.A = A2.p; g.B = A.p; g.C = A3.p;
    if ((rc = launch_zgemm(g, stream))) return rc;                     // A^3
    // p(A) = B0 + A^3 (B1 + A^3 (B2 + A^3 (B3 + A^3 B4))),  B_i = c_3i I + c_3i+1 A + c_3i+2 A^2
    cplx *P = Pt, *Q = U.p;
    wide_poly2_kernel<<<eb, 256, 0, stream>>>(A.p, A2.p, P, M, n, inv_fact[12], inv_fact[13], inv_fact[14]);
    qsx_launch_counter += 1;
    for (int blk = 3; blk >= 0; --blk) {
        g.A = A3.p; g.B = P; g.C = Q; g.X = A.p; g.Y = A2.p;
        g.c0 = inv_fact[3 * blk]; g.c1 = inv_fact[3 * blk + 1]; g.c2 = inv_fact[3 * blk + 2];
        if ((rc = launch_zgemm(g, stream))) return rc;
        std::swap(P, Q);
    }
    g.X = g.Y = nullptr; g.c0 = g.c1 = g.c2 = 0.0;
    for (int q = 0; q < sq; ++q) {
        g.A = P; g.B = P; g.C = Q;
        if ((rc = launch_zgemm(g, stream))) return rc;
        std::swap(P, Q);
    }
    if (P != Pt) QSX_CUDA(cudaMemcpyAsync(Pt, P, n * sizeof(cplx), cudaMemcpyDeviceToDevice, stream));
    QSX_CUDA(cudaStreamSynchronize(stream));       // the work buffers go back to the pool
    if (gemms) *gemms = (unsigned long long)(6 + sq) * n_gen;
    return QSX_OK;
}

// Propagator stepping as tensor-core GEMMs for MANY columns per propagator (the t2 stage of a
// response function: n_t1 columns per unit under one generator and one save operator; reference
// simulate/utils.py:103-109 runs them one ZVODE solve at a time).  Columns come in `n_runs` runs
// of `R` consecutive columns that share a propagator (run_gen) and a save matrix (run_save):
//   per output time:  out[col][it][:] = S_run Y[col]     (R x M)(M x rows)  per run
//   between times:    Y[col] <- P_run Y[col]             (R x M)(M x M)     per run
// Lt is the engine's transposed storage Lt[g][c][r] = P_g[r][c], i.e. exactly the [K][N]
// operand of Y_new^T = Y^T P^T.  Y ping-pongs between two work buffers.
int qsx_dense_map_gemm(const cplx *Lt, int M, int n_runs, int R, const int *run_gen_dev, const int *run_save_dev,
                       const cplx *y0, int nt, const cplx *S, int save_rows, long long S_stride, cplx *out,
                       cudaStream_t stream) {
    const size_t n = (size_t)n_runs * R * M;
    // two state buffers from the stream-ordered allocator (they can be hundreds of MB: the
    // scratch pool would cudaMalloc / cudaFree them, and cudaFree synchronises the device)
    static bool pool_ready = false;
    if (!pool_ready) {
        int dev = 0;
        cudaMemPool_t mp;
        unsigned long long keep = ~0ULL;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess)
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
        pool_ready = true;
    }
    cplx *Ya = nullptr, *Yb = nullptr;
    QSX_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&Ya), n * sizeof(cplx), stream));
    cudaError_t eb = cudaMallocAsync(reinterpret_cast<void **>(&Yb), n * sizeof(cplx), stream);
    if (eb != cudaSuccess) {
        cudaFreeAsync(Ya, stream);
        qsx_set_error("qsx_dense_map_gemm: %s", cudaGetErrorString(eb));
        return QSX_ERR_CUDA;
    }
    struct Release {
        cplx *a, *b; cudaStream_t s;
        ~Release() { cudaFreeAsync(a, s); cudaFreeAsync(b, s); }
    } release{Ya, Yb, stream};
    const cplx *cur = y0;
    cplx *next = Ya;
    ZgemmArgs g;
    g.reduce = 0; g.w = nullptr; g.n_split = 1; g.part = nullptr; g.grp_first = g.grp_count = nullptr;
    g.X = g.Y = nullptr; g.c0 = g.c1 = g.c2 = 0.0;
    g.n_units = n_runs;
    int rc;
    for (int it = 0; it < nt; ++it) {
        if (it > 0) {
            g.A = cur; g.B = Lt; g.C = next; g.ib = run_gen_dev;
            g.M = R; g.N = M; g.K = M; g.lda = M; g.ldb = M; g.ldc = M;
            g.sA = (long long)R * M; g.sB = (long long)M * M; g.sC = (long long)R * M;
            g.b_transposed = 0;
            if ((rc = launch_zgemm(g, stream))) return rc;
            cur = next;
            next = (next == Ya) ? Yb : Ya;
        }
        g.A = cur; g.B = S; g.C = out + (size_t)it * save_rows; g.ib = S_stride ? run_save_dev : nullptr;
        g.M = R; g.N = save_rows; g.K = M; g.lda = M; g.ldb = M; g.ldc = nt * save_rows;
        g.sA = (long long)R * M; g.sB = S_stride; g.sC = (long long)R * nt * save_rows;
        g.b_transposed = 1;
        if ((rc = launch_zgemm(g, stream))) return rc;
    }
    return QSX_OK;
}

static int response_contract_impl(const void *x_dev, const void *y_dev, const void *w_dev, int32_t n_units,
                                  const int32_t *grp_first_host, const int32_t *grp_count_host, int32_t n_x,
                                  int64_t n_ab, int32_t n_c, int32_t K, void *s_dev, cudaStream_t stream) {
    QSX_REQUIRE(x_dev && y_dev && s_dev && n_units > 0 && n_ab > 0 && n_c > 0 && K > 0,
                "qsx_response_contract: bad arguments");
    QSX_REQUIRE(n_ab < ((int64_t)1 << 31) / 64, "qsx_response_contract: signal too large");
    ZgemmArgs g;
    g.A = (const cplx *)x_dev; g.B = (const cplx *)y_dev; g.C = (cplx *)s_dev;
    g.M = (int)n_ab; g.N = n_c; g.K = K;
    g.lda = K; g.ldb = K; g.ldc = n_c;
    g.sA = (long long)n_ab * K; g.sB = (long long)n_c * K; g.sC = 0;
    g.n_units = n_units; g.b_transposed = 1; g.reduce = 1; g.w = (const cplx *)w_dev; g.ib = nullptr;
    g.X = g.Y = nullptr; g.c0 = g.c1 = g.c2 = 0.0;
    g.grp_first = g.grp_count = nullptr;
    DevBuf<int> d_first, d_count;
    if (grp_first_host) {
        QSX_REQUIRE(grp_count_host && w_dev, "qsx_response_contract_grouped: group sizes and weights are required");
        for (int u = 0; u < n_units; ++u)
            QSX_REQUIRE(grp_first_host[u] >= 0 && grp_count_host[u] > 0 &&
                        (long long)grp_first_host[u] + grp_count_host[u] <= n_x,
                        "qsx_response_contract_grouped: group %d out of range", u);
        QSX_CUDA(d_first.upload(grp_first_host, n_units, stream));
        QSX_CUDA(d_count.upload(grp_count_host, n_units, stream));
        g.grp_first = d_first.p; g.grp_count = d_count.p;
    }
    // One CTA per output tile walking every unit leaves most SMs idle for a 197 x 5 x 197 signal
    // (217 tiles): deal the units out to enough CTAs for ~16 resident per SM; the partial sums
    // are added in split order, so the result does not depend on scheduling.
    const long long tiles = ((n_ab + BM - 1) / BM) * ((n_c + BN - 1) / BN);
    int n_split = (int)std::min<long long>(std::min<long long>(n_units, 64), (148LL * 16 + tiles - 1) / tiles);
    n_split = std::max(1, n_split);
    DevBuf<cplx> part;
    if (n_split > 1) QSX_CUDA(part.alloc((size_t)n_split * n_ab * n_c));
    g.n_split = n_split; g.part = part.p;
    int rc = launch_zgemm(g, stream);
    if (rc || n_split == 1) return rc;
    const size_t n = (size_t)n_ab * n_c;
    zgemm_add_partials_kernel<<<(int)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, stream>>>(part.p, n_split, n, g.C);
    qsx_launch_counter += 1;
    QSX_CUDA(cudaGetLastError());
    return QSX_OK;
}

extern "C" int qsx_response_contract(const void *x_dev, const void *y_dev, const void *w_dev, int32_t n_units,
                                     int64_t n_ab, int32_t n_c, int32_t K, void *s_dev, void *stream_) {
    return response_contract_impl(x_dev, y_dev, w_dev, n_units, nullptr, nullptr, n_units, n_ab, n_c, K, s_dev,
                                  (cudaStream_t)stream_);
}

extern "C" int qsx_response_contract_grouped(const void *x_dev, int32_t n_x, const void *y_dev, const void *w_dev,
                                             int32_t n_groups, const int32_t *grp_first_host,
                                             const int32_t *grp_count_host, int64_t n_ab, int32_t n_c, int32_t K,
                                             void *s_dev, void *stream_) {
    QSX_REQUIRE(grp_first_host && grp_count_host, "qsx_response_contract_grouped: null group tables");
    return response_contract_impl(x_dev, y_dev, w_dev, n_groups, grp_first_host, grp_count_host, n_x, n_ab, n_c, K,
                                  s_dev, (cudaStream_t)stream_);
}

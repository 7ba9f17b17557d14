// Hermitian-coordinate ("real form") propagation for dense generators that commute with
// Hermitian conjugation -- every physical Liouvillian on a subspace closed under transposition
// ('ee', 'gg,ee', ...; reference dynamics/liouville_space.py:316-341 builds them as complex
// M x M matrices).  With sigma(k) the position of the transposed ket-bra pair of element k,
// such a generator satisfies L[sigma r][sigma c] = conj L[r][c], so in the coordinates
//     u_k = rho_k                      (sigma k = k:  populations)
//     u_a = Re rho_a,  u_b = Im rho_a  (pair a < b = sigma a)
// it is a REAL M x M matrix G = T L T^-1 and a Hermitian state is a real vector: the
// propagator build exp(G dt) costs one real DMMA product per complex one (instead of three) and
// a stepping y <- P y one DFMA per matrix element (instead of four).  The imaginary residual of
// G and of the packed state is measured on the device (defect[0..3]) and checked by the caller.
//
//   real_expm3_kernel       exp(G dt): the degree-12 Paterson-Stockmeyer series of dense.cu on real DMMA;
//                           FUSED: forms G from the complex generator in shared memory first
//   hermitian_form_kernel   Lt (complex, transposed) -> Gt (real, transposed), inf-norms, defect (two-kernel path)
//   real_map_rows_kernel    u <- P u per output step, one warp per column, two whole rows of P per lane (32 < M <= 50)
//   real_map_kernel         the same with four lanes per row and a shuffle reduction (other M)
//   hermitian_pack/unpack   complex state vectors <-> real coordinates
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

struct HermPerm { unsigned char s[64]; };

__device__ __forceinline__ void dmma884r(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__constant__ double inv_fact_r[13] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040,
                                      1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800,
                                      1.0 / 479001600};

// ------------------------------------------------------------------ L -> G = T L T^-1
__global__ void __launch_bounds__(128)
hermitian_form_kernel(const cplx *__restrict__ Lt, int M, HermPerm perm, double *__restrict__ Gt,
                      double *__restrict__ gnorm, double *__restrict__ defect) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *Ls = reinterpret_cast<cplx *>(smem_raw);              // Ls[c*M + r] = L[r][c]
    __shared__ double rowsum[2][64];
    __shared__ double scratch[32];
    const int gen = blockIdx.x, tid = threadIdx.x;
    const cplx *Lg = Lt + (size_t)gen * M * M;
    for (int i = tid; i < M * M; i += blockDim.x) Ls[i] = Lg[i];
    __syncthreads();
    const int R = tid & 63, half = tid >> 6;
    double rs = 0.0, dmax = 0.0, amax = 0.0;
    if (R < M) {
        const int sR = perm.s[R];
        // (T L)[R][c]
        auto TL = [&](int c) -> cplx {
            const cplx x = Ls[c * M + R];
            if (sR == R) return x;
            const cplx y = Ls[c * M + sR];
            if (R < sR) return cmake(0.5 * (x.x + y.x), 0.5 * (x.y + y.y));       // a = R, b = sR
            return cmake(0.5 * (y.y - x.y), -0.5 * (y.x - x.x));                   // a = sR, b = R: (L_a - L_b) / 2i
        };
        double *Gg = Gt + (size_t)gen * M * M;
        for (int C = half; C < M; C += 2) {
            const int sC = perm.s[C];
            cplx z = TL(C);
            if (sC != C) {
                const cplx w = TL(sC);
                if (C < sC) z = cmake(z.x + w.x, z.y + w.y);                      // column a: X_a + X_b
                else z = cmake(-(w.y - z.y), w.x - z.x);                          // column b: i (X_a - X_b), a = sC
            }
            Gg[C * M + R] = z.x;
            rs += fabs(z.x);
            dmax = fmax(dmax, fabs(z.y));
            amax = fmax(amax, fabs(z.x));
        }
    }
    rowsum[half][R] = rs;
    __syncthreads();
    const double nrm = block_max(tid < 64 ? rowsum[0][tid] + rowsum[1][tid] : 0.0, scratch);
    dmax = block_max(dmax, scratch);
    amax = block_max(amax, scratch);
    if (tid == 0) {
        gnorm[gen] = nrm;
        atomic_max_nonneg(&defect[0], dmax);
        atomic_max_nonneg(&defect[1], amax);
    }
}

// ------------------------------------------------------------------ exp(G dt), real DMMA
// exp(G dt) by the degree-12 Paterson-Stockmeyer series of dense.cu on real DMMA: the element-wise Horner operands A and A^2
// stay in the registers of the thread that owns the matching accumulator fragment (A is loaded once
// per member in that ownership, A^2 is the thread's own result of the first product), the planes
// are filled from those registers, and the start of the Horner recursion is written by the epilogue
// of the A^3 product into a third plane.  One global read of the generator per member and no
// scratch tile (a first version re-read A and A^2 from global memory / L2 in every epilogue, like
// dense_expm2_kernel does: 45 % of its stall samples, 0.98 instead of 0.81 ms per 1e4 FMO members).
// Three planes, two CTAs per SM.
// FUSED: the change of coordinates happens here as well -- the member's complex generator is staged
// in shared memory (aliasing two of the planes), every thread forms its fragment of G = T L T^-1 from
// it, the inf-norm and the Hermiticity defect are reduced in the CTA; hermitian_form_kernel and its
// write + re-read of G (0.30 ms per 1e4 FMO members) drop out of the step.
template <int MT, int KS, bool FUSED>
__global__ void __launch_bounds__(32 * MT, 2)
real_expm3_kernel(const double *__restrict__ Gt, const double *__restrict__ gnorm, const cplx *__restrict__ Lt,
                  HermPerm perm, double *__restrict__ defect, int M, double dt, int n_gen,
                  double *__restrict__ P_out, unsigned long long *__restrict__ status) {
    constexpr int MP = 8 * MT;
    constexpr int LD = (MP % 16 == 12) ? MP : ((MP + 3) / 16 * 16 + 12 >= MP ? (MP + 3) / 16 * 16 + 12 : (MP + 3) / 16 * 16 + 28);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *planes = reinterpret_cast<double *>(smem_raw);
    const int lane = threadIdx.x & 31, rb = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r = rb * 8 + g;
    unsigned long long gemms = 0;
    __shared__ unsigned char sperm[64];
    if (FUSED) {
        if (threadIdx.x < 32) {
            sperm[threadIdx.x] = perm.s[threadIdx.x];
            sperm[threadIdx.x + 32] = perm.s[threadIdx.x + 32];
        }
        __syncthreads();
    }

    for (int gen = blockIdx.x; gen < n_gen; gen += gridDim.x) {
        double *X = planes, *Y = X + MP * LD, *Z = Y + MP * LD;
        double A1e[MT][2], A2e[MT][2], a[KS];
        int sq = 0;
        if (FUSED) {
            __shared__ double red[3][8];
            cplx *Ls = reinterpret_cast<cplx *>(Y);         // Ls[c*M + r] = L[r][c]; aliases Y and Z (M*M <= MP*LD)
            const cplx *Lg = Lt + (size_t)gen * M * M;
            __syncthreads();                                // previous member's output pass is done with the planes
            for (int i = threadIdx.x; i < M * M; i += blockDim.x) Ls[i] = Lg[i];
            // the CTA's next member: pull its generator into L2 while this one is being exponentiated
            // (the staging loads above are the only DRAM round trip of a member)
            if (gen + (int)gridDim.x < n_gen) {
                const cplx *Ln = Lg + (size_t)gridDim.x * M * M;
                for (int i = threadIdx.x * 8; i < M * M; i += blockDim.x * 8)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(Ln + i));
            }
            __syncthreads();
            double rsum = 0.0, dmax = 0.0, amax = 0.0;
            const int sR = r < M ? sperm[r] : 0;
            auto TL = [&](int c) -> cplx {                  // (T L)[r][c]
                const cplx x = Ls[c * M + r];
                if (sR == r) return x;
                const cplx y = Ls[c * M + sR];
                if (r < sR) return cmake(0.5 * (x.x + y.x), 0.5 * (x.y + y.y));
                return cmake(0.5 * (y.y - x.y), -0.5 * (y.x - x.x));
            };
#pragma unroll
            for (int nb = 0; nb < MT; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int C = nb * 8 + 2 * t + e;
                    cplx z = cmake(0, 0);
                    if (r < M && C < M) {
                        const int sC = sperm[C];
                        z = TL(C);
                        if (sC != C) {
                            const cplx w = TL(sC);
                            if (C < sC) z = cmake(z.x + w.x, z.y + w.y);
                            else z = cmake(-(w.y - z.y), w.x - z.x);
                        }
                    }
                    A1e[nb][e] = z.x;
                    rsum += fabs(z.x);
                    dmax = fmax(dmax, fabs(z.y));
                    amax = fmax(amax, fabs(z.x));
                }
            rsum += __shfl_xor_sync(0xffffffffu, rsum, 1);  // the four t-lanes of a row
            rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
            rsum = warp_max(rsum);
            dmax = warp_max(dmax);
            amax = warp_max(amax);
            if (lane == 0) { red[0][rb] = rsum; red[1][rb] = dmax; red[2][rb] = amax; }
            __syncthreads();                                // also: every thread is done with Ls
            double nrm = 0.0;
#pragma unroll
            for (int w8 = 0; w8 < MT; ++w8) nrm = fmax(nrm, red[0][w8]);
            if (threadIdx.x == 0) {
                double d = 0.0, am = 0.0;
                for (int w8 = 0; w8 < MT; ++w8) { d = fmax(d, red[1][w8]); am = fmax(am, red[2][w8]); }
                atomic_max_nonneg(&defect[0], d);
                atomic_max_nonneg(&defect[1], am);
            }
            nrm *= fabs(dt);
            while (nrm > 0.5 && sq < 40) { nrm *= 0.5; ++sq; }
            const double scale = dt / (double)(1ULL << sq);
#pragma unroll
            for (int nb = 0; nb < MT; ++nb) { A1e[nb][0] *= scale; A1e[nb][1] *= scale; }
        } else {
            const double *Gg = Gt + (size_t)gen * M * M;
            double nrm = fabs(dt) * gnorm[gen];
            while (nrm > 0.5 && sq < 40) { nrm *= 0.5; ++sq; }
            const double scale = dt / (double)(1ULL << sq);
#pragma unroll
            for (int nb = 0; nb < MT; ++nb) {
                const int c = nb * 8 + 2 * t;
                A1e[nb][0] = (r < M && c < M) ? scale * __ldg(&Gg[c * M + r]) : 0.0;
                A1e[nb][1] = (r < M && c + 1 < M) ? scale * __ldg(&Gg[(c + 1) * M + r]) : 0.0;
            }
            __syncthreads();                                // previous member's output pass is done with the planes
        }
#pragma unroll
        for (int nb = 0; nb < MT; ++nb) {
            *reinterpret_cast<double2 *>(&X[r * LD + nb * 8 + 2 * t]) = make_double2(A1e[nb][0], A1e[nb][1]);
        }
        __syncthreads();
        auto load_fragments = [&](const double *W) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) a[ks] = W[r * LD + ks * 4 + t];
        };
        auto product = [&](const double *B, int nb, double &v0, double &v1) {
            double p0 = 0, p1 = 0, q0 = 0, q1 = 0;
#pragma unroll
            for (int ks = 0; ks < KS; ks += 2) {
                dmma884r(p0, p1, a[ks], B[(ks * 4 + t) * LD + nb * 8 + g]);
                if (ks + 1 < KS) dmma884r(q0, q1, a[ks + 1], B[((ks + 1) * 4 + t) * LD + nb * 8 + g]);
            }
            v0 = p0 + q0;
            v1 = p1 + q1;
        };
        load_fragments(X);                          // left operand: A
        // A^2 -> Y (operand of the next product) and the registers
#pragma unroll
        for (int nb = 0; nb < MT; ++nb) {
            product(X, nb, A2e[nb][0], A2e[nb][1]);
            *reinterpret_cast<double2 *>(&Y[r * LD + nb * 8 + 2 * t]) = make_double2(A2e[nb][0], A2e[nb][1]);
        }
        __syncthreads();
        // A^3 = A A^2 -> X, and the start of the recursion c9 I + c10 A + c11 A^2 + c12 A^3 -> Z
#pragma unroll
        for (int nb = 0; nb < MT; ++nb) {
            const int c = nb * 8 + 2 * t;
            double v0, v1;
            product(Y, nb, v0, v1);
            *reinterpret_cast<double2 *>(&X[r * LD + c]) = make_double2(v0, v1);
            *reinterpret_cast<double2 *>(&Z[r * LD + c]) = make_double2(
                (r == c ? inv_fact_r[9] : 0.0) + inv_fact_r[10] * A1e[nb][0] + inv_fact_r[11] * A2e[nb][0] + inv_fact_r[12] * v0,
                (r == c + 1 ? inv_fact_r[9] : 0.0) + inv_fact_r[10] * A1e[nb][1] + inv_fact_r[11] * A2e[nb][1] + inv_fact_r[12] * v1);
        }
        __syncthreads();
        load_fragments(X);                          // left operand from here on: A^3
        __syncthreads();                            // X may be overwritten from the second Horner product on
        double *P = Z, *U = Y;
#pragma unroll 1
        for (int blk = 2; blk >= 0; --blk) {
            const double c0 = inv_fact_r[3 * blk], c1 = inv_fact_r[3 * blk + 1], c2 = inv_fact_r[3 * blk + 2];
#pragma unroll
            for (int nb = 0; nb < MT; ++nb) {
                const int c = nb * 8 + 2 * t;
                double v0, v1;
                product(P, nb, v0, v1);
                *reinterpret_cast<double2 *>(&U[r * LD + c]) = make_double2(
                    v0 + c1 * A1e[nb][0] + c2 * A2e[nb][0] + (r == c ? c0 : 0.0),
                    v1 + c1 * A1e[nb][1] + c2 * A2e[nb][1] + (r == c + 1 ? c0 : 0.0));
            }
            __syncthreads();
            { double *x = P; P = U; U = x; }
        }
        for (int q = 0; q < sq; ++q) {
            load_fragments(P);
#pragma unroll
            for (int nb = 0; nb < MT; ++nb) {
                double v0, v1;
                product(P, nb, v0, v1);
                *reinterpret_cast<double2 *>(&U[r * LD + nb * 8 + 2 * t]) = make_double2(v0, v1);
            }
            __syncthreads();
            { double *x = P; P = U; U = x; }
        }
        double *Pg = P_out + (size_t)gen * M * M;           // row-major: P[r*M + c]
        for (int i = threadIdx.x; i < M * M; i += blockDim.x) Pg[i] = P[(i / M) * LD + i % M];
        gemms += 5 + sq;
    }
    if (threadIdx.x == 0) atomicAdd(&status[0], gemms);
}

template <int MT, int KS, bool FUSED>
static cudaError_t launch_real_expm3_ks(const double *Gt, const double *gnorm, const cplx *Lt, const HermPerm &perm,
                                        double *defect, int M, double dt, int n_gen, double *P, unsigned long long *status,
                                        cudaStream_t stream) {
    constexpr int MP = 8 * MT;
    constexpr int LD = (MP % 16 == 12) ? MP : ((MP + 3) / 16 * 16 + 12 >= MP ? (MP + 3) / 16 * 16 + 12 : (MP + 3) / 16 * 16 + 28);
    const size_t smem = (size_t)3 * MP * LD * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(real_expm3_kernel<MT, KS, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, real_expm3_kernel<MT, KS, FUSED>, 32 * MT, smem);
    if (e != cudaSuccess) return e;
    const int grid = std::min(n_gen, sms * std::max(1, per_sm));
    real_expm3_kernel<MT, KS, FUSED><<<grid, 32 * MT, smem, stream>>>(Gt, gnorm, Lt, perm, defect, M, dt, n_gen, P, status);
    return cudaGetLastError();
}

template <int MT>
static cudaError_t launch_fused_expm(const cplx *Lt, const HermPerm &perm, double *defect, int M, double dt, int n_gen,
                                     double *P, unsigned long long *status, cudaStream_t stream) {
    if ((M + 3) / 4 == 2 * MT - 1)
        return launch_real_expm3_ks<MT, 2 * MT - 1, true>(nullptr, nullptr, Lt, perm, defect, M, dt, n_gen, P, status, stream);
    return launch_real_expm3_ks<MT, 2 * MT, true>(nullptr, nullptr, Lt, perm, defect, M, dt, n_gen, P, status, stream);
}

template <int MT>
static cudaError_t launch_real_expm(const double *Gt, const double *gnorm, int M, double dt, int n_gen, double *P,
                                    unsigned long long *status, cudaStream_t stream) {
    // the contraction dimension is padded to a multiple of 4 only (M = 49: 13 k-steps, not 14)
    const HermPerm none = {};
    if ((M + 3) / 4 == 2 * MT - 1)
        return launch_real_expm3_ks<MT, 2 * MT - 1, false>(Gt, gnorm, nullptr, none, nullptr, M, dt, n_gen, P, status, stream);
    return launch_real_expm3_ks<MT, 2 * MT, false>(Gt, gnorm, nullptr, none, nullptr, M, dt, n_gen, P, status, stream);
}

// ------------------------------------------------------------------ u <- P u
// One warp per column: lane (rs, q) = (lane / 4, lane % 4) keeps rows rs + 8h (h < RW) and columns
// q + 4i (i < CQ) of P in registers; the state lives in two per-warp shared buffers, the four
// partial sums of a row meet in two shuffles.  No CTA barrier anywhere.
template <int RW, int CQ>
__global__ void __launch_bounds__(64)
real_map_kernel(const double *__restrict__ P, int M, const int *__restrict__ gen_of, int identity, int n_col,
                const double *__restrict__ u0, int nt, int MS, double *__restrict__ out) {
    __shared__ double xs[2][2][64];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 2 + w;
    if (col >= n_col) return;
    const int q = lane & 3, rs = lane >> 2;
    const int gen = gen_of ? gen_of[col] : (identity ? col : 0);
    const double *Pg = P + (size_t)gen * M * M;
    double p[RW][CQ];
#pragma unroll
    for (int h = 0; h < RW; ++h)
#pragma unroll
        for (int i = 0; i < CQ; ++i) {
            const int r = rs + 8 * h, c = q + 4 * i;
            p[h][i] = (r < M && c < M) ? Pg[r * M + c] : 0.0;
        }
    double *orow = out + (size_t)col * nt * MS;
    for (int i = lane; i < 64; i += 32) {
        const double v = i < M ? u0[(size_t)col * MS + i] : 0.0;
        xs[w][0][i] = v;
        xs[w][1][i] = 0.0;
        if (i < MS) __stcs(&orow[i], v);
    }
    // tails of the output rows (row_stride > M), once, outside the stepping loop
    for (int k = M; k < MS; ++k)
        for (int it = 1 + lane; it < nt; it += 32) __stcs(&orow[(size_t)it * MS + k], 0.0);
    __syncwarp();
    // per-lane base addresses; everything inside a step is a constant offset from them
    const double *ld_a = &xs[w][0][q], *ld_b = &xs[w][1][q];
    double *st_a = &xs[w][0][rs], *st_b = &xs[w][1][rs];
    double *og = orow + rs;
    const bool writer = q == 0, last_row = rs + 8 * (RW - 1) < M;
    auto step = [&](const double *__restrict__ ld, double *__restrict__ st) {
        double acc[RW];
#pragma unroll
        for (int h = 0; h < RW; ++h) acc[h] = 0.0;
#pragma unroll
        for (int i = 0; i < CQ; ++i) {
            const double v = ld[4 * i];
#pragma unroll
            for (int h = 0; h < RW; ++h) acc[h] = fma(p[h][i], v, acc[h]);
        }
#pragma unroll
        for (int h = 0; h < RW; ++h) {
            acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], 1);
            acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], 2);
        }
        if (writer) {
#pragma unroll
            for (int h = 0; h < RW - 1; ++h) {
                st[8 * h] = acc[h];
                __stcs(og + 8 * h, acc[h]);
            }
            if (last_row) {
                st[8 * (RW - 1)] = acc[RW - 1];
                __stcs(og + 8 * (RW - 1), acc[RW - 1]);
            }
        }
        __syncwarp();
    };
    int it = 1;
    for (; it + 1 < nt; it += 2) {
        og += MS;
        step(ld_a, st_b);
        og += MS;
        step(ld_b, st_a);
    }
    if (it < nt) {
        og += MS;
        step(ld_a, st_b);
    }
}

// Wide states (32 < M <= 56): one warp per column, lane l keeps the WHOLE rows l and l + 32 of P in
// registers, so a step needs no cross-lane reduction at all -- the shuffle/DADD tree and the
// predicated stores of the kernel above were almost half of its issue slots (91 DFMA at two issue
// cycles each against ~160 other instructions per step; the FP64 pipe cannot be busier than the
// issue slots left to it).  The state is read as 16-byte broadcast loads; two accumulator chains
// per row.  CP = column pairs (zero padded).
template <int CP>
__global__ void __launch_bounds__(64)
real_map_rows_kernel(const double *__restrict__ P, int M, const int *__restrict__ gen_of, int identity, int n_col,
                     const double *__restrict__ u0, int nt, int MS, double *__restrict__ out) {
    __shared__ __align__(16) double xs[2][2][64];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 2 + w;
    if (col >= n_col) return;
    const int gen = gen_of ? gen_of[col] : (identity ? col : 0);
    const double *Pg = P + (size_t)gen * M * M;
    const int r0 = lane, r1 = lane + 32;
    double p0[2 * CP], p1[2 * CP];
#pragma unroll
    for (int c = 0; c < 2 * CP; ++c) {
        p0[c] = (r0 < M && c < M) ? Pg[r0 * M + c] : 0.0;
        p1[c] = (r1 < M && c < M) ? Pg[r1 * M + c] : 0.0;
    }
    double *orow = out + (size_t)col * nt * MS;
    for (int i = lane; i < 64; i += 32) {
        const double v = i < M ? u0[(size_t)col * MS + i] : 0.0;
        xs[w][0][i] = v;
        xs[w][1][i] = 0.0;
        if (i < MS) __stcs(&orow[i], v);
    }
    for (int k = M; k < MS; ++k)
        for (int it = 1 + lane; it < nt; it += 32) __stcs(&orow[(size_t)it * MS + k], 0.0);
    __syncwarp();
    const bool has1 = r1 < M;
    double *og = orow;
    auto step = [&](const double *__restrict__ ld, double *__restrict__ st) {
        // two accumulator chains per row (four chains per lane cover the 8.4-cycle DFMA latency at
        // ~2.9 issue cycles each); more chains only add DADDs to the issue stream
        double a[2] = {0.0, 0.0}, b[2] = {0.0, 0.0};
#pragma unroll
        for (int cp = 0; cp < CP; ++cp) {
            const double2 v = *reinterpret_cast<const double2 *>(ld + 2 * cp);
            a[0] = fma(p0[2 * cp], v.x, a[0]);
            b[0] = fma(p1[2 * cp], v.x, b[0]);
            a[1] = fma(p0[2 * cp + 1], v.y, a[1]);
            b[1] = fma(p1[2 * cp + 1], v.y, b[1]);
        }
        const double s0 = a[0] + a[1], s1 = b[0] + b[1];
        // row l always exists (M > 32); lanes without a second row park it in the unused slot 63
        st[r0] = s0;
        __stcs(og + r0, s0);
        st[has1 ? r1 : 63] = s1;
        if (has1) __stcs(og + r1, s1);
        __syncwarp();
    };
    int it = 1;
    for (; it + 3 < nt; it += 4) {
        og += MS;
        step(xs[w][0], xs[w][1]);
        og += MS;
        step(xs[w][1], xs[w][0]);
        og += MS;
        step(xs[w][0], xs[w][1]);
        og += MS;
        step(xs[w][1], xs[w][0]);
    }
    for (; it < nt; ++it) {
        og += MS;
        if (it & 1) step(xs[w][0], xs[w][1]);
        else step(xs[w][1], xs[w][0]);
    }
}

template <int CP>
static cudaError_t launch_real_map_rows(const double *P, int M, const int *gen_of, int identity, int n_col, const double *u0,
                                        int nt, int MS, double *out, cudaStream_t stream) {
    real_map_rows_kernel<CP><<<(n_col + 1) / 2, 64, 0, stream>>>(P, M, gen_of, identity, n_col, u0, nt, MS, out);
    return cudaGetLastError();
}

template <int RW, int CQ>
static cudaError_t launch_real_map(const double *P, int M, const int *gen_of, int identity, int n_col, const double *u0,
                                   int nt, int MS, double *out, cudaStream_t stream) {
    real_map_kernel<RW, CQ><<<(n_col + 1) / 2, 64, 0, stream>>>(P, M, gen_of, identity, n_col, u0, nt, MS, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ pack / unpack
__global__ void hermitian_pack_kernel(const cplx *__restrict__ y, int M, long long rows, HermPerm perm, int MS,
                                      double *__restrict__ u, double *__restrict__ defect) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double im = 0.0, re = 0.0;
    if (i < rows * MS) {
        const long long row = i / MS;
        const int k = (int)(i % MS);
        cplx z = cmake(0, 0);
        if (k < M) {
            const int s = perm.s[k];
            const cplx a = y[row * M + k];
            if (s == k) z = a;
            else {
                const cplx b = y[row * M + s];
                if (k < s) z = cmake(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
                else z = cmake(0.5 * (b.y - a.y), -0.5 * (b.x - a.x));            // (y_s - y_k) / 2i, s = a of the pair
            }
        }
        u[i] = z.x;
        re = fabs(z.x);
        im = fabs(z.y);
    }
    re = warp_max(re);
    im = warp_max(im);
    if ((threadIdx.x & 31) == 0 && defect) {
        if (im > 0.0) atomic_max_nonneg(&defect[2], im);
        if (re > 0.0) atomic_max_nonneg(&defect[3], re);
    }
}

__global__ void hermitian_unpack_kernel(const double *__restrict__ u, int M, long long rows, int MS, HermPerm perm,
                                        cplx *__restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * M) return;
    const long long row = i / M;
    const int k = (int)(i % M);
    const int s = perm.s[k];
    const double *ur = u + row * MS;
    y[i] = s == k ? cmake(ur[k], 0.0) : k < s ? cmake(ur[k], ur[s]) : cmake(ur[s], -ur[k]);
}

// ------------------------------------------------------------------ host
static int make_perm(HermPerm &p, const int32_t *perm_host, int M) {
    for (int k = 0; k < M; ++k) {
        const int s = perm_host[k];
        if (s < 0 || s >= M || perm_host[s] != k) return 0;          // must be an involution
        p.s[k] = (unsigned char)s;
    }
    return 1;
}

int qsx_real_form_launch(const cplx *Lt, int M, int n_gen, const int32_t *perm_host, double *Gt, double *gnorm,
                         double *defect, cudaStream_t stream) {
    QSX_REQUIRE(Lt && perm_host && Gt && gnorm && defect && M > 0 && M <= 56 && n_gen > 0,
                "qsx_dense_hermitian_form: bad arguments (state dimension 1..56)");
    HermPerm p;
    QSX_REQUIRE(make_perm(p, perm_host, M), "qsx_dense_hermitian_form: perm is not an involution of 0..M-1");
    const size_t smem = (size_t)M * M * sizeof(cplx);
    QSX_CUDA(cudaFuncSetAttribute(hermitian_form_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hermitian_form_kernel<<<n_gen, 128, smem, stream>>>(Lt, M, p, Gt, gnorm, defect);
    qsx_launch_counter += 1;
    QSX_CUDA(cudaGetLastError());
    return QSX_OK;
}

int qsx_fused_expm_launch(const cplx *Lt, int M, int n_gen, const int32_t *perm_host, double dt, double *P,
                          double *defect, unsigned long long *gemm_count, cudaStream_t stream) {
    QSX_REQUIRE(Lt && perm_host && P && defect && gemm_count && M > 0 && M <= 56 && n_gen > 0,
                "qsx_dense_hermitian_expm: bad arguments (state dimension 1..56)");
    HermPerm p;
    QSX_REQUIRE(make_perm(p, perm_host, M), "qsx_dense_hermitian_expm: perm is not an involution of 0..M-1");
    cudaError_t e;
    switch ((M + 7) / 8) {
        case 1: e = launch_fused_expm<1>(Lt, p, defect, M, dt, n_gen, P, gemm_count, stream); break;
        case 2: e = launch_fused_expm<2>(Lt, p, defect, M, dt, n_gen, P, gemm_count, stream); break;
        case 3: e = launch_fused_expm<3>(Lt, p, defect, M, dt, n_gen, P, gemm_count, stream); break;
        case 4: e = launch_fused_expm<4>(Lt, p, defect, M, dt, n_gen, P, gemm_count, stream); break;
        case 5: e = launch_fused_expm<5>(Lt, p, defect, M, dt, n_gen, P, gemm_count, stream); break;
        case 6: e = launch_fused_expm<6>(Lt, p, defect, M, dt, n_gen, P, gemm_count, stream); break;
        default: e = launch_fused_expm<7>(Lt, p, defect, M, dt, n_gen, P, gemm_count, stream); break;
    }
    qsx_launch_counter += 1;
    if (e != cudaSuccess) {
        qsx_set_error("qsx_dense_hermitian_expm: %s", cudaGetErrorString(e));
        return QSX_ERR_CUDA;
    }
    return QSX_OK;
}

extern "C" int qsx_real_expm(const void *Gt_dev, const void *gnorm_dev, int32_t M, int32_t n_generators, double dt,
                             void *P_dev, void *gemm_count_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(Gt_dev && gnorm_dev && P_dev && gemm_count_dev && M > 0 && M <= 56 && n_generators > 0,
                "qsx_real_expm: bad arguments (state dimension 1..56)");
    const double *Gt = (const double *)Gt_dev, *gn = (const double *)gnorm_dev;
    double *P = (double *)P_dev;
    unsigned long long *st = (unsigned long long *)gemm_count_dev;
    cudaError_t e;
    switch ((M + 7) / 8) {
        case 1: e = launch_real_expm<1>(Gt, gn, M, dt, n_generators, P, st, stream); break;
        case 2: e = launch_real_expm<2>(Gt, gn, M, dt, n_generators, P, st, stream); break;
        case 3: e = launch_real_expm<3>(Gt, gn, M, dt, n_generators, P, st, stream); break;
        case 4: e = launch_real_expm<4>(Gt, gn, M, dt, n_generators, P, st, stream); break;
        case 5: e = launch_real_expm<5>(Gt, gn, M, dt, n_generators, P, st, stream); break;
        case 6: e = launch_real_expm<6>(Gt, gn, M, dt, n_generators, P, st, stream); break;
        default: e = launch_real_expm<7>(Gt, gn, M, dt, n_generators, P, st, stream); break;
    }
    qsx_launch_counter += 1;
    if (e != cudaSuccess) {
        qsx_set_error("qsx_real_expm: %s", cudaGetErrorString(e));
        return QSX_ERR_CUDA;
    }
    return QSX_OK;
}

extern "C" int qsx_real_map(const void *P_dev, int32_t M, int32_t n_generators, const int32_t *generator_of_column_host,
                            int32_t n_columns, const void *u0_dev, int32_t n_times, int32_t row_stride, void *out_dev,
                            void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(P_dev && u0_dev && out_dev && M > 0 && M <= 56 && n_generators > 0 && n_columns > 0 && n_times > 0 &&
                    row_stride >= M && row_stride <= 64,
                "qsx_real_map: bad arguments");
    DevBuf<int> gen;
    int identity = 0;
    if (generator_of_column_host) {
        std::vector<int> g(generator_of_column_host, generator_of_column_host + n_columns);
        for (int v : g) QSX_REQUIRE(v >= 0 && v < n_generators, "qsx_real_map: generator index out of range");
        QSX_CUDA(gen.upload(g, stream));
    } else {
        identity = n_generators == n_columns;
    }
    const double *P = (const double *)P_dev, *u0 = (const double *)u0_dev;
    double *out = (double *)out_dev;
    const int rw = (M + 7) / 8;
    cudaError_t e;
#define QSX_RMAP(RW, CQ) e = launch_real_map<RW, CQ>(P, M, gen.p, identity, n_columns, u0, n_times, row_stride, out, stream)
    switch (rw) {
        case 1: QSX_RMAP(1, 2); break;
        case 2: QSX_RMAP(2, 4); break;
        case 3: QSX_RMAP(3, 6); break;
        case 4: QSX_RMAP(4, 8); break;
        case 5:
            if (getenv("QSX_RMAP_SHUFFLE")) QSX_RMAP(5, 10);
            else e = launch_real_map_rows<20>(P, M, gen.p, identity, n_columns, u0, n_times, row_stride, out, stream);
            break;
        case 6:
            if (getenv("QSX_RMAP_SHUFFLE")) QSX_RMAP(6, 12);
            else e = launch_real_map_rows<24>(P, M, gen.p, identity, n_columns, u0, n_times, row_stride, out, stream);
            break;
        default:
            // (two warps per column -- half the registers, twice the resident warps -- measured slower:
            // 1.0-1.2 vs 0.8 ms per 1e4 FMO members; the per-step overhead doubles)
            // QSX_RMAP_SHUFFLE=1: the four-lanes-per-row kernel for wide states too (A/B runs)
            // (M > 50: two whole rows per lane do not fit the register file without spills)
            if (M <= 50 && !getenv("QSX_RMAP_SHUFFLE")) {
                e = launch_real_map_rows<25>(P, M, gen.p, identity, n_columns, u0, n_times, row_stride, out, stream);
            } else if (M <= 52) {
                QSX_RMAP(7, 13);
            } else {
                QSX_RMAP(7, 14);
            }
            break;
    }
#undef QSX_RMAP
    qsx_launch_counter += 1;
    if (e != cudaSuccess) {
        qsx_set_error("qsx_real_map: %s", cudaGetErrorString(e));
        return QSX_ERR_CUDA;
    }
    return QSX_OK;
}

extern "C" int qsx_hermitian_pack(const void *y_dev, int32_t M, int64_t rows, const int32_t *perm_host, int32_t row_stride,
                                  void *u_dev, void *defect_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(y_dev && u_dev && perm_host && M > 0 && M <= 64 && rows > 0 && row_stride >= M, "qsx_hermitian_pack: bad arguments");
    HermPerm p;
    QSX_REQUIRE(make_perm(p, perm_host, M), "qsx_hermitian_pack: perm is not an involution of 0..M-1");
    const long long n = (long long)rows * row_stride;
    hermitian_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const cplx *)y_dev, M, rows, p, row_stride,
                                                                           (double *)u_dev, (double *)defect_dev);
    qsx_launch_counter += 1;
    QSX_CUDA(cudaGetLastError());
    return QSX_OK;
}

extern "C" int qsx_hermitian_unpack(const void *u_dev, int32_t M, int64_t rows, int32_t row_stride, const int32_t *perm_host,
                                    void *y_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(y_dev && u_dev && perm_host && M > 0 && M <= 64 && rows > 0 && row_stride >= M, "qsx_hermitian_unpack: bad arguments");
    HermPerm p;
    QSX_REQUIRE(make_perm(p, perm_host, M), "qsx_hermitian_unpack: perm is not an involution of 0..M-1");
    const long long n = (long long)rows * M;
    hermitian_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const double *)u_dev, M, rows, row_stride, p,
                                                                             (cplx *)y_dev);
    qsx_launch_counter += 1;
    QSX_CUDA(cudaGetLastError());
    return QSX_OK;
}

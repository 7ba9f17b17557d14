// K5: batched construction of Redfield Liouvillians for a disorder ensemble.
//
// Replaces, per ensemble member, the reference chain
//   redfield_evolve -> redfield_dissipator -> redfield_tensor
// (dynamics/redfield.py:9-104) including the 1000-term Matsubara sum of
// DebyeBath.corr_func_complex (bath.py:84-102) -- 95 % of the reference's
// disorder-ensemble wall time.  One thread block builds one
// member:
//   C[i,j]   = corr(E_i - E_j)                                  (bath.py)
//   K_n      = U^+ V_n U            (V_n diagonal, hamiltonian.py:593-608)
//   Gs[a,c]  = sum_{b,n} K_n[a,b] K_n[b,c] C[c,b]
//   R[abcd]  = conj(d_ac Gs[b,d]) + d_bd Gs[a,c] - conj(G[c,a,b,d]) - G[d,b,a,c],
//              G[a,b,c,d] = sum_n K_n[a,b] K_n[c,d] C[d,c]      (redfield.py:61-69)
//   L        = -i (E_a - E_b) d_ac d_bd - R (.) secular mask    (redfield.py:71-98)
//   L_site   = W^+ L W, W = kron(U^+, U^+)                      (redfield.py:99-100)
// and writes unit_convert * L[idx, idx] for the requested Liouville subspace.
//
// Only the part of the tensor the subspace needs is built: with A the ket
// states and B the bra states the subspace index touches, T[a][b][c][d] is
// formed for a, c in A and b, d in B only, over the Hilbert space A u B --
// provided H couples neither set to its complement, so that the eigenvectors
// (and with them K_n, Gs and the site-basis transform) stay inside the sets.
// For 'fe' of FMO (28 x 7 states) that is 38 416 tensor entries instead of
// 36^4 = 1.7 M.  The host checks the block structure and falls back to the
// full range when it does not hold.
#include "common.cuh"
#include <algorithm>

struct RedfieldBuildArgs {
    int m, N, nb, M;
    int na, nbb;            // ket / bra state counts of the tensor block
    const int *ra, *rb;     // [na], [nbb] their positions in the N-state space
    const int *oa, *ob;     // [M] ket / bra slot (into ra / rb) of each subspace element
    const double *E;        // [m][N]
    const cplx *U;          // [m][N][N] row-major, U[x][a] = <x|a>
    const double *v;        // [nb][N] coupling diagonals
    qsx_bath bath;
    int n_mats;             // Matsubara terms tabulated in shared memory (0 for the real spectrum)
    int secular, eigen_basis;
    double unit_convert;
    cplx *L;                // [m][M][M]
    cplx *scratch;          // [gridDim][2][N^4] when the tensors do not fit in shared memory
    int tensors_in_smem;
    int in_place;           // ket/bra blocks of at most 8 states: site-basis transform in one buffer
    // on-device eigensystems (jacobi != 0): H_m = H0 + diag(sum_j shift[m][j] v[j][.]), lab frame
    int transposed_out;     // write Lt[m][c][r] (storage of qsx_dense_wrap) instead of L[m][r][c]
    int jacobi;
    int u_real;             // the eigenvectors are real (device eigensystems of a real symmetric H)
    const double *H0;       // [N][N] real symmetric
    const double *shifts;   // [m][nb]
    const double *quanta;   // [N] excitation number of each basis state (rotating-frame shift)
    double rw_freq;
};

// Cyclic Jacobi eigensolver for one real symmetric N x N matrix held in shared
// memory (A is overwritten, eigenvalues end on its diagonal in the original
// basis-state order -- no sorting, so block-diagonal manifolds stay in place;
// V receives the eigenvectors as columns).  Round-robin pairing gives N/2
// independent rotations per round.  Called by the whole thread block.
template <bool WARP>
__device__ void jacobi_eigh(double *A, double *V, int N, double *work) {
    // WARP: executed by one warp on its own (small matrices): barriers become __syncwarp
    const int tid = WARP ? (threadIdx.x & 31) : threadIdx.x, nthr = WARP ? 32 : blockDim.x;
    auto sync = [] { if (WARP) __syncwarp(); else __syncthreads(); };
    const int Nn = (N + 1) & ~1, half = Nn / 2;
    double *cs = work, *sn = work + half;
    int *pp = reinterpret_cast<int *>(work + 2 * half), *qq = pp + half;
    for (int i = tid; i < N * N; i += nthr) V[i] = (i / N == i % N) ? 1.0 : 0.0;
    sync();
    for (int sweep = 0; sweep < 16; ++sweep) {
        // convergence: all off-diagonal entries negligible against the diagonal scale
        int big = 0;
        for (int i = tid; i < N * N; i += nthr) {
            int p = i / N, q = i % N;
            if (p < q && fabs(A[i]) > 1e-17 * (fabs(A[p * N + p]) + fabs(A[q * N + q])) && A[i] != 0.0) big = 1;
        }
        if (!(WARP ? __any_sync(0xffffffffu, big) : __syncthreads_or(big))) break;
        for (int r = 0; r < Nn - 1; ++r) {
            if (tid < half) {
                int a = (tid == 0) ? Nn - 1 : (r + tid) % (Nn - 1);
                int b = (r + Nn - 1 - tid) % (Nn - 1);
                int p = min(a, b), q = max(a, b);
                double c = 1.0, s = 0.0;
                if (q < N) {
                    double apq = A[p * N + q];
                    if (apq != 0.0) {
                        double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
                        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                    }
                } else {
                    p = q = -1;       // pairing with the padding slot
                }
                cs[tid] = c; sn[tid] = s; pp[tid] = p; qq[tid] = q;
            }
            sync();
            for (int i = tid; i < half * N; i += nthr) {      // A <- A P, V <- V P
                int pr = i / N, k = i % N;
                int p = pp[pr], q = qq[pr];
                if (p < 0) continue;
                double c = cs[pr], s = sn[pr];
                double ap = A[k * N + p], aq = A[k * N + q];
                A[k * N + p] = c * ap - s * aq;
                A[k * N + q] = s * ap + c * aq;
                double vp = V[k * N + p], vq = V[k * N + q];
                V[k * N + p] = c * vp - s * vq;
                V[k * N + q] = s * vp + c * vq;
            }
            sync();
            for (int i = tid; i < half * N; i += nthr) {      // A <- P^T A
                int pr = i / N, k = i % N;
                int p = pp[pr], q = qq[pr];
                if (p < 0) continue;
                double c = cs[pr], s = sn[pr];
                double ap = A[p * N + k], aq = A[q * N + k];
                A[p * N + k] = c * ap - s * aq;
                A[q * N + k] = s * ap + c * aq;
            }
            sync();
        }
    }
}


__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    double d = b.x * b.x + b.y * b.y;
    return cmake((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

// One-sided correlation spectrum C(x) of the Debye bath (bath.py:17-31, 79-102).  The
// complex form needs the 1000-term Matsubara sum
//   S(x) = sum_m nu_m / ((nu_m^2 - g^2) (nu_m - i x)),  nu_m = 2 pi m T,
// whose x-independent factors live in shared tables (one division per term is left) and
// which satisfies S(-x) = conj(S(x)): a warp sums once per pair of eigenstates.
__device__ __forceinline__ cplx corr_real(const qsx_bath &b, double x) {
    const double T = b.temperature, lam = b.reorg_energy, g = b.cutoff_freq;
    // (n(x)+1) J_anti(x);  T J'(0) at x == 0
    if (x == 0.0) return cmake(T * 2.0 * lam / g, 0.0);
    double ax = fabs(x);
    double J = 2.0 * lam * g * ax / (g * g + ax * ax);
    if (x < 0) J = -J;
    return cmake((1.0 / expm1(x / T) + 1.0) * J, 0.0);
}
__device__ __forceinline__ cplx corr_complex_finish(const qsx_bath &b, double x, double sr, double si) {
    const double T = b.temperature, lam = b.reorg_energy, g = b.cutoff_freq;
    if (x == 0.0) return cmake(lam * 2.0 * T / g, -lam);
    cplx drude = cdiv(cmake(1.0 / tan(g / (2.0 * T)), -1.0), cmake(g, -x));
    return cmake(lam * g * (drude.x + 4.0 * T * sr), lam * g * (drude.y + 4.0 * T * si));
}
// Terms beyond `split` are summed as a power series in x^2 whose coefficients do not depend on x,
//   sum_{m >= split} pre_m (nu_m + i x) / (nu_m^2 + x^2) = sum_k (-x^2)^k (A_k + i x B_k),
//   A_k = sum_m pre_m nu_m^(-1-2k),  B_k = sum_m pre_m nu_m^(-2-2k)        (tail[k], tail[8 + k]),
// tabulated once per CTA: for x^2 <= x2_max = nu_split^2 / 100 eight coefficients leave a relative
// remainder of 1e-16, and a pair of eigenstates costs `split` divisions instead of `cutoff`
// (the 21 x 1000 divisions per FMO member were a quarter of the build kernel).  split = 0 or a larger
// |x|: the plain sum.
#define QSX_MATS_TAIL 8
__device__ __forceinline__ void matsubara_sum(const double *m_nu, const double *m_pre, int cutoff, double x,
                                              const double *tail, int split, double x2_max,
                                              double &sr, double &si) {
    const int lane = threadIdx.x & 31;
    double ar = 0.0, ai = 0.0;
    const double x2 = x * x;
    const bool series = split > 0 && x2 <= x2_max;
    const int stop = series ? split : cutoff;
    for (int mth = lane; mth < stop; mth += 32) {
        const double nu = m_nu[mth], pre = m_pre[mth];
        const double q = pre / (nu * nu + x2);
        ar = fma(q, nu, ar);
        ai = fma(q, x, ai);
    }
    sr = warp_sum(ar);
    si = warp_sum(ai);
    if (series) {
        double pa = 0.0, pb = 0.0;
#pragma unroll
        for (int k = QSX_MATS_TAIL - 1; k >= 0; --k) {
            pa = fma(pa, -x2, tail[k]);
            pb = fma(pb, -x2, tail[QSX_MATS_TAIL + k]);
        }
        sr += pa;
        si = fma(x, pb, si);
    }
}

// Eigensystems of small member Hamiltonians, one WARP per member (N <= 16): the cyclic Jacobi
// sweeps are a chain of dependent rotations that keeps one warp busy; run inside the build
// kernel they left the other seven warps of the member's CTA at a barrier (45 % of the warp
// time in the ncu capture).  Here every warp works on its own member.
__global__ void __launch_bounds__(256) redfield_eig_kernel(int m, int N, int nb, const double *__restrict__ H0,
                                                           const double *__restrict__ shifts,
                                                           const double *__restrict__ v,
                                                           const double *__restrict__ quanta, double rw_freq,
                                                           double *__restrict__ E_out, cplx *__restrict__ U_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int N2 = N * N, per = 2 * N2 + 2 * (((N + 1) & ~1) + 2);
    double *A = reinterpret_cast<double *>(smem_raw) + (size_t)warp * per, *V = A + N2, *work = V + N2;
    for (int mem = blockIdx.x * nwarp + warp; mem < m; mem += gridDim.x * nwarp) {
        for (int i = lane; i < N2; i += 32) {
            const int r = i / N, c = i % N;
            double h = H0[i];
            if (r == c)
                for (int j = 0; j < nb; ++j) h += shifts[(size_t)mem * nb + j] * v[j * N + r];
            A[i] = h;
        }
        __syncwarp();
        jacobi_eigh<true>(A, V, N, work);
        __syncwarp();
        for (int i = lane; i < N2; i += 32) U_out[(size_t)mem * N2 + i] = cmake(V[i], 0.0);
        for (int i = lane; i < N; i += 32) E_out[(size_t)mem * N + i] = A[i * N + i] - quanta[i] * rw_freq;
        __syncwarp();
    }
}

// NC > 0: compile-time state count for full blocks (N = na = nbb = nb = NC, e.g. FMO 'ee'):
// the index arithmetic of the tensor loops -- a fifth of the kernel's instructions with
// run-time divisors -- becomes multiplications by constants and the state loops unroll.
// (NC = 7: 352 threads -- the 343 lines of a transform pass are one round instead of one full and one
// quarter-full round of 256 threads, the 21 eigenstate pairs two rounds of 11 warps instead of three of 8)
template <int NC>
__global__ void __launch_bounds__(NC == 7 ? 352 : 256, NC == 7 ? 2 : 3) redfield_build_kernel(RedfieldBuildArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = NC ? NC : a.N, nb = NC ? NC : a.nb, N2 = N * N;
    const int na = NC ? NC : a.na, nbb = NC ? NC : a.nbb;
    const size_t N4 = (size_t)na * nbb * na * nbb;       // entries of the tensor block
    cplx *Us = reinterpret_cast<cplx *>(smem_raw);       // [N][N]
    cplx *Cs = Us + N2;                                  // [N][N]
    cplx *Gs = Cs + N2;                                  // [N][N]
    cplx *Ks = Gs + N2;                                  // [nb][N][N]
    double *Es = reinterpret_cast<double *>(Ks + (size_t)nb * N2);   // [N]
    const int NE = (N + 1) & ~1;
    double *m_nu = Es + NE, *m_pre = m_nu + a.n_mats;                // [n_mats] Matsubara tables
    cplx *TA, *TB;
    if (a.tensors_in_smem) {
        TA = reinterpret_cast<cplx *>(m_pre + a.n_mats);
        TB = a.in_place ? TA : TA + N4;
    } else {
        TA = a.scratch + (size_t)blockIdx.x * 2 * N4;
        TB = TA + N4;
    }
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, nwarp = nthr >> 5;

    for (int mth = tid; mth < a.n_mats; mth += nthr) {
        const double nu = 2.0 * M_PI * mth * a.bath.temperature, g = a.bath.cutoff_freq;
        m_nu[mth] = nu;
        m_pre[mth] = mth < a.bath.matsubara_cutoff ? nu / (nu * nu - g * g) : 0.0;
    }
    // tail coefficients of the Matsubara sums (see matsubara_sum)
    __shared__ double m_tail[2 * QSX_MATS_TAIL];
    const int m_split = (a.n_mats > 128 && nwarp >= QSX_MATS_TAIL) ? 64 : 0;
    double x2_max = 0.0;
    if (m_split) {
        __syncthreads();
        if (warp < QSX_MATS_TAIL) {
            double sa = 0.0, sb = 0.0;
            for (int mth = m_split + (tid & 31); mth < a.n_mats; mth += 32) {
                const double inv = 1.0 / m_nu[mth], inv2 = inv * inv;
                double pw = m_pre[mth] * inv;
                for (int j = 0; j < warp; ++j) pw *= inv2;
                sa += pw;
                sb = fma(pw, inv, sb);
            }
            sa = warp_sum(sa);
            sb = warp_sum(sb);
            if ((tid & 31) == 0) { m_tail[warp] = sa; m_tail[QSX_MATS_TAIL + warp] = sb; }
        }
        x2_max = 0.01 * m_nu[m_split] * m_nu[m_split];
    }
    for (int mem = blockIdx.x; mem < a.m; mem += gridDim.x) {
        __syncthreads();
        if (a.jacobi) {
            // workspaces: Cs (complex N^2) holds A and V as doubles, Gs the rotation scratch
            double *A = reinterpret_cast<double *>(Cs), *V = A + N2;
            double *work = reinterpret_cast<double *>(Gs);
            for (int i = tid; i < N2; i += nthr) {
                int r = i / N, c = i % N;
                double h = a.H0[i];
                if (r == c)
                    for (int j = 0; j < nb; ++j) h += a.shifts[(size_t)mem * nb + j] * a.v[j * N + r];
                A[i] = h;
            }
            __syncthreads();
            if (N <= 16) {
                if (warp == 0) jacobi_eigh<true>(A, V, N, work);
                __syncthreads();
            } else {
                jacobi_eigh<false>(A, V, N, work);
            }
            for (int i = tid; i < N2; i += nthr) Us[i] = cmake(V[i], 0.0);
            for (int i = tid; i < N; i += nthr) Es[i] = A[i * N + i] - a.quanta[i] * a.rw_freq;
        } else {
            for (int i = tid; i < N2; i += nthr) Us[i] = a.U[(size_t)mem * N2 + i];
            for (int i = tid; i < N; i += nthr) Es[i] = a.E[(size_t)mem * N + i];
        }
        __syncthreads();
        // correlation matrix: one warp per pair of eigenstates (both orderings), diagonal apart
        if (a.bath.kind == QSX_BATH_DEBYE_REAL) {
            for (int p = tid; p < N2; p += nthr) Cs[p] = corr_real(a.bath, Es[p / N] - Es[p % N]);
        } else {
            for (int i = tid; i < N; i += nthr) Cs[i * N + i] = corr_complex_finish(a.bath, 0.0, 0.0, 0.0);
            for (int p = warp; p < N * (N - 1) / 2; p += nwarp) {
                int i = 0, rem = p;
                while (rem >= N - 1 - i) { rem -= N - 1 - i; ++i; }
                const int j = i + 1 + rem;
                const double x = Es[i] - Es[j];
                double sr, si;
                matsubara_sum(m_nu, m_pre, a.n_mats, x, m_tail, m_split, x2_max, sr, si);
                if ((tid & 31) == 0) {
                    Cs[i * N + j] = corr_complex_finish(a.bath, x, sr, si);
                    Cs[j * N + i] = corr_complex_finish(a.bath, -x, sr, -si);
                }
            }
        }
        // couplings in the eigenbasis: K_n[a][b] = sum_x conj(U[x][a]) v_n[x] U[x][b]
        for (int p = tid; p < nb * N2; p += nthr) {
            int n = p / N2, ab = p % N2, aa = ab / N, bb = ab % N;
            cplx acc = cmake(0, 0);
            for (int x = 0; x < N; ++x) {
                cplx ua = Us[x * N + aa];
                ua.y = -ua.y;
                cfma(acc, cscale(a.v[n * N + x], ua), Us[x * N + bb]);
            }
            Ks[p] = acc;
        }
        __syncthreads();
        // Gs[a][c] = sum_b sum_n K_n[a][b] K_n[b][c] C[c][b]
        if ((size_t)N2 * N <= N4) {
            // one thread per (a, c, b) term, staged in the (still unused) tensor buffer: with one
            // thread per (a, c) only N^2 of the CTA's threads worked here (FMO: 49 of 256 on a
            // chain of ~1500 dependent FMAs -- 17 % of the kernel's stall samples)
            for (int p = tid; p < N2 * N; p += nthr) {
                const int bb = p % N, ac = p / N, aa = ac / N, cc = ac % N;
                cplx kk = cmake(0, 0);
                for (int n = 0; n < nb; ++n) cfma(kk, Ks[n * N2 + aa * N + bb], Ks[n * N2 + bb * N + cc]);
                TA[p] = cmul(kk, Cs[cc * N + bb]);
            }
            __syncthreads();
            for (int p = tid; p < N2; p += nthr) {
                cplx acc = cmake(0, 0);
                for (int bb = 0; bb < N; ++bb) acc = cadd(acc, TA[p * N + bb]);
                Gs[p] = acc;
            }
        } else {
            for (int p = tid; p < N2; p += nthr) {
                int aa = p / N, cc = p % N;
                cplx acc = cmake(0, 0);
                for (int bb = 0; bb < N; ++bb) {
                    cplx kk = cmake(0, 0);
                    for (int n = 0; n < nb; ++n) cfma(kk, Ks[n * N2 + aa * N + bb], Ks[n * N2 + bb * N + cc]);
                    cfma(acc, kk, Cs[cc * N + bb]);
                }
                Gs[p] = acc;
            }
        }
        __syncthreads();
        // eigenbasis generator as a 4-index tensor T[a][b][c][d] = L[a + N b, c + N d],
        // a, c over the ket states ra[], b, d over the bra states rb[]
        const int n4 = (int)N4;                  // N <= 64: fits 32 bits (cheap index arithmetic)
        for (int p = tid; p < n4; p += nthr) {
            // full blocks (NC > 0): the ket / bra state lists are the identity
            const int dd = NC ? p % NC : a.rb[p % nbb], cc = NC ? (p / NC) % NC : a.ra[(p / nbb) % na];
            const int bb = NC ? (p / (NC * NC)) % NC : a.rb[(p / (nbb * na)) % nbb];
            const int aa = NC ? p / (NC * NC * NC) : a.ra[p / (nbb * na * nbb)];
            // G[c,a,b,d] = sum_n K_n[c][a] K_n[b][d] and G[d,b,a,c] = sum_n K_n[d][b] K_n[a][c]: the
            // couplings are Hermitian in the eigenbasis (real diagonal system-bath operators), so the
            // second sum is the complex conjugate of the first
            cplx g1 = cmake(0, 0);
            for (int n = 0; n < nb; ++n) cfma(g1, Ks[n * N2 + cc * N + aa], Ks[n * N2 + bb * N + dd]);
            cplx g2 = cmul(cmake(g1.x, -g1.y), Cs[cc * N + aa]);
            g1 = cmul(g1, Cs[dd * N + bb]);
            cplx R = cmake(-g1.x - g2.x, g1.y - g2.y);     // - conj(g1) - g2
            if (aa == cc) { cplx s = Gs[bb * N + dd]; R.x += s.x; R.y -= s.y; }
            if (bb == dd) { cplx s = Gs[aa * N + cc]; R.x += s.x; R.y += s.y; }
            if (a.secular && !((aa == bb && cc == dd) || (aa == cc && bb == dd))) R = cmake(0, 0);
            cplx L = cmake(-R.x, -R.y);
            if (aa == cc && bb == dd) L.y -= (Es[aa] - Es[bb]);
            TA[p] = L;
        }
        __syncthreads();
        cplx *src = TA, *dst = TB;
        if (!a.eigen_basis) {
            // L_site[i,j,k,l] = sum U[i,p] U[j,q] T[p,q,r,s] conj(U[k,r]) conj(U[l,s])
            for (int pos = 0; pos < 4; ++pos) {
                const int stride = pos == 0 ? nbb * na * nbb : pos == 1 ? na * nbb : pos == 2 ? nbb : 1;
                const int n = (pos & 1) ? nbb : na;
                const int *states = (pos & 1) ? a.rb : a.ra;
                if (a.in_place) {
                    // small blocks (n <= 8): a thread owns a whole line of the tensor along the
                    // transformed index, so the pass runs in place and the second tensor
                    // buffer is not needed (twice as many members resident per SM)
                    for (int l = tid; l < n4 / n; l += nthr) {
                        const int base = (l / stride) * (n * stride) + (l % stride);
                        cplx v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = q < n ? src[base + q * stride] : cmake(0, 0);
                        for (int i = 0; i < n; ++i) {
                            const cplx *urow = Us + (NC ? i : states[i]) * N;
                            cplx acc = cmake(0, 0);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                if (q < n) {
                                    cplx u = urow[NC ? q : states[q]];
                                    if (a.u_real) {         // half the multiply-adds
                                        rfma(acc, u.x, v[q]);
                                    } else {
                                        if (pos >= 2) u.y = -u.y;
                                        cfma(acc, u, v[q]);
                                    }
                                }
                            }
                            src[base + i * stride] = acc;
                        }
                    }
                    __syncthreads();
                    continue;
                }
                for (int p = tid; p < n4; p += nthr) {
                    const int i = (p / stride) % n;
                    const int base = p - i * stride;
                    const cplx *urow = Us + states[i] * N;
                    cplx acc = cmake(0, 0);
                    for (int q = 0; q < n; ++q) {
                        cplx u = urow[states[q]];
                        if (pos >= 2) u.y = -u.y;
                        cfma(acc, u, src[base + q * stride]);
                    }
                    dst[p] = acc;
                }
                __syncthreads();
                cplx *tmp = src; src = dst; dst = tmp;
            }
        }
        // restricted, scaled output
        cplx *Lout = a.L + (size_t)mem * a.M * a.M;
        // p runs over the OUTPUT positions (coalesced stores whichever storage order is asked for);
        // the whole operator space of a full block needs no slot tables: element i = ket i % N, bra i / N
        const bool whole = NC && a.M == NC * NC;
        for (int p = tid; p < a.M * a.M; p += nthr) {
            const int hi = p / a.M, lo = p % a.M;
            const int r = a.transposed_out ? lo : hi, c = a.transposed_out ? hi : lo;
            const int ar = whole ? r % (NC ? NC : 1) : a.oa[r], br = whole ? r / (NC ? NC : 1) : a.ob[r];
            const int ac = whole ? c % (NC ? NC : 1) : a.oa[c], bc = whole ? c / (NC ? NC : 1) : a.ob[c];
            const cplx val = src[(((size_t)ar * nbb + br) * na + ac) * nbb + bc];
            Lout[p] = cscale(a.unit_convert, val);
        }
    }
}

static int redfield_build_impl(int32_t n_members, int32_t N, const void *E_dev, const void *U_dev,
                               const double *H0_host, const double *shifts_dev, const double *quanta_host,
                               double rw_freq, int32_t n_baths, const double *coupling_diag_host,
                               const qsx_bath *bath, int32_t secular, int32_t eigen_basis,
                               double unit_convert, int32_t M, const int64_t *subspace_index_host,
                               int32_t transposed_out, void *L_out_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool jacobi = H0_host != nullptr;
    QSX_REQUIRE(n_members > 0 && N > 0 && n_baths > 0 && M > 0 && bath &&
                coupling_diag_host && subspace_index_host && L_out_dev &&
                (jacobi ? (shifts_dev && quanta_host) : (E_dev && U_dev)),
                "qsx_redfield_build: bad arguments");
    QSX_REQUIRE(bath->kind == QSX_BATH_DEBYE_COMPLEX || bath->kind == QSX_BATH_DEBYE_REAL,
                "qsx_redfield_build: unknown bath kind %d", bath->kind);
    QSX_REQUIRE(N <= 64, "qsx_redfield_build: Hilbert dimension %d too large", N);
    const int Nfull = N;
    std::vector<char> in_a(Nfull, 0), in_b(Nfull, 0);
    for (int i = 0; i < M; ++i) {
        QSX_REQUIRE(subspace_index_host[i] >= 0 && subspace_index_host[i] < (int64_t)N * N,
                    "subspace index out of range");
        in_a[subspace_index_host[i] % Nfull] = 1;      // ket state (column-major vec)
        in_b[subspace_index_host[i] / Nfull] = 1;      // bra state
    }
    // the tensor block may be restricted to the ket set A and bra set B only when H
    // does not couple a set to its complement; otherwise widen to A u B, then to all
    auto closed = [&](const std::vector<char> &set) {
        if (!jacobi) return false;                     // eigenvectors supplied: structure unknown
        for (int x = 0; x < Nfull; ++x)
            for (int y = 0; y < Nfull; ++y)
                if (set[x] && !set[y] &&
                    (H0_host[(size_t)x * Nfull + y] != 0.0 || H0_host[(size_t)y * Nfull + x] != 0.0))
                    return false;
        return true;
    };
    if (!closed(in_a) || !closed(in_b)) {
        for (int x = 0; x < Nfull; ++x) in_a[x] = in_b[x] = (char)(in_a[x] | in_b[x]);
        if (!closed(in_a)) std::fill(in_a.begin(), in_a.end(), (char)1), in_b = in_a;
    }
    // Hilbert space of the build: S = A u B, compressed
    std::vector<int> slot(Nfull, -1), states, ra, rb;
    for (int x = 0; x < Nfull; ++x)
        if (in_a[x] || in_b[x]) { slot[x] = (int)states.size(); states.push_back(x); }
    std::vector<int> slot_a(Nfull, -1), slot_b(Nfull, -1);
    for (int x = 0; x < Nfull; ++x) {
        if (in_a[x]) { slot_a[x] = (int)ra.size(); ra.push_back(slot[x]); }
        if (in_b[x]) { slot_b[x] = (int)rb.size(); rb.push_back(slot[x]); }
    }
    std::vector<int> oa(M), ob(M);
    for (int i = 0; i < M; ++i) {
        oa[i] = slot_a[subspace_index_host[i] % Nfull];
        ob[i] = slot_b[subspace_index_host[i] / Nfull];
    }
    N = (int)states.size();
    const int na = (int)ra.size(), nbb = (int)rb.size();
    std::vector<double> v_s((size_t)n_baths * N), H0_s, quanta_s;
    for (int j = 0; j < n_baths; ++j)
        for (int x = 0; x < N; ++x) v_s[(size_t)j * N + x] = coupling_diag_host[(size_t)j * Nfull + states[x]];
    DevBuf<int> d_ra, d_rb, d_oa, d_ob;
    DevBuf<double> d_v, d_H0, d_quanta;
    DevBuf<cplx> scratch;
    if (jacobi) {
        H0_s.resize((size_t)N * N);
        quanta_s.resize(N);
        for (int x = 0; x < N; ++x) {
            quanta_s[x] = quanta_host[states[x]];
            for (int y = 0; y < N; ++y) H0_s[(size_t)x * N + y] = H0_host[(size_t)states[x] * Nfull + states[y]];
        }
        QSX_CUDA(d_H0.upload(H0_s, stream));
        QSX_CUDA(d_quanta.upload(quanta_s, stream));
    }
    QSX_CUDA(d_ra.upload(ra, stream));
    QSX_CUDA(d_rb.upload(rb, stream));
    QSX_CUDA(d_oa.upload(oa, stream));
    QSX_CUDA(d_ob.upload(ob, stream));
    QSX_CUDA(d_v.upload(v_s, stream));

    int dev = 0, smem_limit = 0, sms = 0;
    QSX_CUDA(cudaGetDevice(&dev));
    QSX_CUDA(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    QSX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t N2 = (size_t)N * N, N4 = (size_t)na * nbb * na * nbb;
    const int n_mats = bath->kind == QSX_BATH_DEBYE_COMPLEX ? (bath->matsubara_cutoff > 0 ? bath->matsubara_cutoff : 1000) : 0;
    size_t base = (3 * N2 + (size_t)n_baths * N2) * sizeof(cplx) +
                  ((size_t)((N + 1) & ~1) + 2 * (size_t)((n_mats + 1) & ~1)) * sizeof(double);
    const int in_place = (na <= 8 && nbb <= 8) ? 1 : 0;
    size_t with_t = base + (in_place ? 1 : 2) * N4 * sizeof(cplx);
    QSX_REQUIRE(base <= (size_t)smem_limit, "qsx_redfield_build: too many baths/states for shared memory");
    int tensors_in_smem = with_t <= (size_t)smem_limit;
    size_t smem = tensors_in_smem ? with_t : base;
    int grid;
    const bool fmo7 = N == 7 && na == 7 && nbb == 7 && n_baths == 7;
    const int threads = fmo7 ? 352 : 256;
    if (tensors_in_smem) {
        // resident CTAs only: every CTA loops over its share of the members
        int per_sm = 1;
        if (fmo7) {
            QSX_CUDA(cudaFuncSetAttribute(redfield_build_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            QSX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, redfield_build_kernel<7>, threads, smem));
        } else {
            QSX_CUDA(cudaFuncSetAttribute(redfield_build_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            QSX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, redfield_build_kernel<0>, threads, smem));
        }
        grid = std::min(n_members, sms * std::max(1, per_sm));
    } else {
        grid = std::min(n_members, sms * 2);
        QSX_CUDA(scratch.alloc((size_t)grid * 2 * N4));
    }
    RedfieldBuildArgs a;
    a.m = n_members; a.N = N; a.nb = n_baths; a.M = M;
    a.E = (const double *)E_dev; a.U = (const cplx *)U_dev; a.v = d_v.p;
    a.na = na; a.nbb = nbb; a.ra = d_ra.p; a.rb = d_rb.p; a.oa = d_oa.p; a.ob = d_ob.p;
    a.bath = *bath;
    if (a.bath.matsubara_cutoff <= 0) a.bath.matsubara_cutoff = 1000;
    a.n_mats = (n_mats + 1) & ~1;       // padded entries are zero terms
    a.secular = secular; a.eigen_basis = eigen_basis; a.unit_convert = unit_convert;
    a.L = (cplx *)L_out_dev; a.scratch = scratch.p; a.tensors_in_smem = tensors_in_smem; a.in_place = in_place;
    a.transposed_out = transposed_out;
    a.jacobi = jacobi; a.u_real = jacobi ? 1 : 0; a.H0 = d_H0.p; a.shifts = shifts_dev; a.quanta = d_quanta.p; a.rw_freq = rw_freq;
    DevBuf<double> d_E;
    DevBuf<cplx> d_U;
    if (jacobi && N <= 16) {
        // small systems: eigensystems first, one warp per member, then the build kernel reads them
        QSX_CUDA(d_E.alloc((size_t)n_members * N));
        QSX_CUDA(d_U.alloc((size_t)n_members * N2));
        const int per = 2 * (int)N2 + 2 * (((N + 1) & ~1) + 2);
        const int egrid = std::min((n_members + 7) / 8, sms * 8);
        redfield_eig_kernel<<<egrid, 256, (size_t)8 * per * sizeof(double), stream>>>(
            n_members, N, n_baths, d_H0.p, shifts_dev, d_v.p, d_quanta.p, rw_freq, d_E.p, d_U.p);
        qsx_launch_counter += 1;
        QSX_CUDA(cudaGetLastError());
        a.jacobi = 0; a.E = d_E.p; a.U = d_U.p;
    }
    if (fmo7) {
        QSX_CUDA(cudaFuncSetAttribute(redfield_build_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        redfield_build_kernel<7><<<grid, threads, smem, stream>>>(a);
    } else {
        QSX_CUDA(cudaFuncSetAttribute(redfield_build_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        redfield_build_kernel<0><<<grid, threads, smem, stream>>>(a);
    }
    qsx_launch_counter += 1;
    QSX_CUDA(cudaGetLastError());
    // no host synchronisation: the generators stay on the device, the scratch buffers return
    // to the stream-ordered pool (all calls of a thread share one stream)
    return QSX_OK;
}

extern "C" int qsx_redfield_build(int32_t n_members, int32_t N, const void *E_dev, const void *U_dev,
                                  int32_t n_baths, const double *coupling_diag_host,
                                  const qsx_bath *bath, int32_t secular, int32_t eigen_basis,
                                  double unit_convert, int32_t M, const int64_t *subspace_index_host,
                                  int32_t transposed_out, void *L_out_dev, void *stream) {
    return redfield_build_impl(n_members, N, E_dev, U_dev, nullptr, nullptr, nullptr, 0.0, n_baths,
                               coupling_diag_host, bath, secular, eigen_basis, unit_convert, M,
                               subspace_index_host, transposed_out, L_out_dev, stream);
}

extern "C" int qsx_redfield_build_sampled(int32_t n_members, int32_t N, const double *H0_host,
                                          const void *site_shifts_dev, const double *quanta_host,
                                          double rw_freq, int32_t n_baths,
                                          const double *coupling_diag_host, const qsx_bath *bath,
                                          int32_t secular, int32_t eigen_basis, double unit_convert,
                                          int32_t M, const int64_t *subspace_index_host,
                                          int32_t transposed_out, void *L_out_dev, void *stream) {
    QSX_REQUIRE(H0_host, "qsx_redfield_build_sampled: H0 missing");
    return redfield_build_impl(n_members, N, nullptr, nullptr, H0_host, (const double *)site_shifts_dev,
                               quanta_host, rw_freq, n_baths, coupling_diag_host, bath, secular,
                               eigen_basis, unit_convert, M, subspace_index_host, transposed_out,
                               L_out_dev, stream);
}

// K2 + K4 (grid-resident form): structure-aware HEOM hierarchy application and
// fused propagation.
//
// Replaces HEOMModel.HEOM_tensor + scipy csr_matrix.dot (reference
// dynamics/heom.py:228-244, 298-443).  No sparse matrix is materialised: with
// diagonal system-bath operators V_j (hamiltonian.py:593-608) every inter-ADO
// block of the reference generator is a *diagonal* matrix, so
//
//   d rho_n[e]/dt = sum_e' A[e,e'] rho_n[e']                 (commutator + temperature
//                   - shift_n rho_n[e]                         correction, ELL rows)
//                   + sum_links su(n_jk) gu[e] rho_{n+e_jk}[e] (index-map gather, up)
//                   + sum_links sd(n_jk) gd[e] rho_{n-e_jk}[e] (index-map gather, down)
//
// where the neighbour indices come from the closed-form ADO rank (ado.h).  The
// Heisenberg picture (generator transposed, heom.py:236-237) only swaps the
// link tables and transposes H.
//
// Propagation runs as ONE cooperative kernel for the whole trajectory: CTAs own
// tiles of ADOs, every integrator stage is "tile apply + element-local epilogue"
// followed by a grid-wide barrier; order/convergence control lives in device
// memory, so there is no per-step host round trip.
#include "common.cuh"
#include "ado.h"
#include <cooperative_groups.h>
#include <algorithm>
#include <complex>
#include <memory>
#include <math.h>
#include <string.h>

namespace cg = cooperative_groups;
typedef std::complex<double> zc;

struct HeomDev {
    int M, bins, K1, Lc, R, Lk, n_members;
    long long n_ado;
    const uint8_t *index;     // [n_ado][bins]
    const int *up, *down;     // [n_ado][bins]
    const double *shift;      // [n_ado]
    const int *ccol;          // [M][R]   (-1 padded)
    const cplx *cval;         // [n_members][M][R]
    const int *lbin;          // [M][Lk]  (-1 padded)
    const cplx *gu, *gd;      // [M][Lk]
    const double *su, *sd;    // [K1][Lc]
};

struct qsx_heom_s {
    HeomDev d;
    int n_sites = 0, K = 0, N = 0, heisenberg = 0;
    double lnorm = 0;         // inf-norm bound of the generator
    std::unique_ptr<AdoTables> tabs;
    DevBuf<uint8_t> index;
    DevBuf<int> up, down, ccol, lbin;
    DevBuf<double> shift, su, sd;
    DevBuf<cplx> cval, gu, gd;
};

// ------------------------------------------------------------- tile machinery
struct TileSmem {
    cplx *ys;          // [T][M] own states of the tile
    int *t_up, *t_dn;  // [T][bins]
    uint8_t *t_n;      // [T][bins]
    double *t_shift;   // [T]
    // tables (shared or global)
    const int *ccol;
    const cplx *cval;
    const int *lbin;
    const cplx *gu, *gd;
    const double *su, *sd;
    int cur_member;
    cplx *cval_s;      // shared staging area for cval (or null)
};

__device__ __forceinline__ size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// carve the dynamic shared memory; returns bytes used
__device__ __forceinline__ void tile_smem_setup(const HeomDev &H, unsigned char *base, int T,
                                                int tables_in_smem, TileSmem &s) {
    size_t off = 0;
    s.ys = reinterpret_cast<cplx *>(base + off); off = align16(off + (size_t)T * H.M * sizeof(cplx));
    s.t_up = reinterpret_cast<int *>(base + off); off = align16(off + (size_t)T * H.bins * sizeof(int));
    s.t_dn = reinterpret_cast<int *>(base + off); off = align16(off + (size_t)T * H.bins * sizeof(int));
    s.t_shift = reinterpret_cast<double *>(base + off); off = align16(off + (size_t)T * sizeof(double));
    s.t_n = reinterpret_cast<uint8_t *>(base + off); off = align16(off + (size_t)T * H.bins);
    s.cur_member = -1;
    if (tables_in_smem) {
        cplx *cv = reinterpret_cast<cplx *>(base + off); off = align16(off + (size_t)H.M * H.R * sizeof(cplx));
        cplx *gu = reinterpret_cast<cplx *>(base + off); off = align16(off + (size_t)H.M * H.Lk * sizeof(cplx));
        cplx *gd = reinterpret_cast<cplx *>(base + off); off = align16(off + (size_t)H.M * H.Lk * sizeof(cplx));
        double *su = reinterpret_cast<double *>(base + off); off = align16(off + (size_t)H.K1 * H.Lc * sizeof(double));
        double *sd = reinterpret_cast<double *>(base + off); off = align16(off + (size_t)H.K1 * H.Lc * sizeof(double));
        int *cc = reinterpret_cast<int *>(base + off); off = align16(off + (size_t)H.M * H.R * sizeof(int));
        int *lb = reinterpret_cast<int *>(base + off); off = align16(off + (size_t)H.M * H.Lk * sizeof(int));
        for (int i = threadIdx.x; i < H.M * H.Lk; i += blockDim.x) { gu[i] = H.gu[i]; gd[i] = H.gd[i]; lb[i] = H.lbin[i]; }
        for (int i = threadIdx.x; i < H.K1 * H.Lc; i += blockDim.x) { su[i] = H.su[i]; sd[i] = H.sd[i]; }
        for (int i = threadIdx.x; i < H.M * H.R; i += blockDim.x) cc[i] = H.ccol[i];
        s.cval_s = cv; s.cval = cv; s.gu = gu; s.gd = gd; s.su = su; s.sd = sd; s.ccol = cc; s.lbin = lb;
    } else {
        s.cval_s = nullptr; s.cval = H.cval; s.gu = H.gu; s.gd = H.gd; s.su = H.su; s.sd = H.sd;
        s.ccol = H.ccol; s.lbin = H.lbin;
    }
}

static size_t tile_smem_bytes(const HeomDev &H, int T, int tables_in_smem) {
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    size_t off = 0;
    off = al(off + (size_t)T * H.M * sizeof(cplx));
    off = al(off + (size_t)T * H.bins * sizeof(int));
    off = al(off + (size_t)T * H.bins * sizeof(int));
    off = al(off + (size_t)T * sizeof(double));
    off = al(off + (size_t)T * H.bins);
    if (tables_in_smem) {
        off = al(off + (size_t)H.M * H.R * sizeof(cplx));
        off = al(off + (size_t)H.M * H.Lk * sizeof(cplx));
        off = al(off + (size_t)H.M * H.Lk * sizeof(cplx));
        off = al(off + (size_t)H.K1 * H.Lc * sizeof(double));
        off = al(off + (size_t)H.K1 * H.Lc * sizeof(double));
        off = al(off + (size_t)H.M * H.R * sizeof(int));
        off = al(off + (size_t)H.M * H.Lk * sizeof(int));
    }
    return off;
}

// Apply the hierarchy generator to ADOs [n0, n0+T) of one column.
//   x    : the column's full state [n_ado][M] (global, read through L2)
//   epi  : epi(i, value, own) with i = n*M + e inside the column, value = (L x)[i],
//          own = x[i]; called once per element by its owner thread.
// Contains barriers; must be called by all threads of the CTA.
template <class Epi>
__device__ __forceinline__ void heom_tile(const HeomDev &H, TileSmem &s, const cplx *__restrict__ x,
                                          long long n0, int T, int member, Epi epi) {
    const int M = H.M, bins = H.bins, R = H.R, Lk = H.Lk;
    __syncthreads();          // previous tile fully consumed before its staging area is reused
    if (s.cval_s && member != s.cur_member) {
        const cplx *src = H.cval + (size_t)member * M * R;
        for (int i = threadIdx.x; i < M * R; i += blockDim.x) s.cval_s[i] = src[i];
    }
    s.cur_member = member;
    const cplx *cval = s.cval_s ? s.cval_s : H.cval + (size_t)member * M * R;
    {
        const cplx *xs = x + (size_t)n0 * M;
        for (int i = threadIdx.x; i < T * M; i += blockDim.x) s.ys[i] = __ldcg(&xs[i]);
        const size_t tb = (size_t)n0 * bins;
        for (int i = threadIdx.x; i < T * bins; i += blockDim.x) {
            s.t_up[i] = H.up[tb + i];
            s.t_dn[i] = H.down[tb + i];
            s.t_n[i] = H.index[tb + i];
        }
        for (int i = threadIdx.x; i < T; i += blockDim.x) s.t_shift[i] = H.shift[n0 + i];
    }
    __syncthreads();
    int lanes, la, e0, estride;
    if (M <= (int)blockDim.x) {
        lanes = blockDim.x / M; la = threadIdx.x / M; e0 = threadIdx.x % M; estride = M;
        if (la >= lanes) return;
    } else {
        lanes = 1; la = 0; e0 = threadIdx.x; estride = blockDim.x;
    }
    for (int e = e0; e < M; e += estride) {
        const int *crow = s.ccol + e * R;
        const cplx *vrow = cval + e * R;
        const int *lrow = s.lbin + e * Lk;
        for (int nl = la; nl < T; nl += lanes) {
            const cplx *yn = s.ys + nl * M;
            const cplx own = yn[e];
            cplx acc = cmake(-s.t_shift[nl] * own.x, -s.t_shift[nl] * own.y);
            for (int l = 0; l < R; ++l) {
                int c = crow[l];
                if (c < 0) break;
                cfma(acc, vrow[l], yn[c]);
            }
            for (int l = 0; l < Lk; ++l) {
                int b = lrow[l];
                if (b < 0) break;
                const int k = b % H.K1;
                const int njk = s.t_n[nl * bins + b];
                const int iu = s.t_up[nl * bins + b];
                const int id = s.t_dn[nl * bins + b];
                if (iu >= 0) {
                    cplx v = __ldcg(&x[(size_t)iu * M + e]);
                    cplx g = cscale(s.su[k * H.Lc + njk], s.gu[e * Lk + l]);
                    cfma(acc, g, v);
                }
                if (id >= 0) {
                    cplx v = __ldcg(&x[(size_t)id * M + e]);
                    cplx g = cscale(s.sd[k * H.Lc + njk], s.gd[e * Lk + l]);
                    cfma(acc, g, v);
                }
            }
            epi((n0 + nl) * (long long)M + e, acc, own);
        }
    }
}

// ------------------------------------------------------------------ kernels
struct HeomApplyArgs {
    HeomDev H;
    const cplx *x;
    cplx *y;
    const int *member_of;   // [B] or null
    int B, T, tables_in_smem;
    long long tiles_per_col;
};

__global__ void __launch_bounds__(256) heom_apply_kernel(HeomApplyArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem s;
    tile_smem_setup(a.H, smem_raw, a.T, a.tables_in_smem, s);
    const long long D = a.H.n_ado * a.H.M;
    const long long total = a.tiles_per_col * a.B;
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        int b = (int)(w / a.tiles_per_col);
        long long n0 = (w % a.tiles_per_col) * a.T;
        int T = (int)min((long long)a.T, a.H.n_ado - n0);
        int member = a.member_of ? a.member_of[b] : 0;
        cplx *yb = a.y + (size_t)b * D;
        heom_tile(a.H, s, a.x + (size_t)b * D, n0, T, member,
                  [&](long long i, cplx v, cplx) { yb[i] = v; });
    }
}

struct HeomPropArgs {
    HeomDev H;
    int B, T, tables_in_smem, nt;
    long long tiles_per_col;
    const int *member_of;
    const cplx *y0;
    cplx *Y, *V, *W, *X;        // work vectors [B][D] (X only for RK4)
    const double *t;
    double t0;
    int method;
    double rtol;
    int rk4_sub, kmax;
    double theta, lnorm;
    int save_mode, save_rows;
    const cplx *S;              // [save_rows][M]
    cplx *out;
    long long saved_dim;
    int *flags;                 // [3]
    double *ynorm;              // [3][B]
    unsigned long long *stats;  // rhs, steps, status
};

__device__ __forceinline__ void heom_save(const HeomPropArgs &a, int it) {
    const long long D = a.H.n_ado * a.H.M;
    const int M = a.H.M;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsz = (long long)gridDim.x * blockDim.x;
    if (a.save_mode == QSX_SAVE_STATE) {
        for (long long i = gtid; i < (long long)a.B * D; i += gsz) {
            long long b = i / D, r = i % D;
            a.out[((size_t)b * a.nt + it) * a.saved_dim + r] = __ldcg(&a.Y[i]);
        }
    } else if (a.save_mode == QSX_SAVE_ADO0) {
        for (long long i = gtid; i < (long long)a.B * M; i += gsz) {
            long long b = i / M, r = i % M;
            a.out[((size_t)b * a.nt + it) * a.saved_dim + r] = __ldcg(&a.Y[(size_t)b * D + r]);
        }
    } else {
        const long long per_col = a.H.n_ado * a.save_rows;
        for (long long i = gtid; i < (long long)a.B * per_col; i += gsz) {
            long long b = i / per_col, r = i % per_col;
            long long n = r / a.save_rows;
            int m = (int)(r % a.save_rows);
            const cplx *y = a.Y + (size_t)b * D + (size_t)n * M;
            cplx acc = cmake(0, 0);
            for (int e = 0; e < M; ++e) cfma(acc, __ldg(&a.S[(size_t)m * M + e]), __ldcg(&y[e]));
            a.out[((size_t)b * a.nt + it) * a.saved_dim + r] = acc;
        }
    }
}

__global__ void __launch_bounds__(256) heom_propagate_kernel(HeomPropArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    TileSmem s;
    tile_smem_setup(a.H, smem_raw, a.T, a.tables_in_smem, s);
    const long long D = a.H.n_ado * a.H.M;
    const long long total = a.tiles_per_col * a.B;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsz = (long long)gridDim.x * blockDim.x;
    const int B = a.B;

    // ---- init: Y = y0, reference norms, control words ---------------------
    for (int i = (int)gtid; i < 3 * B; i += (int)gsz) a.ynorm[i] = 0.0;
    if (gtid < 3) a.flags[gtid] = 0;
    grid.sync();
    for (int b = 0; b < B; ++b) {
        double loc = 0.0;
        for (long long i = gtid; i < D; i += gsz) {
            cplx v = a.y0[(size_t)b * D + i];
            a.Y[(size_t)b * D + i] = v;
            loc = fmax(loc, cabs1(v));
        }
        loc = warp_max(loc);
        if ((threadIdx.x & 31) == 0 && loc > 0) atomic_max_nonneg(&a.ynorm[b], loc);
    }
    grid.sync();

    unsigned long long n_rhs = 0, n_steps = 0;
    int status = 0;
    int fslot = 0;      // flag slot of the current convergence check
    int nslot = 0;      // norm slot that holds the latest reference norms
    double tcur = a.t0;

    for (int it = 0; it < a.nt; ++it) {
        const double target = a.t[it];
        if (target != tcur) {
            const double span = target - tcur;
            if (a.method == QSX_METHOD_TAYLOR) {
                int nsub = (int)ceil(fabs(span) * a.lnorm / a.theta);
                if (nsub < 1) nsub = 1;
                const double h = span / nsub;
                for (int sub = 0; sub < nsub; ++sub) {
                    const cplx *src = a.Y;
                    cplx *dst = a.V;
                    bool done = false;
                    for (int k = 1; k <= a.kmax; ++k) {
                        const double fac = h / k;
                        const bool even = (k & 1) == 0;
                        int ok = 1;
                        // block 0 recycles the control slots that come next
                        if (even && blockIdx.x == 0) {
                            if (threadIdx.x == 0) a.flags[(fslot + 1) % 3] = 0;
                            for (int b = threadIdx.x; b < B; b += blockDim.x) a.ynorm[((nslot + 2) % 3) * B + b] = 0.0;
                        }
                        for (long long w = blockIdx.x; w < total; w += gridDim.x) {
                            const int b = (int)(w / a.tiles_per_col);
                            const long long n0 = (w % a.tiles_per_col) * a.T;
                            const int T = (int)min((long long)a.T, a.H.n_ado - n0);
                            const int member = a.member_of ? a.member_of[b] : 0;
                            cplx *db = dst + (size_t)b * D;
                            cplx *Yb = a.Y + (size_t)b * D;
                            if (!even) {
                                heom_tile(a.H, s, src + (size_t)b * D, n0, T, member,
                                          [&](long long i, cplx f, cplx) { db[i] = cscale(fac, f); });
                            } else {
                                const double yref = a.rtol * __ldcg(&a.ynorm[nslot * B + b]);
                                double ymax = 0.0;
                                heom_tile(a.H, s, src + (size_t)b * D, n0, T, member,
                                          [&](long long i, cplx f, cplx own) {
                                              cplx wv = cscale(fac, f);
                                              db[i] = wv;
                                              cplx y = Yb[i];
                                              y.x += own.x + wv.x;
                                              y.y += own.y + wv.y;
                                              Yb[i] = y;
                                              if (cabs1(own) + cabs1(wv) > yref) ok = 0;
                                              ymax = fmax(ymax, cabs1(y));
                                          });
                                ymax = warp_max(ymax);
                                if ((threadIdx.x & 31) == 0 && ymax > 0)
                                    atomic_max_nonneg(&a.ynorm[((nslot + 1) % 3) * B + b], ymax);
                            }
                        }
                        n_rhs += 1;
                        if (even) {
                            int all_ok = __syncthreads_and(ok);
                            if (!all_ok && threadIdx.x == 0) atomicExch(&a.flags[fslot], 1);
                        }
                        grid.sync();
                        src = dst;
                        dst = (dst == a.V) ? a.W : a.V;
                        if (even) {
                            int failed = *((volatile int *)&a.flags[fslot]);
                            fslot = (fslot + 1) % 3;
                            nslot = (nslot + 1) % 3;
                            if (!failed) { done = true; break; }
                        }
                    }
                    if (!done) status = QSX_ERR_INTEGRATOR;
                    n_steps += 1;
                    tcur += h;
                }
            } else {
                // classic RK4 with fixed sub-steps; ACC = V, TA = W, TB = X
                cplx *ACC = a.V, *TA = a.W, *TB = a.X;
                const int nsub = a.rk4_sub > 0 ? a.rk4_sub : 1;
                const double h = span / nsub;
                for (int sub = 0; sub < nsub; ++sub) {
                    for (int stage = 0; stage < 4; ++stage) {
                        const cplx *src = stage == 0 ? a.Y : (stage == 2 ? TB : TA);
                        for (long long w = blockIdx.x; w < total; w += gridDim.x) {
                            const int b = (int)(w / a.tiles_per_col);
                            const long long n0 = (w % a.tiles_per_col) * a.T;
                            const int T = (int)min((long long)a.T, a.H.n_ado - n0);
                            const int member = a.member_of ? a.member_of[b] : 0;
                            const size_t o = (size_t)b * D;
                            cplx *Yb = a.Y + o, *Ab = ACC + o, *TAb = TA + o, *TBb = TB + o;
                            if (stage == 0)
                                heom_tile(a.H, s, src + o, n0, T, member, [&](long long i, cplx k1, cplx own) {
                                    TAb[i] = cadd(own, cscale(0.5 * h, k1));
                                    Ab[i] = cadd(own, cscale(h / 6.0, k1));
                                });
                            else if (stage == 1)
                                heom_tile(a.H, s, src + o, n0, T, member, [&](long long i, cplx k2, cplx) {
                                    TBb[i] = cadd(Yb[i], cscale(0.5 * h, k2));
                                    Ab[i] = cadd(Ab[i], cscale(h / 3.0, k2));
                                });
                            else if (stage == 2)
                                heom_tile(a.H, s, src + o, n0, T, member, [&](long long i, cplx k3, cplx) {
                                    TAb[i] = cadd(Yb[i], cscale(h, k3));
                                    Ab[i] = cadd(Ab[i], cscale(h / 3.0, k3));
                                });
                            else
                                heom_tile(a.H, s, src + o, n0, T, member, [&](long long i, cplx k4, cplx) {
                                    Yb[i] = cadd(Ab[i], cscale(h / 6.0, k4));
                                });
                        }
                        grid.sync();
                    }
                    n_rhs += 4;
                    n_steps += 1;
                    tcur += h;
                }
            }
            tcur = target;
        }
        heom_save(a, it);
        // the next stage that writes Y is separated from this read by >= 1 grid barrier
        // (Taylor: first write of Y happens at k = 2; RK4: at stage 4)
    }
    if (gtid == 0) {
        a.stats[0] = n_rhs * (unsigned long long)B;
        a.stats[1] = n_steps * (unsigned long long)B;
        a.stats[2] = (unsigned long long)(status != 0);
    }
}

// --------------------------------------------------------------------- host
extern "C" int qsx_heom_create(qsx_heom_t *out, const qsx_heom_config *cfg, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(out && cfg, "qsx_heom_create: null argument");
    QSX_REQUIRE(cfg->n_sites > 0 && cfg->K >= 0 && cfg->level_cutoff > 0 && cfg->n_hilbert > 0 &&
                cfg->M > 0 && cfg->n_members > 0, "qsx_heom_create: bad sizes");
    QSX_REQUIRE(cfg->level_cutoff <= 255, "level_cutoff too large");
    const int N = cfg->n_hilbert, M = cfg->M, K1 = cfg->K + 1, bins = cfg->n_sites * K1;
    const int Lc = cfg->level_cutoff;
    const double u = cfg->unit_convert;
    std::unique_ptr<qsx_heom_s> h(new qsx_heom_s());
    h->n_sites = cfg->n_sites; h->K = cfg->K; h->N = N; h->heisenberg = cfg->heisenberg;
    h->tabs.reset(new AdoTables(bins, Lc));
    AdoTables &tb = *h->tabs;
    QSX_REQUIRE(tb.n_ado < ((int64_t)1 << 31), "hierarchy too large");
    tb.enumerate();
    const int64_t n_ado = tb.n_ado;

    std::vector<int> ea(M), eb(M);
    for (int e = 0; e < M; ++e) {
        int64_t f = cfg->subspace_index[e];
        QSX_REQUIRE(f >= 0 && f < (int64_t)N * N, "subspace index out of range");
        ea[e] = (int)(f % N);
        eb[e] = (int)(f / N);
    }
    const zc *Hm = reinterpret_cast<const zc *>(cfg->H);
    const zc *cc = reinterpret_cast<const zc *>(cfg->c);
    const double *v = cfg->coupling_diag;
    const zc mi(0.0, -1.0);

    // ---- commutator rows (ELL), pattern = union over members -----------------
    auto coef = [&](const zc *H, int e, int e2) -> zc {
        int a = ea[e], b = eb[e], c = ea[e2], d = eb[e2];
        zc r = 0;
        if (!cfg->heisenberg) {
            if (d == b) r += H[a * N + c];
            if (a == c) r -= H[d * N + b];
        } else {
            if (b == d) r += H[c * N + a];
            if (c == a) r -= H[b * N + d];
        }
        return r;
    };
    std::vector<std::vector<int>> pattern(M);
    for (int e = 0; e < M; ++e) {
        for (int e2 = 0; e2 < M; ++e2) {
            bool nz = (e2 == e);
            if (!nz && (ea[e] == ea[e2] || eb[e] == eb[e2]))
                for (int m = 0; m < cfg->n_members && !nz; ++m)
                    nz = coef(Hm + (size_t)m * N * N, e, e2) != zc(0);
            if (nz) pattern[e].push_back(e2);
        }
    }
    int R = 1;
    for (auto &p : pattern) R = std::max<int>(R, (int)p.size());
    std::vector<int> ccol((size_t)M * R, -1);
    std::vector<cplx> cval((size_t)cfg->n_members * M * R, cmake(0, 0));
    std::vector<double> dbl(M, 0.0), rowsum(M, 0.0);
    for (int e = 0; e < M; ++e)
        for (int j = 0; j < cfg->n_sites; ++j) {
            double va = v[j * N + ea[e]], vb = v[j * N + eb[e]];
            dbl[e] += va + vb - 2 * va * vb;
        }
    for (int e = 0; e < M; ++e)
        for (size_t l = 0; l < pattern[e].size(); ++l) ccol[(size_t)e * R + l] = pattern[e][l];
    for (int m = 0; m < cfg->n_members; ++m)
        for (int e = 0; e < M; ++e) {
            double rs = 0;
            for (size_t l = 0; l < pattern[e].size(); ++l) {
                int e2 = pattern[e][l];
                zc val = mi * u * coef(Hm + (size_t)m * N * N, e, e2);
                if (e2 == e) val -= u * cfg->temp_corr * dbl[e];
                cval[((size_t)m * M + e) * R + l] = cmake(val.real(), val.imag());
                rs += std::abs(val);
            }
            rowsum[e] = std::max(rowsum[e], rs);
        }

    // ---- link tables ---------------------------------------------------------
    std::vector<std::vector<int>> lpat(M);
    std::vector<std::vector<zc>> lgu(M), lgd(M);
    for (int e = 0; e < M; ++e)
        for (int j = 0; j < cfg->n_sites; ++j) {
            double va = v[j * N + ea[e]], vb = v[j * N + eb[e]];
            for (int k = 0; k < K1; ++k) {
                zc dv = va - vb;
                zc cv = cc[k] * va - std::conj(cc[k]) * vb;
                zc gu = mi * u * (cfg->heisenberg ? cv : dv);
                zc gd = mi * u * (cfg->heisenberg ? dv : cv);
                if (gu != zc(0) || gd != zc(0)) {
                    lpat[e].push_back(j * K1 + k);
                    lgu[e].push_back(gu);
                    lgd[e].push_back(gd);
                }
            }
        }
    int Lk = 1;
    for (auto &p : lpat) Lk = std::max<int>(Lk, (int)p.size());
    std::vector<int> lbin((size_t)M * Lk, -1);
    std::vector<cplx> gu((size_t)M * Lk, cmake(0, 0)), gd((size_t)M * Lk, cmake(0, 0));
    for (int e = 0; e < M; ++e)
        for (size_t l = 0; l < lpat[e].size(); ++l) {
            lbin[(size_t)e * Lk + l] = lpat[e][l];
            gu[(size_t)e * Lk + l] = cmake(lgu[e][l].real(), lgu[e][l].imag());
            gd[(size_t)e * Lk + l] = cmake(lgd[e][l].real(), lgd[e][l].imag());
        }
    std::vector<double> su((size_t)K1 * Lc, 0.0), sd((size_t)K1 * Lc, 0.0);
    for (int k = 0; k < K1; ++k) {
        double ck = std::abs(cc[k]);
        auto modU = [&](int n) { return cfg->modified ? sqrt((n + 1) * ck) : 1.0; };
        auto modD = [&](int n) { return cfg->modified ? sqrt(n / ck) : (double)n; };
        for (int n = 0; n < Lc; ++n) {
            su[(size_t)k * Lc + n] = cfg->heisenberg ? modD(n + 1) : modU(n);
            sd[(size_t)k * Lc + n] = cfg->heisenberg ? (n > 0 ? modU(n - 1) : 0.0) : modD(n);
        }
    }
    std::vector<double> shift(n_ado);
    double lnorm = 0;
    for (int64_t n = 0; n < n_ado; ++n) {
        double sft = 0;
        for (int b = 0; b < bins; ++b) sft += tb.index[(size_t)n * bins + b] * cfg->nu[b % K1];
        shift[n] = u * sft;
    }
    // inf-norm bound: max over (n, e) of the absolute row sum
    for (int64_t n = 0; n < n_ado; ++n) {
        for (int e = 0; e < M; ++e) {
            double rs = rowsum[e] + fabs(shift[n]);
            for (size_t l = 0; l < lpat[e].size(); ++l) {
                int b = lpat[e][l], k = b % K1, njk = tb.index[(size_t)n * bins + b];
                if (tb.up[(size_t)n * bins + b] >= 0) rs += su[(size_t)k * Lc + njk] * std::abs(lgu[e][l]);
                if (tb.down[(size_t)n * bins + b] >= 0) rs += sd[(size_t)k * Lc + njk] * std::abs(lgd[e][l]);
            }
            lnorm = std::max(lnorm, rs);
        }
    }
    h->lnorm = lnorm;

    QSX_CUDA(h->index.upload(tb.index, stream));
    QSX_CUDA(h->up.upload(tb.up, stream));
    QSX_CUDA(h->down.upload(tb.down, stream));
    QSX_CUDA(h->shift.upload(shift, stream));
    QSX_CUDA(h->ccol.upload(ccol, stream));
    QSX_CUDA(h->cval.upload(cval, stream));
    QSX_CUDA(h->lbin.upload(lbin, stream));
    QSX_CUDA(h->gu.upload(gu, stream));
    QSX_CUDA(h->gd.upload(gd, stream));
    QSX_CUDA(h->su.upload(su, stream));
    QSX_CUDA(h->sd.upload(sd, stream));
    QSX_CUDA(cudaStreamSynchronize(stream));   // host vectors go out of scope

    HeomDev &d = h->d;
    d.M = M; d.bins = bins; d.K1 = K1; d.Lc = Lc; d.R = R; d.Lk = Lk; d.n_members = cfg->n_members;
    d.n_ado = n_ado;
    d.index = h->index.p; d.up = h->up.p; d.down = h->down.p; d.shift = h->shift.p;
    d.ccol = h->ccol.p; d.cval = h->cval.p; d.lbin = h->lbin.p; d.gu = h->gu.p; d.gd = h->gd.p;
    d.su = h->su.p; d.sd = h->sd.p;
    *out = h.release();
    return QSX_OK;
}

extern "C" void qsx_heom_destroy(qsx_heom_t h) { delete h; }
extern "C" int64_t qsx_heom_ado_count(qsx_heom_t h) { return h ? h->d.n_ado : -1; }

extern "C" int qsx_heom_index_maps(qsx_heom_t h, int64_t *ado_index, int32_t *up, int32_t *down) {
    QSX_REQUIRE(h, "null handle");
    const AdoTables &t = *h->tabs;
    size_t total = (size_t)t.n_ado * t.bins;
    if (ado_index) for (size_t i = 0; i < total; ++i) ado_index[i] = t.index[i];
    if (up) memcpy(up, t.up.data(), total * sizeof(int32_t));
    if (down) memcpy(down, t.down.data(), total * sizeof(int32_t));
    return QSX_OK;
}

struct HeomLaunchPlan {
    int threads, T, tables_in_smem;
    size_t smem;
    long long tiles_per_col;
};

static int plan_launch(const qsx_heom_s *h, HeomLaunchPlan &p) {
    const HeomDev &d = h->d;
    int dev = 0, smem_limit = 0;
    QSX_CUDA(cudaGetDevice(&dev));
    QSX_CUDA(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int threads;
    if (d.M <= 256) threads = std::max(64, ((256 / d.M) * d.M + 31) / 32 * 32);
    else threads = 256;
    threads = std::min(threads, 256);
    int lanes = d.M <= threads ? threads / d.M : 1;
    int T = std::max(lanes, std::min(64, std::max(1, 2048 / d.M)));
    T = (T + lanes - 1) / lanes * lanes;
    T = (int)std::min<long long>(T, std::max<long long>(1, d.n_ado));
    int tables = 1;
    // keep the per-CTA footprint small enough for several CTAs per SM
    if (tile_smem_bytes(d, T, 1) > (size_t)std::min(smem_limit, 100 * 1024)) tables = 0;
    while (tile_smem_bytes(d, T, tables) > (size_t)smem_limit && T > 1) T = std::max(1, T / 2);
    if (tile_smem_bytes(d, T, tables) > (size_t)smem_limit) {
        qsx_set_error("HEOM subspace dimension %d too large for the tile kernel", d.M);
        return QSX_ERR_UNSUPPORTED;
    }
    p.threads = threads; p.T = T; p.tables_in_smem = tables;
    p.smem = tile_smem_bytes(d, T, tables);
    p.tiles_per_col = (d.n_ado + T - 1) / T;
    return QSX_OK;
}

extern "C" int qsx_heom_apply(qsx_heom_t h, const void *y_dev, void *dy_dev, int32_t n_columns,
                              const int32_t *member_host, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(h && y_dev && dy_dev && n_columns > 0, "qsx_heom_apply: bad arguments");
    QSX_REQUIRE(y_dev != dy_dev, "qsx_heom_apply: in-place application is not supported");
    HeomLaunchPlan p;
    int rc = plan_launch(h, p);
    if (rc) return rc;
    DevBuf<int> member;
    if (member_host) {
        std::vector<int> m(member_host, member_host + n_columns);
        for (int x : m) QSX_REQUIRE(x >= 0 && x < h->d.n_members, "member index out of range");
        QSX_CUDA(member.upload(m, stream));
    }
    HeomApplyArgs a;
    a.H = h->d; a.x = (const cplx *)y_dev; a.y = (cplx *)dy_dev;
    a.member_of = member_host ? member.p : nullptr;
    a.B = n_columns; a.T = p.T; a.tables_in_smem = p.tables_in_smem; a.tiles_per_col = p.tiles_per_col;
    long long total = p.tiles_per_col * n_columns;
    int sms = 148, dev = 0;
    QSX_CUDA(cudaGetDevice(&dev));
    QSX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid = (int)std::min<long long>(total, (long long)sms * 8);
    QSX_CUDA(cudaFuncSetAttribute(heom_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    heom_apply_kernel<<<grid, p.threads, p.smem, stream>>>(a);
    qsx_launch_counter += 1;
    QSX_CUDA(cudaGetLastError());
    QSX_CUDA(cudaStreamSynchronize(stream));
    return QSX_OK;
}

extern "C" int qsx_heom_propagate(qsx_heom_t h, qsx_propagate_args *args, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(h && args, "qsx_heom_propagate: null argument");
    const HeomDev &d = h->d;
    const int B = args->n_columns, nt = args->n_times, M = d.M;
    const long long D = d.n_ado * M;
    QSX_REQUIRE(B > 0 && nt > 0 && args->t_host && args->y0_dev && args->out_dev,
                "qsx_heom_propagate: empty batch or missing buffers");
    QSX_REQUIRE(args->method == QSX_METHOD_TAYLOR || args->method == QSX_METHOD_RK4,
                "HEOM propagation supports the Taylor and RK4 integrators");
    QSX_REQUIRE(args->n_pulses == 0, "pulse-driven HEOM propagation is not available in this build");
    for (int i = 1; i < nt; ++i)
        QSX_REQUIRE(args->t_host[i] >= args->t_host[i - 1], "output times must be non-decreasing");
    QSX_REQUIRE(args->t_host[0] >= args->t0, "first output time precedes t0");
    HeomLaunchPlan p;
    int rc = plan_launch(h, p);
    if (rc) return rc;

    DevBuf<int> member, flags;
    DevBuf<double> d_t, ynorm;
    DevBuf<unsigned long long> stats;
    DevBuf<cplx> Y, V, W, X;
    if (args->generator_of_column_host) {
        std::vector<int> m(args->generator_of_column_host, args->generator_of_column_host + B);
        for (int x : m) QSX_REQUIRE(x >= 0 && x < d.n_members, "member index out of range");
        QSX_CUDA(member.upload(m, stream));
    }
    QSX_CUDA(d_t.upload(args->t_host, nt, stream));
    QSX_CUDA(flags.alloc(3));
    QSX_CUDA(ynorm.alloc((size_t)3 * B));
    QSX_CUDA(stats.alloc(3));
    QSX_CUDA(cudaMemsetAsync(stats.p, 0, 3 * sizeof(unsigned long long), stream));
    QSX_CUDA(Y.alloc((size_t)B * D));
    QSX_CUDA(V.alloc((size_t)B * D));
    QSX_CUDA(W.alloc((size_t)B * D));
    if (args->method == QSX_METHOD_RK4) QSX_CUDA(X.alloc((size_t)B * D));

    HeomPropArgs a;
    a.H = d; a.B = B; a.T = p.T; a.tables_in_smem = p.tables_in_smem; a.nt = nt;
    a.tiles_per_col = p.tiles_per_col;
    a.member_of = args->generator_of_column_host ? member.p : nullptr;
    a.y0 = (const cplx *)args->y0_dev;
    a.Y = Y.p; a.V = V.p; a.W = W.p; a.X = X.p;
    a.t = d_t.p; a.t0 = args->t0; a.method = args->method;
    a.rtol = args->rtol > 0 ? args->rtol : 1e-13;
    a.rk4_sub = args->rk4_substeps > 0 ? args->rk4_substeps : 16;
    a.kmax = 60; a.theta = 2.0; a.lnorm = h->lnorm;
    a.save_mode = args->save_mode; a.save_rows = args->save_rows;
    a.S = (const cplx *)args->save_dev;
    if (a.save_mode == QSX_SAVE_MATRIX) {
        QSX_REQUIRE(a.S && a.save_rows > 0 && args->n_save == 1, "HEOM save matrix must be shared ([rows][M])");
        a.saved_dim = d.n_ado * a.save_rows;
    } else if (a.save_mode == QSX_SAVE_ADO0) {
        a.saved_dim = M;
    } else {
        QSX_REQUIRE(a.save_mode == QSX_SAVE_STATE, "bad save_mode");
        a.saved_dim = D;
    }
    a.out = (cplx *)args->out_dev;
    a.flags = flags.p; a.ynorm = ynorm.p; a.stats = stats.p;

    int dev = 0, sms = 0, per_sm = 0;
    QSX_CUDA(cudaGetDevice(&dev));
    QSX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    QSX_CUDA(cudaFuncSetAttribute(heom_propagate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    QSX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, heom_propagate_kernel, p.threads, p.smem));
    QSX_REQUIRE(per_sm > 0, "heom_propagate_kernel does not fit on an SM");
    long long total = p.tiles_per_col * B;
    int grid = (int)std::min<long long>(total, (long long)sms * per_sm);
    void *kargs[] = {&a};
    cudaEvent_t e0, e1;
    QSX_CUDA(cudaEventCreate(&e0));
    QSX_CUDA(cudaEventCreate(&e1));
    QSX_CUDA(cudaEventRecord(e0, stream));
    cudaError_t e = cudaLaunchCooperativeKernel((void *)heom_propagate_kernel, dim3(grid), dim3(p.threads),
                                                kargs, p.smem, stream);
    qsx_launch_counter += 1;
    if (e != cudaSuccess) {
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        qsx_set_error("heom_propagate launch: %s", cudaGetErrorString(e));
        return QSX_ERR_CUDA;
    }
    QSX_CUDA(cudaEventRecord(e1, stream));
    unsigned long long st[3] = {0, 0, 0};
    QSX_CUDA(cudaMemcpyAsync(st, stats.p, sizeof(st), cudaMemcpyDeviceToHost, stream));
    QSX_CUDA(cudaStreamSynchronize(stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    args->rhs_evaluations = st[0];
    args->accepted_steps = st[1];
    args->kernel_ms = ms;
    if (st[2] != 0) {
        qsx_set_error("HEOM Taylor series did not converge within %d terms", a.kmax);
        return QSX_ERR_INTEGRATOR;
    }
    return QSX_OK;
}

// K2 + K4 (grid-resident form): structure-aware HEOM hierarchy application and
// fused propagation.
//
// Replaces HEOMModel.HEOM_tensor + scipy csr_matrix.dot (reference
// dynamics/heom.py:228-244, 298-443).  No sparse matrix is materialised: with
// diagonal system-bath operators V_j (hamiltonian.py:593-608) every inter-ADO
// block of the reference generator is a *diagonal* matrix, so for an ADO n with
// occupation vector n_jk and a rectangular Liouville block (rows R, cols C)
//
//   d rho_n/dt = Hs_R rho_n - rho_n Hs_C                      Hs = -i u H  (two small GEMMs)
//              - (shift_n + u tc dbl[e]) rho_n[e]              Matsubara shift, temperature corr.
//              + sum_links su(n_jk) gu[e] rho_{n+e_jk}[e]      index-map gather, one level up
//              + sum_links sd(n_jk) gd[e] rho_{n-e_jk}[e]      index-map gather, one level down
//
// with neighbour indices from the closed-form ADO rank (ado.h).  The Heisenberg
// picture (generator transposed, heom.py:236-237) swaps the link tables and
// transposes H.
//
// Device layout ("tile-SoA"): ADOs are grouped in tiles of 32; a column's state
// is stored as [tile][element e][32 ADOs].  A warp lane is an ADO, so
//   * the own-tile load is one contiguous, fully coalesced block;
//   * hierarchy gathers for a fixed element read runs of consecutive ADOs
//     (measured on the depth-8 FMO tables: 0.53-0.61 32-byte sectors per
//     gathered element, ideal 0.5, against 1.0 for an ADO-major layout);
//   * all shared-memory accesses of the small GEMMs are conflict free and the
//     H coefficients are warp-uniform broadcasts.
// The reference ordering [ADO][element] is restored at the API boundary.
//
// Propagation runs as ONE cooperative kernel for the whole trajectory: CTAs own
// tiles, every integrator stage is "tile apply + element-local epilogue"
// followed by a grid-wide barrier; order/convergence control lives in device
// memory, so there is no per-step host round trip.
#include "common.cuh"
#include "ado.h"
#include "heom_row.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <complex>
#include <memory>
#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace cg = cooperative_groups;
typedef std::complex<double> zc;

#define TL 32            // ADOs per tile = warp lanes
#define COMMA ,
#define REG_DIM 8        // small-GEMM register path for block dimensions <= REG_DIM

struct HeomDev {
    int nr, nc, M, bins, K1, Lc, Lk, n_members;
    long long n_ado, n_tiles;
    const double *shift;      // [n_tiles*32]
    const double *scale;      // [n_tiles*32] similarity scale s_n of the balanced hierarchy (error norm)
    const int *up, *down;     // [n_tiles][bins][32]
    const uint8_t *occ;       // [n_tiles][bins][32]
    const int *off_up, *off_dn;   // [n_tiles][bins][32] element offsets of the neighbour (tile*M*32 + lane), -1 = absent
    const int *e_off, *e_stride;  // [M] position of element e inside a tile: e_off[e] + lane * e_stride[e]
    int heis;                 // 1: Heisenberg picture (generator transposed)
    long long top_tile;       // first tile made of top-level ADOs only (no up-neighbours)
    int n_pulse, Rp;          // time-dependent terms: per-ADO operator rows in ELL form
    const int *pcol;          // [n_pulse][M][Rp] (-1 padded)
    const cplx *pval;         // [n_pulse][M][Rp]
    qsx_pulse pulses[QSX_MAX_PULSES];
    int ee;                   // 1: rows/cols are site-projector states (electronic block), TileEE applicable
    int real_h;               // 1: Hs_R, Hs_C purely imaginary (real Hamiltonian)
    cplx GuR[4], GdR[4], GuC[4], GdC[4];   // link coefficients per Matsubara index (row-site / col-site)
    const cplx *HR, *HC;      // [n_members][nr][nr], [n_members][nc][nc]  (pre-scaled by -i u)
    const double *dterm;      // [M]  u * tc * dbl[e]
    const int *lbin;          // [M][Lk]  (-1 padded)
    const cplx *gu, *gd;      // [M][Lk]
    const double *su, *sd;    // [K1][Lc]
};

struct qsx_heom_s {
    HeomDev d;
    int n_sites = 0, K = 0, N = 0, heisenberg = 0;
    double lnorm = 0;         // inf-norm bound of the generator
    std::unique_ptr<AdoTables> tabs;
    DevBuf<uint8_t> occ;
    DevBuf<int> up, down, lbin, off_up, off_dn, e_off, e_stride;
    DevBuf<double> shift, scale, su, sd, dterm;
    DevBuf<cplx> HR, HC, gu, gd;
    // row tile (heom_row.cuh): electronic blocks with a real Hamiltonian, Schroedinger picture
    bool row_ok = false;
    int row_cfg = 0;             // 1: Cfg<7,2> (FMO-like electronic block), 2: Cfg<8,2,4> (vibronic dimer, 4 states per site)
    heom_row::RowDev row;
    DevBuf<unsigned char> row_rec;
    DevBuf<double> row_h, row_g;
    DevBuf<cplx> row_ainv;
    DevBuf<int> row_dep;         // [n_tiles] last tile linked to a tile (flow scheduling, heom_row.cuh)
};

// ------------------------------------------------------------- tile machinery
struct TileSmem {
    cplx *ys;            // [M][32] source tile
    cplx *os;            // [M][32] commutator result
    int *t_up, *t_dn;    // [bins][32]
    uint8_t *t_n;        // [bins][32]
    double *t_shift;     // [32]
    cplx *HR, *HC;       // [nr][nr], [nc][nc] of the current member
    double *hR, *hC;     // imaginary parts of the above (real Hamiltonian, TileLean)
    cplx *tD;            // [K1][Lc] ready-multiplied down-link coefficients (TileLean)
    double *dterm;       // [M]
    int *lbin;           // [M][Lk]
    cplx *gu, *gd;       // [M][Lk]
    double *su, *sd;     // [K1][Lc]
    int cur_member;
    // cross-tile prefetch hints (set by the kernel loops, used by TileEE)
    int cur_col, next_col;
    long long next_tile;
    long long w, wstride, total;   // work index of the current tile, CTA stride, number of (column, tile) units
    int u_tile, u_sdiv, u_smod;    // tile of the current unit; wstride / n_tiles and wstride % n_tiles (no 64-bit
                                   // divisions per tile: a sweep advances (column, tile) incrementally)
    uint64_t *bars;          // mbarriers of the bulk-copy tile pipeline (TileLean OPT 8)
    int q;                   // number of tiles this CTA has processed so far (pipeline sequence number)
    const int *member_of;    // generator of each column (nullptr: member 0), for tiles staged ahead
    const cplx *pf;          // integrator operand (accumulator column) worth prefetching to L2, or nullptr
    const cplx *loaded;      // tile data currently staged (or in flight) in buffer `buf`
    int buf;
    double t_eval;           // time at which the pulse envelopes are evaluated
};

// (column, tile) of the unit `ahead` places after the current one in this CTA's strided sweep
__device__ __forceinline__ void unit_ahead(const TileSmem &s, long long n_tiles, int ahead, int &col, int &tile) {
    col = s.cur_col; tile = s.u_tile;
    for (int i = 0; i < ahead; ++i) {
        tile += s.u_smod; col += s.u_sdiv;
        if (tile >= n_tiles) { tile -= (int)n_tiles; ++col; }
    }
}
// bookkeeping of the unit a sweep is about to process (general form: two 64-bit divisions)
__device__ __forceinline__ void unit_set(TileSmem &s, long long w, long long wstride, long long total, long long n_tiles) {
    s.w = w; s.wstride = wstride; s.total = total;
    s.cur_col = (int)(w / n_tiles); s.u_tile = (int)(w % n_tiles);
    s.u_sdiv = (int)(wstride / n_tiles); s.u_smod = (int)(wstride % n_tiles);
    if (w + wstride < total) {
        int c, t;
        unit_ahead(s, n_tiles, 1, c, t);
        s.next_tile = t; s.next_col = c;
    } else {
        s.next_tile = -1; s.next_col = 0;
    }
}

__host__ __device__ __forceinline__ size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t tile_smem_layout(const HeomDev &H, size_t *offs) {
    size_t off = 0;
    offs[0] = off; off = al16(off + (size_t)H.M * TL * sizeof(cplx));        // ys
    offs[1] = off; off = al16(off + (size_t)H.M * TL * sizeof(cplx));        // os
    offs[2] = off; off = al16(off + (size_t)H.bins * TL * sizeof(int));      // t_up
    offs[3] = off; off = al16(off + (size_t)H.bins * TL * sizeof(int));      // t_dn
    offs[4] = off; off = al16(off + (size_t)H.bins * TL);                    // t_n
    offs[5] = off; off = al16(off + (size_t)2 * TL * sizeof(double));        // t_shift, t_scale
    offs[6] = off; off = al16(off + (size_t)H.nr * H.nr * sizeof(cplx));     // HR
    offs[7] = off; off = al16(off + (size_t)H.nc * H.nc * sizeof(cplx));     // HC
    offs[8] = off; off = al16(off + (size_t)H.M * sizeof(double));           // dterm
    offs[9] = off; off = al16(off + (size_t)H.M * H.Lk * sizeof(int));       // lbin
    offs[10] = off; off = al16(off + (size_t)H.M * H.Lk * sizeof(cplx));     // gu
    offs[11] = off; off = al16(off + (size_t)H.M * H.Lk * sizeof(cplx));     // gd
    offs[12] = off; off = al16(off + (size_t)H.K1 * H.Lc * sizeof(double));  // su
    offs[13] = off; off = al16(off + (size_t)H.K1 * H.Lc * sizeof(double));  // sd
    return off;
}

__device__ __forceinline__ void tile_smem_setup(const HeomDev &H, unsigned char *base, TileSmem &s) {
    size_t o[14];
    tile_smem_layout(H, o);
    s.ys = reinterpret_cast<cplx *>(base + o[0]);
    s.os = reinterpret_cast<cplx *>(base + o[1]);
    s.t_up = reinterpret_cast<int *>(base + o[2]);
    s.t_dn = reinterpret_cast<int *>(base + o[3]);
    s.t_n = reinterpret_cast<uint8_t *>(base + o[4]);
    s.t_shift = reinterpret_cast<double *>(base + o[5]);
    s.HR = reinterpret_cast<cplx *>(base + o[6]);
    s.HC = reinterpret_cast<cplx *>(base + o[7]);
    s.dterm = reinterpret_cast<double *>(base + o[8]);
    s.lbin = reinterpret_cast<int *>(base + o[9]);
    s.gu = reinterpret_cast<cplx *>(base + o[10]);
    s.gd = reinterpret_cast<cplx *>(base + o[11]);
    s.su = reinterpret_cast<double *>(base + o[12]);
    s.sd = reinterpret_cast<double *>(base + o[13]);
    s.cur_member = -1;
    for (int i = threadIdx.x; i < H.M; i += blockDim.x) s.dterm[i] = H.dterm[i];
    for (int i = threadIdx.x; i < H.M * H.Lk; i += blockDim.x) {
        s.lbin[i] = H.lbin[i]; s.gu[i] = H.gu[i]; s.gd[i] = H.gd[i];
    }
    for (int i = threadIdx.x; i < H.K1 * H.Lc; i += blockDim.x) { s.su[i] = H.su[i]; s.sd[i] = H.sd[i]; }
}

// out[a] = sum_c A[a][c] in[c]  (A: n x n in shared memory, warp-uniform reads);
// the result is handed to `put(a, value)`.  Inputs are fetched with get(c).
template <class Get, class Put>
__device__ __forceinline__ void small_gemv(const cplx *A, int n, Get get, Put put) {
    if (n <= REG_DIM) {
        cplx in[REG_DIM];
#pragma unroll
        for (int c = 0; c < REG_DIM; ++c) in[c] = (c < n) ? get(c) : cmake(0, 0);
        for (int a = 0; a < n; ++a) {
            cplx acc = cmake(0, 0);
            const cplx *row = A + a * n;
#pragma unroll
            for (int c = 0; c < REG_DIM; ++c)
                if (c < n) cfma(acc, row[c], in[c]);
            put(a, acc);
        }
    } else {
        for (int a = 0; a < n; ++a) {
            cplx acc = cmake(0, 0);
            const cplx *row = A + a * n;
            for (int c = 0; c < n; ++c) cfma(acc, row[c], get(c));
            put(a, acc);
        }
    }
}

// Apply the hierarchy generator to one tile (32 ADOs) of one column.
//   x    : the column's state in tile-SoA layout (global, read through L2)
//   pre  : p = pre(i) is called early for every owned element (prefetch of
//          integrator data, e.g. the accumulator Y[i]);
//   post : post(i, value, own, p) with i the tile-SoA index inside the column,
//          value = (L x)[i], own = x[i]; called once per element by its owner.
// Contains barriers; must be called by all threads of the CTA.
__device__ __forceinline__ void tile_stage(const HeomDev &H, TileSmem &s, const cplx *__restrict__ x,
                                           long long tile, int member) {
    const int M = H.M, bins = H.bins, nr = H.nr, nc = H.nc;
    __syncthreads();          // previous tile fully consumed before the staging areas are reused
    if (member != s.cur_member) {
        const cplx *hr = H.HR + (size_t)member * nr * nr, *hc = H.HC + (size_t)member * nc * nc;
        for (int i = threadIdx.x; i < nr * nr; i += blockDim.x) s.HR[i] = hr[i];
        for (int i = threadIdx.x; i < nc * nc; i += blockDim.x) s.HC[i] = hc[i];
        s.cur_member = member;
    }
    const cplx *xs = x + (size_t)tile * M * TL;
    for (int i = threadIdx.x; i < M * TL; i += blockDim.x) s.ys[i] = __ldcg(&xs[i]);
    const size_t tb = (size_t)tile * bins * TL;
    for (int i = threadIdx.x; i < bins * TL; i += blockDim.x) {
        s.t_up[i] = __ldg(&H.up[tb + i]);
        s.t_dn[i] = __ldg(&H.down[tb + i]);
        s.t_n[i] = __ldg(&H.occ[tb + i]);
    }
    if (threadIdx.x < TL) {
        s.t_shift[threadIdx.x] = __ldg(&H.shift[tile * TL + threadIdx.x]);
        s.t_shift[TL + threadIdx.x] = __ldg(&H.scale[tile * TL + threadIdx.x]);
    }
    __syncthreads();
}

struct TileGeneric {
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 3;
    static constexpr int UNITS = 1;          // tiles a CTA works on concurrently
    static size_t smem_bytes(const HeomDev &H) { size_t o[14]; return tile_smem_layout(H, o); }
    static __device__ __forceinline__ void setup(const HeomDev &H, unsigned char *base, TileSmem &s) {
        tile_smem_setup(H, base, s);
    }
    template <class Pre, class Post>
    static __device__ __forceinline__ void run(const HeomDev &H, TileSmem &s, const cplx *__restrict__ x,
                                               long long tile, int member, Pre pre, Post post) {
        const int M = H.M, nr = H.nr, nc = H.nc, Lk = H.Lk;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
        tile_stage(H, s, x, tile, member);
        // (Hs_R rho)[:, b]: one task per column b, lane = ADO
        for (int b = warp; b < nc; b += nwarp) {
            const cplx *col = s.ys + (size_t)b * nr * TL + lane;
            cplx *out = s.os + (size_t)b * nr * TL + lane;
            small_gemv(s.HR, nr, [&](int c) { return col[c * TL]; },
                       [&](int a, cplx v) { out[a * TL] = v; });
        }
        __syncthreads();
        // - (rho Hs_C)[a, :]: one task per row a; (rho Hs)[a][b] = sum_c A_C[b][c] rho[a][c]
        for (int a = warp; a < nr; a += nwarp) {
            const cplx *row = s.ys + (size_t)a * TL + lane;
            cplx *out = s.os + (size_t)a * TL + lane;
            small_gemv(s.HC, nc, [&](int c) { return row[(size_t)c * nr * TL]; },
                       [&](int b, cplx v) {
                           cplx o = out[(size_t)b * nr * TL];
                           out[(size_t)b * nr * TL] = cmake(o.x - v.x, o.y - v.y);
                       });
        }
        __syncthreads();
        // diagonal terms, hierarchy links, epilogue: warp = element, lane = ADO
        const double shift = s.t_shift[lane];
        const double wscale = s.t_shift[TL + lane];
        for (int e = warp; e < M; e += nwarp) {
            const long long gi = ((long long)tile * M + e) * TL + lane;
            const cplx p = pre(gi);
            const cplx own = s.ys[e * TL + lane];
            const cplx *yn_base = s.ys + lane;
            cplx acc = s.os[e * TL + lane];
            const double dg = shift + s.dterm[e];
            acc.x -= dg * own.x;
            acc.y -= dg * own.y;
            for (int p = 0; p < H.n_pulse; ++p) {
                // (-i E_p(t)) [V_p, rho_n][e]: the dipole commutator acts on every ADO alike
                const cplx gp = pulse_coefficient(H.pulses[p], s.t_eval);
                const int *pc = H.pcol + ((size_t)p * M + e) * H.Rp;
                const cplx *pw = H.pval + ((size_t)p * M + e) * H.Rp;
                cplx tmp = cmake(0, 0);
                for (int l = 0; l < H.Rp; ++l) {
                    const int c = __ldg(&pc[l]);
                    if (c < 0) break;
                    cfma(tmp, __ldg(&pw[l]), yn_base[c * TL]);
                }
                cfma(acc, gp, tmp);
            }
            const int *lrow = s.lbin + e * Lk;
            for (int l = 0; l < Lk; ++l) {
                const int b = lrow[l];
                if (b < 0) break;
                const int k = b % H.K1;
                const int njk = s.t_n[b * TL + lane];
                const int iu = s.t_up[b * TL + lane];
                const int id = s.t_dn[b * TL + lane];
                if (iu >= 0) {
                    cplx v = __ldcg(&x[((size_t)(iu >> 5) * M + e) * TL + (iu & 31)]);
                    cfma(acc, cscale(s.su[k * H.Lc + njk], s.gu[e * Lk + l]), v);
                }
                if (id >= 0) {
                    cplx v = __ldcg(&x[((size_t)(id >> 5) * M + e) * TL + (id & 31)]);
                    cfma(acc, cscale(s.sd[k * H.Lc + njk], s.gd[e * Lk + l]), v);
                }
            }
            post(gi, acc, own, p, wscale);
        }
    }
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// Electronic-block tile ("ee"-type Liouville blocks whose row/column states are
// site projectors, e.g. FMO 'ee'): warp w owns row w of every ADO matrix of the
// tile.  Structure used:
//   * links of element (a,b): bins of site a ("row-site") and of site b
//     ("col-site") with coefficients that depend only on the Matsubara index;
//     neighbour offsets are precomputed, so a gather is base + immediate;
//   * Hs = -i u H with real H: each complex MAC is 2 DFMA;
//   * no commutator staging buffer: row w of Hs rho and of rho Hs is formed in
//     registers straight from the source tile;
//   * source tiles are double buffered with cp.async: the next tile of this CTA
//     is in flight while the current one is processed (one barrier per tile).
template <int NS, int K1, int MINB, bool REAL_H, bool BATCH = false>
struct TileEE {
    static constexpr int THREADS = 32 * NS;
    static constexpr int MIN_BLOCKS = MINB;
    static constexpr int UNITS = 1;
    static constexpr int M = NS * NS;
    static constexpr int BINS = NS * K1;

    static __host__ __device__ size_t buf_bytes() {
        return al16((size_t)M * TL * sizeof(cplx)) + 2 * al16((size_t)BINS * TL * sizeof(int)) +
               al16((size_t)BINS * TL) + al16((size_t)2 * TL * sizeof(double));
    }
    static __host__ __device__ size_t shared_bytes(const HeomDev &H) {
        return al16((size_t)NS * NS * sizeof(cplx)) * 2 + al16((size_t)M * sizeof(double)) +
               2 * al16((size_t)K1 * H.Lc * sizeof(double));
    }
    static size_t smem_bytes(const HeomDev &H) { return shared_bytes(H) + 2 * buf_bytes(); }

    static __device__ __forceinline__ void setup(const HeomDev &H, unsigned char *base, TileSmem &s) {
        size_t off = 0;
        s.HR = reinterpret_cast<cplx *>(base + off); off += al16((size_t)NS * NS * sizeof(cplx));
        s.HC = reinterpret_cast<cplx *>(base + off); off += al16((size_t)NS * NS * sizeof(cplx));
        s.dterm = reinterpret_cast<double *>(base + off); off += al16((size_t)M * sizeof(double));
        s.su = reinterpret_cast<double *>(base + off); off += al16((size_t)K1 * H.Lc * sizeof(double));
        s.sd = reinterpret_cast<double *>(base + off); off += al16((size_t)K1 * H.Lc * sizeof(double));
        s.os = reinterpret_cast<cplx *>(base + off);       // start of the two tile buffers
        for (int i = threadIdx.x; i < M; i += blockDim.x) s.dterm[i] = H.dterm[i];
        for (int i = threadIdx.x; i < K1 * H.Lc; i += blockDim.x) { s.su[i] = H.su[i]; s.sd[i] = H.sd[i]; }
        s.cur_member = -1;
        s.loaded = nullptr;
        s.buf = 0;
        __syncthreads();
    }

    struct Buf {
        cplx *ys; int *o_up, *o_dn; uint8_t *occ; double *sh;
    };
    static __device__ __forceinline__ Buf buffer(const TileSmem &s, int which) {
        unsigned char *p = reinterpret_cast<unsigned char *>(s.os) + (size_t)which * buf_bytes();
        Buf b;
        b.ys = reinterpret_cast<cplx *>(p); p += al16((size_t)M * TL * sizeof(cplx));
        b.o_up = reinterpret_cast<int *>(p); p += al16((size_t)BINS * TL * sizeof(int));
        b.o_dn = reinterpret_cast<int *>(p); p += al16((size_t)BINS * TL * sizeof(int));
        b.occ = p; p += al16((size_t)BINS * TL);
        b.sh = reinterpret_cast<double *>(p);
        return b;
    }
    // asynchronous copy of one source tile and its tables into a buffer
    static __device__ __forceinline__ void issue(const HeomDev &H, const Buf &b, const cplx *tile_data,
                                                 long long tile) {
        for (int i = threadIdx.x; i < M * TL; i += THREADS) cp_async16(&b.ys[i], &tile_data[i]);
        const size_t tb = (size_t)tile * BINS * TL;
        for (int i = threadIdx.x; i < BINS * TL / 4; i += THREADS) {
            cp_async16(&b.o_up[4 * i], &H.off_up[tb + 4 * i]);
            cp_async16(&b.o_dn[4 * i], &H.off_dn[tb + 4 * i]);
        }
        for (int i = threadIdx.x; i < BINS * TL / 16; i += THREADS) cp_async16(&b.occ[16 * i], &H.occ[tb + 16 * i]);
        if (threadIdx.x < TL / 2) {
            cp_async16(&b.sh[2 * threadIdx.x], &H.shift[tile * TL + 2 * threadIdx.x]);
            cp_async16(&b.sh[TL + 2 * threadIdx.x], &H.scale[tile * TL + 2 * threadIdx.x]);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }

    template <class Pre, class Post>
    static __device__ __forceinline__ void run(const HeomDev &H, TileSmem &s, const cplx *__restrict__ x,
                                               long long tile, int member, Pre pre, Post post) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;     // w = row a
        const size_t Dp = (size_t)H.n_tiles * M * TL;
        const cplx *tile_data = x + (size_t)tile * M * TL;
        if (s.loaded != tile_data) {        // first tile of a phase: nothing in flight yet
            s.buf = 0;
            issue(H, buffer(s, 0), tile_data, tile);
        }
        if (member != s.cur_member) {
            __syncthreads();                // slower warps may still read the previous member's coefficients
            const cplx *hr = H.HR + (size_t)member * NS * NS, *hc = H.HC + (size_t)member * NS * NS;
            for (int i = threadIdx.x; i < NS * NS; i += THREADS) { s.HR[i] = hr[i]; s.HC[i] = hc[i]; }
            s.cur_member = member;
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();                    // tile visible; every warp is done with the other buffer
        const Buf cur = buffer(s, s.buf);
        if (s.next_tile >= 0) {             // prefetch this CTA's next tile into the other buffer
            const cplx *nd = (x - (size_t)s.cur_col * Dp) + (size_t)s.next_col * Dp + (size_t)s.next_tile * M * TL;
            issue(H, buffer(s, s.buf ^ 1), nd, s.next_tile);
            s.loaded = nd;
            s.buf ^= 1;
        } else {
            s.loaded = nullptr;
        }
        const long long gbase = ((long long)tile * M) * TL + lane + (long long)w * TL;   // element (w, 0)
        // integrator prefetch (e.g. the accumulator Y) for the 7 owned elements
        cplx pv[NS];
#pragma unroll
        for (int b = 0; b < NS; ++b) pv[b] = pre(gbase + (long long)b * NS * TL);
        // own row and diagonal terms
        cplx own[NS], acc[NS];
        const double shift = cur.sh[lane];
#pragma unroll
        for (int c = 0; c < NS; ++c) own[c] = cur.ys[(w + NS * c) * TL + lane];
#pragma unroll
        for (int b = 0; b < NS; ++b) {
            const double dg = shift + s.dterm[w + NS * b];
            acc[b] = cmake(-dg * own[b].x, -dg * own[b].y);
        }
        // hierarchy links: row-site bins (site w), then col-site bins (site b)
        const cplx *xw = x + (size_t)w * TL;
        cplx vru[BATCH ? K1 : 1][NS], vrd[BATCH ? K1 : 1][NS], vcu[NS][BATCH ? K1 : 1], vcd[NS][BATCH ? K1 : 1];
        if (BATCH) {
            // issue every gather of this row first (56 independent loads per lane); they
            // are consumed after the commutator, whose arithmetic hides their latency
#pragma unroll
            for (int k = 0; k < K1; ++k) {
                const int bin = w * K1 + k;
                const int ou = cur.o_up[bin * TL + lane], od = cur.o_dn[bin * TL + lane];
#pragma unroll
                for (int b = 0; b < NS; ++b) {
                    vru[k][b] = ou >= 0 ? __ldcg(xw + ou + b * NS * TL) : cmake(0, 0);
                    vrd[k][b] = od >= 0 ? __ldcg(xw + od + b * NS * TL) : cmake(0, 0);
                }
            }
#pragma unroll
            for (int b = 0; b < NS; ++b)
#pragma unroll
                for (int k = 0; k < K1; ++k) {
                    const int bin = b * K1 + k;
                    const int ou = cur.o_up[bin * TL + lane], od = cur.o_dn[bin * TL + lane];
                    vcu[b][k] = ou >= 0 ? __ldcg(xw + ou + b * NS * TL) : cmake(0, 0);
                    vcd[b][k] = od >= 0 ? __ldcg(xw + od + b * NS * TL) : cmake(0, 0);
                }
        } else {
#pragma unroll
        for (int k = 0; k < K1; ++k) {
            const int bin = w * K1 + k;
            const int njk = cur.occ[bin * TL + lane];
            const int ou = cur.o_up[bin * TL + lane], od = cur.o_dn[bin * TL + lane];
            if (__any_sync(0xffffffffu, ou >= 0)) {
                const cplx g = cscale(s.su[k * H.Lc + njk], H.GuR[k]);
                const cplx *base = xw + (ou >= 0 ? ou : 0);
#pragma unroll
                for (int b = 0; b < NS; ++b) {
                    cplx v = ou >= 0 ? __ldcg(base + b * NS * TL) : cmake(0, 0);
                    cfma(acc[b], g, v);
                }
            }
            if (__any_sync(0xffffffffu, od >= 0)) {
                const cplx g = cscale(s.sd[k * H.Lc + njk], H.GdR[k]);
                const cplx *base = xw + (od >= 0 ? od : 0);
#pragma unroll
                for (int b = 0; b < NS; ++b) {
                    cplx v = od >= 0 ? __ldcg(base + b * NS * TL) : cmake(0, 0);
                    cfma(acc[b], g, v);
                }
            }
        }
#pragma unroll
        for (int b = 0; b < NS; ++b) {
#pragma unroll
            for (int k = 0; k < K1; ++k) {
                const int bin = b * K1 + k;
                const int njk = cur.occ[bin * TL + lane];
                const int ou = cur.o_up[bin * TL + lane], od = cur.o_dn[bin * TL + lane];
                if (__any_sync(0xffffffffu, ou >= 0)) {
                    cplx v = ou >= 0 ? __ldcg(xw + ou + b * NS * TL) : cmake(0, 0);
                    cfma(acc[b], cscale(s.su[k * H.Lc + njk], H.GuC[k]), v);
                }
                if (__any_sync(0xffffffffu, od >= 0)) {
                    cplx v = od >= 0 ? __ldcg(xw + od + b * NS * TL) : cmake(0, 0);
                    cfma(acc[b], cscale(s.sd[k * H.Lc + njk], H.GdC[k]), v);
                }
            }
        }
        }
        // commutator: row w of Hs_R rho - rho Hs_C (A_C stored transposed: A_C[b][c] = Hs_C[c][b])
        if (REAL_H) {
            // Hs = i h with real h: (i h) z = h (-z.y, z.x)
#pragma unroll
            for (int c = 0; c < NS; ++c) {
                const double h = s.HR[w * NS + c].y;
#pragma unroll
                for (int b = 0; b < NS; ++b) {
                    const cplx z = cur.ys[(c + NS * b) * TL + lane];
                    acc[b].x = fma(-h, z.y, acc[b].x);
                    acc[b].y = fma(h, z.x, acc[b].y);
                }
            }
#pragma unroll
            for (int b = 0; b < NS; ++b)
#pragma unroll
                for (int c = 0; c < NS; ++c) {
                    const double h = s.HC[b * NS + c].y;
                    acc[b].x = fma(h, own[c].y, acc[b].x);
                    acc[b].y = fma(-h, own[c].x, acc[b].y);
                }
        } else {
#pragma unroll
            for (int c = 0; c < NS; ++c) {
                const cplx h = s.HR[w * NS + c];
#pragma unroll
                for (int b = 0; b < NS; ++b) cfma(acc[b], h, cur.ys[(c + NS * b) * TL + lane]);
            }
#pragma unroll
            for (int b = 0; b < NS; ++b)
#pragma unroll
                for (int c = 0; c < NS; ++c) {
                    const cplx h = s.HC[b * NS + c];
                    cfma(acc[b], cmake(-h.x, -h.y), own[c]);
                }
        }
        if (BATCH) {
#pragma unroll
            for (int k = 0; k < K1; ++k) {
                const int njk = cur.occ[(w * K1 + k) * TL + lane];
                const cplx gu = cscale(s.su[k * H.Lc + njk], H.GuR[k]);
                const cplx gd = cscale(s.sd[k * H.Lc + njk], H.GdR[k]);
#pragma unroll
                for (int b = 0; b < NS; ++b) { cfma(acc[b], gu, vru[k][b]); cfma(acc[b], gd, vrd[k][b]); }
            }
#pragma unroll
            for (int b = 0; b < NS; ++b)
#pragma unroll
                for (int k = 0; k < K1; ++k) {
                    const int njk = cur.occ[(b * K1 + k) * TL + lane];
                    cfma(acc[b], cscale(s.su[k * H.Lc + njk], H.GuC[k]), vcu[b][k]);
                    cfma(acc[b], cscale(s.sd[k * H.Lc + njk], H.GdC[k]), vcd[b][k]);
                }
        }
        const double wscale = cur.sh[TL + lane];
#pragma unroll
        for (int b = 0; b < NS; ++b) post(gbase + (long long)b * NS * TL, acc[b], own[b], pv[b], wscale);
    }
};


// ------------------------------------------------------------ layout kernels
// reference [b][n][e]  <->  tile-SoA [b][tile][e][32]
__global__ void heom_to_internal(const cplx *__restrict__ ref, cplx *__restrict__ internal, int B,
                                 long long n_ado, long long n_tiles, int M,
                                 const int *__restrict__ e_off, const int *__restrict__ e_stride,
                                 const double *__restrict__ g) {
    const long long per = n_tiles * M * TL;
    const long long total = per * B;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long b = i / per, r = i % per;
        long long tile = r / ((long long)M * TL);
        int e = (int)((r / TL) % M), lane = (int)(r % TL);
        long long n = tile * TL + lane;
        cplx v = (n < n_ado) ? ref[((size_t)b * n_ado + n) * M + e] : cmake(0, 0);
        if (g) v = cscale(g[n], v);             // row tile: sigma_n = g_n rho_n (heom_row.cuh)
        internal[(size_t)b * per + tile * M * TL + e_off[e] + lane * e_stride[e]] = v;
    }
}

__global__ void heom_from_internal(const cplx *__restrict__ internal, cplx *__restrict__ ref, int B,
                                   long long n_ado, long long n_tiles, int M,
                                   const int *__restrict__ e_off, const int *__restrict__ e_stride,
                                   const double *__restrict__ g) {
    const long long per = n_ado * M;
    const long long total = per * B;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long b = i / per, r = i % per;
        long long n = r / M;
        int e = (int)(r % M);
        cplx v = internal[((size_t)b * n_tiles * M + (n >> 5) * M) * TL + e_off[e] + (n & 31) * e_stride[e]];
        if (g) v = cscale(1.0 / g[n], v);
        ref[i] = v;
    }
}

// ------------------------------------------------------------------ kernels
struct HeomApplyArgs {
    HeomDev H;
    const cplx *x;          // internal layout
    cplx *y;                // internal layout
    const int *member_of;   // [B] or null
    int B;
};

template <class Tile>
__global__ void __launch_bounds__(Tile::THREADS, Tile::MIN_BLOCKS) heom_apply_kernel(HeomApplyArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem s;
    Tile::setup(a.H, smem_raw, s);
    s.pf = nullptr;
    s.member_of = a.member_of;
    const long long Dp = a.H.n_tiles * a.H.M * TL;
    const long long total = a.H.n_tiles * a.B;
    const long long w0 = Tile::UNITS == 1 ? blockIdx.x : (long long)blockIdx.x * Tile::UNITS + (threadIdx.x >> 5);
    const long long wstride = (long long)gridDim.x * Tile::UNITS;
    for (long long w = w0; w < total; w += wstride) {
        int b = (int)(w / a.H.n_tiles);
        long long tile = w % a.H.n_tiles;
        unit_set(s, w, wstride, total, a.H.n_tiles);
        int member = a.member_of ? a.member_of[b] : 0;
        cplx *yb = a.y + (size_t)b * Dp;
        Tile::run(a.H, s, a.x + (size_t)b * Dp, tile, member,
                  [&](long long) { return cmake(0, 0); },
                  [&](long long i, cplx v, cplx, cplx, double) { yb[i] = v; });
    }
}

struct HeomPropArgs {
    HeomDev H;
    int B, nt;
    const int *member_of;
    const cplx *y0;             // reference layout [B][n_ado][M]
    cplx *Y, *V, *W, *X;        // work vectors, internal layout [B][Dp] (X only for RK4)
    cplx *K[7];                 // DOPRI5 stage derivatives (TA = V, TB = W)
    double atol;
    double *red;                // [4] rotating error accumulators (DOPRI5)
    const double *t;
    double t0;
    double rtol;
    int rk4_sub, kmax;
    double theta, lnorm;
    int save_mode, save_rows;
    const cplx *S;              // [n_save][save_rows][M]
    const int *save_of;         // [B] save matrix of each column, or null (matrix 0)
    cplx *out;
    long long saved_dim;
    int *flags;                 // [3]
    double *ynorm;              // [3][B]
    unsigned long long *stats;  // rhs, steps, status
};

__device__ __forceinline__ void heom_save(const HeomPropArgs &a, int it) {
    const long long n_ado = a.H.n_ado, n_tiles = a.H.n_tiles;
    const int M = a.H.M;
    const long long Dp = n_tiles * M * TL;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsz = (long long)gridDim.x * blockDim.x;
    if (a.save_mode == QSX_SAVE_STATE) {
        const long long per = n_ado * M;
        for (long long i = gtid; i < (long long)a.B * per; i += gsz) {
            long long b = i / per, r = i % per, n = r / M;
            int e = (int)(r % M);
            a.out[((size_t)b * a.nt + it) * a.saved_dim + r] =
                __ldcg(&a.Y[(size_t)b * Dp + ((n >> 5) * M) * TL + a.H.e_off[e] + (n & 31) * a.H.e_stride[e]]);
        }
    } else if (a.save_mode == QSX_SAVE_ADO0) {
        for (long long i = gtid; i < (long long)a.B * M; i += gsz) {
            long long b = i / M, e = i % M;
            a.out[((size_t)b * a.nt + it) * a.saved_dim + e] = __ldcg(&a.Y[(size_t)b * Dp + a.H.e_off[e]]);
        }
    } else {
        const long long per_col = n_ado * a.save_rows;
        for (long long i = gtid; i < (long long)a.B * per_col; i += gsz) {
            long long b = i / per_col, r = i % per_col;
            long long n = r / a.save_rows;
            int m = (int)(r % a.save_rows);
            const cplx *y = a.Y + (size_t)b * Dp + ((n >> 5) * M) * TL;
            const int ln = (int)(n & 31);
            const cplx *Sm = a.S + (a.save_of ? (size_t)a.save_of[b] * a.save_rows * M : 0);
            cplx acc = cmake(0, 0);
            for (int e = 0; e < M; ++e)
                cfma(acc, __ldg(&Sm[(size_t)m * M + e]), __ldcg(&y[a.H.e_off[e] + ln * a.H.e_stride[e]]));
            a.out[((size_t)b * a.nt + it) * a.saved_dim + r] = acc;
        }
    }
}

template <int METHOD, class Tile>
__device__ __forceinline__ void heom_propagate_impl(const HeomPropArgs &a);

template <int METHOD, class Tile>
__global__ void __launch_bounds__(Tile::THREADS, Tile::MIN_BLOCKS) heom_propagate_kernel(const __grid_constant__ HeomPropArgs a) {
    heom_propagate_impl<METHOD, Tile>(a);
}
// the same kernel under an explicit register cap (two 7-warp CTAs per SM fit 144 registers per
// thread; __launch_bounds__(224, 2) rounds down to 128)
template <int METHOD, class Tile, int NREG>
__global__ void __maxnreg__(NREG) heom_propagate_kernel_r(const __grid_constant__ HeomPropArgs a) {
    heom_propagate_impl<METHOD, Tile>(a);
}

template <int METHOD, class Tile>
__device__ __forceinline__ void heom_propagate_impl(const HeomPropArgs &a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    TileSmem s;
    Tile::setup(a.H, smem_raw, s);
    s.pf = nullptr;
    s.member_of = a.member_of;
    const long long w0 = Tile::UNITS == 1 ? blockIdx.x : (long long)blockIdx.x * Tile::UNITS + (threadIdx.x >> 5);
    const long long wstride = (long long)gridDim.x * Tile::UNITS;
    const int M = a.H.M;
    const long long n_tiles = a.H.n_tiles;
    const long long Dp = n_tiles * M * TL;
    const long long total = n_tiles * a.B;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsz = (long long)gridDim.x * blockDim.x;
    const int B = a.B;

    // ---- init: Y = y0 (layout change), reference norms, control words ---------
    for (int i = (int)gtid; i < 3 * B; i += (int)gsz) a.ynorm[i] = 0.0;
    if (gtid < 3) a.flags[gtid] = 0;
    if (gtid < 4 && a.red) a.red[gtid] = 0.0;
    grid.sync();
    for (int b = 0; b < B; ++b) {
        double loc = 0.0;
        for (long long i = gtid; i < Dp; i += gsz) {
            long long tile = i / ((long long)M * TL);
            int e = (int)((i / TL) % M), lane = (int)(i % TL);
            long long n = tile * TL + lane;
            cplx v = (n < a.H.n_ado) ? a.y0[((size_t)b * a.H.n_ado + n) * M + e] : cmake(0, 0);
            a.Y[(size_t)b * Dp + tile * M * TL + a.H.e_off[e] + lane * a.H.e_stride[e]] = v;
            loc = fmax(loc, __ldg(&a.H.scale[tile * TL + lane]) * cabs1(v));
        }
        loc = warp_max(loc);
        if ((threadIdx.x & 31) == 0 && loc > 0) atomic_max_nonneg(&a.ynorm[b], loc);
    }
    grid.sync();

    unsigned long long n_rhs = 0, n_steps = 0;
    int status = 0;
    const int ub0 = (int)(w0 / n_tiles), ut0 = (int)(w0 % n_tiles);
    const int usdiv = (int)(wstride / n_tiles), usmod = (int)(wstride % n_tiles);
    s.t_eval = a.t0;
    int fslot = 0;      // flag slot of the current convergence check
    int nslot = 0;      // norm slot that holds the latest reference norms
    double tcur = a.t0;
    double dp_h = 0.0;          // DOPRI5 step carried across output points
    bool dp_have_k1 = false;    // DOPRI5 FSAL
    int dp_slot = 0;            // rotating error accumulator

    for (int it = 0; it < a.nt; ++it) {
        const double target = a.t[it];
        if (target != tcur) {
            const double span = target - tcur;
            if (METHOD == QSX_METHOD_TAYLOR) {
                // Taylor series of exp(hL) y; terms are accumulated into Y in pairs
                // (even k) so that Y is never read by a neighbour gather while it
                // is being updated and the Y traffic is halved.
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
                        if (even && blockIdx.x == 0) {   // recycle the control slots that come next
                            if (threadIdx.x == 0) a.flags[(fslot + 1) % 3] = 0;
                            for (int b = threadIdx.x; b < B; b += blockDim.x) a.ynorm[((nslot + 2) % 3) * B + b] = 0.0;
                        }
                        int ub = ub0, ut = ut0;        // (column, tile) of the unit, advanced without divisions
                        for (long long w = w0; w < total; w += wstride) {
                            const int b = ub;
                            const long long tile = ut;
                            s.w = w; s.wstride = wstride; s.total = total; s.cur_col = ub; s.u_tile = ut;
                            s.u_sdiv = usdiv; s.u_smod = usmod;
                            ut += usmod; ub += usdiv;
                            if (ut >= n_tiles) { ut -= (int)n_tiles; ++ub; }
                            if (w + wstride < total) { s.next_tile = ut; s.next_col = ub; } else { s.next_tile = -1; s.next_col = 0; }
                            const int member = a.member_of ? a.member_of[b] : 0;
                            cplx *db = dst + (size_t)b * Dp;
                            cplx *Yb = a.Y + (size_t)b * Dp;
                            s.pf = even ? Yb : nullptr;
                            if (!even) {
                                Tile::run(a.H, s, src + (size_t)b * Dp, tile, member,
                                          [&](long long) { return cmake(0, 0); },
                                          [&](long long i, cplx f, cplx, cplx, double) { __stcs(&db[i], cscale(fac, f)); });
                            } else {
                                const double yref = a.rtol * __ldcg(&a.ynorm[nslot * B + b]);
                                double ymax = 0.0;
                                Tile::run(a.H, s, src + (size_t)b * Dp, tile, member,
                                          [&](long long i) { return __ldcs(&Yb[i]); },
                                          [&](long long i, cplx f, cplx own, cplx y, double sc) {
                                              cplx wv = cscale(fac, f);
                                              __stcs(&db[i], wv);
                                              y.x += own.x + wv.x;
                                              y.y += own.y + wv.y;
                                              __stcs(&Yb[i], y);
                                              if (!(sc * (cabs1(own) + cabs1(wv)) <= yref)) ok = 0;      // also catches a non-finite state
                                              ymax = fmax(ymax, sc * cabs1(y));
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
            } else if (METHOD == QSX_METHOD_DOPRI5) {
                // Dormand-Prince 5(4) on the whole hierarchy: 7 tile sweeps per step attempt,
                // error norm through atomics + grid barrier, step size kept in registers
                // (identical on every thread).  TA = V, TB = W.
                cplx *TA = a.V, *TB = a.W;
                cplx *K1 = a.K[0], *K2 = a.K[1], *K3 = a.K[2], *K4 = a.K[3], *K5 = a.K[4], *K6 = a.K[5], *K7 = a.K[6];
                const long long ntot = Dp * B;
                const double dir = span >= 0 ? 1.0 : -1.0;
                auto sweep = [&](const cplx *src, double tt, auto &&post) {
                    s.t_eval = tt;
                    for (long long w = w0; w < total; w += wstride) {
                        const int b = (int)(w / n_tiles);
                        const long long tile = w % n_tiles;
                        unit_set(s, w, wstride, total, n_tiles);
                        const int member = a.member_of ? a.member_of[b] : 0;
                        const size_t o = (size_t)b * Dp;
                        Tile::run(a.H, s, src + o, tile, member, [&](long long) { return cmake(0, 0); },
                                  [&](long long i, cplx f, cplx, cplx, double) { post(o + i, f); });
                    }
                    n_rhs += 1;
                    grid.sync();
                };
                if (!dp_have_k1) {
                    sweep(a.Y, tcur, [&](size_t i, cplx f) { K1[i] = f; });
                    dp_have_k1 = true;
                }
                if (dp_h == 0.0) dp_h = fmin(fabs(span), 0.05 / fmax(a.lnorm, 1e-300));
                int guard = 0;
                while ((target - tcur) * dir > 0) {
                    double h = dir * fabs(dp_h);
                    bool clipped = false;
                    if ((tcur + h - target) * dir >= 0 || fabs(target - (tcur + h)) < 1e-12 * fabs(h)) {
                        h = target - tcur;
                        clipped = true;
                    }
                    for (long long i = gtid; i < ntot; i += gsz) TA[i] = cadd(a.Y[i], cscale(h * DP_A21, K1[i]));
                    if (blockIdx.x == 0 && threadIdx.x == 0) a.red[(dp_slot + 1) & 3] = 0.0;
                    grid.sync();
                    sweep(TA, tcur + DP_C2 * h, [&](size_t i, cplx f) {
                        K2[i] = f;
                        cplx v = a.Y[i]; rfma(v, h * DP_A31, K1[i]); rfma(v, h * DP_A32, f); TB[i] = v;
                    });
                    sweep(TB, tcur + DP_C3 * h, [&](size_t i, cplx f) {
                        K3[i] = f;
                        cplx v = a.Y[i]; rfma(v, h * DP_A41, K1[i]); rfma(v, h * DP_A42, K2[i]); rfma(v, h * DP_A43, f); TA[i] = v;
                    });
                    sweep(TA, tcur + DP_C4 * h, [&](size_t i, cplx f) {
                        K4[i] = f;
                        cplx v = a.Y[i]; rfma(v, h * DP_A51, K1[i]); rfma(v, h * DP_A52, K2[i]);
                        rfma(v, h * DP_A53, K3[i]); rfma(v, h * DP_A54, f); TB[i] = v;
                    });
                    sweep(TB, tcur + DP_C5 * h, [&](size_t i, cplx f) {
                        K5[i] = f;
                        cplx v = a.Y[i]; rfma(v, h * DP_A61, K1[i]); rfma(v, h * DP_A62, K2[i]);
                        rfma(v, h * DP_A63, K3[i]); rfma(v, h * DP_A64, K4[i]); rfma(v, h * DP_A65, f); TA[i] = v;
                    });
                    sweep(TA, tcur + h, [&](size_t i, cplx f) {
                        K6[i] = f;
                        cplx v = a.Y[i]; rfma(v, h * DP_A71, K1[i]); rfma(v, h * DP_A73, K3[i]);
                        rfma(v, h * DP_A74, K4[i]); rfma(v, h * DP_A75, K5[i]); rfma(v, h * DP_A76, f); TB[i] = v;
                    });
                    double errsq = 0.0;
                    s.t_eval = tcur + h;
                    for (long long w = w0; w < total; w += wstride) {
                        const int b = (int)(w / n_tiles);
                        const long long tile = w % n_tiles;
                        unit_set(s, w, wstride, total, n_tiles);
                        const int member = a.member_of ? a.member_of[b] : 0;
                        const size_t o = (size_t)b * Dp;
                        Tile::run(a.H, s, TB + o, tile, member, [&](long long) { return cmake(0, 0); },
                                  [&](long long il, cplx f, cplx ynew, cplx, double) {
                                      const size_t i = o + il;
                                      K7[i] = f;
                                      cplx e = cmake(0, 0);
                                      rfma(e, DP_E1, K1[i]); rfma(e, DP_E3, K3[i]); rfma(e, DP_E4, K4[i]);
                                      rfma(e, DP_E5, K5[i]); rfma(e, DP_E6, K6[i]); rfma(e, DP_E7, f);
                                      const double sc = a.atol + a.rtol * sqrt(fmax(cabs2(a.Y[i]), cabs2(ynew)));
                                      errsq += (h * h) * cabs2(e) / (sc * sc);
                                  });
                    }
                    n_rhs += 1;
                    errsq = warp_sum(errsq);
                    if ((threadIdx.x & 31) == 0 && errsq != 0.0) atomicAdd(&a.red[dp_slot & 3], errsq);
                    grid.sync();
                    const double err = sqrt(__ldcg(&a.red[dp_slot & 3]) / (double)((long long)B * a.H.n_ado * M));
                    dp_slot += 1;
                    const bool finite = (err == err) && err < 1e300;
                    double fac;
                    if (finite && err <= 1.0) {
                        for (long long i = gtid; i < ntot; i += gsz) { a.Y[i] = TB[i]; K1[i] = K7[i]; }
                        grid.sync();
                        tcur = clipped ? target : tcur + h;
                        n_steps += 1;
                        fac = (err < 1e-10) ? 5.0 : fmin(5.0, fmax(0.2, 0.9 * pow(err, -0.2)));
                        if (!clipped || fac < 1.0) dp_h = fabs(h) * fac;
                        else dp_h = fmax(fabs(dp_h), fabs(h) * fmin(fac, 1.0));
                    } else {
                        fac = finite ? fmax(0.2, 0.9 * pow(err, -0.2)) : 0.2;
                        dp_h = fabs(h) * fmin(fac, 1.0);
                    }
                    if (dp_h < 1e-14 * fmax(1.0, fabs(tcur)) || ++guard > 1000000) {
                        status = QSX_ERR_INTEGRATOR;
                        break;
                    }
                }
            } else {
                // classic RK4 with fixed sub-steps; ACC = V, TA = W, TB = X
                cplx *ACC = a.V, *TA = a.W, *TB = a.X;
                const int nsub = a.rk4_sub > 0 ? a.rk4_sub : 1;
                const double h = span / nsub;
                for (int sub = 0; sub < nsub; ++sub) {
                    for (int stage = 0; stage < 4; ++stage) {
                        const cplx *src = stage == 0 ? a.Y : (stage == 2 ? TB : TA);
                        for (long long w = w0; w < total; w += wstride) {
                            const int b = (int)(w / n_tiles);
                            const long long tile = w % n_tiles;
                            unit_set(s, w, wstride, total, n_tiles);
                            const int member = a.member_of ? a.member_of[b] : 0;
                            const size_t o = (size_t)b * Dp;
                            cplx *Yb = a.Y + o, *Ab = ACC + o, *TAb = TA + o, *TBb = TB + o;
                            if (stage == 0)
                                Tile::run(a.H, s, src + o, tile, member,
                                          [&](long long) { return cmake(0, 0); },
                                          [&](long long i, cplx k1, cplx own, cplx, double) {
                                    TAb[i] = cadd(own, cscale(0.5 * h, k1));
                                    Ab[i] = cadd(own, cscale(h / 6.0, k1));
                                });
                            else if (stage == 1)
                                Tile::run(a.H, s, src + o, tile, member,
                                          [&](long long i) { return Yb[i]; },
                                          [&](long long i, cplx k2, cplx, cplx y, double) {
                                    TBb[i] = cadd(y, cscale(0.5 * h, k2));
                                    Ab[i] = cadd(Ab[i], cscale(h / 3.0, k2));
                                });
                            else if (stage == 2)
                                Tile::run(a.H, s, src + o, tile, member,
                                          [&](long long i) { return Yb[i]; },
                                          [&](long long i, cplx k3, cplx, cplx y, double) {
                                    TAb[i] = cadd(y, cscale(h, k3));
                                    Ab[i] = cadd(Ab[i], cscale(h / 3.0, k3));
                                });
                            else
                                Tile::run(a.H, s, src + o, tile, member,
                                          [&](long long i) { return Ab[i]; },
                                          [&](long long i, cplx k4, cplx, cplx acc, double) {
                                    Yb[i] = cadd(acc, cscale(h / 6.0, k4));
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
    QSX_REQUIRE(tb.n_ado < ((int64_t)1 << 31) - 64, "hierarchy too large");
    tb.enumerate();
    const int64_t n_ado = tb.n_ado;
    const int64_t n_tiles = (n_ado + TL - 1) / TL;

    // ---- the Liouville subspace must be one rectangular block rows x cols -----
    std::vector<int> ea(M), eb(M), rows, cols;
    for (int e = 0; e < M; ++e) {
        int64_t f = cfg->subspace_index[e];
        QSX_REQUIRE(f >= 0 && f < (int64_t)N * N, "subspace index out of range");
        ea[e] = (int)(f % N);
        eb[e] = (int)(f / N);
        rows.push_back(ea[e]);
        cols.push_back(eb[e]);
    }
    std::sort(rows.begin(), rows.end()); rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    std::sort(cols.begin(), cols.end()); cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    const int nr = (int)rows.size(), nc = (int)cols.size();
    bool rect = (nr * nc == M);
    for (int e = 0; e < M && rect; ++e) rect = (ea[e] == rows[e % nr] && eb[e] == cols[e / nr]);
    if (!rect) {
        qsx_set_error("HEOM kernel needs a rectangular Liouville block (rows x cols), e.g. 'ee', 'eg', "
                      "'fe' or 'gg,ge,eg,ee'; propagate the blocks of a union separately");
        return QSX_ERR_UNSUPPORTED;
    }
    const zc *Hm = reinterpret_cast<const zc *>(cfg->H);
    const zc *cc = reinterpret_cast<const zc *>(cfg->c);
    const double *v = cfg->coupling_diag;
    const zc mi(0.0, -1.0);

    // ---- Hs_R and (Hs_C)^T per member, pre-scaled by -i u -----------------------
    // Schroedinger: d rho = Hs_R rho - rho Hs_C; Heisenberg (generator transposed):
    // d X = Hs_R^T X - X Hs_C^T.  The kernel wants A_R[a][c] and A_C[b][c] with
    // (rho Hs_C)[a][b] = sum_c A_C[b][c] rho[a][c]  ->  A_C = Hs_C^T.
    std::vector<cplx> HR((size_t)cfg->n_members * nr * nr), HC((size_t)cfg->n_members * nc * nc);
    double hnorm = 0;
    for (int m = 0; m < cfg->n_members; ++m) {
        const zc *H = Hm + (size_t)m * N * N;
        double rs_max = 0, cs_max = 0;
        for (int a = 0; a < nr; ++a) {
            double rs = 0;
            for (int c = 0; c < nr; ++c) {
                zc val = mi * u * (cfg->heisenberg ? H[rows[c] * N + rows[a]] : H[rows[a] * N + rows[c]]);
                HR[((size_t)m * nr + a) * nr + c] = cmake(val.real(), val.imag());
                rs += std::abs(val);
            }
            rs_max = std::max(rs_max, rs);
        }
        for (int b = 0; b < nc; ++b) {
            double cs = 0;
            for (int c = 0; c < nc; ++c) {
                // A_C[b][c] = Hs_C[c][b] (Schroedinger) or Hs_C^T[c][b] = Hs_C[b][c] (Heisenberg)
                zc val = mi * u * (cfg->heisenberg ? H[cols[b] * N + cols[c]] : H[cols[c] * N + cols[b]]);
                HC[((size_t)m * nc + b) * nc + c] = cmake(val.real(), val.imag());
                cs += std::abs(val);
            }
            cs_max = std::max(cs_max, cs);
        }
        hnorm = std::max(hnorm, rs_max + cs_max);
    }
    std::vector<double> dterm(M, 0.0);
    for (int e = 0; e < M; ++e) {
        double dbl = 0;
        for (int j = 0; j < cfg->n_sites; ++j) {
            double va = v[j * N + ea[e]], vb = v[j * N + eb[e]];
            dbl += va + vb - 2 * va * vb;
        }
        dterm[e] = u * cfg->temp_corr * dbl;
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
    // ---- balancing similarity transform (error norm and sub-step count) ---------
    // The raw hierarchy couples level n to n+1 with O(1) and n+1 to n with O(n |c_k|)
    // coefficients, so its inf-norm grossly over-estimates the spectral radius.  With
    // s_n = prod_b prod_{m<n_b} 1/rho_k(m), rho_k(m) = sqrt(|down(m+1)| / |up(m)|), the
    // transformed generator S L S^-1 has symmetric couplings; Taylor sub-steps are sized
    // by its norm and truncation errors are measured in the norm |S y|_inf.  (For
    // modified_HEOM rho == 1: that option is exactly this rescaling, heom.py:423-437.)
    std::vector<double> mu(K1, 0.0), md(K1, 0.0);
    for (int e = 0; e < M; ++e)
        for (size_t l = 0; l < lpat[e].size(); ++l) {
            int k = lpat[e][l] % K1;
            mu[k] = std::max(mu[k], std::abs(lgu[e][l]));
            md[k] = std::max(md[k], std::abs(lgd[e][l]));
        }
    std::vector<double> rho((size_t)K1 * Lc, 1.0);
    for (int k = 0; k < K1; ++k)
        for (int n = 0; n + 1 < Lc; ++n) {
            double upc = su[(size_t)k * Lc + n] * mu[k], dnc = sd[(size_t)k * Lc + n + 1] * md[k];
            rho[(size_t)k * Lc + n] = (upc > 0 && dnc > 0) ? sqrt(dnc / upc) : 1.0;
        }
    // ---- per-ADO tables in tile layout [tile][bin][32] ---------------------------
    std::vector<double> shift((size_t)n_tiles * TL, 0.0), scale((size_t)n_tiles * TL, 1.0);
    std::vector<int> up((size_t)n_tiles * bins * TL, -1), down((size_t)n_tiles * bins * TL, -1);
    std::vector<uint8_t> occ((size_t)n_tiles * bins * TL, 0);
    std::vector<int> off_up((size_t)n_tiles * bins * TL, -1), off_dn((size_t)n_tiles * bins * TL, -1);
    double lnorm = 0;
    for (int64_t n = 0; n < n_ado; ++n) {
        double sft = 0, logs = 0;
        const int64_t tile = n / TL;
        const int lane = (int)(n % TL);
        for (int b = 0; b < bins; ++b) {
            const int njk = tb.index[(size_t)n * bins + b];
            sft += njk * cfg->nu[b % K1];
            for (int m = 0; m < njk; ++m) logs -= log(rho[(size_t)(b % K1) * Lc + m]);
            const size_t o = ((size_t)tile * bins + b) * TL + lane;
            up[o] = tb.up[(size_t)n * bins + b];
            down[o] = tb.down[(size_t)n * bins + b];
            occ[o] = (uint8_t)njk;
            if (up[o] >= 0) off_up[o] = (int)(((int64_t)(up[o] / TL) * M) * TL + up[o] % TL);
            if (down[o] >= 0) off_dn[o] = (int)(((int64_t)(down[o] / TL) * M) * TL + down[o] % TL);
        }
        shift[n] = u * sft;
        scale[n] = exp(logs);
    }
    // inf-norm of the balanced generator: max over (n, e) of the absolute row sum
    for (int64_t n = 0; n < n_ado; ++n)
        for (int e = 0; e < M; ++e) {
            double rs = hnorm + fabs(shift[n]) + fabs(dterm[e]);
            for (size_t l = 0; l < lpat[e].size(); ++l) {
                int b = lpat[e][l], k = b % K1, njk = tb.index[(size_t)n * bins + b];
                if (tb.up[(size_t)n * bins + b] >= 0)
                    rs += su[(size_t)k * Lc + njk] * std::abs(lgu[e][l]) * rho[(size_t)k * Lc + njk];
                if (tb.down[(size_t)n * bins + b] >= 0)
                    rs += sd[(size_t)k * Lc + njk] * std::abs(lgd[e][l]) / rho[(size_t)k * Lc + njk - 1];
            }
            lnorm = std::max(lnorm, rs);
        }
    h->lnorm = lnorm;

    QSX_CUDA(h->occ.upload(occ, stream));
    QSX_CUDA(h->up.upload(up, stream));
    QSX_CUDA(h->down.upload(down, stream));
    QSX_CUDA(h->off_up.upload(off_up, stream));
    QSX_CUDA(h->off_dn.upload(off_dn, stream));
    QSX_CUDA(h->shift.upload(shift, stream));
    QSX_CUDA(h->scale.upload(scale, stream));
    QSX_CUDA(h->HR.upload(HR, stream));
    QSX_CUDA(h->HC.upload(HC, stream));
    QSX_CUDA(h->dterm.upload(dterm, stream));
    QSX_CUDA(h->lbin.upload(lbin, stream));
    QSX_CUDA(h->gu.upload(gu, stream));
    QSX_CUDA(h->gd.upload(gd, stream));
    QSX_CUDA(h->su.upload(su, stream));
    QSX_CUDA(h->sd.upload(sd, stream));
    QSX_CUDA(cudaStreamSynchronize(stream));   // host vectors go out of scope

    HeomDev &d = h->d;
    d.nr = nr; d.nc = nc; d.M = M; d.bins = bins; d.K1 = K1; d.Lc = Lc; d.Lk = Lk;
    d.n_members = cfg->n_members; d.n_ado = n_ado; d.n_tiles = n_tiles;
    d.shift = h->shift.p; d.scale = h->scale.p; d.up = h->up.p; d.down = h->down.p; d.occ = h->occ.p;
    d.off_up = h->off_up.p; d.off_dn = h->off_dn.p;
    d.heis = cfg->heisenberg ? 1 : 0;
    d.top_tile = (tb.level_offset[Lc - 1] + TL - 1) / TL;
    d.n_pulse = 0; d.Rp = 0; d.pcol = nullptr; d.pval = nullptr;
    // electronic-block structure: row state a <-> site a, column state b <-> site b
    d.ee = (nr == cfg->n_sites && nc == cfg->n_sites && K1 <= 4 &&
            (int64_t)n_tiles * M * TL < ((int64_t)1 << 31));
    for (int j = 0; j < cfg->n_sites && d.ee; ++j)
        for (int a = 0; a < nr && d.ee; ++a)
            d.ee = (v[j * N + rows[a]] == (j == a ? 1.0 : 0.0)) && (v[j * N + cols[a]] == (j == a ? 1.0 : 0.0));
    for (int k = 0; k < K1 && k < 4; ++k) {
        zc guR = mi * u * (cfg->heisenberg ? cc[k] : zc(1.0));
        zc gdR = mi * u * (cfg->heisenberg ? zc(1.0) : cc[k]);
        zc guC = mi * u * (cfg->heisenberg ? -std::conj(cc[k]) : zc(-1.0));
        zc gdC = mi * u * (cfg->heisenberg ? zc(-1.0) : -std::conj(cc[k]));
        d.GuR[k] = cmake(guR.real(), guR.imag()); d.GdR[k] = cmake(gdR.real(), gdR.imag());
        d.GuC[k] = cmake(guC.real(), guC.imag()); d.GdC[k] = cmake(gdC.real(), gdC.imag());
    }
    d.real_h = 1;
    for (auto &z : HR) if (z.x != 0.0) d.real_h = 0;
    for (auto &z : HC) if (z.x != 0.0) d.real_h = 0;
    {
        std::vector<int> e_off(M), e_stride(M);
        for (int e = 0; e < M; ++e) { e_off[e] = e * TL; e_stride[e] = 1; }
        QSX_CUDA(h->e_off.upload(e_off, stream));
        QSX_CUDA(h->e_stride.upload(e_stride, stream));
        QSX_CUDA(cudaStreamSynchronize(stream));
        d.e_off = h->e_off.p; d.e_stride = h->e_stride.p;
    }
    d.HR = h->HR.p; d.HC = h->HC.p; d.dterm = h->dterm.p; d.lbin = h->lbin.p;
    d.gu = h->gu.p; d.gd = h->gd.p; d.su = h->su.p; d.sd = h->sd.p;
    size_t offs[14];
    int dev = 0, smem_limit = 0;
    QSX_CUDA(cudaGetDevice(&dev));
    QSX_CUDA(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (tile_smem_layout(d, offs) > (size_t)smem_limit) {
        qsx_set_error("HEOM subspace dimension %d too large for the tile kernel", M);
        return QSX_ERR_UNSUPPORTED;
    }
    // ---- row tile (heom_row.cuh): sigma_n = g_n rho_n variables, tile records -----------
    {
        // block structure: both sides of the block are the same `nr` states, grouped per site
        // (vib states each: 1 for an electronic block, > 1 for a vibronic one), every coupling
        // operator the projector onto its site's states; real symmetric H
        const int vib = (cfg->n_sites > 0 && nr % cfg->n_sites == 0) ? nr / cfg->n_sites : 0;
        bool ok = vib > 0 && d.real_h && !cfg->heisenberg && nr == nc && K1 == 2 &&
                  (int64_t)n_tiles * M * TL < ((int64_t)1 << 31);
        for (int a = 0; a < nr && ok; ++a) ok = rows[a] == cols[a];
        for (int j = 0; j < cfg->n_sites && ok; ++j)
            for (int a = 0; a < nr && ok; ++a) ok = v[j * N + rows[a]] == (a / vib == j ? 1.0 : 0.0);
        for (int m = 0; m < cfg->n_members && ok; ++m)
            for (int a = 0; a < nr && ok; ++a)
                for (int c = 0; c < nr && ok; ++c)
                    ok = HR[((size_t)m * nr + a) * nr + c].y == HR[((size_t)m * nr + c) * nr + a].y &&
                         HC[((size_t)m * nr + a) * nr + c].y == HR[((size_t)m * nr + a) * nr + c].y;
        auto build_row = [&](auto C_) -> int {
            typedef decltype(C_) C;
            static_assert(C::M <= 64, "RowDev::hc holds 64 coefficients");
            heom_row::RowDev &r = h->row;
            std::vector<unsigned char> rec((size_t)n_tiles * C::REC_BYTES, 0);
            std::vector<double> g((size_t)n_tiles * TL, 0.0);
            for (int64_t n = 0; n < n_tiles * TL; ++n) {
                const int64_t tile = n / TL;
                const int lane = (int)(n % TL);
                unsigned char *base = rec.data() + (size_t)tile * C::REC_BYTES;
                int *dn = reinterpret_cast<int *>(base + C::OFF_DN) + lane * C::LD;
                int *up_ = reinterpret_cast<int *>(base + C::OFF_UP) + lane * C::LD;
                double *upc = reinterpret_cast<double *>(base + C::OFF_UPC) + lane * C::UD;
                for (int b = 0; b < C::LD; ++b) { dn[b] = -1; up_[b] = -1; }
                if (n >= n_ado) continue;
                double logg = 0;
                for (int b = 0; b < bins; ++b) {
                    const int njk = tb.index[(size_t)n * bins + b];
                    const size_t o = ((size_t)tile * bins + b) * TL + lane;
                    dn[b] = off_dn[o];
                    up_[b] = off_up[o];
                    upc[b] = u * (njk + 1);
                    const double ck = std::abs(cc[b % K1]);
                    logg += cfg->modified ? 0.5 * (njk * log(ck) - lgamma(njk + 1.0)) : -lgamma(njk + 1.0);
                }
                g[n] = exp(logg);
                reinterpret_cast<double *>(base + C::OFF_SHIFT)[lane] = shift[n];
                reinterpret_cast<double *>(base + C::OFF_SCALE)[lane] = scale[n] / g[n];
            }
            // last tile any ADO of a tile links to (links are symmetric: it is also the last tile that reads it)
            std::vector<int> dep((size_t)n_tiles);
            for (int64_t tile = 0; tile < n_tiles; ++tile) {
                int hi = (int)tile;
                for (int b = 0; b < bins; ++b)
                    for (int lane = 0; lane < TL; ++lane) {
                        const size_t o = ((size_t)tile * bins + b) * TL + lane;
                        if (off_dn[o] >= 0) hi = std::max(hi, (int)(off_dn[o] / (M * TL)));
                        if (off_up[o] >= 0) hi = std::max(hi, (int)(off_up[o] / (M * TL)));
                    }
                dep[tile] = hi;
            }
            QSX_CUDA(h->row_dep.upload(dep, stream));
            std::vector<double> hm((size_t)cfg->n_members * C::MH, 0.0);
            for (int m = 0; m < cfg->n_members; ++m)
                for (int i = 0; i < M; ++i) hm[(size_t)m * C::MH + i] = HR[(size_t)m * M + i].y;
            std::vector<cplx> ainv(sizeof(qsx_taylor_ainv) / sizeof(qsx_taylor_ainv[0]));
            for (size_t i = 0; i < ainv.size(); ++i) ainv[i] = cmake(qsx_taylor_ainv[i][0], qsx_taylor_ainv[i][1]);
            QSX_CUDA(h->row_rec.upload(rec, stream));
            QSX_CUDA(h->row_g.upload(g, stream));
            QSX_CUDA(h->row_h.upload(hm, stream));
            QSX_CUDA(h->row_ainv.upload(ainv, stream));
            QSX_CUDA(cudaStreamSynchronize(stream));
            r.n_members = cfg->n_members;
            r.n_ado = n_ado; r.n_tiles = n_tiles; r.top_tile = d.top_tile;
            r.rec = h->row_rec.p; r.hmem = h->row_h.p; r.gscale = h->row_g.p;
            for (int i = 0; i < 64; ++i) r.hc[i] = i < M ? hm[i] : 0.0;
            for (int k = 0; k < 4; ++k) r.cd[k] = k < K1 ? d.GdR[k] : cmake(0, 0);
            r.d2 = 2.0 * u * cfg->temp_corr;
            r.const_h = cfg->n_members == 1;
            r.dbg = 0;
            r.stream_rec = (size_t)n_tiles * C::REC_BYTES >= ((size_t)8 << 20);
            return QSX_OK;
        };
        int rc_row = QSX_OK;
        if (ok && nr == 7 && vib == 1) { rc_row = build_row(heom_row::Cfg<7, 2>()); h->row_cfg = 1; }
        else if (ok && nr == 8 && vib == 4) { rc_row = build_row(heom_row::Cfg<8, 2, 4>()); h->row_cfg = 2; }
        if (rc_row) return rc_row;
        h->row_ok = h->row_cfg != 0;
    }
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

static int upload_members(DevBuf<int> &buf, const int32_t *host, int n, int n_members, cudaStream_t s) {
    std::vector<int> m(host, host + n);
    for (int x : m) QSX_REQUIRE(x >= 0 && x < n_members, "member index out of range");
    QSX_CUDA(buf.upload(m, s));
    return QSX_OK;
}

// ---- tile selection ---------------------------------------------------------------
// QSX_HEOM_VARIANT (tests / A-B runs): 'r' row tile also for the smallest hierarchies, 'b' batch tile
// (TileEE), 'g' generic tile.  Default: the row tile from 64 (column, tile) units on -- measured on
// FMO 'ee' (us per RHS, batch tile vs row tile): depth 4 (22 tiles) 6.5 / 6.2, depth 6 (364) 14.9 / 10.3,
// depth 7 (1212) 41.2 / 20.8 -- otherwise the batch tile.
static char heom_variant() {
    const char *v = getenv("QSX_HEOM_VARIANT");
    return v ? v[0] : ' ';
}
static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}
static bool use_row_tile(const qsx_heom_s *h, long long units) {
    const char v = heom_variant();
    if (!h->row_ok || v == 'b' || v == 'g') return false;
    // the vibronic block has no batch tile: its alternative is the table-driven generic tile
    return v == 'r' || units >= 64 || h->row_cfg == 2;
}

// Row-tile launch configurations.  FMO-like blocks: two tile buffers per CTA, two CTAs per SM at
// 128 registers (three buffers leave too little L1 for the local-memory traffic and ran 25 % slower;
// three CTAs do not fit 128 registers).  Vibronic-dimer blocks (8 x 8 matrices: 16 more registers
// per gather batch) run one CTA per SM without a register cap.
template <class Fn>
static int row_dispatch(const qsx_heom_s *h, bool const_h, Fn &&fn) {
    typedef std::integral_constant<int, 1> I1;
    typedef std::integral_constant<int, 2> I2;
    if (h->row_cfg == 2) {
        typedef heom_row::Cfg<8, 2, 4> C;
        return const_h ? fn(C(), std::true_type(), I2(), I1()) : fn(C(), std::false_type(), I2(), I1());
    }
    typedef heom_row::Cfg<7, 2> C;
    return const_h ? fn(C(), std::true_type(), I2(), I2()) : fn(C(), std::false_type(), I2(), I2());
}

extern "C" int qsx_heom_apply(qsx_heom_t h, const void *y_dev, void *dy_dev, int32_t n_columns,
                              const int32_t *member_host, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(h && y_dev && dy_dev && n_columns > 0, "qsx_heom_apply: bad arguments");
    const HeomDev &d = h->d;
    const long long Dp = d.n_tiles * d.M * TL;
    DevBuf<int> member;
    DevBuf<cplx> xi, yi;
    int rc;
    if (member_host && (rc = upload_members(member, member_host, n_columns, d.n_members, stream))) return rc;
    QSX_CUDA(xi.alloc((size_t)n_columns * Dp));
    QSX_CUDA(yi.alloc((size_t)n_columns * Dp));
    int dev = 0, sms = 148;
    QSX_CUDA(cudaGetDevice(&dev));
    QSX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long total = d.n_tiles * n_columns;
    const bool row = use_row_tile(h, total);
    heom_to_internal<<<sms * 4, 256, 0, stream>>>((const cplx *)y_dev, xi.p, n_columns, d.n_ado, d.n_tiles, d.M,
                                                   d.e_off, d.e_stride, row ? h->row.gscale : nullptr);
    if (row) {
        QSX_REQUIRE(total < ((long long)1 << 31), "too many (column, tile) units");
        heom_row::RowApplyArgs a;
        a.R = h->row; a.R.dbg = env_int("QSX_ROW_DBGMASK", 0); a.x = xi.p; a.y = yi.p; a.member_of = member_host ? member.p : nullptr; a.B = n_columns;
        rc = row_dispatch(h, h->row.const_h && !member_host, [&](auto C_, auto CH, auto NB, auto MB) -> int {
            typedef decltype(C_) C;
            auto kernel = heom_row::heom_row_apply_kernel<C, decltype(CH)::value, decltype(NB)::value, decltype(MB)::value>;
            const size_t smem = C::smem_bytes(decltype(NB)::value);
            int per_sm = 0;
            QSX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            QSX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, C::THREADS, smem));
            QSX_REQUIRE(per_sm > 0, "heom_row_apply_kernel does not fit on an SM");
            int grid = (int)std::min<long long>(total, (long long)sms * per_sm);
            if (const char *gs = getenv("QSX_HEOM_GRID")) grid = std::min(grid, std::max(1, atoi(gs)));   // tests / experiments
            kernel<<<grid, C::THREADS, smem, stream>>>(a);
            if (const int reps = env_int("QSX_HEOM_TIME", 0)) {      // diagnostics: mean time of the bare kernel
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0, stream);
                for (int r = 0; r < reps; ++r) kernel<<<grid, C::THREADS, smem, stream>>>(a);
                cudaEventRecord(e1, stream);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                cudaEventDestroy(e0); cudaEventDestroy(e1);
                fprintf(stderr, "heom_row_apply: grid %d, %d CTA/SM, smem %zu: %.2f us per launch\n", grid, per_sm, smem,
                        1e3 * ms / reps);
            }
            return QSX_OK;
        });
        if (rc) return rc;
    } else {
        HeomApplyArgs a;
        a.H = d; a.x = xi.p; a.y = yi.p; a.member_of = member_host ? member.p : nullptr; a.B = n_columns;
        size_t smem;
        int grid, apply_occ = 0;
#define QSX_APPLY(TILE)                                                                                  \
    {                                                                                                    \
        typedef TILE T;                                                                                  \
        smem = T::smem_bytes(d);                                                                         \
        QSX_CUDA(cudaFuncSetAttribute(heom_apply_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                       \
        QSX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&apply_occ, heom_apply_kernel<T>,         \
                                                               T::THREADS, smem));                       \
        grid = (int)std::min<long long>((total + T::UNITS - 1) / T::UNITS,                               \
                                        (long long)sms * std::max(1, apply_occ));                        \
        heom_apply_kernel<T><<<grid, T::THREADS, smem, stream>>>(a);                                     \
    }
        const bool ee7 = d.ee && d.nr == 7 && d.K1 == 2;
        const char vsel = heom_variant();
        if (ee7 && d.real_h && vsel != 'g') QSX_APPLY(TileEE<7 COMMA 2 COMMA 1 COMMA true COMMA true>)
        else if (ee7 && !d.real_h && vsel != 'g') QSX_APPLY(TileEE<7 COMMA 2 COMMA 2 COMMA false>)
        else QSX_APPLY(TileGeneric)
#undef QSX_APPLY
    }
    heom_from_internal<<<sms * 4, 256, 0, stream>>>(yi.p, (cplx *)dy_dev, n_columns, d.n_ado, d.n_tiles, d.M,
                                                     d.e_off, d.e_stride, row ? h->row.gscale : nullptr);
    qsx_launch_counter += 3;
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
    const long long Dp = d.n_tiles * M * TL;
    QSX_REQUIRE(B > 0 && nt > 0 && args->t_host && args->y0_dev && args->out_dev,
                "qsx_heom_propagate: empty batch or missing buffers");
    QSX_REQUIRE(args->method == QSX_METHOD_TAYLOR || args->method == QSX_METHOD_RK4 ||
                args->method == QSX_METHOD_DOPRI5 || args->method == QSX_METHOD_POLY,
                "HEOM propagation supports taylor, poly, rk4 and dopri5");
    QSX_REQUIRE(args->n_pulses >= 0 && args->n_pulses <= QSX_MAX_PULSES, "too many pulses");
    QSX_REQUIRE(!(args->n_pulses > 0 && (args->method == QSX_METHOD_TAYLOR || args->method == QSX_METHOD_POLY)),
                "Taylor / product-form propagation needs a time-independent generator");
    const bool dopri = args->method == QSX_METHOD_DOPRI5;
    for (int i = 1; i < nt; ++i)
        QSX_REQUIRE(args->t_host[i] >= args->t_host[i - 1], "output times must be non-decreasing");
    QSX_REQUIRE(args->t_host[0] >= args->t0, "first output time precedes t0");

    const long long total = d.n_tiles * B;
    const bool lti = args->method == QSX_METHOD_TAYLOR || args->method == QSX_METHOD_POLY;
    const bool row = lti && args->n_pulses == 0 && use_row_tile(h, total);
    // the product form lives in the row kernel; other tiles run the adaptive Taylor series instead
    const int method = (args->method == QSX_METHOD_POLY && !row) ? (int)QSX_METHOD_TAYLOR : args->method;

    DevBuf<int> member, flags;
    DevBuf<double> d_t, ynorm;
    DevBuf<unsigned long long> stats;
    DevBuf<cplx> Y, V, W, X;
    int rc;
    if (args->generator_of_column_host &&
        (rc = upload_members(member, args->generator_of_column_host, B, d.n_members, stream)))
        return rc;
    QSX_CUDA(d_t.upload(args->t_host, nt, stream));
    QSX_CUDA(flags.alloc(3));
    QSX_CUDA(ynorm.alloc((size_t)3 * B));
    QSX_CUDA(stats.alloc(5));
    QSX_CUDA(cudaMemsetAsync(stats.p, 0, 5 * sizeof(unsigned long long), stream));
    QSX_CUDA(Y.alloc((size_t)B * Dp));
    QSX_CUDA(V.alloc((size_t)B * Dp));
    QSX_CUDA(W.alloc((size_t)B * Dp));
    if (method == QSX_METHOD_RK4) QSX_CUDA(X.alloc((size_t)B * Dp));
    DevBuf<cplx> Kbuf, pval;
    DevBuf<int> pcol;
    DevBuf<double> red;
    QSX_CUDA(red.alloc(4));
    if (dopri) QSX_CUDA(Kbuf.alloc((size_t)7 * B * Dp));

    long long saved_dim;
    DevBuf<int> save_of;
    if (args->save_mode == QSX_SAVE_MATRIX) {
        // per-ADO blocks [n_save][rows][M]: one shared matrix, or one per column through
        // save_of_column (e.g. the dipole operator of each polarisation configuration)
        QSX_REQUIRE(args->save_dev && args->save_rows > 0 && args->n_save >= 1 &&
                    (args->n_save == 1 || args->save_of_column_host),
                    "HEOM save matrices need save_of_column when n_save > 1");
        if (args->save_of_column_host) {
            std::vector<int> so(args->save_of_column_host, args->save_of_column_host + B);
            for (int x : so) QSX_REQUIRE(x >= 0 && x < args->n_save, "save index out of range");
            QSX_CUDA(save_of.upload(so, stream));
        }
        saved_dim = d.n_ado * args->save_rows;
    } else if (args->save_mode == QSX_SAVE_ADO0) {
        saved_dim = M;
    } else {
        QSX_REQUIRE(args->save_mode == QSX_SAVE_STATE, "bad save_mode");
        saved_dim = D;
    }
    const double rtol = args->rtol > 0 ? args->rtol : (dopri ? 1e-10 : 1e-13);
    int dev = 0, sms = 0, per_sm = 0;
    QSX_CUDA(cudaGetDevice(&dev));
    QSX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

    const void *kernel;
    int threads, grid;
    size_t smem;
    HeomPropArgs a;
    heom_row::RowPropArgs ra;
    void *kargs[1];
    if (row) {
        QSX_REQUIRE(total < ((long long)1 << 31), "too many (column, tile) units");
        ra.R = h->row; ra.R.dbg = 0;
        ra.B = B; ra.nt = nt;
        ra.use_flow = 1;      // decided below once the grid is known
        ra.member_of = args->generator_of_column_host ? member.p : nullptr;
        ra.y0 = (const cplx *)args->y0_dev;
        ra.Y = Y.p; ra.V = V.p; ra.W = W.p;
        ra.t = d_t.p; ra.t0 = args->t0; ra.rtol = rtol; ra.theta = 2.0; ra.lnorm = h->lnorm;
        ra.kmax = 60; ra.method = method;
        ra.repilot = std::max(1, env_int("QSX_HEOM_REPILOT", 24));
        ra.ainv = h->row_ainv.p;
        ra.save_mode = args->save_mode; ra.save_rows = args->save_rows;
        ra.S = (const cplx *)args->save_dev;
        ra.save_of = save_of.p;
        ra.out = (cplx *)args->out_dev; ra.saved_dim = saved_dim;
        ra.flags = flags.p; ra.ynorm = ynorm.p; ra.stats = stats.p;
        rc = row_dispatch(h, h->row.const_h && !args->generator_of_column_host, [&](auto C_, auto CH, auto NB, auto MB) -> int {
            typedef decltype(C_) C;
            kernel = (const void *)heom_row::heom_row_propagate_kernel<C, decltype(CH)::value, decltype(NB)::value, decltype(MB)::value>;
            threads = C::THREADS;
            smem = C::smem_bytes(decltype(NB)::value) + (size_t)env_int("QSX_HEOM_PADSMEM", 0);   // experiments: less L1
            return QSX_OK;
        });
        if (rc) return rc;
        kargs[0] = &ra;
    } else {
        a.H = d; a.B = B; a.nt = nt;
        for (int k = 0; k < 7; ++k) a.K[k] = dopri ? Kbuf.p + (size_t)k * B * Dp : nullptr;
        a.atol = args->atol > 0 ? args->atol : 1e-12;
        a.red = red.p;
        if (args->n_pulses > 0) {
            // per-ADO pulse operators: dense [n_pulses][M][M] on the device -> ELL rows
            QSX_REQUIRE(args->pulse_ops_dev && args->n_pulse_sets == 1,
                        "HEOM pulse operators must be one shared set of [n_pulses][M][M] matrices");
            const int np = args->n_pulses;
            std::vector<cplx> dense((size_t)np * M * M);
            qsx_d2h_counter += dense.size() * sizeof(cplx);
            QSX_CUDA(cudaMemcpyAsync(dense.data(), args->pulse_ops_dev, dense.size() * sizeof(cplx),
                                     cudaMemcpyDeviceToHost, stream));
            QSX_CUDA(cudaStreamSynchronize(stream));
            int Rp = 1;
            for (int p = 0; p < np; ++p)
                for (int e = 0; e < M; ++e) {
                    int nz = 0;
                    for (int c = 0; c < M; ++c) {
                        const cplx v = dense[((size_t)p * M + e) * M + c];
                        nz += (v.x != 0.0 || v.y != 0.0);
                    }
                    Rp = std::max(Rp, nz);
                }
            std::vector<int> hcol((size_t)np * M * Rp, -1);
            std::vector<cplx> hval((size_t)np * M * Rp, cmake(0, 0));
            for (int p = 0; p < np; ++p)
                for (int e = 0; e < M; ++e) {
                    int l = 0;
                    for (int c = 0; c < M; ++c) {
                        const cplx v = dense[((size_t)p * M + e) * M + c];
                        if (v.x != 0.0 || v.y != 0.0) {
                            hcol[((size_t)p * M + e) * Rp + l] = c;
                            hval[((size_t)p * M + e) * Rp + l] = v;
                            ++l;
                        }
                    }
                }
            QSX_CUDA(pcol.upload(hcol, stream));
            QSX_CUDA(pval.upload(hval, stream));
            QSX_CUDA(cudaStreamSynchronize(stream));
            a.H.n_pulse = np; a.H.Rp = Rp; a.H.pcol = pcol.p; a.H.pval = pval.p;
            for (int p = 0; p < np; ++p) a.H.pulses[p] = args->pulses[p];
        }
        a.member_of = args->generator_of_column_host ? member.p : nullptr;
        a.y0 = (const cplx *)args->y0_dev;
        a.Y = Y.p; a.V = V.p; a.W = W.p; a.X = X.p;
        a.t = d_t.p; a.t0 = args->t0;
        a.rtol = rtol;
        a.rk4_sub = args->rk4_substeps > 0 ? args->rk4_substeps : 16;
        a.kmax = 60; a.theta = 2.0; a.lnorm = h->lnorm;
        a.save_mode = args->save_mode; a.save_rows = args->save_rows;
        a.S = (const cplx *)args->save_dev;
        a.save_of = save_of.p;
        a.saved_dim = saved_dim;
        a.out = (cplx *)args->out_dev;
        a.flags = flags.p; a.ynorm = ynorm.p; a.stats = stats.p;
        const bool taylor = method == QSX_METHOD_TAYLOR;
        const char vsel = heom_variant();
#define QSX_PICK(TILE)                                                                        \
    {                                                                                         \
        typedef TILE T;                                                                       \
        threads = T::THREADS; smem = T::smem_bytes(d);                                        \
        kernel = taylor ? (const void *)heom_propagate_kernel<QSX_METHOD_TAYLOR, T>           \
                        : (const void *)heom_propagate_kernel<QSX_METHOD_RK4, T>;             \
    }
        const bool ee7 = d.ee && d.nr == 7 && d.K1 == 2;
        if (dopri || args->n_pulses > 0) {
            // time-dependent right-hand sides run on the generic tile (any rectangular block)
            typedef TileGeneric T;
            threads = T::THREADS; smem = T::smem_bytes(d);
            kernel = dopri ? (const void *)heom_propagate_kernel<QSX_METHOD_DOPRI5, T>
                           : (const void *)heom_propagate_kernel<QSX_METHOD_RK4, T>;
        }
        else if (ee7 && d.real_h && vsel != 'g') QSX_PICK(TileEE<7 COMMA 2 COMMA 1 COMMA true COMMA true>)
        else if (ee7 && !d.real_h && vsel != 'g') QSX_PICK(TileEE<7 COMMA 2 COMMA 2 COMMA false>)
        else QSX_PICK(TileGeneric)
#undef QSX_PICK
        kargs[0] = &a;
    }
    QSX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QSX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    QSX_REQUIRE(per_sm > 0, "heom_propagate_kernel does not fit on an SM");
    grid = (int)std::min<long long>(total, (long long)sms * per_sm);
    if (const char *gs = getenv("QSX_HEOM_GRID")) grid = std::min(grid, std::max(1, atoi(gs)));   // tests / experiments
    DevBuf<unsigned long long> flow_cnt;
    if (row) {
        // completion counters of the barrier-free product-form stages, one per round of `grid` units
        const size_t rounds = (size_t)((total + grid - 1) / grid);
        QSX_CUDA(flow_cnt.alloc(rounds + 1));
        QSX_CUDA(cudaMemsetAsync(flow_cnt.p, 0, (rounds + 1) * sizeof(unsigned long long), stream));
        ra.F.cnt = flow_cnt.p;
        ra.F.dep_hi = h->row_dep.p;
        ra.F.rot = (unsigned)(total % grid);      // the short last round moves on by its own length every stage
        // barrier-free stages pay off when a stage is a few rounds long (depth 8: 12.3 rounds, -12 %);
        // a 512-member depth-4 batch (38 rounds) runs 2 % faster behind grid barriers
        ra.use_flow = env_int("QSX_HEOM_FLOW", rounds <= 24 ? 1 : 0);
    }
    if (getenv("QSX_HEOM_VERBOSE"))
        fprintf(stderr, "heom_propagate: %s tile, threads %d smem %zu B, %d CTA/SM, grid %d\n",
                row ? "row" : "batch/generic", threads, smem, per_sm, grid);
    cudaEvent_t e0, e1;
    QSX_CUDA(cudaEventCreate(&e0));
    QSX_CUDA(cudaEventCreate(&e1));
    QSX_CUDA(cudaEventRecord(e0, stream));
    cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(threads), kargs, smem, stream);
    qsx_launch_counter += 1;
    if (e != cudaSuccess) {
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        qsx_set_error("heom_propagate launch: %s", cudaGetErrorString(e));
        return QSX_ERR_CUDA;
    }
    QSX_CUDA(cudaEventRecord(e1, stream));
    unsigned long long st[5] = {0, 0, 0, 0, 0};
    qsx_d2h_counter += sizeof(st);
    QSX_CUDA(cudaMemcpyAsync(st, stats.p, sizeof(st), cudaMemcpyDeviceToHost, stream));
    QSX_CUDA(cudaStreamSynchronize(stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    args->rhs_evaluations = st[0];
    args->accepted_steps = st[1];
    args->kernel_ms = ms;
    if (st[2] != 0 || st[4] != 0) {
        qsx_set_error("HEOM integration failed (Taylor series not converged within 60 terms, DOPRI5 "
                      "step-size underflow, or a non-finite state)");
        return QSX_ERR_INTEGRATOR;
    }
    return QSX_OK;
}

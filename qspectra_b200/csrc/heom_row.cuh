// K2 + K4 for electronic-block hierarchies (rows = cols = site-projector states, real
// Hamiltonian, Schroedinger picture): the row tile and its grid-resident integrators.
//
// Replaces HEOM_tensor + csr_matrix.dot (reference dynamics/heom.py:228-244, 298-443)
// for the stress configuration (FMO depth 8) and the depth-4 ensembles.
//
// Internal variables.  The kernel propagates sigma_n = g_n rho_n with
//     g_n = prod_jk 1/n_jk!                      (plain hierarchy, heom.py:411-421)
//     g_n = prod_jk sqrt(|c_k|^n_jk / n_jk!)     (modified_HEOM variables, heom.py:423-437)
// In these variables every down-link carries the occupation-independent coefficient
// -i u c_k (row site) / its conjugate (column site) and every up-link -+i u (n_jk + 1):
// tiles of the top level (two thirds of all tiles at depth 8) need no occupation data at
// all.  g is applied where states enter and leave the device layout.
//
// Tile = 32 consecutive ADOs, state layout [tile][element][32] as in heom.cu.  A CTA is NS
// warps; a thread owns one row of one ADO matrix.  Rows are paired: a warp holds rows
// (2p, 2p+1) of 16 ADOs -- lanes 0-15 row 2p, lanes 16-31 row 2p+1 of the same ADOs -- so the
// shared-memory reads of H sigma (every thread needs the whole matrix of its ADO) are the same
// address in both half-warps and cost half the wavefronts; an odd last row takes a full warp
// of 32 ADOs.  (L1TEX wavefronts bound this kernel: 84 % busy before the pairing.)  The source tile, its
// neighbour-offset record and (ensembles) the member's H are staged by bulk asynchronous
// copies (cp.async.bulk -> SASS UBLKCP) that complete on an mbarrier; buffers are released
// through a second mbarrier, so the row-warps never meet at a CTA barrier.  Hierarchy
// gathers are plain 16-byte loads of runs of consecutive ADOs.
//
// Integrators (one cooperative launch per trajectory, one grid barrier per stage):
//   * QSX_METHOD_POLY: exp(hL) y ~= T_m(hL) y = prod_j (I + h a_j L) y with a_j = -1/z_j the
//     reciprocal roots of the degree-m Taylor polynomial (taylor_roots.h).  Each stage reads
//     the state once and writes it once: 32 D bytes per RHS application, the algorithmic
//     minimum (the paired Taylor update of the round-1 kernel moved 48 D).  The degree is
//     taken from an adaptive Taylor pilot interval and refreshed periodically.
//   * QSX_METHOD_TAYLOR: adaptive-order Taylor series with paired accumulation (as heom.cu).
#pragma once
#include "common.cuh"
#include "taylor_roots.h"
#include <cooperative_groups.h>

namespace heom_row {
namespace cg = cooperative_groups;

// hierarchy gathers: L2-only loads by default (-DQSX_ROW_LDCA: allocate in L1 as well)
#ifdef QSX_ROW_LDCA
#define ROW_LD(p) (*(p))
#else
#define ROW_LD(p) __ldcg(p)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, int parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// the same with the L2 evict_first priority: lines nobody reads again (top-level tiles of the source)
__device__ __forceinline__ void bulk_g2s_stream(void *dst, const void *src, int bytes, uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 pol;\n\t"
                 "createpolicy.fractional.L2::evict_first.b64 pol, 1.0;\n\t"
                 "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], pol;\n\t}"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, int bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- compile-time layout
// NS states per side of the block, K1 exponentials per site, VIB states per site (1: electronic
// block, state = site; > 1: vibronic block, state a belongs to site a / VIB).
template <int NS_, int K1_, int VIB_ = 1>
struct Cfg {
    static constexpr int NS = NS_, K1 = K1_, VIB = VIB_;
    static constexpr int M = NS * NS, BINS = (NS / VIB) * K1;
    static constexpr int E4 = (BINS + 3) / 4 * 4;
    // int32 words per lane of an offset table: 16-byte rows whose quarter-warp reads
    // (LDS.128) fall into distinct banks need an odd number of 16-byte units
    static constexpr int LD = ((E4 / 4) & 1) ? E4 : E4 + 4;
    static constexpr int UB = (BINS + 1) / 2 * 2;
    static constexpr int UD = ((UB / 2) & 1) ? UB : UB + 2;      // doubles per lane of the up-coefficient table
    static constexpr int MH = (M + 1) / 2 * 2;                   // doubles of one member's coefficient block
    // tile record (global memory, one per tile, REC_BYTES apart)
    static constexpr int OFF_DN = 0;                             // int32 [32][LD]  down-neighbour offsets, -1 absent
    static constexpr int OFF_SHIFT = OFF_DN + 32 * LD * 4;       // double [32]     u sum_jk n_jk nu_k
    static constexpr int OFF_SCALE = OFF_SHIFT + 32 * 8;         // double [32]     error-norm weight
    static constexpr int TOP_BYTES = OFF_SCALE + 32 * 8;         // top-level tiles stop here
    static constexpr int OFF_UP = TOP_BYTES;                     // int32 [32][LD]  up-neighbour offsets
    static constexpr int OFF_UPC = OFF_UP + 32 * LD * 4;         // double [32][UD] u (n_jk + 1)
    static constexpr int REC_BYTES = OFF_UPC + 32 * UD * 8;
    static constexpr int YS_BYTES = M * 32 * 16;
    static constexpr int OFF_H = YS_BYTES + REC_BYTES;           // member coefficients behind the record
    static constexpr int BUF_BYTES = OFF_H + MH * 8;
    static constexpr int THREADS = 32 * (NS + 1);                 // NS row-warps + the producer warp
    static constexpr int HDR_BYTES = 128 + MH * 8;               // mbarriers, coefficients of member 0
    static __host__ __device__ constexpr size_t smem_bytes(int nbuf) { return HDR_BYTES + (size_t)nbuf * BUF_BYTES; }
};

struct RowDev {
    int n_members;
    long long n_ado, n_tiles, top_tile;     // tiles >= top_tile hold top-level ADOs only
    const unsigned char *rec;               // [n_tiles][REC_BYTES]
    const double *hmem;                     // [n_members][MH]  h = Im(-i u H) = -u H, row-major (symmetric)
    const double *gscale;                   // [n_tiles * 32]   g_n (0 for padding lanes)
    double hc[64];                          // member 0 (constant bank path of single-member handles)
    cplx cd[4];                             // -i u c_k
    double d2;                              // u * temp_corr * 2: Ishizaki-Tanimura term of off-diagonal elements
    int const_h;
    int stream_rec;                         // large single hierarchy: top-level tiles and tile records are read once per stage (L2 evict_first)
    int dbg;                                // diagnostics (QSX_ROW_DBG builds): 1 no gathers, 2 gathers from the own tile, 4 no stores
};

// ------------------------------------------------------------------------ tile body
// acc[b] = (L sigma)[w, b] for the ADO of this lane;
// epi(integral_constant<bool, UP> (false: top-level tile), index within the column, value, own,
// error-norm weight of the ADO).
// `w` = row and `lane` = ADO (within the tile) of this thread, see row_of() / ado_of().
template <class C, bool UP, bool CONSTH, class Epi>
__device__ __forceinline__ void row_body(const RowDev &R, const unsigned char *buf, const double *hs0,
                                         const cplx *__restrict__ xc, int w, int lane, int tile, Epi &&epi) {
    constexpr int NS = C::NS, K1 = C::K1, LD = C::LD, UD = C::UD, E4 = C::E4, UB = C::UB, VIB = C::VIB;
    const int ws = w / VIB;                // site of the row state
    const cplx *ys = reinterpret_cast<const cplx *>(buf) + lane;            // element e at ys[e * 32]
    const unsigned char *rec = buf + C::YS_BYTES;
    const int *dn = reinterpret_cast<const int *>(rec + C::OFF_DN) + lane * LD;
    const double *hm = CONSTH ? hs0 : reinterpret_cast<const double *>(buf + C::OFF_H);
    const double *hrow = hm + w * NS;      // row w of h: per half-warp, so it comes from shared memory
    const cplx *xw = xc + w * 32;          // element (w, b) of the ADO at offset o: xw[o + b * NS * 32]
    const cplx zero = cmake(0.0, 0.0);
    cplx acc[NS];
#ifdef QSX_ROW_DBG
    auto dbg_off = [&](int o) { return (R.dbg & 1) ? -1 : ((R.dbg & 2) && o >= 0) ? tile * C::M * 32 + lane : o; };
#else
    auto dbg_off = [&](int o) { return o; };
#endif

    // ---- batch 1: row-site down-links (the whole row shares neighbour and coefficient)
    cplx g[K1][NS];
    {
        int o[K1];
#pragma unroll
        for (int k = 0; k < K1; ++k) o[k] = dbg_off(dn[ws * K1 + k]);
#pragma unroll
        for (int k = 0; k < K1; ++k) {
            const cplx *p = xw + (o[k] >= 0 ? o[k] : 0);
#pragma unroll
            for (int b = 0; b < NS; ++b) g[k][b] = o[k] >= 0 ? ROW_LD(p + b * NS * 32) : zero;
        }
    }
    // ---- slice 1: diagonal terms and - sigma Hs from the own row
    {
        cplx own[NS];
#pragma unroll
        for (int c = 0; c < NS; ++c) own[c] = ys[(w + NS * c) * 32];
        const double shift = reinterpret_cast<const double *>(rec + C::OFF_SHIFT)[lane];
#pragma unroll
        for (int b = 0; b < NS; ++b) {
            const double dg = shift + (b / VIB == ws ? 0.0 : R.d2);
            acc[b] = cmake(-dg * own[b].x, -dg * own[b].y);
#pragma unroll
            for (int c = 0; c < NS; ++c) {
                const double h = CONSTH ? R.hc[b * NS + c] : hm[b * NS + c];
                acc[b].x = fma(h, own[c].y, acc[b].x);
                acc[b].y = fma(-h, own[c].x, acc[b].y);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K1; ++k)
#pragma unroll
        for (int b = 0; b < NS; ++b) cfma(acc[b], R.cd[k], g[k][b]);

    // Hs sigma for source rows [c0, c1): (i h) z = h (-z.y, z.x)
    auto left = [&](auto c0, auto c1) {
#pragma unroll
        for (int c = decltype(c0)::value; c < decltype(c1)::value; ++c) {
            const double h = hrow[c];
#pragma unroll
            for (int b = 0; b < NS; ++b) {
                const cplx z = ys[(c + NS * b) * 32];
                acc[b].x = fma(-h, z.y, acc[b].x);
                acc[b].y = fma(h, z.x, acc[b].y);
            }
        }
    };
    constexpr int CA = UP ? (NS + 2) / 3 : NS, CB = UP ? (2 * NS + 2) / 3 : NS;
    using I0 = std::integral_constant<int, 0>;
    using IA = std::integral_constant<int, CA>;
    using IB = std::integral_constant<int, CB>;
    using IN = std::integral_constant<int, NS>;

    // ---- batch 2: column-site down-links.  The diagonal element (b == w) re-reads the
    // neighbour of batch 1: cd v + conj(cd) v = 2 Re(cd) v without a special case.
    {
        int o[E4];
#pragma unroll
        for (int j = 0; j < E4 / 4; ++j) *reinterpret_cast<int4 *>(&o[4 * j]) = reinterpret_cast<const int4 *>(dn)[j];
#ifdef QSX_ROW_DBG
#pragma unroll
        for (int j = 0; j < E4; ++j) o[j] = dbg_off(o[j]);
#endif
        cplx v[NS][K1];
#pragma unroll
        for (int b = 0; b < NS; ++b)
#pragma unroll
            for (int k = 0; k < K1; ++k)
                v[b][k] = o[(b / VIB) * K1 + k] >= 0 ? ROW_LD(xw + o[(b / VIB) * K1 + k] + b * NS * 32) : zero;
        left(I0(), IA());
#pragma unroll
        for (int b = 0; b < NS; ++b)
#pragma unroll
            for (int k = 0; k < K1; ++k) cfma(acc[b], cmake(R.cd[k].x, -R.cd[k].y), v[b][k]);
    }
    if (UP) {
        const int *up = reinterpret_cast<const int *>(rec + C::OFF_UP) + lane * LD;
        const double *upc = reinterpret_cast<const double *>(rec + C::OFF_UPC) + lane * UD;
        // ---- batch 3: row-site up-links, coefficient -i u (n + 1)
        {
            int o[K1];
#pragma unroll
            for (int k = 0; k < K1; ++k) o[k] = dbg_off(up[ws * K1 + k]);
#pragma unroll
            for (int k = 0; k < K1; ++k) {
                const cplx *p = xw + (o[k] >= 0 ? o[k] : 0);
#pragma unroll
                for (int b = 0; b < NS; ++b) g[k][b] = o[k] >= 0 ? ROW_LD(p + b * NS * 32) : zero;
            }
            left(IA(), IB());
#pragma unroll
            for (int k = 0; k < K1; ++k) {
                const double t = upc[ws * K1 + k];
#pragma unroll
                for (int b = 0; b < NS; ++b) {
                    acc[b].x = fma(t, g[k][b].y, acc[b].x);
                    acc[b].y = fma(-t, g[k][b].x, acc[b].y);
                }
            }
        }
        // ---- batch 4: column-site up-links, coefficient +i u (n + 1); on the diagonal the two cancel
        {
            int o[E4];
#pragma unroll
            for (int j = 0; j < E4 / 4; ++j) *reinterpret_cast<int4 *>(&o[4 * j]) = reinterpret_cast<const int4 *>(up)[j];
#ifdef QSX_ROW_DBG
#pragma unroll
            for (int j = 0; j < E4; ++j) o[j] = dbg_off(o[j]);
#endif
            cplx v[NS][K1];
#pragma unroll
            for (int b = 0; b < NS; ++b)
#pragma unroll
                for (int k = 0; k < K1; ++k)
                    v[b][k] = o[(b / VIB) * K1 + k] >= 0 ? ROW_LD(xw + o[(b / VIB) * K1 + k] + b * NS * 32) : zero;
            left(IB(), IN());
            double t[UB];
#pragma unroll
            for (int j = 0; j < UB / 2; ++j) *reinterpret_cast<double2 *>(&t[2 * j]) = reinterpret_cast<const double2 *>(upc)[j];
#pragma unroll
            for (int b = 0; b < NS; ++b)
#pragma unroll
                for (int k = 0; k < K1; ++k) {
                    acc[b].x = fma(-t[(b / VIB) * K1 + k], v[b][k].y, acc[b].x);
                    acc[b].y = fma(t[(b / VIB) * K1 + k], v[b][k].x, acc[b].y);
                }
        }
    } else {
        left(IA(), IN());
    }
    const int base = (tile * C::M + w) * 32 + lane;
    const double sc = reinterpret_cast<const double *>(rec + C::OFF_SCALE)[lane];
#pragma unroll
    for (int b = 0; b < NS; ++b) epi(std::integral_constant<bool, UP>(), base + b * NS * 32, acc[b], ys[(w + NS * b) * 32], sc);
}

// ------------------------------------------------------------------ tile pipeline
// A CTA is NS row-warps plus one producer warp (its lane 0 stages tiles; the register
// allocation of a 7-warp CTA is that of 8 warps anyway).  Work units u = column * n_tiles +
// tile are dealt out round by round: round r = units [r G, (r + 1) G), one per CTA.
//
// sweep(): one stage behind a grid barrier (RHS application, the adaptive Taylor pilot).
//
// flow(): a run of product-form stages WITHOUT grid barriers between them.  Stage S reads what
// stage S - 1 wrote, but only from the neighbourhood of a tile: the producer stages unit u of
// stage S as soon as every unit up to dep_hi(u) -- the last tile any ADO of the tile links to --
// has been completed in stage S - 1 (per-round completion counters in global memory, published
// by the producers, cumulative over the stages of a launch; by symmetry of the links the same
// condition covers the tiles of stage S - 1 that still read what the unit overwrites).  The
// row-warps only ever wait for their tile's mbarrier, so a CTA that finishes a stage early
// runs on into the low hierarchy levels of the next one: no pipeline drain, no barrier
// latency, and the uneven last round (3634 tiles over 296 CTAs) is spread by rotating the
// unit -> CTA assignment from stage to stage.
__device__ __forceinline__ bool mbar_test(uint64_t *bar, int parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct FlowDev {
    unsigned long long *cnt; // [rounds] units of a round completed, cumulative over flow stages
    const int *dep_hi;       // [n_tiles] last tile linked to a tile
    unsigned rot;            // units the CTA assignment advances per stage
};

template <class C, int NBUF>
struct Pipe {
    uint64_t *full, *empty;
    unsigned char *bufs;
    double *hs0;             // coefficients of member 0 (single-member handles)
    unsigned q;              // tiles this CTA has staged/consumed so far (same value in every thread)
    int lane, wid, w, ado;   // thread -> (row, ADO): warps below 2 P hold the row pair (2p, 2p+1) of 16 ADOs,
                             // the odd last row takes a whole warp; warp NS is the producer

    __device__ __forceinline__ void init(const RowDev &R, unsigned char *smem) {
        full = reinterpret_cast<uint64_t *>(smem);
        empty = full + NBUF;
        hs0 = reinterpret_cast<double *>(smem + 128);
        bufs = smem + C::HDR_BYTES;
        q = 0;
        lane = threadIdx.x & 31; wid = threadIdx.x >> 5;
        constexpr int P = C::NS / 2;
        const bool paired = wid < 2 * P;
        w = paired ? 2 * (wid % P) + (lane >> 4) : C::NS - 1;
        ado = paired ? 16 * (wid / P) + (lane & 15) : lane;
        for (int i = threadIdx.x; i < C::MH; i += blockDim.x) hs0[i] = i < C::M ? R.hc[i] : 0.0;
        if (threadIdx.x == 0) {
            for (int i = 0; i < NBUF; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], C::NS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
    }
    __device__ __forceinline__ bool producer() const { return wid == C::NS; }
    // producer: stage (column, tile) of `src` as the qq-th tile of this CTA (its buffer is free)
    template <bool CONSTH>
    __device__ __forceinline__ void fill(const RowDev &R, unsigned qq, const cplx *src, long long Dp, int col, int tile,
                                         const int *member_of) {
        const int bi = qq % NBUF;
        unsigned char *b = bufs + (size_t)bi * C::BUF_BYTES;
        const int recb = tile >= R.top_tile ? C::TOP_BYTES : C::REC_BYTES;
        mbar_expect_tx(&full[bi], C::YS_BYTES + recb + (CONSTH ? 0 : C::MH * 8));
        // in a large hierarchy a top-level tile is read once per stage (nothing links down into it
        // from above, and the up-links of the level below were served earlier in the stage); in a
        // batch of small hierarchies all tiles of a column are in flight together
        if (R.stream_rec && tile >= R.top_tile) bulk_g2s_stream(b, src + (size_t)col * Dp + (size_t)tile * C::M * 32, C::YS_BYTES, &full[bi]);
        else bulk_g2s(b, src + (size_t)col * Dp + (size_t)tile * C::M * 32, C::YS_BYTES, &full[bi]);
        // records: re-read a whole stage later in a large hierarchy, but shared by every column of a batch
        if (R.stream_rec) bulk_g2s_stream(b + C::YS_BYTES, R.rec + (size_t)tile * C::REC_BYTES, recb, &full[bi]);
        else bulk_g2s(b + C::YS_BYTES, R.rec + (size_t)tile * C::REC_BYTES, recb, &full[bi]);
        if (!CONSTH) {
            const int m = member_of ? member_of[col] : 0;
            bulk_g2s(b + C::OFF_H, R.hmem + (size_t)m * C::MH, C::MH * 8, &full[bi]);
        }
    }
    // row-warps: one staged tile
    template <bool CONSTH, class Epi>
    __device__ __forceinline__ void tile_body(const RowDev &R, const cplx *srcb, int tile, Epi &&epi) {
        mbar_wait(&full[q % NBUF], (q / NBUF) & 1);
        const unsigned char *buf = bufs + (size_t)(q % NBUF) * C::BUF_BYTES;
        if (tile >= R.top_tile) row_body<C, false, CONSTH>(R, buf, hs0, srcb, w, ado, tile, epi);
        else row_body<C, true, CONSTH>(R, buf, hs0, srcb, w, ado, tile, epi);
    }
    __device__ __forceinline__ void release() {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[q % NBUF]);
    }

    // All threads.  make_epi(col) returns the epilogue functor of a column; done(col) runs
    // after every tile.  Callers separate sweeps by grid barriers.
    template <bool CONSTH, class MakeEpi, class Done>
    __device__ __forceinline__ void sweep(const RowDev &R, const cplx *src, int B, const int *member_of,
                                          MakeEpi &&make_epi, Done &&done) {
        const long long Dp = R.n_tiles * C::M * 32;
        const unsigned n_tiles = (unsigned)R.n_tiles, total = n_tiles * (unsigned)B, G = gridDim.x;
        if (producer()) {
            unsigned qq = q;
            if (lane == 0) {
                // writes of other CTAs (previous stage, generic proxy) before the bulk reads below
                asm volatile("fence.proxy.async.global;" ::: "memory");
                for (unsigned u = blockIdx.x; u < total; u += G, ++qq) {
                    if (qq >= NBUF) mbar_wait(&empty[qq % NBUF], ((qq / NBUF) - 1) & 1);
                    const unsigned c = u / n_tiles;
                    fill<CONSTH>(R, qq, src, Dp, (int)c, (int)(u - c * n_tiles), member_of);
                }
            }
            if (blockIdx.x < total) q += (total - 1 - blockIdx.x) / G + 1;
            return;
        }
        for (unsigned u = blockIdx.x; u < total; u += G, ++q) {
            const unsigned c = u / n_tiles;
            const int col = (int)c, tile = (int)(u - c * n_tiles);
            auto epi = make_epi(col);
            tile_body<CONSTH>(R, src + (size_t)col * Dp, tile, epi);
            done(col);
            release();
        }
    }

    // All threads: stages S0 .. S0 + n_stages - 1 (cumulative flow-stage numbers of the launch),
    // stage j reads X[j & 1] and writes X[(j & 1) ^ 1]; make_stage(j)(col) returns the epilogue.
    // The caller puts a grid barrier before and after the run.
    template <bool CONSTH, class MakeEpi>
    __device__ __forceinline__ void flow(const RowDev &R, const FlowDev &F, cplx *X0, cplx *X1, int B, const int *member_of,
                                         unsigned S0, int n_stages, MakeEpi &&make_stage) {
        const long long Dp = R.n_tiles * C::M * 32;
        const unsigned n_tiles = (unsigned)R.n_tiles, total = n_tiles * (unsigned)B, G = gridDim.x;
        auto first = [&](unsigned S) { return (unsigned)((blockIdx.x + (unsigned long long)S * F.rot) % G); };
        if (producer()) {
            unsigned n_mine = 0;
            for (int j = 0; j < n_stages; ++j) {
                const unsigned f = first(S0 + j);
                if (f < total) n_mine += (total - 1 - f) / G + 1;
            }
            if (lane == 0) {
                // fill cursor (jf, uf, kf) and publish cursor (jp, up, kp) over the units of this CTA
                int jf = 0, jp = 0;
                unsigned uf = first(S0), up = uf, kf = 0, kp = 0;
                unsigned frontier = 0;      // rounds [0, frontier) of stage S0 + jf - 1 are known to be complete
                auto skip = [&](int &j, unsigned &u) {      // move the cursor to its next existing unit
                    while (j < n_stages && u >= total) { ++j; u = first(S0 + j); }
                };
                skip(jf, uf);
                skip(jp, up);
                int spins = 0;
                while (kp < n_mine) {
                    if (kp < kf) {
                        uint64_t *eb = &empty[(q + kp) % NBUF];
                        const int par = ((q + kp) / NBUF) & 1;
                        bool left;
                        if (kf - kp == NBUF || kf == n_mine) { mbar_wait(eb, par); left = true; }   // nothing else to do
                        else left = mbar_test(eb, par);
                        if (left) {
                            // every row-warp has left unit kp: make its stores visible, then count it
                            __threadfence();
                            atomicAdd(&F.cnt[up / G], 1ULL);
                            ++kp;
                            up += G;
                            if (up >= total) { ++jp; up = first(S0 + jp); skip(jp, up); }
                        }
                    }
                    if (kf < n_mine && kf - kp < NBUF) {
                        const unsigned c = uf / n_tiles, tile = uf - c * n_tiles;
                        bool ok = true;
                        if (jf > 0) {
                            const unsigned need = (c * n_tiles + (unsigned)__ldg(&F.dep_hi[tile])) / G + 1;   // rounds of the previous stage
                            const unsigned S = S0 + jf;            // completed stages a full counter shows
                            while (frontier < need) {
                                const unsigned in_round = min(G, total - frontier * G);
                                if (ld_acquire(&F.cnt[frontier]) >= (unsigned long long)S * in_round) ++frontier;
                                else { ok = false; break; }
                            }
                        }
                        if (ok) {
                            asm volatile("fence.proxy.async.global;" ::: "memory");
                            fill<CONSTH>(R, q + kf, (jf & 1) ? X1 : X0, Dp, (int)c, (int)tile, member_of);
                            ++kf;
                            uf += G;
                            if (uf >= total) { ++jf; uf = first(S0 + jf); frontier = 0; skip(jf, uf); }
                            spins = 0;
                        } else {
                            __nanosleep(200);
                            if (++spins > (1 << 23)) __trap();      // seconds: a broken dependency table must not hang the GPU
                        }
                    }
                }
            }
            q += n_mine;
            return;
        }
        for (int j = 0; j < n_stages; ++j) {
            const cplx *src = (j & 1) ? X1 : X0;
            auto make_epi = make_stage(j);
            for (unsigned u = first(S0 + j); u < total; u += G, ++q) {
                const unsigned c = u / n_tiles;
                const int col = (int)c, tile = (int)(u - c * n_tiles);
                auto epi = make_epi(col);
                tile_body<CONSTH>(R, src + (size_t)col * Dp, tile, epi);
                release();
            }
        }
    }
};

// ------------------------------------------------------------------------- kernels
struct RowApplyArgs {
    RowDev R;
    const cplx *x;          // internal layout, sigma variables
    cplx *y;
    const int *member_of;   // [B] or null
    int B;
};

template <class C, bool CONSTH, int NBUF, int MINB>
__global__ void __launch_bounds__(C::THREADS, MINB) heom_row_apply_kernel(const __grid_constant__ RowApplyArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Pipe<C, NBUF> pipe;
    pipe.init(a.R, smem_raw);
    const long long Dp = a.R.n_tiles * C::M * 32;
    pipe.template sweep<CONSTH>(a.R, a.x, a.B, a.member_of, [&](int col) {
        cplx *yb = a.y + (size_t)col * Dp;
#ifdef QSX_ROW_DBG
        const bool nost = a.R.dbg & 4;
        return [=](auto, int i, cplx f, cplx, double) { if (!nost || f.x == 1.2345e-300) __stcs(&yb[i], f); };
#else
        return [=](auto, int i, cplx f, cplx, double) { __stcs(&yb[i], f); };
#endif
    }, [](int) {});
}

struct RowPropArgs {
    RowDev R;
    FlowDev F;
    int B, nt;
    int use_flow;               // product-form stages without grid barriers (flow())
    const int *member_of;
    const cplx *y0;             // reference layout [B][n_ado][M]
    cplx *Y, *V, *W;            // work vectors, internal layout [B][Dp]
    const double *t;
    double t0, rtol, theta, lnorm;
    int kmax, method;           // QSX_METHOD_TAYLOR or QSX_METHOD_POLY
    int repilot;                // POLY: product-form intervals between two Taylor pilot intervals
    const cplx *ainv;           // device copy of qsx_taylor_ainv
    int save_mode, save_rows;
    const cplx *S;              // [n_save][save_rows][M]
    const int *save_of;         // [B] save matrix of each column, or null (matrix 0)
    cplx *out;
    long long saved_dim;
    int *flags;                 // [3]
    double *ynorm;              // [3][B]
    unsigned long long *stats;  // rhs, steps, status, degree of the last product step, non-finite saves
};

template <class C>
__device__ __forceinline__ void row_save(const RowPropArgs &a, const cplx *Y, int it) {
    const long long n_ado = a.R.n_ado;
    constexpr int M = C::M;
    const long long Dp = a.R.n_tiles * M * 32;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsz = (long long)gridDim.x * blockDim.x;
    if (a.save_mode == QSX_SAVE_STATE) {
        const long long per = n_ado * M;
        for (long long i = gtid; i < (long long)a.B * per; i += gsz) {
            const long long b = i / per, r = i % per, n = r / M;
            const int e = (int)(r % M);
            const cplx v = __ldcg(&Y[(size_t)b * Dp + ((n >> 5) * M + e) * 32 + (n & 31)]);
            a.out[((size_t)b * a.nt + it) * a.saved_dim + r] = cscale(1.0 / a.R.gscale[n], v);
        }
    } else if (a.save_mode == QSX_SAVE_ADO0) {
        for (long long i = gtid; i < (long long)a.B * M; i += gsz) {
            const long long b = i / M, e = i % M;
            const cplx v = __ldcg(&Y[(size_t)b * Dp + e * 32]);        // g_0 = 1
            a.out[((size_t)b * a.nt + it) * a.saved_dim + e] = v;
            // the product form has no convergence test that a non-finite state would fail
            if (!(fabs(v.x) + fabs(v.y) < 1e300)) atomicAdd(&a.stats[4], 1ULL);
        }
    } else {
        const long long per_col = n_ado * a.save_rows;
        for (long long i = gtid; i < (long long)a.B * per_col; i += gsz) {
            const long long b = i / per_col, r = i % per_col, n = r / a.save_rows;
            const int m = (int)(r % a.save_rows);
            const cplx *y = Y + (size_t)b * Dp + ((n >> 5) * M) * 32 + (n & 31);
            const cplx *Sm = a.S + (a.save_of ? (size_t)a.save_of[b] * a.save_rows * M : 0);
            cplx acc = cmake(0, 0);
            for (int e = 0; e < M; ++e) cfma(acc, __ldg(&Sm[(size_t)m * M + e]), __ldcg(&y[e * 32]));
            a.out[((size_t)b * a.nt + it) * a.saved_dim + r] = cscale(1.0 / a.R.gscale[n], acc);
        }
    }
}

template <class C, bool CONSTH, int NBUF, int MINB>
__global__ void __launch_bounds__(C::THREADS, MINB) heom_row_propagate_kernel(const __grid_constant__ RowPropArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    Pipe<C, NBUF> pipe;
    pipe.init(a.R, smem_raw);
    constexpr int M = C::M;
    const RowDev &R = a.R;
    const long long Dp = R.n_tiles * M * 32;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsz = (long long)gridDim.x * blockDim.x;
    const int B = a.B;

    // ---- init: Y = g y0 (layout change), reference norms, control words ---------
    for (int i = (int)gtid; i < 3 * B; i += (int)gsz) a.ynorm[i] = 0.0;
    if (gtid < 3) a.flags[gtid] = 0;
    grid.sync();
    for (int b = 0; b < B; ++b) {
        double loc = 0.0;
        for (long long i = gtid; i < Dp; i += gsz) {
            const long long tile = i / (M * 32);
            const int e = (int)((i / 32) % M), lane = (int)(i % 32);
            const long long n = tile * 32 + lane;
            cplx v = cmake(0, 0);
            if (n < R.n_ado) v = cscale(R.gscale[n], a.y0[((size_t)b * R.n_ado + n) * M + e]);
            a.Y[(size_t)b * Dp + i] = v;
            const double sc = reinterpret_cast<const double *>(R.rec + (size_t)tile * C::REC_BYTES + C::OFF_SCALE)[lane];
            loc = fmax(loc, sc * cabs1(v));
        }
        loc = warp_max(loc);
        if ((threadIdx.x & 31) == 0 && loc > 0) atomic_max_nonneg(&a.ynorm[b], loc);
    }
    grid.sync();

    unsigned long long n_rhs = 0, n_steps = 0;
    int status = 0;
    int fslot = 0;          // flag slot of the current convergence check
    int nslot = 0;          // norm slot that holds the latest reference norms
    double tcur = a.t0;
    cplx *cur = a.Y;        // vector that holds the state
    int degree = 0;         // POLY: degree found by the last pilot (0: none yet, -1: series too long for the table)
    int since_pilot = 0;
    unsigned flow_stage = 0;    // flow stages completed so far in this launch

    // adaptive Taylor interval: cur <- exp(h L) cur; the other two vectors hold the terms.
    // Returns the number of terms used (even), or -1 if not converged within kmax.
    auto taylor_step = [&](double h) -> int {
        cplx *Yv = cur;
        cplx *ta = (cur == a.Y) ? a.V : a.Y;
        cplx *tb = (cur == a.W) ? a.V : a.W;
        const cplx *src = Yv;
        cplx *dst = ta;
        for (int k = 1; k <= a.kmax; ++k) {
            const double fac = h / k;
            const bool even = (k & 1) == 0;
            int ok = 1;
            if (even && blockIdx.x == 0) {   // recycle the control slots that come next
                if (threadIdx.x == 0) a.flags[(fslot + 1) % 3] = 0;
                for (int b = threadIdx.x; b < B; b += blockDim.x) a.ynorm[((nslot + 2) % 3) * B + b] = 0.0;
            }
            if (!even) {
                pipe.template sweep<CONSTH>(R, src, B, a.member_of, [&](int col) {
                    cplx *db = dst + (size_t)col * Dp;
                    return [=](auto, int i, cplx f, cplx, double) { __stcs(&db[i], cscale(fac, f)); };
                }, [](int) {});
            } else {
                double ymax = 0.0;
                pipe.template sweep<CONSTH>(R, src, B, a.member_of, [&](int col) {
                    cplx *db = dst + (size_t)col * Dp;
                    cplx *Yb = Yv + (size_t)col * Dp;
                    const double yref = a.rtol * __ldcg(&a.ynorm[nslot * B + col]);
                    return [&ok, &ymax, db, Yb, yref, fac](auto, int i, cplx f, cplx own, double sc) {
                        const cplx wv = cscale(fac, f);
                        __stcs(&db[i], wv);
                        cplx y = __ldcs(&Yb[i]);
                        y.x += own.x + wv.x;
                        y.y += own.y + wv.y;
                        __stcs(&Yb[i], y);
                        if (!(sc * (cabs1(own) + cabs1(wv)) <= yref)) ok = 0;      // also catches a non-finite state
                        ymax = fmax(ymax, sc * cabs1(y));
                    };
                }, [&](int col) {
                    const double m = warp_max(ymax);
                    if ((threadIdx.x & 31) == 0 && m > 0) atomic_max_nonneg(&a.ynorm[((nslot + 1) % 3) * B + col], m);
                    ymax = 0.0;
                });
            }
            n_rhs += 1;
            if (even) {
                const int all_ok = __syncthreads_and(ok);
                if (!all_ok && threadIdx.x == 0) atomicExch(&a.flags[fslot], 1);
            }
            grid.sync();
            src = dst;
            dst = (dst == ta) ? tb : ta;
            if (even) {
                const int failed = *((volatile int *)&a.flags[fslot]);
                fslot = (fslot + 1) % 3;
                nslot = (nslot + 1) % 3;
                if (!failed) return k;
            }
        }
        return -1;
    };

    // nsub product-form steps of degree m: cur <- [prod_j (I + h a_j L)]^nsub cur, ping-pong with one
    // other vector; a grid barrier precedes and follows the run
    auto poly_run = [&](double h, int m, int nsub) {
        const cplx *tab = a.ainv + (size_t)m * (m - 1) / 2;
        cplx *other = (cur == a.Y) ? a.V : a.Y;
        const int n_stages = m * nsub;
        if (a.use_flow) {
            cplx *X0 = cur, *X1 = other;
            pipe.template flow<CONSTH>(R, a.F, X0, X1, B, a.member_of, flow_stage, n_stages, [&](int j) {
                const cplx al = cscale(h, __ldg(&tab[j % m]));
                cplx *dst = (j & 1) ? X0 : X1;
                return [=](int col) {
                    cplx *db = dst + (size_t)col * Dp;
                    return [=](auto, int i, cplx f, cplx own, double) {
                        cfma(own, al, f);
                        db[i] = own;
                    };
                };
            });
            flow_stage += (unsigned)n_stages;
            if (n_stages & 1) cur = other;
            grid.sync();
        } else {
            for (int j = 0; j < n_stages; ++j) {
                const cplx al = cscale(h, __ldg(&tab[j % m]));
                cplx *dst = other;
                pipe.template sweep<CONSTH>(R, cur, B, a.member_of, [&](int col) {
                    cplx *db = dst + (size_t)col * Dp;
                    return [=](auto, int i, cplx f, cplx own, double) {
                        cfma(own, al, f);
                        db[i] = own;
                    };
                }, [](int) {});
                grid.sync();
                other = cur;
                cur = dst;
            }
        }
        n_rhs += (unsigned long long)n_stages;
    };

    for (int it = 0; it < a.nt; ++it) {
        const double target = a.t[it];
        if (target != tcur) {
            const double span = target - tcur;
            int nsub = (int)ceil(fabs(span) * a.lnorm / a.theta);
            if (nsub < 1) nsub = 1;
            const double h = span / nsub;
            int sub = 0;
            while (sub < nsub) {
                const bool pilot = a.method == QSX_METHOD_TAYLOR || degree <= 0 || since_pilot >= a.repilot;
                if (pilot) {
                    const int k = taylor_step(h);
                    if (k < 0) status = QSX_ERR_INTEGRATOR;
                    degree = (k > 0 && k <= QSX_POLY_MMAX) ? k : -1;
                    since_pilot = 0;
                    sub += 1;
                    n_steps += 1;
                } else {
                    // all remaining steps of the interval up to the next pilot in one run
                    const int n = min(nsub - sub, a.repilot - since_pilot);
                    poly_run(h, degree, n);
                    since_pilot += n;
                    sub += n;
                    n_steps += (unsigned long long)n;
                }
            }
            tcur = target;
        }
        row_save<C>(a, cur, it);
        // the next stage that writes `cur` is separated from this read by >= 1 grid barrier
        // (Taylor: first write of the accumulator at k = 2; product form: the dependency counters
        // of stage 2 require stage 1 complete, which follows this save in every CTA -- see below)
        if (a.use_flow) grid.sync();
    }
    if (gtid == 0) {
        a.stats[0] = n_rhs * (unsigned long long)B;
        a.stats[1] = n_steps * (unsigned long long)B;
        a.stats[2] = (unsigned long long)(status != 0);
        a.stats[3] = (unsigned long long)(degree > 0 ? degree : 0);
    }
}

}  // namespace heom_row

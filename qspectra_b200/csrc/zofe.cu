// K3 + K4: ZOFE master equation (nonlinear auxiliary-operator ODE) as a fused
// elementwise + small-GEMM right-hand side inside the CTA-resident integrators.
//
// Replaces ZOFEModel.rhodot_oopdot_vec / equation_of_motion (reference
// dynamics/zofe.py:121-234).  State = [vec_F(rho) ; vec_F(O)], O of shape
// (P pseudomodes, S sites, n, n) flattened column-major: O[p,s,a,b] sits at
// n^2 + p + P (s + S (a + n b)).  With L_s = -V_s (V_s diagonal) and
// Sg_s = sum_p O[p,s]:
//   a   = sum_s L_s^+ Sg_s            b = -i H - a
//   c   = sum_s L_s rho Sg_s^+        big = sum_s Sg_s rho L_s^+
//   rho' = b rho + c + rho (i H~ - a^+) + big          (general branch)
//   O'[p,s] = Gamma[p,s] L_s - w[p,s] O[p,s] + b O[p,s] - O[p,s] b
// (rho_hermit / ham_hermit select the reference's algebraic shortcuts so that
// results agree with it flag for flag), everything times unit_convert.
// One thread block integrates one trajectory; the integrator work vectors live
// in an L2-resident global workspace, the n x n operators in shared memory.
#include "cta_integrator.cuh"
#include <algorithm>
#include <memory>

struct ZofeDev {
    int n, S, P, n_members;
    long long dim;              // n^2 (1 + P S)
    const cplx *H;              // [n_members][n][n] row-major
    const double *v;            // [S][n] diagonal of V_s
    const cplx *Gamma, *w;      // [P][S]
    double u;
    int ham_hermit, rho_hermit;
    // the auxiliary operators of a stage input are copied to shared memory once per RHS
    // (every O' element reads 2 n of them) when they fit next to the n x n operators
    int stage_O;
    unsigned mP, mS;            // 2^32 / P + 1, 2^32 / S + 1: division by multiplication (0: plain division)
};

__device__ __forceinline__ unsigned zofe_div(unsigned i, int d, unsigned m) {
    return m ? __umulhi(i, m) : i / (unsigned)d;
}

struct qsx_zofe_s {
    ZofeDev d;
    DevBuf<cplx> H, Gamma, w;
    DevBuf<double> v;
};

struct ZofeRhs {
    ZofeDev Z;
    const cplx *Hm;         // this trajectory's Hamiltonian (global)
    cplx *Sg;               // [S][n][n] shared (row-major per site)
    cplx *bop, *aop, *rho;  // [n][n] shared (row-major)
    cplx *cop;              // [n][n]
    cplx *Os;               // [P S n n] shared copy of the auxiliary operators (stage_O) or null
    int n_pulse;
    const qsx_pulse *pulses;
    const cplx *Vp;         // [n_pulse][n][n] row-major Hilbert-space dipole operators (global)

    template <class Epi>
    __device__ __forceinline__ void apply(const cplx *x, double t, Epi epi) {
        // compile-time state count for the FMO-sized systems (unrolled n-loops, constant divisions)
        if (Z.n == 7) apply_impl<7>(x, t, epi);
        else apply_impl<0>(x, t, epi);
    }

    template <int NC, class Epi>
    __device__ __forceinline__ void apply_impl(const cplx *x, double t, Epi epi) {
        const int n = NC ? NC : Z.n, S = Z.S, P = Z.P, nn = n * n;
        cplx gp[QSX_MAX_PULSES];
        for (int p = 0; p < n_pulse; ++p) gp[p] = pulse_coefficient(pulses[p], t);
        const int tid = threadIdx.x, nthr = blockDim.x;
        const long long no = (long long)P * S * nn;
        __syncthreads();        // previous users of the shared operators are done
        const cplx *O = x + nn;
        if (Os) {
            for (int i = tid; i < (int)no; i += nthr) Os[i] = O[i];
            O = Os;
            __syncthreads();
        }
        // rho (row-major copy) and Sg_s = sum_p O[p,s]
        for (int i = tid; i < nn; i += nthr) {
            int a = i / n, b = i % n;
            rho[i] = x[a + n * b];
        }
        // sixteen lanes per sum: they read consecutive pseudomodes (contiguous, conflict-free
        // in shared memory) and meet in four shuffles; the trip count is uniform per warp
        for (int i0 = (tid >> 5) * 2; i0 < S * nn; i0 += (nthr >> 5) * 2) {
            const int i = i0 + ((tid >> 4) & 1);
            const bool live = i < S * nn;
            cplx acc = cmake(0, 0);
            if (live) {
                int s = i / nn, ab = i % nn, a = ab / n, b = ab % n;
                const cplx *col = O + P * (s + S * (a + n * b));
                for (int p = tid & 15; p < P; p += 16) acc = cadd(acc, col[p]);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o, 16);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o, 16);
            }
            if (live && (tid & 15) == 0) Sg[i] = acc;
        }
        __syncthreads();
        // a[x][y] = sum_s (-v_s[x]) Sg_s[x][y];  b = -i H - a
        for (int i = tid; i < nn; i += nthr) {
            int xr = i / n;
            cplx acc = cmake(0, 0);
            for (int s = 0; s < S; ++s) rfma(acc, -Z.v[s * n + xr], Sg[s * nn + i]);
            aop[i] = acc;
            cplx h = Hm[i];
            bop[i] = cmake(h.y - acc.x, -h.x - acc.y);
        }
        __syncthreads();
        // c[x][y] = sum_s (-v_s[x]) sum_z rho[x][z] conj(Sg_s[y][z])
        for (int i = tid; i < nn; i += nthr) {
            int xr = i / n, yc = i % n;
            cplx acc = cmake(0, 0);
            for (int s = 0; s < S; ++s) {
                cplx t = cmake(0, 0);
                for (int z = 0; z < n; ++z) {
                    cplx sg = Sg[s * nn + yc * n + z];
                    sg.y = -sg.y;
                    cfma(t, rho[xr * n + z], sg);
                }
                rfma(acc, -Z.v[s * n + xr], t);
            }
            cop[i] = acc;
        }
        __syncthreads();
        // rho' elements
        for (int i = tid; i < nn; i += nthr) {
            int xr = i / n, yc = i % n;
            // d = b rho + c
            cplx d = cop[i];
            for (int z = 0; z < n; ++z) cfma(d, bop[xr * n + z], rho[z * n + yc]);
            cplx f;
            if (Z.rho_hermit && Z.ham_hermit) {
                // f = d^+ : f[x][y] = conj(d[y][x])
                cplx dt = cop[yc * n + xr];
                for (int z = 0; z < n; ++z) cfma(dt, bop[yc * n + z], rho[z * n + xr]);
                f = cmake(dt.x, -dt.y);
            } else {
                // rho (i H~ - a^+): H~ = H^+ when ham_hermit (b^+), else H
                f = cmake(0, 0);
                for (int z = 0; z < n; ++z) {
                    cplx h = Z.ham_hermit ? Hm[yc * n + z] : Hm[z * n + yc];
                    if (Z.ham_hermit) h.y = -h.y;
                    cplx ad = aop[yc * n + z];                     // a^+[z][y] = conj(a[y][z])
                    cplx m = cmake(-h.y - ad.x, h.x + ad.y);       // i h - conj(a)
                    cfma(f, rho[xr * n + z], m);
                }
                if (Z.rho_hermit) {
                    cplx ct = cop[yc * n + xr];                    // c^+
                    f.x += ct.x;
                    f.y -= ct.y;
                } else {
                    // big[x][y] = sum_s sum_z Sg_s[x][z] rho[z][y] (-v_s[y])
                    for (int s = 0; s < S; ++s) {
                        cplx t = cmake(0, 0);
                        for (int z = 0; z < n; ++z) cfma(t, Sg[s * nn + xr * n + z], rho[z * n + yc]);
                        rfma(f, -Z.v[s * n + yc], t);
                    }
                }
            }
            cplx tot = cscale(Z.u, cadd(d, f));
            // field terms (-i E_p(t)) [V_p, rho]  (eom.py:87-94; not scaled by unit_convert)
            for (int p = 0; p < n_pulse; ++p) {
                const cplx *V = Vp + (size_t)p * nn;
                cplx c2 = cmake(0, 0);
                for (int z = 0; z < n; ++z) {
                    cfma(c2, __ldg(&V[xr * n + z]), rho[z * n + yc]);
                    cplx vz = __ldg(&V[z * n + yc]), rz = rho[xr * n + z];
                    c2.x -= rz.x * vz.x - rz.y * vz.y;
                    c2.y -= rz.x * vz.y + rz.y * vz.x;
                }
                cfma(tot, gp[p], c2);
            }
            epi(xr + n * yc, tot);
        }
        // O' elements; consecutive threads = consecutive pseudomodes (coalesced)
        for (unsigned i = tid; i < (unsigned)no; i += nthr) {
            unsigned r = zofe_div(i, P, Z.mP);
            const int p = (int)(i - r * P);
            const unsigned r2 = zofe_div(r, S, Z.mS);
            const int s = (int)(r - r2 * S);
            const int xr = (int)(r2 % n), yc = (int)(r2 / n);
            const cplx o = O[i];
            const cplx wv = Z.w[p * S + s];
            cplx acc = cmake(-(wv.x * o.x - wv.y * o.y), -(wv.x * o.y + wv.y * o.x));
            if (xr == yc) {
                cplx g = Z.Gamma[p * S + s];
                rfma(acc, -Z.v[s * n + xr], g);
            }
            const cplx *Ops = O + p + P * s;               // O[p,s,a,b] at Ops[P S (a + n b)]
            const int st = P * S;
#pragma unroll
            for (int z = 0; z < n; ++z) {
                cfma(acc, bop[xr * n + z], Ops[st * (z + n * yc)]);
                cplx bz = bop[z * n + yc];
                cplx oz = Ops[st * (xr + n * z)];
                acc.x -= oz.x * bz.x - oz.y * bz.y;
                acc.y -= oz.x * bz.y + oz.y * bz.x;
            }
            cplx tot = cscale(Z.u, acc);
            for (int p = 0; p < n_pulse; ++p) {
                // the dipole operator multiplies every auxiliary operator too (zofe.py:24-35)
                const cplx *V = Vp + (size_t)p * nn;
                cplx c2 = cmake(0, 0);
                for (int z = 0; z < n; ++z) {
                    cfma(c2, __ldg(&V[xr * n + z]), Ops[st * (z + n * yc)]);
                    cplx vz = __ldg(&V[z * n + yc]), oz = Ops[st * (xr + n * z)];
                    c2.x -= oz.x * vz.x - oz.y * vz.y;
                    c2.y -= oz.x * vz.y + oz.y * vz.x;
                }
                cfma(tot, gp[p], c2);
            }
            epi(nn + (int)i, tot);
        }
    }
};

struct ZofeSaver {
    int nt, mode, save_rows, head;
    long long dim, saved_dim;
    const cplx *S;
    cplx *out;
    __device__ __forceinline__ void operator()(int it, const cplx *Y) {
        cplx *o = out + (size_t)it * saved_dim;
        if (mode == QSX_SAVE_MATRIX) {
            for (int m = threadIdx.x; m < save_rows; m += blockDim.x) {
                cplx acc = cmake(0, 0);
                for (int r = 0; r < head; ++r) cfma(acc, __ldg(&S[(size_t)m * head + r]), Y[r]);
                o[m] = acc;
            }
        } else {
            const long long count = mode == QSX_SAVE_ADO0 ? head : dim;
            for (long long i = threadIdx.x; i < count; i += blockDim.x) o[i] = Y[i];
        }
    }
};

struct ZofeKernelArgs {
    ZofeDev Z;
    int nt, n_vec;
    const int *member_of;
    const cplx *y0;
    cplx *work;             // [B][n_vec][dim]
    const double *t;
    double t0;
    int method;
    double rtol, atol;
    int rk4_sub;
    int save_mode, save_rows;
    const cplx *S;
    cplx *out;
    long long saved_dim;
    unsigned long long *stats;
    int n_pulse;
    qsx_pulse pulses[QSX_MAX_PULSES];
    const cplx *Vp;
};

__device__ __forceinline__ void zofe_rhs_setup(const ZofeDev &Z, unsigned char *smem, int member, ZofeRhs &r,
                                               double *&scratch) {
    const int nn = Z.n * Z.n;
    scratch = reinterpret_cast<double *>(smem);
    cplx *p = reinterpret_cast<cplx *>(smem) + 16;
    r.Z = Z;
    r.Hm = Z.H + (size_t)member * nn;
    r.Sg = p; p += (size_t)Z.S * nn;
    r.bop = p; p += nn;
    r.aop = p; p += nn;
    r.rho = p; p += nn;
    r.cop = p; p += nn;
    r.Os = Z.stage_O ? p : nullptr;
    r.n_pulse = 0; r.pulses = nullptr; r.Vp = nullptr;
}

#ifndef QSX_ZOFE_MINB
#define QSX_ZOFE_MINB 2
#endif
__global__ void __launch_bounds__(256, QSX_ZOFE_MINB) zofe_propagate_kernel(ZofeKernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int col = blockIdx.x;
    ZofeRhs rhs;
    double *scratch;
    zofe_rhs_setup(a.Z, smem_raw, a.member_of ? a.member_of[col] : 0, rhs, scratch);
    rhs.n_pulse = a.n_pulse; rhs.pulses = a.pulses; rhs.Vp = a.Vp;
    cplx *vec = a.work + (size_t)col * a.n_vec * a.Z.dim;
    for (long long i = threadIdx.x; i < a.Z.dim; i += blockDim.x) vec[i] = a.y0[(size_t)col * a.Z.dim + i];
    __syncthreads();
    ZofeSaver saver;
    saver.nt = a.nt; saver.mode = a.save_mode; saver.save_rows = a.save_rows;
    saver.head = a.Z.n * a.Z.n; saver.dim = a.Z.dim; saver.saved_dim = a.saved_dim;
    saver.S = a.S; saver.out = a.out + (size_t)col * a.nt * a.saved_dim;
    CtaProp P;
    P.n = (int)a.Z.dim; P.method = a.method; P.rtol = a.rtol; P.atol = a.atol;
    P.rk4_sub = a.rk4_sub; P.kmax = 0; P.theta = 1.0; P.lnorm = 0.0;
    P.nt = a.nt; P.t = a.t; P.t0 = a.t0;
    CtaStats st;
    cta_propagate<1>(rhs, saver, P, vec, scratch, st);
    if (threadIdx.x == 0) {
        atomicAdd(&a.stats[0], st.rhs);
        atomicAdd(&a.stats[1], st.steps);
        if (st.status != 0) atomicAdd(&a.stats[2], 1ULL);
    }
}

struct ZofeApplyArgs {
    ZofeDev Z;
    const int *member_of;
    const cplx *x;
    cplx *y;
};

__global__ void __launch_bounds__(256) zofe_apply_kernel(ZofeApplyArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int col = blockIdx.x;
    ZofeRhs rhs;
    double *scratch;
    zofe_rhs_setup(a.Z, smem_raw, a.member_of ? a.member_of[col] : 0, rhs, scratch);
    cplx *yb = a.y + (size_t)col * a.Z.dim;
    rhs.apply(a.x + (size_t)col * a.Z.dim, 0.0, [&](int i, cplx v) { yb[i] = v; });
}

static size_t zofe_smem_base(const ZofeDev &Z) { return (16 + (size_t)(Z.S + 4) * Z.n * Z.n) * sizeof(cplx); }
static size_t zofe_smem(const ZofeDev &Z) {
    return zofe_smem_base(Z) + (Z.stage_O ? (size_t)Z.P * Z.S * Z.n * Z.n * sizeof(cplx) : 0);
}

extern "C" int qsx_zofe_create(qsx_zofe_t *out, const qsx_zofe_config *cfg, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(out && cfg && cfg->H && cfg->coupling_diag && cfg->Gamma && cfg->w, "qsx_zofe_create: null argument");
    QSX_REQUIRE(cfg->n_states > 0 && cfg->n_sites > 0 && cfg->n_pm > 0 && cfg->n_members > 0,
                "qsx_zofe_create: bad sizes");
    std::unique_ptr<qsx_zofe_s> h(new qsx_zofe_s());
    const int n = cfg->n_states, S = cfg->n_sites, P = cfg->n_pm;
    QSX_CUDA(h->H.upload((const cplx *)cfg->H, (size_t)cfg->n_members * n * n, stream));
    QSX_CUDA(h->v.upload(cfg->coupling_diag, (size_t)S * n, stream));
    QSX_CUDA(h->Gamma.upload((const cplx *)cfg->Gamma, (size_t)P * S, stream));
    QSX_CUDA(h->w.upload((const cplx *)cfg->w, (size_t)P * S, stream));
    QSX_CUDA(cudaStreamSynchronize(stream));
    ZofeDev &d = h->d;
    d.n = n; d.S = S; d.P = P; d.n_members = cfg->n_members;
    d.dim = (long long)n * n * (1 + (long long)P * S);
    d.H = h->H.p; d.v = h->v.p; d.Gamma = h->Gamma.p; d.w = h->w.p;
    d.u = cfg->unit_convert; d.ham_hermit = cfg->ham_hermit; d.rho_hermit = cfg->rho_hermit;
    int dev = 0, smem_limit = 0;
    QSX_CUDA(cudaGetDevice(&dev));
    QSX_CUDA(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // staged auxiliary operators: two CTAs per SM must still fit
    d.stage_O = 2 * (zofe_smem_base(d) + (size_t)P * S * n * n * sizeof(cplx) + 1024) <= (size_t)228 * 1024 ? 1 : 0;
    const unsigned long long no = (unsigned long long)P * S * n * n;
    const bool magic = no * (unsigned long long)std::max(P, S) < (1ULL << 31);
    d.mP = magic ? (unsigned)((1ULL << 32) / P + 1) : 0;
    d.mS = magic ? (unsigned)((1ULL << 32) / S + 1) : 0;
    if (zofe_smem(d) > (size_t)smem_limit || d.dim > 0x7fffffffLL) {
        qsx_set_error("ZOFE system too large for the CTA-resident kernel (n=%d, sites=%d)", n, S);
        return QSX_ERR_UNSUPPORTED;
    }
    *out = h.release();
    return QSX_OK;
}

extern "C" void qsx_zofe_destroy(qsx_zofe_t h) { delete h; }
extern "C" int64_t qsx_zofe_state_dim(qsx_zofe_t h) { return h ? h->d.dim : -1; }

static int zofe_members(DevBuf<int> &buf, const int32_t *host, int n, int n_members, cudaStream_t s) {
    std::vector<int> m(host, host + n);
    for (int x : m) QSX_REQUIRE(x >= 0 && x < n_members, "member index out of range");
    QSX_CUDA(buf.upload(m, s));
    return QSX_OK;
}

extern "C" int qsx_zofe_apply(qsx_zofe_t h, const void *y_dev, void *dy_dev, int32_t n_columns,
                              const int32_t *member_host, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(h && y_dev && dy_dev && n_columns > 0 && y_dev != dy_dev, "qsx_zofe_apply: bad arguments");
    DevBuf<int> member;
    int rc;
    if (member_host && (rc = zofe_members(member, member_host, n_columns, h->d.n_members, stream))) return rc;
    ZofeApplyArgs a;
    a.Z = h->d; a.member_of = member_host ? member.p : nullptr;
    a.x = (const cplx *)y_dev; a.y = (cplx *)dy_dev;
    size_t smem = zofe_smem(h->d);
    QSX_CUDA(cudaFuncSetAttribute(zofe_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    zofe_apply_kernel<<<n_columns, 256, smem, stream>>>(a);
    qsx_launch_counter += 1;
    QSX_CUDA(cudaGetLastError());
    QSX_CUDA(cudaStreamSynchronize(stream));
    return QSX_OK;
}

extern "C" int qsx_zofe_propagate(qsx_zofe_t h, qsx_propagate_args *args, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(h && args, "qsx_zofe_propagate: null argument");
    const ZofeDev &d = h->d;
    const int B = args->n_columns, nt = args->n_times;
    QSX_REQUIRE(B > 0 && nt > 0 && args->t_host && args->y0_dev && args->out_dev,
                "qsx_zofe_propagate: empty batch or missing buffers");
    QSX_REQUIRE(args->method == QSX_METHOD_RK4 || args->method == QSX_METHOD_DOPRI5,
                "the ZOFE equation is nonlinear: use the RK4 or DOPRI5 integrator");
    QSX_REQUIRE(args->n_pulses >= 0 && args->n_pulses <= QSX_MAX_PULSES, "too many pulses");
    QSX_REQUIRE(args->n_pulses == 0 || (args->pulse_ops_dev && args->n_pulse_sets == 1),
                "ZOFE pulse operators must be one shared set of [n_pulses][n][n] Hilbert-space matrices");
    for (int i = 1; i < nt; ++i)
        QSX_REQUIRE(args->t_host[i] >= args->t_host[i - 1], "output times must be non-decreasing");
    QSX_REQUIRE(args->t_host[0] >= args->t0, "first output time precedes t0");
    const int n_vec = args->method == QSX_METHOD_RK4 ? 4 : 10;
    DevBuf<int> member;
    DevBuf<double> d_t;
    DevBuf<unsigned long long> stats;
    DevBuf<cplx> work;
    int rc;
    if (args->generator_of_column_host &&
        (rc = zofe_members(member, args->generator_of_column_host, B, d.n_members, stream)))
        return rc;
    QSX_CUDA(d_t.upload(args->t_host, nt, stream));
    QSX_CUDA(stats.alloc(3));
    QSX_CUDA(cudaMemsetAsync(stats.p, 0, 3 * sizeof(unsigned long long), stream));
    QSX_CUDA(work.alloc((size_t)B * n_vec * d.dim));
    ZofeKernelArgs a;
    a.Z = d; a.nt = nt; a.n_vec = n_vec;
    a.member_of = args->generator_of_column_host ? member.p : nullptr;
    a.y0 = (const cplx *)args->y0_dev; a.work = work.p; a.t = d_t.p; a.t0 = args->t0;
    a.method = args->method;
    a.rtol = args->rtol > 0 ? args->rtol : 1e-10;
    a.atol = args->atol > 0 ? args->atol : 1e-12;
    a.rk4_sub = args->rk4_substeps > 0 ? args->rk4_substeps : 16;
    a.save_mode = args->save_mode; a.save_rows = args->save_rows; a.S = (const cplx *)args->save_dev;
    const long long head = (long long)d.n * d.n;
    if (a.save_mode == QSX_SAVE_MATRIX) {
        QSX_REQUIRE(a.S && a.save_rows > 0 && args->n_save == 1, "ZOFE save matrix must be [rows][n^2]");
        a.saved_dim = a.save_rows;
    } else {
        a.saved_dim = a.save_mode == QSX_SAVE_ADO0 ? head : d.dim;
    }
    a.out = (cplx *)args->out_dev;
    a.stats = stats.p;
    a.n_pulse = args->n_pulses;
    for (int p = 0; p < QSX_MAX_PULSES; ++p) a.pulses[p] = args->pulses[p];
    a.Vp = (const cplx *)args->pulse_ops_dev;
    size_t smem = zofe_smem(d);
    QSX_CUDA(cudaFuncSetAttribute(zofe_propagate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    QSX_CUDA(cudaEventCreate(&e0));
    QSX_CUDA(cudaEventCreate(&e1));
    QSX_CUDA(cudaEventRecord(e0, stream));
    zofe_propagate_kernel<<<B, 256, smem, stream>>>(a);
    qsx_launch_counter += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        qsx_set_error("zofe_propagate launch: %s", cudaGetErrorString(e));
        return QSX_ERR_CUDA;
    }
    QSX_CUDA(cudaEventRecord(e1, stream));
    unsigned long long st[3] = {0, 0, 0};
    qsx_d2h_counter += sizeof(st);
    QSX_CUDA(cudaMemcpyAsync(st, stats.p, sizeof(st), cudaMemcpyDeviceToHost, stream));
    QSX_CUDA(cudaStreamSynchronize(stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    args->rhs_evaluations = st[0];
    args->accepted_steps = st[1];
    args->kernel_ms = ms;
    if (st[2] != 0) {
        qsx_set_error("ZOFE integration failed (step-size underflow or non-finite state)");
        return QSX_ERR_INTEGRATOR;
    }
    return QSX_OK;
}

// K4 (CTA-resident form): fused time integrators for systems small enough that
// one thread block owns a whole batch of state vectors.  Replaces the Python
// loop around scipy's ZVODE in the reference (simulate/utils.py:16-50): the
// whole trajectory, including the save_func epilogue, runs inside one kernel
// with no host round trip.
//
// The right-hand side is a functor with
//     template <class Epi> void apply(const cplx *x, double t, Epi epi)
// which evaluates f(t, x) and calls epi(idx, value) exactly once per output
// element, from the thread that produced it.  Every integrator stage is thus
// "apply + element-local epilogue + one barrier".
#pragma once
#include "common.cuh"

struct CtaProp {
    int n;            // complex elements per vector owned by the CTA (rows * NB; element idx is column idx % NB)
    int method;
    double rtol, atol;
    int rk4_sub;
    int kmax;         // Taylor: maximum order per sub-step
    double theta;     // Taylor: |h| * lnorm <= theta
    double lnorm;     // inf-norm bound of the generator (Taylor sub-stepping, DP5 first step)
    int nt;
    const double *t;  // device, [nt]
    double t0;
};

struct CtaStats {
    unsigned long long rhs;
    unsigned long long steps;
    int status;
};

template <int NB>
__device__ __forceinline__ void block_max_cols(double (&loc)[NB], double *scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int j = 0; j < NB; ++j) loc[j] = warp_max(loc[j]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NB; ++j) scratch[w * NB + j] = loc[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        double r = scratch[j];
        for (int i = 1; i < nw; ++i) r = fmax(r, scratch[i * NB + j]);
        loc[j] = r;
    }
}

// vec: base of the CTA's vector workspace (shared or global memory), laid out
// as consecutive vectors of P.n elements; needs 3 (Taylor), 4 (RK4) or 10
// (DOPRI5) vectors.  vec[0..n) holds the state on entry and on every save.
// scratch: >= 32*NB doubles of shared memory.
template <int NB, class Rhs, class Saver>
__device__ void cta_propagate(Rhs &rhs, Saver &save, const CtaProp &P, cplx *vec,
                              double *scratch, CtaStats &st) {
    const int n = P.n;
    const int tid = threadIdx.x, nthr = blockDim.x;
    cplx *Y = vec;
    double tcur = P.t0;
    double h_dp = 0.0;          // DOPRI5 step carried across output points
    bool have_k1 = false;       // DOPRI5 FSAL
    st.rhs = 0; st.steps = 0; st.status = 0;

    for (int it = 0; it < P.nt; ++it) {
        const double target = P.t[it];
        if (target != tcur) {
            const double span = target - tcur;
            if (P.method == QSX_METHOD_TAYLOR) {
                // ---- adaptive-order Taylor expansion of exp(h L) y ----------
                cplx *V = vec + n, *W = vec + 2 * n;
                int nsub = (int)ceil(fabs(span) * P.lnorm / P.theta);
                if (nsub < 1) nsub = 1;
                const double h = span / nsub;
                for (int s = 0; s < nsub; ++s) {
                    double ynorm[NB];
#pragma unroll
                    for (int j = 0; j < NB; ++j) ynorm[j] = 0.0;
                    for (int i = tid; i < n; i += nthr) {
                        cplx y0 = Y[i];
                        W[i] = y0;              // zero-th term; Y itself is only accumulated into
                        double a = cabs1(y0);
                        int j = i % NB;
#pragma unroll
                        for (int jj = 0; jj < NB; ++jj) if (jj == j) ynorm[jj] = fmax(ynorm[jj], a);
                    }
                    block_max_cols<NB>(ynorm, scratch);
                    const cplx *src = W;
                    cplx *dst = V;
                    bool done = false;
                    int k = 1;
                    for (; k <= P.kmax; ++k) {
                        const double fac = h / k;
                        int ok = 1;
                        rhs.apply(src, tcur, [&](int i, cplx f) {
                            cplx w = cscale(fac, f);
                            dst[i] = w;
                            cplx y = Y[i];
                            Y[i] = cadd(y, w);
                            // previous term: zero-th term is y itself, never "small"
                            double prev = (k == 1) ? 1e300 : cabs1(src[i]);
                            int j = i % NB;
                            double yn = 0.0;
#pragma unroll
                            for (int jj = 0; jj < NB; ++jj) if (jj == j) yn = ynorm[jj];
                            if (!(prev + cabs1(w) <= P.rtol * yn)) ok = 0;      // also catches a non-finite state
                        });
                        st.rhs += 1;
                        int all_ok = __syncthreads_and(ok);
                        src = dst;
                        dst = (dst == V) ? W : V;
                        if (all_ok) { done = true; break; }
                    }
                    if (!done) st.status = QSX_ERR_INTEGRATOR;
                    st.steps += 1;
                    tcur += h;
                }
            } else if (P.method == QSX_METHOD_MAP) {
                // the "generator" is the one-step propagator exp(L dt): y <- P y
                cplx *TA = vec + n;
                rhs.apply(Y, tcur, [&](int i, cplx f) { TA[i] = f; });
                __syncthreads();
                for (int i = tid; i < n; i += nthr) Y[i] = TA[i];
                __syncthreads();
                st.rhs += 1;
                st.steps += 1;
            } else if (P.method == QSX_METHOD_RK4) {
                // ---- classic RK4, fixed sub-steps -----------------------------
                cplx *ACC = vec + n, *TA = vec + 2 * n, *TB = vec + 3 * n;
                const int nsub = P.rk4_sub > 0 ? P.rk4_sub : 1;
                const double h = span / nsub;
                for (int s = 0; s < nsub; ++s) {
                    rhs.apply(Y, tcur, [&](int i, cplx k1) {
                        cplx y = Y[i];
                        TA[i] = cadd(y, cscale(0.5 * h, k1));
                        ACC[i] = cadd(y, cscale(h / 6.0, k1));
                    });
                    __syncthreads();
                    rhs.apply(TA, tcur + 0.5 * h, [&](int i, cplx k2) {
                        TB[i] = cadd(Y[i], cscale(0.5 * h, k2));
                        ACC[i] = cadd(ACC[i], cscale(h / 3.0, k2));
                    });
                    __syncthreads();
                    rhs.apply(TB, tcur + 0.5 * h, [&](int i, cplx k3) {
                        TA[i] = cadd(Y[i], cscale(h, k3));
                        ACC[i] = cadd(ACC[i], cscale(h / 3.0, k3));
                    });
                    __syncthreads();
                    rhs.apply(TA, tcur + h, [&](int i, cplx k4) {
                        Y[i] = cadd(ACC[i], cscale(h / 6.0, k4));    // Y is not an input of this stage
                    });
                    __syncthreads();
                    st.rhs += 4;
                    st.steps += 1;
                    tcur += h;
                }
            } else {
                // ---- Dormand-Prince 5(4) with on-device step-size control ------
                cplx *K1 = vec + n, *K2 = vec + 2 * n, *K3 = vec + 3 * n, *K4 = vec + 4 * n,
                     *K5 = vec + 5 * n, *K6 = vec + 6 * n, *K7 = vec + 7 * n,
                     *TA = vec + 8 * n, *TB = vec + 9 * n;
                const double dir = span >= 0 ? 1.0 : -1.0;
                if (!have_k1) {
                    rhs.apply(Y, tcur, [&](int i, cplx f) { K1[i] = f; });
                    __syncthreads();
                    st.rhs += 1;
                    have_k1 = true;
                }
                if (h_dp == 0.0) {
                    // first-step guess (Hairer II.4): 0.01 * |y| / |f| in the scaled norm
                    double d0 = 0.0, d1 = 0.0;
                    for (int i = tid; i < n; i += nthr) {
                        double sc = P.atol + P.rtol * sqrt(cabs2(Y[i]));
                        d0 += cabs2(Y[i]) / (sc * sc);
                        d1 += cabs2(K1[i]) / (sc * sc);
                    }
                    d0 = sqrt(block_sum(d0, scratch) / n);
                    d1 = sqrt(block_sum(d1, scratch) / n);
                    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
                    if (P.lnorm > 0) h0 = fmin(h0, 0.1 / P.lnorm);
                    h_dp = h0;
                }
                int guard = 0;
                while ((target - tcur) * dir > 0) {
                    double h = dir * fabs(h_dp);
                    bool clipped = false;
                    if ((tcur + h - target) * dir >= 0 || fabs(target - (tcur + h)) < 1e-12 * fabs(h)) {
                        h = target - tcur;
                        clipped = true;
                    }
                    // stage 2 input
                    for (int i = tid; i < n; i += nthr)
                        TA[i] = cadd(Y[i], cscale(h * DP_A21, K1[i]));
                    __syncthreads();
                    rhs.apply(TA, tcur + DP_C2 * h, [&](int i, cplx f) {
                        K2[i] = f;
                        cplx a = Y[i];
                        rfma(a, h * DP_A31, K1[i]); rfma(a, h * DP_A32, f);
                        TB[i] = a;
                    });
                    __syncthreads();
                    rhs.apply(TB, tcur + DP_C3 * h, [&](int i, cplx f) {
                        K3[i] = f;
                        cplx a = Y[i];
                        rfma(a, h * DP_A41, K1[i]); rfma(a, h * DP_A42, K2[i]); rfma(a, h * DP_A43, f);
                        TA[i] = a;
                    });
                    __syncthreads();
                    rhs.apply(TA, tcur + DP_C4 * h, [&](int i, cplx f) {
                        K4[i] = f;
                        cplx a = Y[i];
                        rfma(a, h * DP_A51, K1[i]); rfma(a, h * DP_A52, K2[i]);
                        rfma(a, h * DP_A53, K3[i]); rfma(a, h * DP_A54, f);
                        TB[i] = a;
                    });
                    __syncthreads();
                    rhs.apply(TB, tcur + DP_C5 * h, [&](int i, cplx f) {
                        K5[i] = f;
                        cplx a = Y[i];
                        rfma(a, h * DP_A61, K1[i]); rfma(a, h * DP_A62, K2[i]);
                        rfma(a, h * DP_A63, K3[i]); rfma(a, h * DP_A64, K4[i]); rfma(a, h * DP_A65, f);
                        TA[i] = a;
                    });
                    __syncthreads();
                    rhs.apply(TA, tcur + h, [&](int i, cplx f) {
                        K6[i] = f;
                        cplx a = Y[i];
                        rfma(a, h * DP_A71, K1[i]); rfma(a, h * DP_A73, K3[i]);
                        rfma(a, h * DP_A74, K4[i]); rfma(a, h * DP_A75, K5[i]); rfma(a, h * DP_A76, f);
                        TB[i] = a;                                     // 5th-order solution
                    });
                    __syncthreads();
                    double errsq = 0.0;
                    rhs.apply(TB, tcur + h, [&](int i, cplx f) {
                        K7[i] = f;
                        cplx e = cmake(0, 0);
                        rfma(e, DP_E1, K1[i]); rfma(e, DP_E3, K3[i]); rfma(e, DP_E4, K4[i]);
                        rfma(e, DP_E5, K5[i]); rfma(e, DP_E6, K6[i]); rfma(e, DP_E7, f);
                        double sc = P.atol + P.rtol * sqrt(fmax(cabs2(Y[i]), cabs2(TB[i])));
                        errsq += (h * h) * cabs2(e) / (sc * sc);
                    });
                    st.rhs += 6;
                    double err = sqrt(block_sum(errsq, scratch) / n);   // barriers inside
                    bool finite = (err == err) && err < 1e300;
                    double fac;
                    if (finite && err <= 1.0) {
                        for (int i = tid; i < n; i += nthr) { Y[i] = TB[i]; K1[i] = K7[i]; }
                        __syncthreads();
                        tcur = clipped ? target : tcur + h;
                        st.steps += 1;
                        fac = (err < 1e-10) ? 5.0 : fmin(5.0, fmax(0.2, 0.9 * pow(err, -0.2)));
                        if (!clipped || fac < 1.0) h_dp = fabs(h) * fac;
                        else h_dp = fmax(fabs(h_dp), fabs(h) * fmin(fac, 1.0));
                    } else {
                        fac = finite ? fmax(0.2, 0.9 * pow(err, -0.2)) : 0.2;
                        h_dp = fabs(h) * fmin(fac, 1.0);
                        __syncthreads();
                    }
                    if (h_dp < 1e-14 * fmax(1.0, fabs(tcur)) || ++guard > 20000000) {
                        st.status = QSX_ERR_INTEGRATOR;
                        tcur = target;
                        break;
                    }
                }
            }
            tcur = target;
        }
        save(it, Y);
        __syncthreads();
    }
}

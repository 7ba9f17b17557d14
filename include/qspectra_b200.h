/*
 * qspectra_b200 -- C ABI of the B200-native Liouville-space propagation engine.
 *
 * The reference (whaley-group-berkeley/qspectra) is pure Python and has no FFI;
 * its boundary for this path is the DynamicalModel plugin protocol.  Each entry
 * point below names the reference call site it replaces (paths relative to the
 * reference root).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - complex128 is passed as interleaved (re, im) doubles ("double2").
 *   - a density operator is vectorised column-major (liouville_space.py:46-50)
 *     and restricted to a Liouville subspace index list (:9-29) on the host.
 *   - state batches are row-major [column][state_dim]; trajectories are
 *     [column][time][saved_dim].
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).
 *   - every function returns 0 on success or a negative qsx_status; the message
 *     is available from qsx_last_error() (thread local).
 *   - pointers named *_dev are device pointers owned by the caller (torch
 *     tensors on the Python side); handles own their staged generators/tables.
 */
#ifndef QSPECTRA_B200_H
#define QSPECTRA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    QSX_OK = 0,
    QSX_ERR_INVALID = -1,      /* bad argument                           -> ValueError       */
    QSX_ERR_CUDA = -2,         /* CUDA runtime failure                   -> RuntimeError     */
    QSX_ERR_INTEGRATOR = -3,   /* step-size underflow / non-finite state -> IntegratorError  */
    QSX_ERR_UNSUPPORTED = -4   /* configuration outside the kernel limits-> NotImplementedError */
} qsx_status;

typedef enum {
    QSX_METHOD_TAYLOR = 0,     /* adaptive-order Taylor of exp(hL), LTI generators only */
    QSX_METHOD_RK4 = 1,        /* classic RK4, fixed sub-steps per output interval      */
    QSX_METHOD_DOPRI5 = 2,     /* Dormand-Prince 5(4), on-device step-size control      */
    QSX_METHOD_MAP = 3,        /* y <- P y per output interval; the handle holds the
                                  propagators P = exp(L dt) made by qsx_dense_expm      */
    QSX_METHOD_POLY = 4        /* HEOM, LTI generators: the Taylor polynomial of exp(hL) in product
                                  form, prod_j (I + h a_j L) y -- one state read and one write per RHS
                                  application; its degree comes from adaptive Taylor pilot intervals
                                  (rtol), refreshed periodically.  Handles without the row tile run
                                  QSX_METHOD_TAYLOR instead                              */
} qsx_method;

typedef enum {
    QSX_SAVE_STATE = 0,        /* save_func = identity                                  */
    QSX_SAVE_MATRIX = 1,       /* save_func = S . y  (commutator / left / right / expectation) */
    QSX_SAVE_ADO0 = 2          /* HEOM: state_vector_to_density_matrix, first M entries */
} qsx_save_mode;

#define QSX_MAX_PULSES 4

/* A Gaussian pulse in the rotating frame (pulse.py:110-114):
 *   E(t) = scale * exp(i*detuning*(t - t_peak) - (t - t_peak)^2 * inv_two_sigma_sq)
 * The RHS adds (-i*E) * C.y, with E conjugated when `conjugate` != 0 (eom.py:87-94). */
typedef struct {
    double scale, detuning, t_peak, inv_two_sigma_sq;
    int32_t conjugate;
    int32_t _pad;
} qsx_pulse;

/* Arguments of a propagation = one call of simulate/utils.py:53-109 `integrate`
 * for a batch of initial states that share one output grid. */
typedef struct {
    int32_t n_columns;            /* number of initial states (leading axis of y0)   */
    int32_t n_times;              /* len(t)                                          */
    const double *t_host;         /* output times, host pointer, increasing          */
    double t0;                    /* start time (utils.py:17-18)                     */
    const void *y0_dev;           /* [n_columns][D] complex128                       */
    const int32_t *generator_of_column_host; /* [n_columns] or NULL (all use 0)      */
    int32_t method;               /* qsx_method                                      */
    double rtol, atol;            /* tolerances (Taylor: rtol = per-step truncation) */
    int32_t rk4_substeps;         /* RK4 sub-steps per output interval               */
    int32_t save_mode;            /* qsx_save_mode                                   */
    int32_t save_rows;            /* rows of S when save_mode == QSX_SAVE_MATRIX     */
    const void *save_dev;         /* S: [n_save][save_rows][M_from] complex128       */
    int32_t n_save;               /* 1 (shared) or n_generators (per member)         */
    const int32_t *save_of_column_host; /* [n_columns] or NULL: column c is saved through
                                     S[save_of_column[c]] (any n_save; dense generators, and HEOM
                                     where S is a stack of per-ADO blocks) -- e.g. one dipole
                                     operator per polarisation configuration */
    int32_t n_pulses;             /* time-dependent terms, <= QSX_MAX_PULSES         */
    qsx_pulse pulses[QSX_MAX_PULSES];
    const void *pulse_ops_dev;    /* C_p: [n_pulse_sets][n_pulses][D][D] commutator blocks (dense),
                                     [1][n_pulses][M][M] per-ADO commutator blocks (HEOM),
                                     [1][n_pulses][n][n] Hilbert-space dipole operators (ZOFE) */
    int32_t n_pulse_sets;         /* 1 (shared) or n_generators                      */
    void *out_dev;                /* [n_columns][n_times][saved_dim] complex128      */
    /* results */
    uint64_t rhs_evaluations;     /* RHS applications summed over columns            */
    uint64_t accepted_steps;      /* integrator steps summed over columns            */
    double kernel_ms;             /* device time of the propagation kernel (CUDA events) */
} qsx_propagate_args;

const char *qsx_last_error(void);
int qsx_version(void);
/* number of kernels this library has launched since load (bench.py: gpu_launches) */
uint64_t qsx_kernel_launches(void);
/* bytes the library itself has copied host->device / device->host since load (handle tables,
 * per-call column maps, status words); the Python side adds its own tensor copies
 * (qspectra_b200._capi.transfer_bytes) for bench.py's e2e.h2d/d2h_bytes_per_step */
void qsx_transfer_bytes(uint64_t *h2d, uint64_t *d2h);
/* device properties the host side needs: SM count, L2 bytes, max smem per block */
int qsx_device_info(int32_t *sm_count, int64_t *l2_bytes, int32_t *smem_per_block);

/* ------------------------------------------------------------------------
 * Dense Liouvillians (RedfieldModel / UnitaryModel).
 * Replaces `evolve_matrix.dot(rho)` of LiouvilleSpaceModel.equation_of_motion
 * (dynamics/liouville_space.py:316-341) and the ZVODE loop around it
 * (simulate/utils.py:45-49).
 * ---------------------------------------------------------------------- */
typedef struct qsx_dense_s *qsx_dense_t;

/* L: [n_generators][M][M] complex128 row-major, already restricted to the
 * subspace (np.ix_(index, index)) and scaled by unit_convert.  `on_device` says
 * whether L is a device pointer.  `transpose` != 0 stores L^T: the Heisenberg
 * picture of liouville_space.py:325-330. */
int qsx_dense_create(qsx_dense_t *out, int32_t M, int32_t n_generators,
                     const void *L, int32_t on_device, int32_t transpose,
                     void *stream);
/* dy[c] = L[gen(c)] . y[c] for a batch of columns: the function returned by
 * equation_of_motion.  y/dy: [n_columns][M] device. */
int qsx_dense_apply(qsx_dense_t h, const void *y_dev, void *dy_dev,
                    int32_t n_columns, const int32_t *generator_of_column_host,
                    void *stream);
int qsx_dense_propagate(qsx_dense_t h, qsx_propagate_args *args, void *stream);
/* New handle holding P_g = exp(L_g * dt) for every generator of `h`, computed on the FP64
 * tensor cores (|A| <= 1/2 scaling, degree-12 Taylor polynomial in Paterson-Stockmeyer form: five
 * products, remainder < 2e-14; degree 14 in the tiled path for M > 56,
 * squarings; one CTA per generator up to M = 56, tiled GEMM launches up to M = 1024).  Use it
 * with QSX_METHOD_MAP on a uniform output grid of spacing dt: the exact counterpart of the
 * reference's ZVODE loop for a constant generator (simulate/utils.py:45-49). */
int qsx_dense_expm(qsx_dense_t h, double dt, void *Pt_dev, void *lnorm_dev,
                   qsx_dense_t *out, void *stream);
/* Handle over caller-owned device storage: Lt_dev [n_generators][M][M] in the engine's
 * transposed storage (Lt[g][c][r] = L_g[r][c]; what qsx_redfield_build* write with
 * transposed_out = 1 and qsx_dense_expm writes to Pt_dev) and lnorm_dev [n_generators]
 * float64 scratch that receives the inf-norms (formed by the first call that needs them).  No copy is made; the caller keeps the
 * buffers alive for the lifetime of the handle. */
int qsx_dense_wrap(qsx_dense_t *out, int32_t M, int32_t n_generators, void *Lt_dev,
                   void *lnorm_dev, void *stream);
/* Device time (CUDA events) and number of M x M complex GEMMs of the qsx_dense_expm call
 * that produced `h` (bench.py: FP64 tensor roofline). */
int qsx_dense_build_stats(qsx_dense_t h, double *kernel_ms, uint64_t *complex_gemms);
/* qsx_dense_expm and qsx_dense_propagate with QSX_METHOD_MAP return as soon as their kernel is
 * queued (no host synchronisation; neither can fail at run time): qsx_dense_build_stats waits for
 * the build, and a propagation that reports kernel_ms < 0 has its device time collected here. */
int qsx_dense_last_kernel_ms(qsx_dense_t h, double *kernel_ms);
/* 1 when the two calls above would not block (the recorded events have completed). */
int qsx_dense_events_ready(qsx_dense_t h);
void qsx_dense_destroy(qsx_dense_t h);

/* ------------------------------------------------------------------------
 * Hermitian-coordinate ("real form") propagation of dense generators.
 * Every physical generator on a Liouville subspace that is closed under transposition
 * ('ee', 'gg,ee', ... -- what simulate_dynamics propagates, reference simulate/eom.py:11-26 with
 * dynamics/liouville_space.py:316-341) commutes with Hermitian conjugation:
 * L[perm r][perm c] = conj L[r][c], perm[k] = position of the transposed ket-bra pair of element
 * k.  In the coordinates u_k = rho_k (perm k = k), u_a = Re rho_a, u_b = Im rho_a (pair a < b =
 * perm a) it is a REAL M x M matrix G and a Hermitian state a real vector, so exp(G dt) costs one
 * real tensor-core product per complex one and a step u <- P u a quarter of the complex
 * multiply-adds.  M <= 56.  defect_dev: four float64 the caller zeroes once,
 * [0] max |Im G|, [1] max |Re G|, [2] max |Im u0|, [3] max |Re u0| -- the caller checks
 * [0] <= eps [1] and [2] <= eps [3] after its next synchronisation (a generator or state that is
 * not Hermiticity-compatible must go through qsx_dense_propagate instead).
 * ---------------------------------------------------------------------- */
/* Gt_dev [n_generators][M][M] float64, transposed storage like the handle's generators;
 * gnorm_dev [n_generators] inf-norms of G. */
int qsx_dense_hermitian_form(qsx_dense_t h, const int32_t *perm_host, void *Gt_dev, void *gnorm_dev,
                             void *defect_dev, void *stream);
/* Both steps in one kernel (the default of the Python layer): every CTA forms the real generator
 * of its member from the handle's complex one in shared memory and goes straight into the series;
 * G never exists in global memory.  P_dev, gemm_count_dev as for qsx_real_expm; defect_dev [0..1]
 * as for qsx_dense_hermitian_form. */
int qsx_dense_hermitian_expm(qsx_dense_t h, const int32_t *perm_host, double dt, void *P_dev,
                             void *defect_dev, void *gemm_count_dev, void *stream);
/* P_dev [n_generators][M][M] float64 row-major = exp(G_g dt), the series of qsx_dense_expm on real
 * DMMA; gemm_count_dev: one uint64 counter (incremented by the number of M x M real products). */
int qsx_real_expm(const void *Gt_dev, const void *gnorm_dev, int32_t M, int32_t n_generators, double dt,
                  void *P_dev, void *gemm_count_dev, void *stream);
/* out[c][i][:] = P_gen(c)^i u0[c], i < n_times; rows of u0 and out are row_stride >= M float64 apart
 * (the tail of every output row is zero-filled).  generator_of_column_host == NULL: column c uses
 * generator c when n_columns == n_generators, else generator 0. */
int qsx_real_map(const void *P_dev, int32_t M, int32_t n_generators, const int32_t *generator_of_column_host,
                 int32_t n_columns, const void *u0_dev, int32_t n_times, int32_t row_stride, void *out_dev,
                 void *stream);
/* complex128 state vectors y [rows][M] <-> real coordinates u [rows][row_stride] */
int qsx_hermitian_pack(const void *y_dev, int32_t M, int64_t rows, const int32_t *perm_host, int32_t row_stride,
                       void *u_dev, void *defect_dev, void *stream);
int qsx_hermitian_unpack(const void *u_dev, int32_t M, int64_t rows, int32_t row_stride, const int32_t *perm_host,
                         void *y_dev, void *stream);

/* ------------------------------------------------------------------------
 * HEOM hierarchy (HEOMModel).  Replaces HEOM_tensor + csr_matrix.dot
 * (dynamics/heom.py:228-244, 298-443) with an index-map driven structured
 * apply; no CSR matrix is ever built.
 * ---------------------------------------------------------------------- */
typedef struct qsx_heom_s *qsx_heom_t;

typedef struct {
    int32_t n_sites;              /* independent baths (heom.py:221)                 */
    int32_t K;                    /* Matsubara terms beyond the Drude pole           */
    int32_t level_cutoff;         /* ADOs with sum(n) < level_cutoff (heom.py:92-152)*/
    int32_t n_hilbert;            /* N = states of the Hilbert subspace              */
    int32_t M;                    /* size of the Liouville subspace                  */
    const int64_t *subspace_index;/* [M] flat column-major positions (host)          */
    int32_t n_members;            /* distinct Hamiltonians (disorder ensemble), >= 1 */
    const void *H;                /* [n_members][N][N] complex128 row-major (host), rotating frame */
    const double *coupling_diag;  /* [n_sites][N]: diagonal of V_j (hamiltonian.py:593-608) */
    const double *nu;             /* [K+1] Matsubara frequencies (heom.py:61-67)     */
    const void *c;                /* [K+1] complex coefficients (heom.py:69-89)      */
    double temp_corr;             /* sum_{k>K} c_k/nu_k, 0 if low_temp_corr is off (heom.py:387-393) */
    double unit_convert;
    int32_t modified;             /* modified_HEOM scaling (heom.py:423-437)         */
    int32_t heisenberg;           /* generator transposed (heom.py:236-237)          */
} qsx_heom_config;

int qsx_heom_create(qsx_heom_t *out, const qsx_heom_config *cfg, void *stream);
int64_t qsx_heom_ado_count(qsx_heom_t h);
/* integer artefacts for bit-exact comparison with ADO_mappings (heom.py:92-152):
 * ado_index [n_ado][n_sites*(K+1)], up/down [n_ado][bins] (-1 = absent). Host buffers. */
int qsx_heom_index_maps(qsx_heom_t h, int64_t *ado_index, int32_t *up,
                        int32_t *down);
/* dy = L_heom . y for [n_columns][n_ado*M] states; member_of_column selects H. */
int qsx_heom_apply(qsx_heom_t h, const void *y_dev, void *dy_dev,
                   int32_t n_columns, const int32_t *member_of_column_host,
                   void *stream);
int qsx_heom_propagate(qsx_heom_t h, qsx_propagate_args *args, void *stream);
void qsx_heom_destroy(qsx_heom_t h);

/* Closed-form ADO enumeration without a handle (host only; used by the
 * bit-exact index-map tests and by HEOMModel.ado_indices). */
int64_t qsx_ado_count(int32_t bins, int32_t level_cutoff);
int qsx_ado_enumerate(int32_t bins, int32_t level_cutoff, int64_t *ado_index,
                      int32_t *up, int32_t *down);


/* ------------------------------------------------------------------------
 * ZOFE master equation (ZOFEModel).  Replaces rhodot_oopdot_vec and the
 * closure returned by equation_of_motion (dynamics/zofe.py:121-234).
 * State layout = reference: [vec_F(rho) ; vec_F(O)], O[p,s,a,b] at
 * n^2 + p + P (s + S (a + n b))  (zofe.py:84, 110-119).
 * QSX_SAVE_ADO0 saves the first n^2 entries (rho), QSX_SAVE_MATRIX applies a
 * [rows][n^2] matrix to them (expectation values, zofe.py:38-41).
 * ---------------------------------------------------------------------- */
typedef struct qsx_zofe_s *qsx_zofe_t;

typedef struct {
    int32_t n_states;             /* n: states of the Hilbert subspace              */
    int32_t n_sites;              /* S                                              */
    int32_t n_pm;                 /* P pseudomodes (bath.py:105-144)                */
    int32_t n_members;            /* distinct Hamiltonians (ensemble), >= 1         */
    const void *H;                /* [n_members][n][n] complex128 row-major (host)  */
    const double *coupling_diag;  /* [S][n] diagonals of V_s; L_s = -V_s (zofe.py:216-217) */
    const void *Gamma;            /* [P][S] complex128: Omega^2 * huang (zofe.py:224) */
    const void *w;                /* [P][S] complex128: i Omega + gamma (zofe.py:225) */
    double unit_convert;
    int32_t ham_hermit, rho_hermit; /* reference shortcut flags (zofe.py:71-87)     */
} qsx_zofe_config;

int qsx_zofe_create(qsx_zofe_t *out, const qsx_zofe_config *cfg, void *stream);
int64_t qsx_zofe_state_dim(qsx_zofe_t h);
int qsx_zofe_apply(qsx_zofe_t h, const void *y_dev, void *dy_dev,
                   int32_t n_columns, const int32_t *member_of_column_host,
                   void *stream);
int qsx_zofe_propagate(qsx_zofe_t h, qsx_propagate_args *args, void *stream);
void qsx_zofe_destroy(qsx_zofe_t h);

/* ------------------------------------------------------------------------
 * Batched Redfield generator construction for disorder ensembles (K5).
 * Replaces redfield_evolve / redfield_tensor (dynamics/redfield.py:9-104) and
 * DebyeBath.corr_func_complex / Bath.corr_func_real (bath.py:17-31, 84-102),
 * which the reference re-runs for every ensemble member (base.py:120-128).
 * ---------------------------------------------------------------------- */
typedef enum {
    QSX_BATH_DEBYE_COMPLEX = 0,  /* DebyeBath.corr_func_complex, Matsubara sum   */
    QSX_BATH_DEBYE_REAL = 1      /* Bath.corr_func_real (discard_imag_corr=True) */
} qsx_bath_kind;

typedef struct {
    int32_t kind;                /* qsx_bath_kind                                */
    int32_t matsubara_cutoff;    /* 1000 in the reference (bath.py:84)           */
    double temperature, reorg_energy, cutoff_freq;
} qsx_bath;

/* E_dev [n_members][N] (float64) and U_dev [n_members][N][N] (complex128,
 * U[x][a] = <site x | eigenstate a>) are the members' eigen-systems in the
 * rotating frame (hamiltonian.py:310-328); coupling_diag_host [n_baths][N] the
 * diagonals of the system-bath operators; subspace_index_host [M] the Liouville
 * subspace.  Writes unit_convert * L[idx, idx] to L_out_dev [n_members][M][M]
 * (row-major, or the transposed storage of qsx_dense_wrap when transposed_out != 0). */
int qsx_redfield_build(int32_t n_members, int32_t N, const void *E_dev,
                       const void *U_dev, int32_t n_baths,
                       const double *coupling_diag_host, const qsx_bath *bath,
                       int32_t secular, int32_t eigen_basis, double unit_convert,
                       int32_t M, const int64_t *subspace_index_host,
                       int32_t transposed_out, void *L_out_dev, void *stream);

/* Same, with the members' eigensystems computed on the device too (cyclic Jacobi):
 * member m has the real symmetric lab-frame Hamiltonian
 *   H_m = H0 + diag(sum_j site_shifts[m][j] * coupling_diag[j][.])
 * (static diagonal disorder, hamiltonian.py:458-461); its eigen-energies are moved to the
 * rotating frame as E[a] - quanta[a] * rw_freq (hamiltonian.py:310-328).
 * H0_host [N][N], quanta_host [N] (0/1/2 per basis state), site_shifts_dev [n_members][n_baths]. */
int qsx_redfield_build_sampled(int32_t n_members, int32_t N, const double *H0_host,
                               const void *site_shifts_dev, const double *quanta_host,
                               double rw_freq, int32_t n_baths,
                               const double *coupling_diag_host, const qsx_bath *bath,
                               int32_t secular, int32_t eigen_basis, double unit_convert,
                               int32_t M, const int64_t *subspace_index_host,
                               int32_t transposed_out, void *L_out_dev, void *stream);

/* ------------------------------------------------------------------------
 * K6 (single-GPU part): weighted sum over ensemble members / columns,
 *   out[i] = scale * sum_m in[m][i]   (complex128, i < n)
 * Replaces `total_signal += signal; total_signal /= ensemble_size`
 * (simulate/decorators.py:55-61).  The cross-GPU part is one NCCL reduce issued
 * from the host side (torch.distributed).
 * ---------------------------------------------------------------------- */
int qsx_reduce_members(const void *in_dev, int32_t n_members, int64_t n,
                       double scale, void *out_dev, void *stream);

/* ------------------------------------------------------------------------
 * K6 (signal contraction of the third-order response):
 *   S[ab][c] += sum_u w[u] * sum_i X[u][ab][i] * Y[u][c][i]        (complex128, no conjugation)
 * X: [n_units][n_ab][K] = V_rho2 of every unit (ensemble member x polarisation configuration),
 * ab = (t1, t2) flattened; Y: [n_units][n_c][K] = the Heisenberg-propagated detection vectors
 * over t3; w: [n_units] complex weights (isotropic-average weight of the unit's configuration).
 * Replaces `np.einsum('ci,abi', V_Gt3, V_rho2)` (simulate/response.py:336) and the weighted
 * sums around it (decorators.py:55-61, 86-92) with one tensor-core GEMM launch whose CTAs walk
 * all units (deterministic summation order).
 * ---------------------------------------------------------------------- */
int qsx_response_contract(const void *x_dev, const void *y_dev, const void *w_dev,
                          int32_t n_units, int64_t n_ab, int32_t n_c, int32_t K,
                          void *s_dev, void *stream);
/* The same with units that share their Y operand multiplied once: group g contracts
 *   X'_g = sum_{j < grp_count[g]} w[grp_first[g] + j] * X[grp_first[g] + j]   (summed while the tile is staged)
 * with Y[g] -- e.g. the polarisation configurations of the isotropic average that end in the same
 * detection polarisation (3 products per member instead of 21).  X: [n_x][n_ab][K], Y: [n_groups][n_c][K]. */
int qsx_response_contract_grouped(const void *x_dev, int32_t n_x, const void *y_dev, const void *w_dev,
                                  int32_t n_groups, const int32_t *grp_first_host,
                                  const int32_t *grp_count_host, int64_t n_ab, int32_t n_c, int32_t K,
                                  void *s_dev, void *stream);

/* ------------------------------------------------------------------------
 * K7: Fourier transform of a response function sampled on t = 0, dt, ..., (n-1) dt,
 * along the middle axis of x[outer][n][inner] (complex128):
 *   out[o][k][i] = dt * sum_j x[o][j][i] * exp(sign * 2 pi i * j * (k - n + 1) / (2n - 1)),
 *   k = 0 .. 2n - 2.
 * This is exactly what `fourier_transform` (simulate/utils.py:154-219) computes
 * through zero-padding to a grid symmetric around t = 0 (`_symmetrize`,
 * utils.py:128-151), ifftshift -> fft -> fftshift and the flip for sign = +1; the
 * frequency axis (fftfreq, rw_freq shift) stays on the host.  Called twice by
 * `two_dimensional_spectra` (response.py:430-455).  n <= 4096.
 * ---------------------------------------------------------------------- */
int qsx_fourier_transform(const void *x_dev, int64_t outer, int32_t n, int64_t inner,
                          double dt, int32_t sign, void *out_dev, void *stream);


/* ------------------------------------------------------------------------
 * Seeded disorder streams (host only).  Bit-exact replay of
 *   rng = numpy.random.RandomState(list(seed) + [n]); rng.randn(n_gauss); rng.rand(n_uniform)
 * for members n = member0 .. member0+n_members-1: the draws of
 * ElectronicHamiltonian._sample (hamiltonian.py:458-461, 552-578) and
 * random_rotation_matrix (polarization.py:96).  gauss_out [n_members][n_gauss],
 * uniform_out [n_members][n_uniform].
 * ---------------------------------------------------------------------- */
int qsx_sample_streams(const uint32_t *seed_prefix, int32_t n_prefix,
                       int64_t member0, int32_t n_members, int32_t n_gauss,
                       int32_t n_uniform, double *gauss_out, double *uniform_out);

/* The Gaussian part of the same streams generated on the GPU, one thread per member:
 *   out_dev[m][i] = scale * RandomState(list(seed) + [member0 + m]).randn(n_gauss)[i]
 * (integer stream and uniform doubles bit-identical with numpy; the Box-Muller
 * log/sqrt are the device functions, <= 1 ulp from the host libm).  Used for the
 * static-disorder shifts of an ensemble (hamiltonian.py:458-461, 566-573) so that
 * neither the host replay nor an H2D copy sits on the end-to-end path. */
int qsx_sample_gauss_device(const uint32_t *seed_prefix, int32_t n_prefix, int64_t member0,
                            int32_t n_members, int32_t n_gauss, double scale, void *out_dev,
                            void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QSPECTRA_B200_H */

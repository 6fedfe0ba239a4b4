/*
 * sddc_b200.h -- C ABI of the B200-native time-stepping / JVP hot path of the axisymmetric spherical-shell
 * double-diffusive convection solver (reference: mannixp/SpectralDoubleDiffusiveConvection).
 *
 * The reference has no FFI: its hot path is a set of Python free functions resolved by module name at call
 * time.  Every entry point below names the reference function(s) it replaces.
 * Conventions:
 *   - all arrays are float64, C-contiguous;
 *   - a member state is X = [psi | T | S], each field K = N_fm blocks of n = N_r-1 radial values
 *     (Matrix_Operators.py:529-575); an ensemble is [B][3*K*n];
 *   - device entry points take CALLER-OWNED DEVICE pointers and a cudaStream_t (passed as void*); they are
 *     stream-ordered, non-blocking and never allocate after plan creation (CUDA-graph capturable);
 *   - *_host entry points take HOST pointers and include the host<->device copies (blocking);
 *   - every function returns 0 on success, a negative sddc_status otherwise; nothing throws across the ABI;
 *   - a plan is not thread-safe; distinct plans are independent.
 */
#ifndef SDDC_B200_H
#define SDDC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sddc_plan sddc_plan;

enum sddc_status {
    SDDC_OK = 0,
    SDDC_ERR_INVALID = -1,     /* bad argument (reference raises ValueError, e.g. odd N_fm: Matrix_Operators.py:758) */
    SDDC_ERR_CUDA = -2,        /* CUDA runtime error, see sddc_last_error */
    SDDC_ERR_UNSUPPORTED = -3, /* shape outside what the kernels are instantiated for */
    SDDC_ERR_NO_DEVICE = -4
};

/* Scalar configuration: the arguments of Main.Build_Matrix_Operators (Main.py:179) + symmetric flag
 * (Main.py:242-245) + capacity. */
typedef struct {
    int N_fm;      /* number of latitudinal modes K (multiple of 4) */
    int N_r;       /* Chebyshev order; n = N_r - 1 interior points */
    int symmetric; /* 0/1: equatorial symmetry (Main.Eq_SYM, Main.py:137-176) */
    int max_batch; /* largest B any call will use; scratch is sized for it */
    int device;    /* CUDA device ordinal */
    double dt, Pr, Tau, d;
    int flags;     /* SDDC_FLAG_* (0: defaults) */
} sddc_config;

/* Keep the dense DMMA transforms where the FFT formulation of the nonlinear term would be used (N_fm = 128, 256,
 * 512): lets the tests hold the path every other N_fm takes against the golden vectors of the headline shape. */
#define SDDC_FLAG_DENSE_TRANSFORMS 1

/* Host-built radial operators, computed with the reference's formulas (cheb_radial, Nabla2, Nabla4,
 * A4_TSTEP_MATS, NAB2_TSTEP_MATS: Matrix_Operators.py:10-76,1014-1030,1089-1112; Main.py:196-222).
 * They are copied to the device at plan creation; the plan owns the copies. */
typedef struct {
    const double* Dr;      /* [n*n]  D[1:-1,1:-1]                                  (Matrix_Operators.py:764) */
    const double* Dsq;     /* [n*n]  (D@D)[1:-1,1:-1]                              (Matrix_Operators.py:217) */
    const double* D2r;     /* [n*n]  (diag(1/R^2) D D)[1:-1,1:-1]                  (Matrix_Operators.py:498) */
    const double* D2;      /* [n*n]  (IR2 (2 D^2 - 4 IR D + 6 IR2))[1:-1,1:-1]     (Main.py:212) */
    const double* r2;      /* [n]    R[1:-1]^2                                     (Matrix_Operators.py:85) */
    const double* ir2;     /* [n]    1/R[1:-1]^2                                   (Matrix_Operators.py:216) */
    const double* ir4;     /* [n]    1/R[1:-1]^4                                   (Matrix_Operators.py:496) */
    const double* a4_ir2;  /* [n]    diag of IR2 passed to A4_BSub_TSTEP_V2        (Main.py:213) */
    const double* a4_ir4;  /* [n]    diag of IR4                                   (Main.py:214) */
    const double* dT0;     /* [n]    A_T / R[1:-1]^2                               (Main.py:201-202) */
    const double* gbuoy;   /* [n]    (1/d)^2 / R[1:-1]^2                           (Matrix_Operators.py:113) */
    const double* ir;      /* [n]    1/R[1:-1]                                     (Main.py:102) */
    const double* nu_in;   /* [n]    (R[0]^2/A_T)  * D[0,1:-1]                     (Main.py:58-62) */
    const double* nu_out;  /* [n]    (R[-1]^2/A_T) * D[-1,1:-1] */
    const double* r;       /* [n]    R[1:-1] */
    double R_in, R_out;    /*        R[0], R[-1] */
    const double* Linv_A4; /* [K*n*n] descending-mode order as returned by A4_TSTEP_MATS  */
    const double* Linv_T;  /* [K*n*n] NAB2_TSTEP_MATS(dt)      */
    const double* Linv_S;  /* [K*n*n] NAB2_TSTEP_MATS(Tau*dt)  */
} sddc_operators;

/* linear operator codes for sddc_linear_op */
enum sddc_linop {
    SDDC_OP_J_THETA = 0,    /* J_theta_RT      (Matrix_Operators.py:436-472) */
    SDDC_OP_DT0_THETA = 1,  /* DT0_theta       (Matrix_Operators.py:131-189) */
    SDDC_OP_A2_SINE = 2,    /* A2_SINE         (Matrix_Operators.py:192-245) */
    SDDC_OP_A2_SINE_R2 = 3, /* A2_SINE_R2      (Matrix_Operators.py:475-526) */
    SDDC_OP_KGR = 4,        /* kGR_RT(...).dot (Matrix_Operators.py:97-128)  */
    SDDC_OP_R2 = 5          /* R2(...).dot     (Matrix_Operators.py:79-94)   */
};

/* transform kinds for sddc_transform (Transforms.py:73-129) */
enum sddc_transform_kind { SDDC_T_IDCT = 0, SDDC_T_IDST = 1, SDDC_T_DCT = 2, SDDC_T_DST = 3 };

int sddc_version(void);
int sddc_device_count(void);

/* replaces Main.Build_Matrix_Operators (device side of it) */
int sddc_plan_create(sddc_plan** out, const sddc_config* cfg, const sddc_operators* ops);
void sddc_plan_destroy(sddc_plan* plan);
const char* sddc_last_error(const sddc_plan* plan); /* plan may be NULL: error of the last failed create */
/* number of this library's kernel launches issued through the plan so far */
long long sddc_launch_count(const sddc_plan* plan);
/* kernel-selection facts of a plan: what = 0 second mirror level (quarter-wave split) of the dense transforms active,
 * 1 dense synthesis kernel (0 generic, 1 persistent warp-specialised), 2 dense two-state JVP synthesis available,
 * 3 padded radial size, 4 grid size M of the FFT formulation of the nonlinear term (N_fm = 128, 256, 512; 0: dense
 * DMMA transforms), 5 FFT formulation also used for the two-state (JVP) products, 6 grid size of the kinetic-energy
 * FFT (0: dense synthesis), 7 direct-summation row kernel (1 every product, 2 two-state products only), 8 the hot
 * back-substitution reads the analysed products of the row kernels itself (three kernels per member-step instead of
 * four; experimental builds with -DSDDC_EXPERIMENTAL_GATHER only, 0 in the shipped library: DESIGN.md section 4) */
int sddc_plan_info(const sddc_plan* plan, int what);

/* Replace one pre-inverted operator stack (which: 0 = A4 / psi, 1 = NAB2 / T, 2 = NAB2 / S) and the effective
 * time step the matching back-substitution uses (the `dt` argument of A4_BSub_TSTEP_V2 / NAB2_BSub_TSTEP_V2).
 * Blocking; used by the drop-in Matrix_Operators module when the caller supplies its own L_inv lists. */
int sddc_plan_set_linv(sddc_plan* plan, int which, const double* Linv /* [K*n*n] host */, double dt_eff);

/* Replace the auxiliary arrays of the A4 back-substitution: D2 [n*n], diag(IR2) [n], diag(IR4) [n] -- the
 * positional arguments args_A4 of A4_BSub_TSTEP_V2 (Main.py:222, Matrix_Operators.py:1116). Blocking. */
int sddc_plan_set_a4_aux(sddc_plan* plan, const double* D2, const double* ir2, const double* ir4);

/* F(X) for B members: Matrix_Operators.NLIN_FX (743-804). X, F: [B][3Kn] */
int sddc_nlin_fx(sddc_plan* plan, const double* X, double* F, int B, void* stream);
/* DF(X) dv: Matrix_Operators.NLIN_DFX (807-898) */
int sddc_nlin_dfx(sddc_plan* plan, const double* dv, const double* X, double* F, int B, void* stream);
/* one of the theta-coupling / diagonal operators on one field: in, out: [B][Kn] */
int sddc_linear_op(sddc_plan* plan, int op, const double* in, double* out, int B, void* stream);
/* implicit solves: Matrix_Operators.A4_BSub_TSTEP_V2 (1115-1194) and NAB2_BSub_TSTEP_V2 (1033-1086);
 * g, f: [B][Kn]; which = 0 uses the dt stack (T), 1 the Tau*dt stack (S) */
int sddc_solve_a4(sddc_plan* plan, const double* g, double* f, int B, void* stream);
int sddc_solve_nab2(sddc_plan* plan, int which, const double* g, double* f, int B, void* stream);

/* nsteps IMEX-Euler member-steps: the loop body of Main._Time_Step (Step_Python, Main.py:255-283, and
 * X = X_SYM*X_new, Main.py:323). Ra, Ra_s: [B] per-member Rayleigh numbers (device). Xin is not modified;
 * Xout may not alias Xin. linear != 0 drops the nonlinear term (Main.py:261-264). */
int sddc_step(sddc_plan* plan, const double* Xin, double* Xout, const double* Ra, const double* Ra_s, int B,
              int nsteps, int linear, void* stream);
/* Step(X) - X: PFX (Main.py:473-496, 779-802) */
int sddc_residual(sddc_plan* plan, const double* X, double* out, const double* Ra, const double* Ra_s, int B,
                  void* stream);
/* PDFX(dv, X) (Main.py:498-521, 804-827) */
int sddc_jvp(sddc_plan* plan, const double* dv, const double* X, double* out, const double* Ra,
             const double* Ra_s, int B, void* stream);
/* The same JVP split for Krylov solves, where X is fixed over many products (Main.py:528-534, 905-920):
 * sddc_jvp_set_base synthesises and caches the base state's grid fields once, sddc_jvp_apply then costs one
 * synthesis (of dv) instead of two.  B must match between the two calls. */
int sddc_jvp_set_base(sddc_plan* plan, const double* X, int B, void* stream);
int sddc_jvp_apply(sddc_plan* plan, const double* dv, double* out, const double* Ra, const double* Ra_s, int B,
                   void* stream);
/* PDFX(dv, X) + dv: the linearised member-step applied to dv, i.e. PDFX without the "- delta" that ends each of its
 * three solves (Main.py:511, 515, 519).  For Krylov solvers that work with the shifted operator A + I and correct the
 * Hessenberg diagonal themselves (krylov.py): the back-substitution then needs no subtrahend (its scattered state-layout
 * loads cost a quarter of the solve).  SDDC_ERR_UNSUPPORTED on the dense paths that fall back to sddc_jvp. */
int sddc_jvp_apply_plus(sddc_plan* plan, const double* dv, double* out, const double* Ra, const double* Ra_s, int B,
                        void* stream);
/* PDFmu(X) (Main.py:829-837) */
int sddc_dF_dRa(sddc_plan* plan, const double* X, double* out, int B, void* stream);
/* per member [ ||X||_2, KE, Nu_T, Nu_S, Nu_T(outer wall), Nu_S(outer wall) ]: Main.py:292-295, Kinetic_Energy
 * (71-134), Nusselt (41-68).  out: [B][6] */
int sddc_diagnostics(sddc_plan* plan, const double* X, double* out, int B, void* stream);

/* Transforms.IDCT / IDST / DCT / DST on `rows` rows of length n_in -> n_out (device pointers).
 * Synthesis kinds zero-pad / truncate to n_out like scipy's n= argument; analysis kinds truncate. */
int sddc_transform(int kind, const double* in, double* out, int rows, int n_in, int n_out, void* stream);

/* Resolution transfer, the step every driver runs right before the hot path (Main.py:414-416, 599-601, 1093-1095;
 * Gap_Continuation.py:27-31), batched over members (device pointers).
 * sddc_interp_radial: INTERP_RADIAL (Matrix_Operators.py:901-941) as out[row][i] = sum_j W[i][j] in[row][j] over
 *   rows = B * 3 * N_fm radial profiles; W [nr_n][nr_o] is the reference's polyfit / polyval map, built on the host.
 * sddc_interp_thetas: INTERP_THETAS (Matrix_Operators.py:944-1011): [B][3][K_o][nr] -> [B][3][K_n][nr]. */
int sddc_interp_radial(const double* in, double* out, const double* W, long long rows, int nr_o, int nr_n, void* stream);
int sddc_interp_thetas(const double* in, double* out, int B, int K_o, int K_n, int nr, void* stream);

/* Batched Arnoldi orthogonalisation for the lock-step Newton / pseudo-arc-length drivers: the Krylov algebra SciPy's
 * LGMRES performs inside Main._Newton (Main.py:530-534) and Main._ContinC (Main.py:917-920, 948), for B independent
 * solves at once.  V: [B][.][n] bases (member_stride doubles apart), w: [B][n] new directions, device pointers.
 * Classical Gram-Schmidt, applied twice, in three passes over the basis:
 *     sddc_gs_dots  (V, w)                         -> part1
 *     sddc_gs_update(V, w, part1, h1, part2, 1)    w -= V h1, part2 = V^T w
 *     sddc_gs_update(V, w, part2, h2, part3, 0)    w -= V h2, part3[.][chunk][nvec] = |w_chunk|^2
 * part buffers: [B][sddc_gs_chunks(n)][ldp] with ldp >= nvec + 1 (slot nvec of sddc_gs_update's output holds the
 * squared norm of the chunk of the updated w); h buffers: [B][ldp].  Reductions run in a fixed order.
 * member_mask (optional, [B] ints): members with 0 are skipped by sddc_gs_update -- the third pass only runs for the
 * members whose second pass found something left. */
int sddc_gs_chunks(int n);
int sddc_gs_dots(const double* V, long long member_stride, int n, int nvec, const double* w, double* part, int ldp, int B,
                 void* stream);
int sddc_gs_update(const double* V, long long member_stride, int n, int nvec, double* w, const double* part_in,
                   double* h_out, double* part_out, int ldp, int want_dots, const int* member_mask, int B, void* stream);

/* Hessenberg column of one Arnoldi step for all members of a batched GMRES (krylov.py): previous Givens rotations applied
 * to (h[0..j], hn), new rotation, rotated right-hand side g, residual estimate, retirement of the members that reached
 * tol (live[b] -> 0; any_live = OR of the new flags).  shifted != 0: the operator applied was A + I, so 1 is subtracted
 * from h[j].  H [B][m+1][m], cs / sn [B][m], g [B][m+1], device pointers; one small kernel instead of two dozen tensor
 * operations per Arnoldi step. */
int sddc_gmres_column(const double* h, int ldh, const double* hn, double* H, double* cs, double* sn, double* g, double* resid,
                      const double* tol, int* live, int* any_live, int B, int j, int m, int shifted, void* stream);

/* Per-stage device timing with CUDA events recorded on the caller's stream around each kernel launch.
 * sddc_profile_begin switches recording on; sddc_profile_end synchronises the device, switches it off and
 * returns the summed milliseconds and launch counts per stage (arrays of SDDC_STAGE_COUNT). */
enum sddc_stage {
    SDDC_STAGE_SCAN = 0, SDDC_STAGE_PREP = 1, SDDC_STAGE_SYNTH = 2, SDDC_STAGE_ANALYSIS = 3, SDDC_STAGE_SOLVE = 4,
    SDDC_STAGE_KE_PREP = 5, SDDC_STAGE_KE_SYNTH = 6, SDDC_STAGE_DIAG = 7, SDDC_STAGE_COUNT = 8
};
int sddc_profile_begin(sddc_plan* plan);
int sddc_profile_end(sddc_plan* plan, double* ms, int* counts);

/* The loop of Main._Time_Step (Main.py:286-329) device resident: nsteps member-steps from Xin to Xout (caller-owned
 * device buffers, Xout must not alias Xin) and the diagnostics {|X|, KE, Nu_T, Nu_S, Nu_T(outer), Nu_S(outer)} of every
 * diag_every-th step into the device buffer diag_hist[nsteps/diag_every][B][6] (diag_every = 0: none). Stream ordered,
 * allocation free. With the FFT formulation the kinetic energy of step s comes from the spectral rows the prep stage
 * of step s+1 produces anyway, and steps 2.. skip the scan launch. */
int sddc_time_step(sddc_plan* plan, const double* Xin, double* Xout, const double* Ra, const double* Ra_s, int B,
                   int nsteps, int linear, int diag_every, double* diag_hist, void* stream);

/* Host-buffer variants (blocking; copies included): what a ctypes / NumPy caller binds. */
int sddc_step_host(sddc_plan* plan, const double* Xin, double* Xout, const double* Ra, const double* Ra_s, int B,
                   int nsteps, int linear, double* diag_out /* [B][6] or NULL */);
/* Main._Time_Step for an ensemble from host buffers (Main.py:286-329): state H2D once, nsteps member-steps, the
 * diagnostics of every diag_every-th step copied back to diag_hist[nsteps/diag_every][B][6] and the state of every
 * ckpt_every-th step to ckpt[nsteps/ckpt_every][B][3Kn] while later steps run (0 disables either), final state in
 * Xout. Blocking. */
int sddc_time_step_host(sddc_plan* plan, const double* Xin, double* Xout, const double* Ra, const double* Ra_s, int B,
                        int nsteps, int linear, int diag_every, double* diag_hist, int ckpt_every, double* ckpt);
/* Step of the first checkpoint of sddc_time_step_host: checkpoints are taken after the steps first, first + ckpt_every, ...
 * (ckpt needs (nsteps - first) / ckpt_every + 1 records).  0 (default) = ckpt_every: after every ckpt_every-th step.
 * The reference's X_DATA is saved after the steps 1, 1 + N_save, 1 + 2 N_save, ... (`iteration % N_save == 0` with the
 * iteration counter starting at 0, Main.py:286-321): first_step = 1 reproduces exactly those checkpoints. */
int sddc_plan_set_ckpt_phase(sddc_plan* plan, int first_step);

int sddc_jvp_host(sddc_plan* plan, const double* dv, const double* X, double* out, const double* Ra,
                  const double* Ra_s, int B);

#ifdef __cplusplus
}
#endif
#endif /* SDDC_B200_H */

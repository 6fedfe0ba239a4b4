"""Call sequences of the reference's drivers, written against whatever modules are registered as
`Matrix_Operators` / `Transforms` in sys.modules (the compat drop-ins in these tests).  They mirror, call for
call, what Main.Build_Matrix_Operators (Main.py:179-225), Step_Python (255-283), PFX / PDFX (473-521) and PDFmu
(829-837) ask of the operator layer -- same function names, positional arguments and return conventions -- so a
pass here means Main.py itself can run on the drop-ins unchanged."""
import numpy as np


def build_matrix_operators(MO, N_fm, N_r, d, dt, Pr, Tau):
    D, R = MO.cheb_radial(N_r, d)
    Rsq = MO.R2(R, N_fm)
    gr_K = MO.kGR_RT(R, N_fm, d)
    R_1, R_2 = 1.0 / d, (1.0 + d) / d
    A_T = (R_1 * R_2) / (R_1 - R_2)
    DT0 = A_T / (R[1:-1] ** 2)
    nr = len(R[1:-1])
    L4 = MO.A4_TSTEP_MATS(Pr * dt, N_fm, nr, D, R)
    LT = MO.NAB2_TSTEP_MATS(dt, N_fm, nr, D, R)
    LS = MO.NAB2_TSTEP_MATS(Tau * dt, N_fm, nr, D, R)
    IR = np.diag(1.0 / R)
    IR2 = IR @ IR
    D2 = np.ascontiguousarray(np.matmul(IR2, 2 * (D @ D) - 4 * (IR @ D) + 6 * IR2)[1:-1, 1:-1])
    IR2 = np.ascontiguousarray(IR2[1:-1, 1:-1])
    IR4 = np.ascontiguousarray(IR2 @ IR2)
    return np.ascontiguousarray(D), R, Rsq, DT0, gr_K, L4, LT, LS, (D2, IR4, IR2, N_fm, nr), (D, R, N_fm, nr)


def make_closures(MO, ops, Ra, Ra_s, dt, Pr, Tau, symmetric):
    D, R, Rsq, DT0, gr_k, L4, LT, LS, args_A4, args_FX = ops
    N_fm, nr = args_FX[2], args_FX[3]
    N = N_fm * nr

    def advance(NX, Y, sub):
        y0, y1, y2 = Y[0:N], Y[N:2 * N], Y[2 * N:3 * N]
        y_T0 = MO.DT0_theta(y0, DT0, N_fm, nr, symmetric)
        Om = MO.A2_SINE(y0, D, R, N_fm, nr, symmetric)
        NX[0:N] += Om + dt * Pr * gr_k.dot(Ra * y1 - Ra_s * y2)
        a = MO.A4_BSub_TSTEP_V2(NX[0:N], L4, *args_A4, Pr * dt, symmetric)
        NX[N:2 * N] += Rsq.dot(y1) - dt * y_T0
        b = MO.NAB2_BSub_TSTEP_V2(NX[N:2 * N], LT, N_fm, nr, dt, symmetric)
        NX[2 * N:3 * N] += Rsq.dot(y2) - dt * y_T0
        c = MO.NAB2_BSub_TSTEP_V2(NX[2 * N:3 * N], LS, N_fm, nr, Tau * dt, symmetric)
        out = np.hstack((a, b, c))
        return out - Y if sub else out

    def step(Xn):
        return advance(-1 * dt * MO.NLIN_FX(Xn, *args_FX, symmetric), Xn, False)

    def residual(Xn):
        return advance(-1.0 * dt * MO.NLIN_FX(Xn, *args_FX, symmetric), Xn, True)

    def jvp(dv, Xn):
        return advance(-1.0 * dt * MO.NLIN_DFX(dv, Xn, *args_FX, symmetric), dv, True)

    def dmu(Xn):
        out = 0.0 * Xn
        out[0:N] = MO.A4_BSub_TSTEP_V2(dt * Pr * gr_k.dot(Xn[N:2 * N]), L4, *args_A4, Pr * dt, symmetric)
        return out

    return step, residual, jvp, dmu


def kinetic_energy(MO, TR, X_hat, R, D, N_fm, nr, symmetric):
    """Main.Kinetic_Energy (Main.py:71-134) on the drop-in J_theta_RT / IDCT / IDST / grid."""
    N = N_fm * nr
    Dr = D[1:-1, 1:-1]
    IR = np.diag(1.0 / R[1:-1])
    psi = X_hat[0:N]
    JPSI = MO.J_theta_RT(psi, nr, N_fm, symmetric)
    Jh = np.zeros((nr, N_fm))
    Dh = np.zeros((nr, N_fm))
    for k in range(N_fm):
        Dh[:, k] = Dr @ psi[k * nr:(1 + k) * nr]
        Jh[:, k] = IR @ JPSI[k * nr:(1 + k) * nr]
    Dh[:, 1:] = Dh[:, 0:-1]
    Dh[:, 0] = 0.0
    th = TR.grid(3 * N_fm)
    KE_rt = TR.IDCT(Jh, n=3 * N_fm) ** 2 + TR.IDST(Dh, n=3 * N_fm) ** 2
    trap = np.trapezoid if hasattr(np, "trapezoid") else np.trapz
    KE_t = trap(KE_rt, x=R[1:-1], axis=0)
    KE = trap(KE_t * np.sin(th), x=th, axis=-1)
    V = (2.0 / 3.0) * (R[-1] ** 3 - R[0] ** 3)
    return (0.5 / V) * KE


def newton(MO, X, Ra, Ra_s, Tau, Pr, d, N_fm, N_r, symmetric, dt=1.0, tol_newton=1e-8, tol_gmres=1e-4,
           Krylov_Space_Size=300, max_it=5):
    """The iteration of Main._Newton (Main.py:430-554): X <- X - lgmres(PDFX(.,X), PFX(X)); returns the final X and
    the printed history error_k = |dv| / |X|.  SciPy's LGMRES stays on the host exactly as in the reference."""
    import scipy.sparse.linalg as spla
    ops = build_matrix_operators(MO, N_fm, N_r, d, dt, Pr, Tau)
    _, residual, jvp, _ = make_closures(MO, ops, Ra, Ra_s, dt, Pr, Tau, symmetric)
    nr = N_r - 1
    mask = np.ones(3 * N_fm * nr)
    if symmetric:
        m3 = mask.reshape(3, N_fm, nr)
        m3[0, 0::2] = 0.0
        m3[1:, 1::2] = 0.0
    history, n_matvec = [], [0]
    error, it = 1.0, 0
    X = X.copy()
    while error > tol_newton and it < max_it:
        X = mask * X
        fx = residual(X)

        def DF(v, X=X):
            n_matvec[0] += 1
            return jvp(v, X)

        A = spla.LinearOperator((X.shape[0], X.shape[0]), matvec=DF, dtype="float64")
        b_norm = np.linalg.norm(fx, 2)
        dv, info = spla.lgmres(A, fx, maxiter=250, inner_m=Krylov_Space_Size, atol=tol_gmres * b_norm)
        X = X - dv
        error = np.linalg.norm(dv, 2) / np.linalg.norm(X, 2)
        history.append(error)
        it += 1
    return X, np.array(history), n_matvec[0]


def continc(MO, Y_0, sign, ds, Ra_s, Tau, Pr, d, N_fm, N_r, symmetric, dt=1.0, tol_newton=1e-8, tol_gmres=1e-4,
            Krylov_Space_Size=300):
    """One pseudo-arc-length step as Main._ContinC makes it (Main.py:742-955): Predict (839-876), the bordered Newton
    corrector with the ds halving rule (885-930), ds doubling (933-935).  SciPy's LGMRES on the host as in the
    reference.  Returns (Y, ds, history [(err_X, err_mu)], n_matvec)."""
    import scipy.sparse.linalg as spla
    ops = build_matrix_operators(MO, N_fm, N_r, d, dt, Pr, Tau)
    nr = N_r - 1
    N = N_fm * nr
    delta = 1.0 / (3.0 * N)
    mask = np.ones(3 * N)
    if symmetric:
        m3 = mask.reshape(3, N_fm, nr)
        m3[0, 0::2] = 0.0
        m3[1:, 1::2] = 0.0
    n_matvec = [0]

    def closures(mu):
        return make_closures(MO, ops, mu, Ra_s, dt, Pr, Tau, symmetric)

    X_0 = mask * Y_0[0:-1].copy()
    mu_0 = float(Y_0[-1])
    _, _, jvp0, dmu0 = closures(mu_0)
    dfmu = -1.0 * dmu0(X_0)

    def DF0(v):
        n_matvec[0] += 1
        return jvp0(v, X_0)

    xi, info = spla.lgmres(spla.LinearOperator((3 * N, 3 * N), matvec=DF0, dtype="float64"), dfmu, maxiter=250,
                           inner_m=Krylov_Space_Size, atol=tol_newton * np.linalg.norm(dfmu, 2))
    assert info == 0
    mu_dot = sign / np.sqrt(1.0 + delta * (np.linalg.norm(xi, 2) - 1.0))
    X_dot = mu_dot * xi
    Y = np.hstack((X_0 + X_dot * ds, mu_0 + mu_dot * ds))
    G = 0.0 * Y
    err_X = err_mu = 1.0
    it = 0
    history = []
    while (err_X > tol_newton or err_mu > tol_newton) or it < 2:
        if it >= 5:
            it = 0
            ds *= 0.5
            assert ds >= tol_newton
            Y = np.hstack((X_0 + X_dot * ds, mu_0 + mu_dot * ds))
        X = mask * Y[0:-1].copy()
        mu = float(Y[-1])
        _, residual, jvp, dmu = closures(mu)
        DF_mu = dmu(X)
        G[0:-1] = residual(X)
        G[-1] = delta * np.dot(X_dot, X - X_0) + (1.0 - delta) * mu_dot * (mu - mu_0) - ds

        def DG(dY, X=X, jvp=jvp, DF_mu=DF_mu):
            n_matvec[0] += 1
            return np.hstack((jvp(dY[0:-1], X) + DF_mu * dY[-1],
                              delta * np.dot(X_dot, dY[0:-1]) + (1.0 - delta) * mu_dot * dY[-1]))

        b_norm = np.sqrt(delta * np.dot(G[0:-1], G[0:-1]) + (1.0 - delta) * (G[-1] ** 2))
        dY, info = spla.lgmres(spla.LinearOperator((Y.shape[0],) * 2, matvec=DG, dtype="float64"), G, maxiter=250,
                               inner_m=Krylov_Space_Size, atol=tol_gmres * b_norm)
        assert info == 0
        Y = Y - dY
        err_X = np.linalg.norm(dY[0:-1], 2) / np.linalg.norm(X, 2)
        err_mu = abs(dY[-1]) / abs(mu)
        history.append((err_X, err_mu))
        it += 1
    if it <= 4:
        ds *= 2
    return Y, ds, np.array(history), n_matvec[0]


class OracleOperators:
    """The same operator-layer surface as Matrix_Operators, backed by the NumPy oracle (CPU) -- the comparison arm
    for driver-level parity tests."""

    def __init__(self, orc, N_fm, N_r, d, dt, Pr, Tau):
        self.orc = orc
        self.op = orc.Operators(N_fm, N_r, d, dt, Pr, Tau)
        self.K, self.n = N_fm, N_r - 1

    def _f(self, v):
        return np.asarray(v, dtype=np.float64).reshape(self.K, self.n)

    def cheb_radial(self, N, d):
        return self.orc.cheb_radial(N, d)

    def R2(self, R, N_fm):
        r2 = (R[1:-1] ** 2)[None, :]
        return type("Dot", (), {"dot": lambda s, v: (r2 * self._f(v)).reshape(-1)})()

    def kGR_RT(self, R, N_fm, d):
        return type("Dot", (), {"dot": lambda s, v: self.orc.buoyancy(self._f(v), self.op).reshape(-1)})()

    def A4_TSTEP_MATS(self, dt, N_fm, nr, D, R):
        return self.orc.a4_tstep_mats(dt, N_fm, nr, D, R)

    def NAB2_TSTEP_MATS(self, dt, N_fm, nr, D, R):
        return self.orc.nab2_tstep_mats(dt, N_fm, nr, D, R)

    def DT0_theta(self, g, dT0, N_fm, nr, symmetric):
        return self.orc.DT0_theta(self._f(g), dT0, symmetric).reshape(-1)

    def A2_SINE(self, g, D, R, N_fm, nr, symmetric):
        return self.orc.A2_SINE(self._f(g), self.op, symmetric).reshape(-1)

    def NLIN_FX(self, X, D, R, N_fm, nr, symmetric):
        return self.orc.NLIN_FX(X, self.op, symmetric)

    def NLIN_DFX(self, dv, X, D, R, N_fm, nr, symmetric):
        return self.orc.NLIN_DFX(dv, X, self.op, symmetric)

    def A4_BSub_TSTEP_V2(self, g, L_inv, D2, IR4, IR2, N_fm, nr, dt, symmetric):
        return self.orc.A4_BSub(self._f(g), L_inv, self.op, dt, symmetric).reshape(-1)

    def NAB2_BSub_TSTEP_V2(self, g, L_inv, N_fm, nr, dt, symmetric):
        return self.orc.NAB2_BSub(self._f(g), L_inv, dt, symmetric).reshape(-1)

"""Host-side construction of the radial operators (NumPy, once per plan).

These are the one-off builds of the reference (cheb_radial, Nabla2, Nabla4, A4_TSTEP_MATS, NAB2_TSTEP_MATS:
Matrix_Operators.py:10-76, 1014-1030, 1089-1112; Main.Build_Matrix_Operators: Main.py:179-225).  They are
explicit inverses of matrices with condition numbers up to ~1e6, so they are formed on the host with the same
LAPACK call (np.linalg.inv) as the reference and uploaded verbatim; the GPU only ever applies them.
"""
from __future__ import annotations

import numpy as np


def cheb_radial(N: int, d: float):
    """Differentiation matrix D[(N+1),(N+1)] and collocation radii R[N+1] on [1/d, (1+d)/d]
    (Matrix_Operators.py:10-28); R[0] is the inner wall."""
    r_in, r_out = 1.0 / d, (1.0 + d) / d
    if N == 0:
        return 0.0, np.array([1.0])
    idx = np.arange(0, N + 1)
    x = np.cos(np.pi * idx / N).reshape(N + 1, 1)
    x = 0.5 * (r_in + r_out) + 0.5 * (r_in - r_out) * x
    return diff_matrix(x), x.reshape(N + 1)


def diff_matrix(R):
    """The differentiation matrix of cheb_radial for given collocation radii (Matrix_Operators.py:22-26: weights
    c_i / c_j over the node differences, diagonal by the negative-sum trick)."""
    N = len(R) - 1
    x = np.asarray(R, dtype=np.float64).reshape(N + 1, 1)
    idx = np.arange(0, N + 1)
    c = (np.hstack(([2.0], np.ones(N - 1), [2.0])) * (-1) ** idx).reshape(N + 1, 1)
    X = np.tile(x, (1, N + 1))
    dX = X - X.T
    D = np.dot(c, 1.0 / c.T) / (dX + np.eye(N + 1))
    D -= np.diag(np.sum(D.T, axis=0))
    return D


def nabla2(D, R):
    """Interior block of r^2 d_rr + 2 r d_r (Matrix_Operators.py:31-45)."""
    A = np.diag(R[:] ** 2) @ (D @ D) + np.diag(2.0 * R[:]) @ D
    return A[1:-1, 1:-1]


def nabla4(D, R):
    """Interior block of the clamped-boundary fourth derivative (Matrix_Operators.py:48-76)."""
    ones = np.ones(len(R))
    r_i, r_o = R[0], R[-1]
    b = -(r_i + r_o)
    c = r_i * r_o
    with np.errstate(divide="ignore"):
        S = np.diag(1.0 / ((R ** 2) + b * R + c * ones))
    S[0, 0] = 0.0
    S[-1, -1] = 0.0
    D2 = D @ D
    D3 = D @ D2
    D4 = D2 @ D2
    L4 = np.diag(R ** 2 + b * R + c * ones) @ D4 + 4.0 * np.diag(2.0 * R + b * ones) @ D3 + 12.0 * D2
    return (L4 @ S)[1:-1, 1:-1]


def nab2_tstep_mats(dt, N_fm, nr, D, R):
    """[N_fm, nr, nr] stack, entry jj = inv(r^2 - dt (r^2 Lap_r + b_j I)), j = N_fm-1-jj
    (Matrix_Operators.py:1014-1030)."""
    eye = np.eye(nr)
    R2 = np.diag(R[1:-1] ** 2)
    R2_Nab2 = nabla2(D, R)
    out = np.empty((N_fm, nr, nr))
    for jj in range(N_fm):
        j = N_fm - (jj + 1)
        bj = -j * (j + 1)
        out[jj] = np.linalg.inv(R2 - dt * (R2_Nab2 + bj * eye))
    return out


def a4_parts(D, R):
    """D2, IR2, IR4 (dense diag) and Dsq interior blocks (Matrix_Operators.py:1093-1101, Main.py:209-214)."""
    IR = np.diag(1.0 / R)
    IR2 = IR @ IR
    D_sq = D @ D
    D2 = (IR2 @ (2 * D_sq - 4 * (IR @ D) + 6 * IR2))[1:-1, 1:-1]
    A2 = D_sq[1:-1, 1:-1]
    IR2 = IR2[1:-1, 1:-1]
    IR4 = IR2 @ IR2
    return D2, IR2, IR4, A2


def a4_tstep_mats(dt, N_fm, nr, D, R):
    """[N_fm, nr, nr] stack, entry jj = inv((A2 + b_j IR2) - dt (D4 + b_j (D2 + b_j IR4))), j = N_fm-jj
    (Matrix_Operators.py:1089-1112); dt is Pr*dt."""
    D2, IR2, IR4, A2 = a4_parts(D, R)
    D4 = nabla4(D, R)
    out = np.empty((N_fm, nr, nr))
    for jj in range(N_fm):
        j = N_fm - jj
        bj = -j * (j + 1)
        L1 = D2 + bj * IR4
        out[jj] = np.linalg.inv((A2 + bj * IR2) - dt * (D4 + bj * L1))
    return out


def base_state_coeffs(d):
    """A_T, B_T of T_0 = -A_T/r + B_T (Main.py:23-38)."""
    R_1 = 1.0 / d
    R_2 = (1.0 + d) / d
    return (R_1 * R_2) / (R_1 - R_2), R_1 / (R_1 - R_2)


class RadialOperators:
    """All host-built arrays a plan needs, for one (N_fm, N_r, d, dt, Pr, Tau)."""

    def __init__(self, N_fm, N_r, d, dt, Pr, Tau, L_inv_A4=None, L_inv_T=None, L_inv_S=None, D=None, R=None):
        self.N_fm, self.N_r, self.nr = int(N_fm), int(N_r), int(N_r) - 1
        self.d, self.dt, self.Pr, self.Tau = float(d), float(dt), float(Pr), float(Tau)
        if D is None or R is None:
            D, R = cheb_radial(self.N_r, self.d)
        self.D = np.ascontiguousarray(D, dtype=np.float64)
        self.R = np.ascontiguousarray(R, dtype=np.float64)
        nr = self.nr
        r = self.R[1:-1]
        c = np.ascontiguousarray
        self.r = c(r)
        self.Dr = c(self.D[1:-1, 1:-1])
        self.Dsq = c((self.D @ self.D)[1:-1, 1:-1])
        self.D2r = c((np.diag(1.0 / self.R ** 2) @ (self.D @ self.D))[1:-1, 1:-1])
        D2, IR2, IR4, _ = a4_parts(self.D, self.R)
        self.D2 = c(D2)
        self.IR2 = c(IR2)
        self.IR4 = c(IR4)
        self.a4_ir2 = c(np.diag(IR2))
        self.a4_ir4 = c(np.diag(IR4))
        self.r2 = c(r ** 2)
        self.ir2 = c(1.0 / (r ** 2))
        self.ir4 = c(1.0 / (r ** 4))
        self.ir = c(1.0 / r)
        A_T = base_state_coeffs(self.d)[0]
        self.A_T = A_T
        self.dT0 = c(A_T / (r ** 2))
        self.gbuoy = c(((1.0 / self.d) ** 2) / (r ** 2))
        self.nu_in = c((self.R[0] ** 2 / A_T) * self.D[0, 1:-1])
        self.nu_out = c((self.R[-1] ** 2 / A_T) * self.D[-1, 1:-1])

        def stack(L, builder, *a):
            if L is None:
                return builder(*a)
            return np.ascontiguousarray(np.stack([np.asarray(m, dtype=np.float64) for m in L]))

        self.L_inv_A4 = stack(L_inv_A4, a4_tstep_mats, self.Pr * self.dt, self.N_fm, nr, self.D, self.R)
        self.L_inv_T = stack(L_inv_T, nab2_tstep_mats, self.dt, self.N_fm, nr, self.D, self.R)
        self.L_inv_S = stack(L_inv_S, nab2_tstep_mats, self.Tau * self.dt, self.N_fm, nr, self.D, self.R)
        for name in ("L_inv_A4", "L_inv_T", "L_inv_S"):
            if getattr(self, name).shape != (self.N_fm, nr, nr):
                raise ValueError("%s has shape %s, expected %s" % (name, getattr(self, name).shape, (self.N_fm, nr, nr)))

"""Initial conditions from linear theory: the step every Newton / continuation run of the reference starts from.

The reference seeds its l = 10 / l = 11 branches with the marginal eigenvector of the conduction state
(`Linear_Problem.Eig_Vals` / `Eig_Vec` / `Critical_Eigval`, Linear_Problem.py:16-116, built from the per-degree blocks of
`Linear_Matrix_Operators.py:44-259`) extended over latitude by `Full_Eig_Vec` (Linear_Problem.py:295-317).  These are
small dense host problems (3 nr x 3 nr generalised eigenproblems), so they stay on the host, in NumPy, like the
operator build; what they produce is the [3 nr N_fm] state vector the GPU path consumes.

Two conventions of the reference are kept on purpose, because a seeded branch must be reproducible against it:
  * the linear code orders the radial grid from the OUTER wall inwards (Linear_Matrix_Operators.py:44-61), the
    time-stepper from the inner wall outwards (Matrix_Operators.py:10-28); `Full_Eig_Vec` does not flip the profile;
  * `Critical_Eigval` refines Ra with the default Ra_s / Pr / Tau / Nr of `Eig_Vals`, whatever the caller's values.
"""
from __future__ import annotations

import numpy as np


def cheb_radial_linear(N, d):
    """Chebyshev points on [1/d, (1+d)/d] ordered from the outer wall inwards, and the differentiation matrix
    (Linear_Matrix_Operators.py:44-61)."""
    r_i, r_o = 1.0 / d, (1.0 + d) / d
    n = np.arange(0, N + 1)
    x = np.cos(np.pi * n / N).reshape(N + 1, 1)
    x = 0.5 * (r_o + r_i) + 0.5 * (r_o - r_i) * x
    c = (np.hstack(([2.0], np.ones(N - 1), [2.0])) * (-1) ** n).reshape(N + 1, 1)
    X = np.tile(x, (1, N + 1))
    dX = X - X.T
    D = np.dot(c, 1.0 / c.T) / (dX + np.eye(N + 1))
    D -= np.diag(np.sum(D.T, axis=0))
    return D, x.reshape(N + 1)


def laplacian_block(D, r, l):
    """Interior block of d2/dr2 + (2/r) d/dr - l(l+1)/r^2 (Linear_Matrix_Operators.py:64-73)."""
    A = D @ D + np.diag(2.0 / r) @ D - l * (l + 1.0) * np.diag(1.0 / r ** 2)
    return A[1:-1, 1:-1]


def stokes_d2_block(D, r, l):
    """Interior block of d2/dr2 - l(l+1)/r^2 (Linear_Matrix_Operators.py:76-84)."""
    A = D @ D - l * (l + 1.0) * np.diag(1.0 / r ** 2)
    return A[1:-1, 1:-1]


def stokes_d2d2_block(D, r, l):
    """Interior block of the squared Stokes operator with clamped boundary conditions
    (Linear_Matrix_Operators.py:87-111): the fourth derivative acts on q(r) v with q = (r - r_i)(r - r_o)."""
    r_i, r_o = r[-1], r[0]
    b, c = -(r_i + r_o), r_i * r_o
    q = r ** 2 + b * r + c
    with np.errstate(divide="ignore"):
        S = np.diag(1.0 / q)            # q vanishes at both walls; those entries are zeroed (clamped conditions)
    S[0, 0] = 0.0
    S[-1, -1] = 0.0
    D2 = D @ D
    D3 = D @ D2
    D4 = D2 @ D2
    L = np.diag(q) @ D4 + (4.0 * np.diag(2.0 * r + b)) @ D3 + 12.0 * D2
    ll = l * (l + 1.0)
    A = (L @ S - 2.0 * ll * (np.diag(1.0 / r ** 2) @ D2) + 4.0 * ll * (np.diag(1.0 / r ** 3) @ D)
         + (ll ** 2 - 6.0 * ll) * np.diag(1.0 / r ** 4))
    return A[1:-1, 1:-1]


def buoyancy_block(r, d):
    """diag(R_1^2 / r^2), interior (Linear_Matrix_Operators.py:133-155)."""
    return np.diag((1.0 / d) ** 2 / r ** 2)[1:-1, 1:-1]


def base_gradient_block(r, l, d):
    """l(l+1)/r^2 * A_T/r^2, interior: advection of the conduction profile (Linear_Matrix_Operators.py:158-185)."""
    R_1, R_2 = 1.0 / d, (1.0 + d) / d
    A_T = (R_1 * R_2) / (R_1 - R_2)
    return ((l * (l + 1.0)) * np.diag(1.0 / r ** 2) @ np.diag(A_T / r ** 2))[1:-1, 1:-1]


def mass_matrix(D, r, l):
    """M_l = blockdiag(Stokes D2, I, I) (Linear_Matrix_Operators.py:192-200)."""
    n = len(r) - 2
    M = np.zeros((3 * n, 3 * n))
    M[:n, :n] = stokes_d2_block(D, r, l)
    M[n:2 * n, n:2 * n] = np.eye(n)
    M[2 * n:, 2 * n:] = np.eye(n)
    return M


def linear_operator(D, r, d, l, Ra, Ra_s, Pr, Tau):
    """L_l of the conduction state for spherical-harmonic degree l (Linear_Matrix_Operators.py:231-259)."""
    n = len(r) - 2
    L = np.zeros((3 * n, 3 * n))
    nab = laplacian_block(D, r, l)
    L[:n, :n] = Pr * stokes_d2d2_block(D, r, l)
    L[n:2 * n, n:2 * n] = nab
    L[2 * n:, 2 * n:] = Tau * nab
    if l != 0:
        g = buoyancy_block(r, d)
        tt = base_gradient_block(r, l, d)
        L[:n, n:2 * n] = Pr * Ra * g
        L[:n, 2 * n:] = -Pr * Ra_s * g
        L[n:2 * n, :n] = tt
        L[2 * n:, :n] = tt
    return L


def _eig_problem(Ra, l, d, Ra_s, Pr, Tau, Nr):
    D, r = cheb_radial_linear(Nr, d)
    M = mass_matrix(D, r, l)
    A = linear_operator(D, r, d, l, Ra, Ra_s, Pr, Tau)
    return np.matmul(np.linalg.inv(M), A)


def eig_vals(Ra, l, d, Nvals, Ra_s=150, Pr=1, Tau=1.0 / 15.0, Nr=20):
    """Growth rates sorted by descending real part (Linear_Problem.Eig_Vals, Linear_Problem.py:16-54): Nvals in
    {0, 1, 2} returns the real part of that eigenvalue (0: Hopf pair, 1: first steady mode), larger Nvals the first
    Nvals eigenvalues."""
    ev = np.linalg.eigvals(_eig_problem(Ra, l, d, Ra_s, Pr, Tau, Nr))
    ev = ev[ev.real.argsort()[::-1]]
    if Nvals in (0, 1, 2):
        return ev[Nvals].real
    return ev[0:Nvals]


def eig_vec(Ra, l, d, k, Ra_s=150, Pr=1, Tau=1.0 / 15.0, Nr=20):
    """Real part of the k-th eigenvector [psi | T | S] (3 (Nr-1) radial values) (Linear_Problem.Eig_Vec, 56-85)."""
    w, v = np.linalg.eig(_eig_problem(Ra, l, d, Ra_s, Pr, Tau, Nr))
    idx = w.real.argsort()[::-1]
    return v[:, idx][:, k].real


def critical_rayleigh(Ra_guess, l, d, Nvals=1):
    """Secant/Newton refinement of Ra so that the Nvals-th growth rate vanishes (Linear_Problem.Critical_Eigval,
    87-116; like the reference it uses the defaults of eig_vals for Ra_s, Pr, Tau, Nr)."""
    import scipy.optimize as scp
    return scp.newton(eig_vals, x0=Ra_guess, args=(l, d, Nvals), tol=1e-05, maxiter=30).real


def _theta_grid(N):
    return np.pi * (2.0 * np.arange(N) + 1.0) / (2.0 * N)       # Transforms.grid (Transforms.py:8-13)


def _dct(f):
    """Transforms.DCT on the midpoint grid, sinusoid convention (Transforms.py:28-39,102-114): closed form."""
    M = f.shape[-1]
    th = _theta_grid(M)
    k = np.arange(M)
    out = (2.0 / M) * (f @ np.cos(np.outer(th, k)))
    out[..., 0] *= 0.5
    return out


def _dst(g):
    """Transforms.DST, sinusoid convention: coefficient k of sin(k theta), entry 0 is zero (Transforms.py:56-85)."""
    M = g.shape[-1]
    th = _theta_grid(M)
    k = np.arange(M)
    out = (2.0 / M) * (g @ np.sin(np.outer(th, k)))
    out[..., 0] = 0.0
    return out


def full_eig_vec(f, l, N_fm, nr, symmetric=False):
    """Extend the radial eigenfunctions f = [psi | T | S] over latitude: psi ~ -sin(theta) C^{3/2}_{l-1}(cos theta),
    T, S ~ P_l(cos theta), analysed on the N_fm-point grid and packed into the flat state layout
    (Linear_Problem.Full_Eig_Vec, Linear_Problem.py:295-317)."""
    from scipy.special import eval_gegenbauer, eval_legendre
    f = np.asarray(f, dtype=np.float64)
    if f.shape != (3 * nr,):
        raise ValueError("eigenvector has %s entries, expected %d" % (f.shape, 3 * nr))
    th = _theta_grid(N_fm)
    Gl_hat = _dst(-np.sin(th) * eval_gegenbauer(l - 1, 1.5, np.cos(th)))
    Pl_hat = _dct(eval_legendre(l, np.cos(th)))
    Gl_hat[0:-1] = Gl_hat[1:].copy()      # sinusoid index k -> code block k-1
    Gl_hat[-1] = 0.0
    X = np.stack([np.outer(Gl_hat, f[0:nr]), np.outer(Pl_hat, f[nr:2 * nr]), np.outer(Pl_hat, f[2 * nr:3 * nr])])
    if symmetric:                          # Vecs_to_X keeps odd psi blocks and even T / S blocks (Matrix_Operators.py:540-556)
        X[0, 0::2, :] = 0.0
        X[1:, 1::2, :] = 0.0
    return X.reshape(-1)


def seed_state(l, d, Ra, Ra_s, Pr, Tau, N_fm, N_r, k=1, amplitude=None):
    """The state a branch is seeded with (Linear_Problem.main_program, 319-360): eigenvector k (1: steady mode,
    0: Hopf pair) at N_r radial points extended to N_fm latitudinal modes; equatorially symmetric for even l
    (Main.py:632-635).  `amplitude` rescales to the given 2-norm (Main.Newton multiplies by fac * ||X||, Main.py:630)."""
    f = eig_vec(Ra, l, d, k, Ra_s=Ra_s, Pr=Pr, Tau=Tau, Nr=N_r)
    X = full_eig_vec(f, l, N_fm, N_r - 1, symmetric=(int(l) % 2 == 0))
    if amplitude is not None:
        X = X * (float(amplitude) / np.linalg.norm(X))
    return X

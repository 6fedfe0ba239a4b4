"""CPU oracle for the time-stepping / JVP hot path (TEST INFRASTRUCTURE ONLY).

This module is a NumPy restatement of the arithmetic of the reference
(mannixp/SpectralDoubleDiffusiveConvection) for the path named in
BASELINE.json: nonlinear term, radial operators, implicit block
back-substitutions, the IMEX-Euler member-step, the Newton residual / JVP and
the per-step diagnostics.  It is the *checker*: only tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke() may import it.
The product path (spectraldoublediffusiveconvection_b200) never does.

Pinning: every function below is checked in tests/test_oracle_golden.py against
fixtures produced by the *unmodified* reference running in the build container
(tests/golden/make_golden.py imports /root/reference and dumps inputs/outputs),
and against the reference's own transform known-answer tests
(Transforms.py:132-374).  Parity is therefore pinned.

State layout (reference Matrix_Operators.py:529-575): X = [psi | T | S], each
field K blocks of n radial values; psi block k multiplies sin((k+1) theta),
T/S block k multiplies cos(k theta).  Here fields are handled as (K, n) arrays
("mode-major"), i.e. X.reshape(3, K, n).

The transforms are restated in closed form (dense cosine/sine sums) rather than
through scipy.fftpack, so that the oracle is an independent statement of what
Transforms.py:73-129 computes.
"""
from __future__ import annotations

import numpy as np

# Transform back-end.  "dense": closed-form cosine/sine sums (independent restatement, default).
# "fft": scipy.fft dct/dst (pocketfft, the same engine behind the reference's scipy.fftpack calls) with the
# scalings of Transforms.py:16-70 -- used when the oracle is *timed* as the CPU baseline so that the baseline
# is not handicapped by O(K*M) transforms.  tests/test_oracle_golden.py checks both against the golden vectors.
_BACKEND = "dense"


def set_transform_backend(name: str) -> None:
    global _BACKEND
    if name not in ("dense", "fft"):
        raise ValueError(name)
    _BACKEND = name


# --------------------------------------------------------------------------------------
# Latitudinal grid and transforms (reference Transforms.py)
# --------------------------------------------------------------------------------------


def grid(M: int) -> np.ndarray:
    """Midpoint grid theta_j = pi (2j+1) / (2M)  (Transforms.py:8-13)."""
    return np.pi * (2.0 * np.arange(M) + 1.0) / (2.0 * M)


def _trig_tables(K: int, M: int):
    """cos(k theta_j), sin(k theta_j) for k < K, j < M with exact integer argument reduction."""
    k = np.arange(K, dtype=np.int64)[:, None]
    j = np.arange(M, dtype=np.int64)[None, :]
    q = (k * (2 * j + 1)) % (4 * M)  # angle = pi q / (2M), period 4M
    ang = np.pi * q.astype(np.float64) / (2.0 * M)
    return np.cos(ang), np.sin(ang)


def IDCT(f_hat: np.ndarray, n: int | None = None) -> np.ndarray:
    """f(theta_j) = sum_{k<K} f_hat[...,k] cos(k theta_j) on the n-point grid
    (Transforms.py:16-26,116-129: scaled DCT-III, zero padded / truncated to n)."""
    K = f_hat.shape[-1]
    M = K if n is None else n
    if _BACKEND == "fft":
        from scipy.fft import dct
        a = np.array(f_hat, dtype=np.float64, copy=True)
        a[..., 1:] *= 0.5
        return dct(a, type=3, n=M, axis=-1)
    Ku = min(K, M)
    C, _ = _trig_tables(Ku, M)
    return f_hat[..., :Ku] @ C


def IDST(g_hat: np.ndarray, n: int | None = None) -> np.ndarray:
    """g(theta_j) = sum_{1<=k<K} g_hat[...,k] sin(k theta_j); entry 0 is ignored
    (Transforms.py:41-54,87-100: shift-left + DST-III; Nyquist dropped)."""
    K = g_hat.shape[-1]
    M = K if n is None else n
    if _BACKEND == "fft":
        from scipy.fft import dst
        a = np.zeros_like(g_hat, dtype=np.float64)
        a[..., :-1] = 0.5 * g_hat[..., 1:]
        return dst(a, type=3, n=M, axis=-1)
    # after the shift the scaled array holds modes 1..K-1 in slots 0..K-2; truncation to n keeps slots < n
    # A TRUNCATING call (n < K) keeps mode n in the last slot, where the DST-III takes its input with half the weight
    # of the others: + 0.5 g_hat[n] sin(n theta_j) = 0.5 (-1)^j g_hat[n]  (checked against Transforms.IDST)
    Ku = min(K, M + 1)
    _, S = _trig_tables(Ku, M)
    S = S.copy()
    S[0, :] = 0.0
    if M < K:
        S[M, :] *= 0.5
    return g_hat[..., :Ku] @ S


def DCT(f: np.ndarray, n: int | None = None) -> np.ndarray:
    """f_hat_k = (2/M) sum_j f_j cos(k theta_j), k=0 halved; optionally truncated to n
    (Transforms.py:28-39,102-114)."""
    M = f.shape[-1]
    if _BACKEND == "fft":
        from scipy.fft import dct
        out = dct(f, type=2, axis=-1) * (1.0 / M)
    else:
        C, _ = _trig_tables(M, M)
        out = (f @ C.T) * (2.0 / M)
    out[..., 0] *= 0.5
    return out if n is None else out[..., :n]


def DST(g: np.ndarray, n: int | None = None) -> np.ndarray:
    """g_hat_k = (2/M) sum_j g_j sin(k theta_j) for 1<=k<=M-1, g_hat_0 = 0 (the sin(M theta)
    output of the DST-II is shifted out) (Transforms.py:56-70,73-85)."""
    M = g.shape[-1]
    if _BACKEND == "fft":
        from scipy.fft import dst
        raw = dst(g, type=2, axis=-1) * (1.0 / M)
        out = np.zeros_like(raw)
        out[..., 1:] = raw[..., :-1]
    else:
        _, S = _trig_tables(M, M)
        out = (g @ S.T) * (2.0 / M)
        out[..., 0] = 0.0
    return out if n is None else out[..., :n]


# --------------------------------------------------------------------------------------
# Radial operators (reference Matrix_Operators.py:10-76, 1014-1030, 1089-1112; Main.py:179-225)
# --------------------------------------------------------------------------------------


def cheb_radial(N: int, d: float):
    """Chebyshev-Gauss-Lobatto points mapped to [1/d, (1+d)/d] and the differentiation matrix
    with the negative-sum diagonal (Matrix_Operators.py:10-28)."""
    r_i = 1.0 / d
    r_o = (1.0 + d) / d
    idx = np.arange(N + 1)
    x = np.cos(np.pi * idx / N)
    r = 0.5 * (r_i + r_o) + 0.5 * (r_i - r_o) * x
    c = np.ones(N + 1)
    c[0] = c[-1] = 2.0
    c = c * (-1.0) ** idx
    dX = r[:, None] - r[None, :]
    D = np.outer(c, 1.0 / c) / (dX + np.eye(N + 1))
    D = D - np.diag(D.sum(axis=1))
    return D, r


def nabla2_interior(D: np.ndarray, R: np.ndarray) -> np.ndarray:
    """Interior block of r^2 d_rr + 2 r d_r (Matrix_Operators.py:31-45)."""
    A = np.diag(R ** 2) @ (D @ D) + np.diag(2.0 * R) @ D
    return A[1:-1, 1:-1]


def nabla4_interior(D: np.ndarray, R: np.ndarray) -> np.ndarray:
    """Interior block of the clamped fourth-derivative operator (Matrix_Operators.py:48-76)."""
    r_i, r_o = R[0], R[-1]
    b = -(r_i + r_o)
    c = r_i * r_o
    q = R ** 2 + b * R + c
    qi = np.zeros_like(q)
    qi[1:-1] = 1.0 / q[1:-1]          # boundary rows/cols of diag(1/q) are zeroed (clamped BCs)
    S = np.diag(qi)
    D2 = D @ D
    D3 = D @ D2
    D4 = D2 @ D2
    L4 = np.diag(q) @ D4 + 4.0 * np.diag(2.0 * R + b) @ D3 + 12.0 * D2
    return (L4 @ S)[1:-1, 1:-1]


def nab2_tstep_mats(dt: float, K: int, n: int, D: np.ndarray, R: np.ndarray) -> np.ndarray:
    """Stack [K, n, n]; entry jj is inv(diag(r^2) - dt (Nabla2 - j(j+1) I)), j = K-1-jj
    (Matrix_Operators.py:1014-1030; descending-mode order kept)."""
    R2 = np.diag(R[1:-1] ** 2)
    N2 = nabla2_interior(D, R)
    eye = np.eye(n)
    out = np.empty((K, n, n))
    for jj in range(K):
        j = K - (jj + 1)
        bj = -j * (j + 1)
        out[jj] = np.linalg.inv(R2 - dt * (N2 + bj * eye))
    return out


def a4_aux(D: np.ndarray, R: np.ndarray):
    """D2, IR2, IR4, Dsq interior blocks (Matrix_Operators.py:1093-1101, Main.py:209-214)."""
    IR = np.diag(1.0 / R)
    IR2f = IR @ IR
    Dsq = D @ D
    D2 = (IR2f @ (2 * Dsq - 4 * (IR @ D) + 6 * IR2f))[1:-1, 1:-1]
    IR2 = IR2f[1:-1, 1:-1]
    IR4 = IR2 @ IR2
    return np.ascontiguousarray(D2), np.ascontiguousarray(IR2), np.ascontiguousarray(IR4), Dsq[1:-1, 1:-1]


def a4_tstep_mats(dt: float, K: int, n: int, D: np.ndarray, R: np.ndarray) -> np.ndarray:
    """Stack [K, n, n]; entry jj is inv((Dsq + b_j IR2) - dt (Nabla4 + b_j (D2 + b_j IR4))), j = K - jj
    (Matrix_Operators.py:1089-1112). dt here is Pr*dt."""
    D2, IR2, IR4, A2 = a4_aux(D, R)
    D4 = nabla4_interior(D, R)
    out = np.empty((K, n, n))
    for jj in range(K):
        j = K - jj
        bj = -j * (j + 1)
        L1 = D2 + bj * IR4
        out[jj] = np.linalg.inv((A2 + bj * IR2) - dt * (D4 + bj * L1))
    return out


def base_state_AT(d: float) -> float:
    """A_T of the conductive base state T0' = A_T / r^2 (Main.py:23-38)."""
    R1 = 1.0 / d
    R2 = (1.0 + d) / d
    return (R1 * R2) / (R1 - R2)


class Operators:
    """Everything Main.Build_Matrix_Operators (Main.py:179-225) assembles, as dense arrays."""

    def __init__(self, K: int, N_r: int, d: float, dt: float, Pr: float, Tau: float):
        self.K, self.N_r, self.n = K, N_r, N_r - 1
        self.d, self.dt, self.Pr, self.Tau = d, dt, Pr, Tau
        self.D, self.R = cheb_radial(N_r, d)
        n = self.n
        self.r = self.R[1:-1]
        self.Dr = np.ascontiguousarray(self.D[1:-1, 1:-1])
        self.Dsq = np.ascontiguousarray((self.D @ self.D)[1:-1, 1:-1])
        # (1/r^2) D^2 interior, used by A2_SINE_R2 (Matrix_Operators.py:498)
        self.D2r = np.ascontiguousarray((np.diag(1.0 / self.R ** 2) @ (self.D @ self.D))[1:-1, 1:-1])
        self.D2, self.IR2, self.IR4, _ = a4_aux(self.D, self.R)
        self.dT0 = base_state_AT(d) / self.r ** 2
        self.Linv_A4 = a4_tstep_mats(Pr * dt, K, n, self.D, self.R)
        self.Linv_T = nab2_tstep_mats(dt, K, n, self.D, self.R)
        self.Linv_S = nab2_tstep_mats(Tau * dt, K, n, self.D, self.R)


# --------------------------------------------------------------------------------------
# theta-coupling sums (Matrix_Operators.py:131-245, 436-526)
# --------------------------------------------------------------------------------------


def sym_mask(K: int, n: int) -> np.ndarray:
    """0/1 mask [3, K, n] of Main.Eq_SYM (Main.py:137-176): psi keeps odd blocks, T,S keep even blocks."""
    m = np.ones((3, K, n))
    m[0, 0::2, :] = 0.0
    m[1, 1::2, :] = 0.0
    m[2, 1::2, :] = 0.0
    return m


def _parity_suffix_sums(psi: np.ndarray) -> np.ndarray:
    """Ssum[m] = sum_{p=m+2,m+4,..<=K} psi^{(p)} for m = 0..K (psi^{(p)} = block p-1), accumulated
    from the highest mode downward like the reference's running b / f_e vectors (np.cumsum adds sequentially)."""
    K, n = psi.shape
    S = np.zeros((K + 1, n))
    ev = np.cumsum(psi[K - 1::-2], axis=0)      # blocks K-1, K-3, .., 1  -> Ssum[K-2], Ssum[K-4], .., Ssum[0]
    S[K - 2::-2] = ev
    od = np.cumsum(psi[K - 2:0:-2], axis=0)     # blocks K-2, K-4, .., 2  -> Ssum[K-3], .., Ssum[1]
    S[K - 3:0:-2] = od
    return S


def J_theta_RT(psi: np.ndarray, symmetric: bool = False) -> np.ndarray:
    """cos block j: (j+1) psi^{(j)} + 2 Ssum[j] (j>=1), Ssum[0] (j=0) (Matrix_Operators.py:436-472)."""
    K, n = psi.shape
    if symmetric:
        psi = psi * sym_mask(K, n)[0]
    S = _parity_suffix_sums(psi)
    out = np.empty_like(psi)
    out[0] = S[0]
    jj = np.arange(1, K)[:, None]
    out[1:] = (jj + 1.0) * psi[:-1] + 2.0 * S[1:K]
    if symmetric:
        out[1::2] = 0.0
    return out


def DT0_theta(psi: np.ndarray, dT0: np.ndarray, symmetric: bool = False) -> np.ndarray:
    """r^2 J(psi, T0) (Matrix_Operators.py:131-189)."""
    return dT0[None, :] * J_theta_RT(psi, symmetric)


def _a2_sine_generic(psi, Dmat, w, symmetric):
    K, n = psi.shape
    if symmetric:
        psi = psi * sym_mask(K, n)[0]
    S = _parity_suffix_sums(psi)
    m = np.arange(1, K + 1)[:, None].astype(np.float64)
    out = psi @ Dmat.T - m * w[None, :] * ((m + 1.0) * psi + 2.0 * S[1:K + 1])
    if symmetric:
        out[0::2] = 0.0
    return out


def A2_SINE(psi: np.ndarray, op: Operators, symmetric: bool = False) -> np.ndarray:
    """A^2 psi, sine mode m: Dsq psi^{(m)} - (m/r^2)[(m+1) psi^{(m)} + 2 Ssum[m]] (Matrix_Operators.py:192-245)."""
    return _a2_sine_generic(psi, op.Dsq, 1.0 / op.r ** 2, symmetric)


def A2_SINE_R2(psi: np.ndarray, op: Operators, symmetric: bool = False) -> np.ndarray:
    """(1/r^2) A^2 psi with (1/r^2)Dsq and 1/r^4 (Matrix_Operators.py:475-526)."""
    return _a2_sine_generic(psi, op.D2r, 1.0 / op.r ** 4, symmetric)


def buoyancy(f: np.ndarray, op: Operators) -> np.ndarray:
    """kGR_RT.dot: sine block jj <- -(jj+1) (1/d^2)/r^2 * cos block jj+1; last block 0
    (Matrix_Operators.py:97-128)."""
    K, n = f.shape
    g = (1.0 / op.d) ** 2 / op.r ** 2
    out = np.zeros_like(f)
    k = np.arange(1, K)[:, None].astype(np.float64)
    out[:-1] = -k * g[None, :] * f[1:]
    return out


# --------------------------------------------------------------------------------------
# Nonlinear term and its Jacobian-vector product (Matrix_Operators.py:630-898)
# --------------------------------------------------------------------------------------


def _spectral_fields(X3: np.ndarray, op: Operators, symmetric: bool):
    """The nine (n, K) sinusoid-indexed arrays of Derivatives (Matrix_Operators.py:630-740)."""
    K, n = op.K, op.n
    if symmetric:
        X3 = X3 * sym_mask(K, n)
    psi, T, S = X3
    Jpsi = J_theta_RT(psi, symmetric)          # cos blocks 0..K-1
    omega = A2_SINE_R2(psi, op, symmetric)     # sine modes 1..K (block m-1)
    Dpsi = psi @ op.Dr.T
    k = np.arange(K)[:, None].astype(np.float64)

    def shift(a):  # sine block m-1 -> sinusoid column m ; mode K dropped, column 0 zero
        out = np.zeros_like(a)
        out[1:] = a[:-1]
        return out

    Dpsi_s = shift(Dpsi)
    om_s = shift(omega)
    cos_type = dict(JT=Jpsi, kDpsi=k * Dpsi_s, kom=k * om_s, DT=T @ op.Dr.T, DS=S @ op.Dr.T)
    sin_type = dict(om=om_s, Dpsi=Dpsi_s, kT=-k * T, kS=-k * S)
    return {a: v.T for a, v in cos_type.items()}, {a: v.T for a, v in sin_type.items()}


def _grid_fields(X3, op, symmetric):
    M = (3 * op.K) // 2
    c, s = _spectral_fields(X3, op, symmetric)
    g = {a: IDCT(v, n=M) for a, v in c.items()}
    g.update({a: IDST(v, n=M) for a, v in s.items()})
    return g


def _analyse_pack(Npsi, NT, NS, op, symmetric):
    K, n = op.K, op.n
    Fpsi = DST(Npsi)[:, :K]
    FT = DCT(NT)[:, :K]
    FS = DCT(NS)[:, :K]
    Fpsi_code = np.zeros_like(Fpsi)
    Fpsi_code[:, :-1] = Fpsi[:, 1:]      # sinusoid k -> code block k-1; last block 0 (Matrix_Operators.py:802)
    out = np.stack([Fpsi_code.T, FT.T, FS.T])
    if symmetric:
        out = out * sym_mask(K, n)
    return out


def NLIN_FX(X: np.ndarray, op: Operators, symmetric: bool = False) -> np.ndarray:
    """F(X) (Matrix_Operators.py:743-804)."""
    K, n = op.K, op.n
    if K % 2:
        raise ValueError("The number of Fourier modes is not even %d" % K)
    g = _grid_fields(X.reshape(3, K, n), op, symmetric)
    Npsi = op.Dr @ (g["JT"] * g["om"]) - (g["kDpsi"] * g["om"] + g["Dpsi"] * g["kom"])
    NT = g["JT"] * g["DT"] - g["Dpsi"] * g["kT"]
    NS = g["JT"] * g["DS"] - g["Dpsi"] * g["kS"]
    return _analyse_pack(Npsi, NT, NS, op, symmetric).reshape(-1)


def NLIN_DFX(dv: np.ndarray, X: np.ndarray, op: Operators, symmetric: bool = False) -> np.ndarray:
    """DF(X) dv = F(X,dv) + F(dv,X) (Matrix_Operators.py:807-898)."""
    K, n = op.K, op.n
    if K % 2:
        raise ValueError("The number of Fourier modes is not even %d" % K)
    g = _grid_fields(X.reshape(3, K, n), op, symmetric)
    h = _grid_fields(dv.reshape(3, K, n), op, symmetric)
    Npsi = op.Dr @ (g["JT"] * h["om"]) - (g["kDpsi"] * h["om"] + g["Dpsi"] * h["kom"])
    Npsi += op.Dr @ (h["JT"] * g["om"]) - (h["kDpsi"] * g["om"] + h["Dpsi"] * g["kom"])
    NT = (h["JT"] * g["DT"] - h["Dpsi"] * g["kT"]) + (g["JT"] * h["DT"] - g["Dpsi"] * h["kT"])
    NS = (h["JT"] * g["DS"] - h["Dpsi"] * g["kS"]) + (g["JT"] * h["DS"] - g["Dpsi"] * h["kS"])
    return _analyse_pack(Npsi, NT, NS, op, symmetric).reshape(-1)


# --------------------------------------------------------------------------------------
# Implicit block back-substitutions (Matrix_Operators.py:1033-1086, 1115-1194)
# --------------------------------------------------------------------------------------


_ACCEL = False
_jit_cache = {}


def set_accel(on: bool) -> None:
    """Compile the two sequential back-substitution loops with numba (no fastmath) -- used only when the oracle
    is timed as the CPU baseline, to put the port on the same footing as the reference's numba-JIT loops
    (Matrix_Operators.py:1033,1115).  Same arithmetic, same order."""
    global _ACCEL
    _ACCEL = bool(on)


def _jit(fn):
    if fn.__name__ not in _jit_cache:
        from numba import njit
        _jit_cache[fn.__name__] = njit(cache=False, fastmath=False)(fn)
    return _jit_cache[fn.__name__]


def _nab2_core(g, Linv, dt, symmetric):
    K, n = g.shape
    f = np.zeros_like(g)
    for c in range(2):
        if c == 0 and symmetric:
            continue
        j0 = K - 1 if c == 0 else K - 2
        b = np.zeros(n)
        j = j0
        while j >= 0:
            if j < K - 2:
                b = b + (2.0 * dt * (j + 2.0)) * f[j + 2]
            if j == 0:
                rhs = g[j] - 0.5 * b
            else:
                rhs = g[j] - b
            f[j] = Linv[K - 1 - j] @ rhs
            j -= 2
    return f


def _a4_core(g, Linv, D2, ir2, ir4, dt, symmetric):
    K, n = g.shape
    f = np.zeros_like(g)
    for c in range(2):
        if c == 1 and symmetric:
            continue
        j0 = K - c
        f_e = np.zeros(n)
        bf_e = np.zeros(n)
        j = j0
        while j >= 1:
            row = j - 1
            bj = -1.0 * j * (j + 1)
            bjt = -2.0 * j
            if j == j0:
                f[row] = Linv[K - j] @ g[row]
                bf_e = bf_e + bj * f[row]
            else:
                f_e = f_e + f[row + 2]
                L1f = D2 @ f_e + bj * (ir4 * f_e)
                rhs = g[row] + dt * bjt * (L1f + ir4 * bf_e) - bjt * (ir2 * f_e)
                f[row] = Linv[K - j] @ rhs
                bf_e = bf_e + bj * f[row] + bjt * f_e
            j -= 2
    return f


def NAB2_BSub(g: np.ndarray, Linv: np.ndarray, dt: float, symmetric: bool = False) -> np.ndarray:
    """Solve for T or S. g, result: (K, n). dt is dt or Tau*dt."""
    if _ACCEL:
        return _jit(_nab2_core)(np.ascontiguousarray(g), Linv, float(dt), bool(symmetric))
    K, n = g.shape
    f = np.zeros_like(g)
    starts = [K - 2] if symmetric else [K - 1, K - 2]
    for j0 in starts:
        b = np.zeros(n)
        for j in range(j0, -1, -2):
            if j < K - 2:
                b = b + (2.0 * dt * (j + 2.0)) * f[j + 2]
            rhs = g[j] - (0.5 * b if j == 0 else b)
            f[j] = Linv[K - 1 - j] @ rhs
    return f


def A4_BSub(g: np.ndarray, Linv: np.ndarray, op: Operators, dt: float, symmetric: bool = False) -> np.ndarray:
    """Solve for psi. g, result: (K, n) with row = sine mode - 1. dt is Pr*dt."""
    if _ACCEL:
        return _jit(_a4_core)(np.ascontiguousarray(g), Linv, op.D2, np.ascontiguousarray(np.diag(op.IR2)),
                              np.ascontiguousarray(np.diag(op.IR4)), float(dt), bool(symmetric))
    K, n = g.shape
    f = np.zeros_like(g)
    ir2 = np.diag(op.IR2)
    ir4 = np.diag(op.IR4)
    starts = [K] if symmetric else [K, K - 1]
    for j0 in starts:
        f_e = np.zeros(n)
        bf_e = np.zeros(n)
        for j in range(j0, 0, -2):
            row = j - 1
            bj = -j * (j + 1)
            bjt = -2 * j
            if j == j0:
                f[row] = Linv[K - j] @ g[row]
                bf_e = bf_e + bj * f[row]
            else:
                f_e = f_e + f[row + 2]
                L1f = op.D2 @ f_e + bj * (ir4 * f_e)
                rhs = g[row] + dt * bjt * (L1f + ir4 * bf_e) - bjt * (ir2 * f_e)
                f[row] = Linv[K - j] @ rhs
                bf_e = bf_e + bj * f[row] + bjt * f_e
    return f


def NAB2_BSub_solve(g: np.ndarray, op: Operators, dt: float, symmetric: bool = False) -> np.ndarray:
    """Second oracle for the T / S solve: the reference's solve-based NAB2_BSub_TSTEP (Matrix_Operators.py:248-323),
    which factorises A_j = r^2 - dt (Nabla2 + b_j I) per mode instead of applying a pre-computed inverse."""
    K, n = g.shape
    f = np.zeros_like(g)
    R2 = np.diag(op.r ** 2)
    N2 = nabla2_interior(op.D, op.R)
    eye = np.eye(n)
    starts = [K - 2] if symmetric else [K - 1, K - 2]
    for j0 in starts:
        b = np.zeros(n)
        for j in range(j0, -1, -2):
            A = R2 - dt * (N2 + (-j * (j + 1.0)) * eye)
            if j < K - 2:
                b = b + (2.0 * dt * (j + 2.0)) * f[j + 2]
            f[j] = np.linalg.solve(A, g[j] - (0.5 * b if j == 0 else b))
    return f


def A4_BSub_solve(g: np.ndarray, op: Operators, dt: float, symmetric: bool = False) -> np.ndarray:
    """Second oracle for the psi solve: the reference's solve-based A4_BSub_TSTEP (Matrix_Operators.py:327-433).
    dt is Pr*dt."""
    K, n = g.shape
    f = np.zeros_like(g)
    D4 = nabla4_interior(op.D, op.R)
    starts = [K] if symmetric else [K, K - 1]
    for j0 in starts:
        f_e = np.zeros(n)
        bf_e = np.zeros(n)
        for j in range(j0, 0, -2):
            row = j - 1
            bj = -j * (j + 1.0)
            bjt = -2.0 * j
            L1 = op.D2 + bj * op.IR4
            L = (op.Dsq + bj * op.IR2) - dt * (D4 + bj * L1)
            if j == j0:
                f[row] = np.linalg.solve(L, g[row])
                bf_e = bf_e + bj * f[row]
            else:
                f_e = f_e + f[row + 2]
                b_test = dt * bjt * (L1 @ f_e + op.IR4 @ bf_e) - bjt * (op.IR2 @ f_e)
                f[row] = np.linalg.solve(L, g[row] + b_test)
                bf_e = bf_e + bj * f[row] + bjt * f_e
    return f


# --------------------------------------------------------------------------------------
# Member-step, residual, JVP, d/dRa (Main.py:255-283, 473-521, 779-837)
# --------------------------------------------------------------------------------------


def _advance(NX3, X3, op, Ra, Ra_s, symmetric):
    """Linear right-hand sides + three solves shared by step / residual / JVP.
    NX3 = -dt * (nonlinear part) as [3, K, n]; X3 the state (or perturbation) the linear terms act on."""
    dt, Pr, Tau = op.dt, op.Pr, op.Tau
    psi, T, S = X3
    psi_T0 = DT0_theta(psi, op.dT0, symmetric)
    Om = A2_SINE(psi, op, symmetric)
    r2 = (op.r ** 2)[None, :]
    rhs_psi = NX3[0] + Om + dt * Pr * buoyancy(Ra * T - Ra_s * S, op)
    rhs_T = NX3[1] + r2 * T - dt * psi_T0
    rhs_S = NX3[2] + r2 * S - dt * psi_T0
    psi_n = A4_BSub(rhs_psi, op.Linv_A4, op, Pr * dt, symmetric)
    T_n = NAB2_BSub(rhs_T, op.Linv_T, dt, symmetric)
    S_n = NAB2_BSub(rhs_S, op.Linv_S, Tau * dt, symmetric)
    return np.stack([psi_n, T_n, S_n])


def step(X: np.ndarray, op: Operators, Ra: float, Ra_s: float, symmetric: bool = False,
         linear: bool = False) -> np.ndarray:
    """One IMEX-Euler member-step = Step_Python (Main.py:255-283)."""
    K, n = op.K, op.n
    X3 = X.reshape(3, K, n)
    if linear:
        NX3 = np.zeros_like(X3)
    else:
        NX3 = (-1.0 * op.dt * NLIN_FX(X, op, symmetric)).reshape(3, K, n)
    return _advance(NX3, X3, op, Ra, Ra_s, symmetric).reshape(-1)


def time_step(X: np.ndarray, op: Operators, Ra: float, Ra_s: float, n_steps: int,
              symmetric: bool = False, linear: bool = False) -> np.ndarray:
    """The loop of Main._Time_Step (Main.py:286-329) without I/O: X <- X_SYM * Step(X)."""
    mask = sym_mask(op.K, op.n).reshape(-1) if symmetric else 1.0
    Xn = X
    for _ in range(n_steps):
        Xn = step(X, op, Ra, Ra_s, symmetric, linear)
        X = mask * Xn
    return Xn


def residual(X: np.ndarray, op: Operators, Ra: float, Ra_s: float, symmetric: bool = False) -> np.ndarray:
    """PFX = Step(X) - X (Main.py:473-496)."""
    return step(X, op, Ra, Ra_s, symmetric) - X


def jvp(dv: np.ndarray, X: np.ndarray, op: Operators, Ra: float, Ra_s: float,
        symmetric: bool = False) -> np.ndarray:
    """PDFX(dv, X) (Main.py:498-521)."""
    K, n = op.K, op.n
    NX3 = (-1.0 * op.dt * NLIN_DFX(dv, X, op, symmetric)).reshape(3, K, n)
    return _advance(NX3, dv.reshape(3, K, n), op, Ra, Ra_s, symmetric).reshape(-1) - dv


def dF_dRa(X: np.ndarray, op: Operators, symmetric: bool = False) -> np.ndarray:
    """PDFmu(X): A4 solve of dt Pr G(T); zero in T, S (Main.py:829-837)."""
    K, n = op.K, op.n
    T = X.reshape(3, K, n)[1]
    out = np.zeros((3, K, n))
    out[0] = A4_BSub(op.dt * op.Pr * buoyancy(T, op), op.Linv_A4, op, op.Pr * op.dt, symmetric)
    return out.reshape(-1)


# --------------------------------------------------------------------------------------
# Diagnostics (Main.py:41-134)
# --------------------------------------------------------------------------------------


def nusselt(T: np.ndarray, op: Operators, outer: bool = False) -> float:
    """Nu - 1 at the inner wall (outer=True: at the outer wall) (Main.py:41-68). T: (K, n)."""
    K = op.K
    row = op.D[-1, 1:-1] if outer else op.D[0, 1:-1]
    Rw = op.R[-1] if outer else op.R[0]
    k = np.arange(0, K, 2).astype(np.float64)
    s = np.sum((T[0::2] @ row) / (1.0 - k ** 2))
    return (Rw ** 2 / base_state_AT(op.d)) * s


def kinetic_energy(X: np.ndarray, op: Operators, symmetric: bool = False) -> float:
    """Volume-averaged kinetic energy on the 3K-point grid with interior-node trapezoid rules
    (Main.py:71-134)."""
    K, n = op.K, op.n
    psi = X.reshape(3, K, n)[0]
    Jp = J_theta_RT(psi, symmetric) / op.r[None, :]
    Dp = psi @ op.Dr.T
    Dp_s = np.zeros_like(Dp)
    Dp_s[1:] = Dp[:-1]
    M3 = 3 * K
    th = grid(M3)
    ke = IDCT(Jp.T, n=M3) ** 2 + IDST(Dp_s.T, n=M3) ** 2          # (n, M3)
    wr = np.zeros(n)
    dr = np.diff(op.r)
    wr[:-1] += 0.5 * dr
    wr[1:] += 0.5 * dr
    wt = np.zeros(M3)
    dth = np.diff(th)
    wt[:-1] += 0.5 * dth
    wt[1:] += 0.5 * dth
    ke_th = wr @ ke
    V = (2.0 / 3.0) * (op.R[-1] ** 3 - op.R[0] ** 3)
    return (0.5 / V) * float(np.sum(wt * np.sin(th) * ke_th))


def diagnostics(X: np.ndarray, op: Operators, symmetric: bool = False) -> np.ndarray:
    """[||X||_2, KE, Nu_T, Nu_S] as appended per step by Main._Time_Step (Main.py:292-295)."""
    K, n = op.K, op.n
    X3 = X.reshape(3, K, n)
    return np.array([np.linalg.norm(X), kinetic_energy(X, op, symmetric),
                     nusselt(X3[1], op), nusselt(X3[2], op)])

"""GPU-backed drop-in for the reference's Matrix_Operators.py: same function names, argument order and meaning,
NumPy float64 arrays in and out, inputs never mutated, ValueError on an odd number of modes.

This is the single-member "correctness mode" of SURVEY.md section 8(b): every call crosses host -> device ->
host.  Throughput comes from the batched EnsemblePlan API; this module exists so that the reference's own
drivers (Main._Time_Step, _Newton, _ContinC, Gap_Continuation.Gap_Vary) run unchanged on the CUDA path.

Plans are cached per (N_fm, N_r, symmetric, radial grid); operator lists returned by A4_TSTEP_MATS /
NAB2_TSTEP_MATS are opaque to Main.py and carry the host-built inverse stack plus the dt they were built for.
"""
import numpy as np

from collections import OrderedDict

from ..operators import (RadialOperators, a4_tstep_mats, cheb_radial, diff_matrix, nab2_tstep_mats)  # noqa: F401  (cheb_radial re-exported)

_PLANS = OrderedDict()   # least recently used first; evicted plans are closed (a gap sweep builds one per gap)
MAX_PLANS = 8


class OperatorStack(list):
    """Return type of A4_TSTEP_MATS / NAB2_TSTEP_MATS: a list of nr x nr inverses in the reference's descending-mode
    order (so indexing / len() work as for numba.typed.List) with the dense stack attached for uploading."""

    def __init__(self, stack, dt, kind):
        super().__init__(stack)
        self.stack = np.ascontiguousarray(stack)
        self.dt = float(dt)
        self.kind = kind


def _check_even(N_fm):
    if N_fm % 2 != 0:
        raise ValueError('The number of Fourier modes is not even %d' % N_fm)


def _plan(N_fm, nr, symmetric, D=None, R=None, d=None):
    """One cached EnsemblePlan per (N_fm, nr, symmetric, grid).  Calls that carry no grid (the back-substitutions, whose
    operator stacks are uploaded per call and identified by object) take the most recently used plan of that shape.
    d is the caller's gap where the reference passes one, else recovered from the inner radius R[0] = 1/d."""
    from ..plan import EnsemblePlan
    _check_even(N_fm)
    if R is None:
        for k_ in reversed(_PLANS):
            if k_[:3] == (N_fm, nr, bool(symmetric)):
                _PLANS.move_to_end(k_)
                return _PLANS[k_]
        D, R = cheb_radial(nr + 1, 1.0)
    R = np.asarray(R, dtype=np.float64)
    key = (N_fm, nr, bool(symmetric), float(R[0]), float(R[-1]))
    pl = _PLANS.get(key)
    if pl is None:
        if D is None:
            D = diff_matrix(R)
        d = 1.0 / float(R[0]) if d is None else float(d)
        eye = np.broadcast_to(np.eye(nr), (N_fm, nr, nr))
        ops = RadialOperators(N_fm, nr + 1, d, 1.0, 1.0, 1.0, L_inv_A4=eye, L_inv_T=eye, L_inv_S=eye, D=D, R=R)
        pl = EnsemblePlan(N_fm, nr + 1, d, 1.0, 1.0, 1.0, symmetric=bool(symmetric), max_batch=1, operators=ops)
        pl._loaded, pl._lru, pl._keep, pl._tick, pl._aux = [None] * 3, [0] * 3, {}, 0, None
        _PLANS[key] = pl
        while len(_PLANS) > MAX_PLANS:
            _, old = _PLANS.popitem(last=False)
            old.close()
    else:
        _PLANS.move_to_end(key)
    return pl


def _dev(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _host(t):
    return t.cpu().numpy().reshape(-1)


# ----------------------------------------------------------------------------------------------- operators
class _DotOperator:
    """Stand-in for the scipy.sparse matrices returned by R2 / kGR_RT: Main.py only ever calls .dot(vector)."""

    def __init__(self, op, N_fm, R, d=None):
        self.op, self.N_fm, self.R, self.d = op, N_fm, np.array(R, dtype=np.float64), d
        self.shape = (N_fm * (len(R) - 2),) * 2

    def dot(self, v):
        # the grid this operator was built on, as given (no 1 / (1 / x) round trip through d)
        pl = _plan(self.N_fm, len(self.R) - 2, False, None, self.R, self.d)
        return _host(pl.linear_op(self.op, _dev(v)))


def R2(R, N_fm):
    from ..plan import OP_R2
    return _DotOperator(OP_R2, N_fm, R)


def kGR_RT(R, N_fm, d):
    from ..plan import OP_KGR
    return _DotOperator(OP_KGR, N_fm, R, d)


def NAB2_TSTEP_MATS(dt, N_fm, nr, D, R):
    return OperatorStack(nab2_tstep_mats(dt, N_fm, nr, D, R), dt, "nab2")


def A4_TSTEP_MATS(dt, N_fm, nr, D, R):
    return OperatorStack(a4_tstep_mats(dt, N_fm, nr, D, R), dt, "a4")


def J_theta_RT(g, nr, N_fm, symmetric):
    from ..plan import OP_J_THETA
    return _host(_plan(N_fm, nr, symmetric).linear_op(OP_J_THETA, _dev(g)))


def DT0_theta(g, dT0, N_fm, nr, symmetric):
    """r^2 J(psi, T0) = dT0 * J_theta(psi) blockwise (Matrix_Operators.py:131-189)."""
    J = J_theta_RT(g, nr, N_fm, symmetric).reshape(N_fm, nr)
    return (np.asarray(dT0)[None, :] * J).reshape(-1)


def A2_SINE(g, D, R, N_fm, nr, symmetric):
    from ..plan import OP_A2_SINE
    return _host(_plan(N_fm, nr, symmetric, D, R).linear_op(OP_A2_SINE, _dev(g)))


def A2_SINE_R2(g, N_fm, nr, D, R, symmetric):
    from ..plan import OP_A2_SINE_R2
    return _host(_plan(N_fm, nr, symmetric, D, R).linear_op(OP_A2_SINE_R2, _dev(g)))


def NLIN_FX(X_hat, D, R, N_fm, nr, symmetric):
    return _host(_plan(N_fm, nr, symmetric, D, R).nlin_fx(_dev(X_hat)))


def NLIN_DFX(dv_hat, X_hat, D, R, N_fm, nr, symmetric):
    return _host(_plan(N_fm, nr, symmetric, D, R).nlin_dfx(_dev(dv_hat), _dev(X_hat)))


def _stack_of(L_inv):
    return L_inv.stack if isinstance(L_inv, OperatorStack) else np.stack([np.asarray(m) for m in L_inv])


def _ensure_stack(pl, slots, L_inv, dt):
    """Make sure `L_inv` (with its dt) is resident in one of the plan's operator `slots`; returns the slot.
    Stacks are recognised by object identity, so alternating T / S solves do not re-upload anything."""
    tag = (id(L_inv), float(dt))
    for s in slots:
        if pl._loaded[s] == tag:
            pl._lru[s] = pl._tick = pl._tick + 1
            return s
    s = min(slots, key=lambda k: pl._lru[k])
    pl.set_linv(s, _stack_of(L_inv), dt)
    pl._loaded[s] = tag
    pl._keep[s] = L_inv          # keep the caller's list alive so that id() stays unique
    pl._lru[s] = pl._tick = pl._tick + 1
    return s


def A4_BSub_TSTEP_V2(g, L_inv, D2, IR4, IR2, N_fm, nr, dt, symmetric):
    pl = _plan(N_fm, nr, symmetric)
    aux = (id(D2), id(IR4), id(IR2))
    if pl._aux != aux:
        pl.set_a4_aux(D2, np.diag(np.asarray(IR2)), np.diag(np.asarray(IR4)))
        pl._aux, pl._aux_keep = aux, (D2, IR4, IR2)
    _ensure_stack(pl, (0,), L_inv, dt)
    return _host(pl.solve_a4(_dev(g)))


def NAB2_BSub_TSTEP_V2(g, L_inv, N_fm, nr, dt, symmetric):
    pl = _plan(N_fm, nr, symmetric)
    slot = _ensure_stack(pl, (1, 2), L_inv, dt)
    return _host(pl.solve_nab2(_dev(g), slot - 1))


# ----------------------------------------------------------------------------------------------- layout helpers (host)
def Vecs_to_X(PSI, T, C, N_fm, nr, symmetric):
    """(nr, N_fm) x 3 -> flat state; symmetric keeps odd psi blocks and even T/C blocks (Matrix_Operators.py:529-575)."""
    X = np.stack([np.asarray(PSI).T, np.asarray(T).T, np.asarray(C).T]).astype(np.float64)
    if symmetric:
        X[0, 0::2, :] = 0.0
        X[1:, 1::2, :] = 0.0
    return X.reshape(-1)


def X_to_Vecs(X, N_fm, nr, symmetric):
    X3 = np.array(X, dtype=np.float64).reshape(3, N_fm, nr)
    if symmetric:
        X3[0, 0::2, :] = 0.0
        X3[1:, 1::2, :] = 0.0
    return X3[0].T.copy(), X3[1].T.copy(), X3[2].T.copy()


def INTERP_RADIAL(N_n, N_o, X_o, d):
    """Polynomial re-interpolation of every mode onto a new radial grid (Matrix_Operators.py:901-941). Host-side."""
    if N_n == N_o:
        return X_o
    _, R_n = cheb_radial(N_n, d)
    _, R_o = cheb_radial(N_o, d)
    nr_n, nr_o = N_n - 1, N_o - 1
    K = len(X_o) // (3 * nr_o)
    Xo = np.asarray(X_o, dtype=np.float64).reshape(3 * K, nr_o)
    Xn = np.empty((3 * K, nr_n))
    import warnings
    with warnings.catch_warnings():
        # one unknown more than equations: np.polyfit warns about the rank on every call; the reference silences the
        # warning for the whole module (Matrix_Operators.py:5-6)
        warnings.simplefilter("ignore")
        for row in range(3 * K):
            coeff = np.polyfit(R_o, np.hstack(([0.0], Xo[row], [0.0])), len(R_o))
            Xn[row] = np.polyval(coeff, R_n[1:-1])
    return Xn.reshape(-1)


def INTERP_THETAS(N_fm_n, N_fm_o, X_o):
    """Re-sample in latitude through grid space (Matrix_Operators.py:944-1011). Uses the GPU transforms."""
    if N_fm_n == N_fm_o:
        return X_o
    from . import Transforms as TR
    nr = len(X_o) // (3 * N_fm_o)
    Kw = max(N_fm_n, N_fm_o)
    F = np.zeros((3, nr, Kw))
    F[:, :, :N_fm_o] = np.asarray(X_o, dtype=np.float64).reshape(3, N_fm_o, nr).transpose(0, 2, 1)
    psi = TR.DST(TR.IDST(F[0]), n=N_fm_n)
    T = TR.DCT(TR.IDCT(F[1]), n=N_fm_n)
    S = TR.DCT(TR.IDCT(F[2]), n=N_fm_n)
    return np.stack([psi.T, T.T, S.T]).reshape(-1)

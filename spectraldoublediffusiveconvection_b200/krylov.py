"""Lock-step batched Krylov drivers on the device: many independent Newton / pseudo-arc-length problems (one per
ensemble member: parameter points, branch points) advance together, every Krylov vector of every member costing one
batched GPU Jacobian-vector product (EnsemblePlan.jvp_apply = PDFX, Main.py:498-521).

Reference driver                      ->  batched counterpart here
  Main._Newton      (Main.py:430-554) ->  newton_batched        X <- X - DF(X)^-1 F(X), error = |dv| / |X|, <= 5 iterations
  Predict           (Main.py:839-876) ->  predict_batched       DF_X xi = -DF_mu, mu_dot = sign / sqrt(1 + delta (|xi| - 1))
  Main._ContinC     (Main.py:742-955) ->  continc_batched       predictor, bordered Newton corrector with the ds halving
                                                                 (Main.py:888-899) / doubling (933-935) rules, new tangent
  Main._NewtonC     (Main.py:717-739) ->  newtonc_batched       natural-parameter step, ds doubled / halved (734-739)
  Main._Continuation(Main.py:958-1045)->  continuation_batched  the branch loop: Newton steps while ds > ds_min, arc-length
                                                                 steps otherwise, fold detection, sign of the next step
Every member carries its own iteration counter, step size ds, sign and status; members that have converged or failed are
masked out of the Krylov recurrences (their vectors are frozen) while the others continue.

The linear solver is a restarted GMRES (batched Arnoldi, classical Gram-Schmidt applied twice) in place of SciPy's
single-vector LGMRES, which cannot interleave the matvec requests of several solves.  Inexact-Newton tolerances are the
reference's (tol_gmres * |F| for the corrector solves, tol_newton for tangents and the outer iteration), so converged
states agree with the reference's to the Newton tolerance; iteration histories agree to the linear-solve tolerance.
The orthogonalisation - the only part of the Krylov algebra that touches O(B m n) data - runs in the library's fused
Gram-Schmidt kernels (csrc/k_krylov.cuh, three passes over the basis per Arnoldi step); the Hessenberg / Givens
bookkeeping is O(B m) and stays in torch.  CPU tensors (host-logic tests only) use a torch.bmm orthogonalisation.
"""
from __future__ import annotations

import ctypes as C

import torch


# ------------------------------------------------------------------------------------------------ optional device timing
class _Profile:
    """CUDA-event timing of the two device-heavy pieces of a solve (operator applications, orthogonalisation)."""
    enabled = False
    events = []

    @classmethod
    def timed(cls, kind, fn, *args):
        if not cls.enabled:
            return fn(*args)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn(*args)
        b.record()
        cls.events.append((kind, a, b))
        return out


def profile_begin():
    _Profile.enabled, _Profile.events = True, []


def profile_end():
    """{"matvec_ms", "ortho_ms", "matvec_calls", "ortho_calls"} since profile_begin() (synchronises the device)."""
    torch.cuda.synchronize()
    out = {"matvec_ms": 0.0, "ortho_ms": 0.0, "matvec_calls": 0, "ortho_calls": 0}
    for kind, a, b in _Profile.events:
        out[kind + "_ms"] += a.elapsed_time(b)
        out[kind + "_calls"] += 1
    _Profile.enabled, _Profile.events = False, []
    return out


# ------------------------------------------------------------------------------------------------ orthogonalisation
class _TorchOrtho:
    """CGS2 with torch.bmm: CPU tensors (tests of the host logic)."""

    def __init__(self, V):
        self.V = V

    def __call__(self, j, w):
        Vj = self.V[:, :j + 1]
        h = torch.bmm(Vj, w.unsqueeze(2)).squeeze(2)
        w = w - torch.bmm(h.unsqueeze(1), Vj).squeeze(1)
        h2 = torch.bmm(Vj, w.unsqueeze(2)).squeeze(2)
        w = w - torch.bmm(h2.unsqueeze(1), Vj).squeeze(1)
        return h + h2, torch.linalg.vector_norm(w, dim=1), w


class _FusedOrtho:
    """Classical Gram-Schmidt with re-orthogonalisation in at most three passes over the basis with the library's
    kernels (sddc_gs_dots / sddc_gs_update).  The second pass already measures what the first one left behind
    (part2 = V^T w'): when that is below reorth_tol |w'| for every member the third pass is skipped ("twice is enough"
    applied only where once was not).  reorth_tol follows the accuracy the solve asks for (batched_gmres sets it to a
    thousandth of the smallest relative residual tolerance, between 1e-13 and 1e-8)."""
    reorth_tol = 1e-11
    calls = third_passes = 0      # process-wide counters (tools/run_configs45.py reports them)

    def __init__(self, V):
        from . import _lib
        self.lib = _lib.load()
        self.V = V
        B, ldv, n = V.shape
        self.B, self.n, self.stride = B, n, ldv * n
        self.nchunk = int(self.lib.sddc_gs_chunks(n))
        self.ldp = ldv + 1
        mk = lambda *s: torch.empty(s, dtype=torch.float64, device=V.device)
        self.p1, self.p2, self.p3 = (mk(B, self.nchunk, self.ldp) for _ in range(3))
        self.h1, self.h2 = mk(B, self.ldp), mk(B, self.ldp)

    def __call__(self, j, w):
        lib, nvec = self.lib, j + 1
        w = w.contiguous()
        st = C.c_void_p(torch.cuda.current_stream(w.device).cuda_stream)
        args = (self.V.data_ptr(), self.stride, self.n, nvec)
        rc = lib.sddc_gs_dots(*args, w.data_ptr(), self.p1.data_ptr(), self.ldp, self.B, st)
        rc = rc or lib.sddc_gs_update(*args, w.data_ptr(), self.p1.data_ptr(), self.h1.data_ptr(), self.p2.data_ptr(),
                                      self.ldp, 1, None, self.B, st)
        if rc:
            raise RuntimeError("libsddc_b200 Gram-Schmidt kernels failed (%d)" % rc)
        s2 = self.p2.sum(dim=1)                                  # [B, ldp]: V^T w' and |w'|^2 in slot nvec (fixed order)
        left = torch.linalg.vector_norm(s2[:, :nvec], dim=1)
        need = left > self.reorth_tol * torch.sqrt(s2[:, nvec])   # members whose first pass left something behind
        _FusedOrtho.calls += 1
        hn2 = torch.sqrt(s2[:, nvec])
        if not bool(need.any()):
            return self.h1[:, :nvec].clone(), hn2, w
        _FusedOrtho.third_passes += 1
        mask = need.to(torch.int32)
        rc = lib.sddc_gs_update(*args, w.data_ptr(), self.p2.data_ptr(), self.h2.data_ptr(), self.p3.data_ptr(),
                                self.ldp, 0, mask.data_ptr(), self.B, st)
        if rc:
            raise RuntimeError("libsddc_b200 Gram-Schmidt kernels failed (%d)" % rc)
        h = self.h1[:, :nvec] + torch.where(need[:, None], self.h2[:, :nvec], torch.zeros_like(self.h2[:, :nvec]))
        hn = torch.where(need, torch.sqrt(self.p3[:, :, nvec].sum(dim=1)), hn2)
        return h, hn, w


def _make_ortho(V):
    return _FusedOrtho(V) if V.is_cuda else _TorchOrtho(V)


# ------------------------------------------------------------------------------------------------ GMRES
LGMRES_RTOL = 1e-5   # SciPy's default rtol: the reference passes only atol, so every solve stops at max(atol, 1e-5 |b|)


def batched_gmres(matvec, b, rtol=1e-4, atol=None, m=60, max_restarts=8, x0=None, active=None, shifted=False):
    """Solve A_k x_k = b_k for every row k of b [B, n] in lock step.  matvec maps [B, n] -> [B, n] (row-wise
    independent operators).  Member k stops when |r_k| <= max(atol_k, rtol * |b_k|) (SciPy's convention; atol = None
    means 0); from then on it is masked: its Krylov vectors are zero, its Hessenberg columns the identity, its solution untouched.  Members
    with active[k] == False are never touched (x_k = x0_k or 0).
    shifted=True: matvec applies A + I instead of A (EnsemblePlan.jvp_apply(plus_identity=True): the back-substitution
    then carries no subtrahend); orthogonalising (A + I) v_j against a basis that contains v_j leaves the same new
    direction, and the Hessenberg column of A is the one of A + I minus e_j.
    Returns (x, info), info = {"iters": batched matvec calls, "member_iters": Krylov vectors per member [B],
    "converged": bool [B], "resid": |r_k| [B]}."""
    B, n = b.shape
    dev, dt = b.device, b.dtype
    bnorm = torch.linalg.vector_norm(b, dim=1)
    tol = rtol * bnorm
    if atol is not None:
        tol = torch.maximum(tol, torch.as_tensor(atol, dtype=dt, device=dev).expand(B))
    tol = torch.clamp(tol, min=torch.finfo(dt).tiny)
    act = torch.ones(B, dtype=torch.bool, device=dev) if active is None else active.clone()
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    V = torch.zeros((B, m + 1, n), dtype=dt, device=dev)
    ortho = _make_ortho(V)
    fused = isinstance(ortho, _FusedOrtho)
    lib = ortho.lib if fused else None
    if fused:
        tol = tol.contiguous()
    if isinstance(ortho, _FusedOrtho):
        rel = torch.where(act & (bnorm > 0), tol / torch.where(bnorm > 0, bnorm, torch.ones_like(bnorm)), torch.ones_like(bnorm))
        ortho.reorth_tol = float(min(1e-8, max(1e-13, 1e-3 * float(rel.min()))))
    total = 0
    member_iters = torch.zeros(B, dtype=torch.long, device=dev)
    resid = torch.where(act, bnorm, torch.zeros_like(bnorm))
    one, zero = torch.ones_like(bnorm), torch.zeros_like(bnorm)
    for _ in range(max_restarts):
        if x0 is not None or total > 0:
            r = b - _Profile.timed("matvec", matvec, x)
            if shifted:
                r = r + x
            total += 1
        else:
            r = b.clone()
        beta = torch.linalg.vector_norm(r, dim=1)
        resid = torch.where(act, beta, resid)
        live = act & (resid > tol)                 # members still iterating
        if not bool(live.any()):
            break
        safe_beta = torch.where(beta > 0, beta, one)
        V[:, 0] = torch.where(live[:, None], r / safe_beta[:, None], torch.zeros_like(r))
        H = torch.zeros((B, m + 1, m), dtype=dt, device=dev)
        cs = torch.zeros((B, m), dtype=dt, device=dev)
        sn = torch.zeros((B, m), dtype=dt, device=dev)
        gvec = torch.zeros((B, m + 1), dtype=dt, device=dev)
        gvec[:, 0] = torch.where(live, beta, zero)
        jdone = 0
        if fused:
            # the Hessenberg / Givens bookkeeping of a step is one kernel over the members (sddc_gmres_column)
            live_i = live.to(torch.int32)
            any_live = torch.zeros(1, dtype=torch.int32, device=dev)
            resid = resid.contiguous()
        else:
            # Q = G_{j-1} .. G_0, the product of the Givens rotations so far: one small bmm applies them all to a new column
            # (a Python loop over the rotations would cost O(m^2) kernel launches per solve)
            Q = torch.eye(m + 1, dtype=dt, device=dev).repeat(B, 1, 1)
        for j in range(m):
            w = _Profile.timed("matvec", matvec, V[:, j].contiguous())
            total += 1
            member_iters += live.long()
            w = torch.where(live[:, None], w, torch.zeros_like(w))
            h, hn, w = _Profile.timed("ortho", ortho, j, w)
            V[:, j + 1] = w / torch.where(hn > 0, hn, one)[:, None]
            jdone = j + 1
            if fused:
                hn = hn.contiguous()
                rc = lib.sddc_gmres_column(h.data_ptr(), h.stride(0), hn.data_ptr(), H.data_ptr(), cs.data_ptr(), sn.data_ptr(),
                                           gvec.data_ptr(), resid.data_ptr(), tol.data_ptr(), live_i.data_ptr(),
                                           any_live.data_ptr(), B, j, m, int(shifted),
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
                if rc:
                    raise RuntimeError("sddc_gmres_column failed (%d)" % rc)
                live = live_i.bool()
                if int(any_live.item()) == 0:
                    break
                continue
            if shifted:
                h = h.clone()
                h[:, j] -= live.to(dt)
            col = torch.cat([h, hn[:, None]], dim=1)          # [B, j+2]
            col = torch.bmm(Q[:, :j + 2, :j + 2], col.unsqueeze(2)).squeeze(2)
            den = torch.sqrt(col[:, j] ** 2 + col[:, j + 1] ** 2)
            ok = live & (den > 0)
            den = torch.where(ok, den, one)
            # masked members: identity column, zero right-hand side entry -> y_j = 0
            cs[:, j] = torch.where(ok, col[:, j] / den, one)
            sn[:, j] = torch.where(ok, col[:, j + 1] / den, zero)
            col[:, j] = torch.where(ok, cs[:, j] * col[:, j] + sn[:, j] * col[:, j + 1], one)
            col[:, :j] = torch.where(live[:, None], col[:, :j], torch.zeros_like(col[:, :j]))
            col[:, j + 1] = 0.0
            H[:, :j + 2, j] = col
            qj, qj1 = Q[:, j, :].clone(), Q[:, j + 1, :].clone()
            Q[:, j, :] = cs[:, j, None] * qj + sn[:, j, None] * qj1
            Q[:, j + 1, :] = -sn[:, j, None] * qj + cs[:, j, None] * qj1
            gvec[:, j + 1] = -sn[:, j] * gvec[:, j]
            gvec[:, j] = torch.where(live, cs[:, j] * gvec[:, j], zero)
            resid = torch.where(live, gvec[:, j + 1].abs(), resid)
            gvec[:, j + 1] = torch.where(live & (resid > tol), gvec[:, j + 1], zero)   # a member that stops here
            live = live & (resid > tol)
            if not bool(live.any()):
                break
        R = H[:, :jdone, :jdone]
        y = torch.linalg.solve_triangular(R, gvec[:, :jdone].unsqueeze(2), upper=True).squeeze(2)
        x = x + torch.bmm(y.unsqueeze(1), V[:, :jdone]).squeeze(1)
        if not bool((act & (resid > tol)).any()):
            break
    return x, {"iters": total, "member_iters": member_iters, "converged": (resid <= tol) | ~act, "resid": resid}


# ------------------------------------------------------------------------------------------------ helpers
def symmetry_mask(plan, device=None):
    """Main.Eq_SYM (Main.py:137-176) as a [3N] 0/1 vector, or None for a non-symmetric plan."""
    if not plan.symmetric:
        return None
    m = torch.ones((3, plan.N_fm, plan.nr), dtype=torch.float64, device=device)
    m[0, 0::2, :] = 0.0
    m[1:, 1::2, :] = 0.0
    return m.reshape(-1)


def _masked(X, mask):
    return X if mask is None else X * mask


def _wnorm(Yx, Ymu, delta):
    """sqrt(delta |X|^2 + (1 - delta) mu^2): the norm of the extended system (Main.py:919, 953)."""
    return torch.sqrt(delta * (Yx ** 2).sum(dim=1) + (1.0 - delta) * Ymu ** 2)


# ------------------------------------------------------------------------------------------------ Newton
def newton_batched(plan, X, Ra, Ra_s, tol_newton=1e-8, tol_gmres=1e-4, krylov=100, max_it=5, max_restarts=8,
                   lgmres_rtol=LGMRES_RTOL):
    """B concurrent matrix-free Newton solves for steady states (Main._Newton, Main.py:430-554; use a plan built with
    dt = 1 like the reference's default).  X: [B, 3N].  Per member: iterate while error > tol_newton and fewer than
    max_it iterations were made; a member succeeds when it stopped before max_it iterations with a converged linear
    solve (Main.py:541).  Returns (X, info) with info = {"history": [iterations, B] (error = |dv| / |X|, NaN once a
    member has stopped), "converged": bool [B], "iterations": [B], "member_jvps": [B], "jvps": batched JVP calls}."""
    X = X.clone()
    B = X.shape[0]
    dev = X.device
    Ra, Ra_s = plan._param(Ra, B), plan._param(Ra_s, B)
    mask = symmetry_mask(plan, dev)
    active = torch.ones(B, dtype=torch.bool, device=dev)
    lin_ok = torch.ones(B, dtype=torch.bool, device=dev)
    its = torch.zeros(B, dtype=torch.long, device=dev)
    mj = torch.zeros(B, dtype=torch.long, device=dev)
    hist, njvp = [], 0
    nan = torch.full((B,), float("nan"), dtype=X.dtype, device=dev)
    last = torch.ones(B, dtype=X.dtype, device=dev)
    for _ in range(max_it):
        X = torch.where(active[:, None], _masked(X, mask), X)           # X = X_SYM * X (Main.py:525)
        fx = plan.residual(X, Ra, Ra_s)                                  # PFX (Main.py:473-496)
        plan.jvp_set_base(X)                                             # X is fixed during the linear solve
        b_norm = torch.linalg.vector_norm(fx, dim=1)
        plus = bool(getattr(plan, "has_jvp_plus", False))
        mv = (lambda v: plan.jvp_apply(v, Ra, Ra_s, plus_identity=True)) if plus else (lambda v: plan.jvp_apply(v, Ra, Ra_s))
        dv, info = batched_gmres(mv, fx, rtol=lgmres_rtol, atol=tol_gmres * b_norm, m=krylov, max_restarts=max_restarts,
                                 active=active, shifted=plus)                                  # Main.py:533-534
        njvp += info["iters"]
        mj += info["member_iters"]
        dv = torch.where(active[:, None], dv, torch.zeros_like(dv))
        X = X - dv
        err = torch.linalg.vector_norm(dv, dim=1) / torch.linalg.vector_norm(X, dim=1)
        hist.append(torch.where(active, err, nan))
        its += active.long()
        lin_ok = torch.where(active, info["converged"], lin_ok)
        last = torch.where(active, err, last)
        active = active & ~(err <= tol_newton)                           # a NaN error keeps the member "not converged"
        if not bool((active & torch.isfinite(last)).any()):
            break
    converged = (last <= tol_newton) & lin_ok & (its < max_it)
    return X, {"history": torch.stack(hist), "converged": converged, "iterations": its, "member_jvps": mj, "jvps": njvp}


# ------------------------------------------------------------------------------------------------ arc-length continuation
def predict_batched(plan, X0, mu0, sign, ds, Ra_s, tol_newton=1e-8, krylov=100, max_restarts=8, lgmres_rtol=LGMRES_RTOL):
    """Predict of Main._ContinC (Main.py:839-876) for B branch points: tangent from DF_X xi = -DF_mu at (X0, mu0),
    mu_dot = sign / sqrt(1 + delta (|xi| - 1)), X_dot = mu_dot xi, and the prediction (X0 + ds X_dot, mu0 + ds mu_dot).
    (With the reference's dt = 1 the stored tangent is never reused: the `dt < 10` branch always recomputes it.)
    Returns (X, mu, X0_masked, X_dot, mu_dot, ok [B], jvps)."""
    B, n = X0.shape
    delta = 1.0 / n
    mask = symmetry_mask(plan, X0.device)
    X0 = _masked(X0, mask)
    dfmu = -plan.dF_dRa(X0)                                              # (-1) PDFmu (Main.py:848)
    plan.jvp_set_base(X0)
    plus = bool(getattr(plan, "has_jvp_plus", False))
    mv = (lambda v: plan.jvp_apply(v, mu0, Ra_s, plus_identity=True)) if plus else (lambda v: plan.jvp_apply(v, mu0, Ra_s))
    xi, info = batched_gmres(mv, dfmu, rtol=lgmres_rtol, atol=tol_newton * torch.linalg.vector_norm(dfmu, dim=1), m=krylov,
                             max_restarts=max_restarts, shifted=plus)
    mu_dot = sign / torch.sqrt(1.0 + delta * (torch.linalg.vector_norm(xi, dim=1) - 1.0))
    X_dot = mu_dot[:, None] * xi
    return X0 + X_dot * ds[:, None], mu0 + mu_dot * ds, X0, X_dot, mu_dot, info["converged"], info["iters"]


def continc_batched(plan, X0, mu0, sign, ds, Ra_s, tol_newton=1e-8, tol_gmres=1e-4, krylov=100, max_restarts=8,
                    max_rounds=60, lgmres_rtol=LGMRES_RTOL):
    """One pseudo-arc-length step for B branch points in lock step (Main._ContinC, Main.py:742-955): predictor, then the
    bordered Newton corrector on
        G(X, mu) = [ P F(X, mu) ;  delta X_dot.(X - X0) + (1 - delta) mu_dot (mu - mu0) - ds ],   delta = 1/(3N),
    with the reference's step-size rules per member - after 5 corrector iterations without convergence ds is halved and
    the prediction restarted (Main.py:888-899; a member whose ds falls below tol_newton fails), a member that needed at
    most 4 iterations leaves with ds doubled (933-935) - and finally the new unit tangent from DG Y_dot = e_last
    (944-953; zero if that solve does not converge).  mu is the thermal Rayleigh number.
    X0: [B, 3N]; mu0, sign, ds: [B].  Returns a dict: X, mu, X_dot, mu_dot (new tangent), ds (updated), ok [B] (corrector
    converged), tangent_ok [B], history [rounds, B, 2] (err_X, err_mu; NaN for members not iterating), iterations [B],
    halvings [B], jvps (batched JVP calls), member_jvps [B]."""
    B, n = X0.shape
    dev, dt = X0.device, X0.dtype
    delta = 1.0 / n
    Ra_s = plan._param(Ra_s, B)
    mu0, sign, ds = (plan._param(v, B).clone() for v in (mu0, sign, ds))
    mask = symmetry_mask(plan, dev)
    X, mu, X0, X_dot, mu_dot, pred_ok, njvp = predict_batched(plan, X0, mu0, sign, ds, Ra_s, tol_newton, krylov, max_restarts,
                                                              lgmres_rtol)
    Yx, Ymu = X.clone(), mu.clone()                  # the iterate Y = (X, mu); X is re-masked before every evaluation
    active = pred_ok.clone()                         # LGMRES failure in Predict raises in the reference
    failed = ~pred_ok
    it = torch.zeros(B, dtype=torch.long, device=dev)
    its_total = torch.zeros(B, dtype=torch.long, device=dev)
    halvings = torch.zeros(B, dtype=torch.long, device=dev)
    mj = torch.zeros(B, dtype=torch.long, device=dev)
    err_X = torch.ones(B, dtype=dt, device=dev)
    err_mu = torch.ones(B, dtype=dt, device=dev)
    nan = torch.full((B,), float("nan"), dtype=dt, device=dev)
    hist = []
    Xe = mue = dfmu = None

    plus = bool(getattr(plan, "has_jvp_plus", False))

    def DG(dY):
        """The bordered operator (Main.py:914-915); with `plus`, DG + I (every batched_gmres call below is shifted then)."""
        dX, dmu = dY[:, :n].contiguous(), dY[:, n]
        top = plan.jvp_apply(dX, mue, Ra_s, plus_identity=True) if plus else plan.jvp_apply(dX, mue, Ra_s)
        top = top + dfmu * dmu[:, None]
        bot = delta * (X_dot * dX).sum(dim=1) + (1.0 - delta) * mu_dot * dmu
        if plus:
            bot = bot + dmu
        return torch.cat([top, bot[:, None]], dim=1)

    for _ in range(max_rounds):
        if not bool(active.any()):
            break
        # ds control (Main.py:888-899)
        halve = active & (it >= 5)
        if bool(halve.any()):
            ds = torch.where(halve, 0.5 * ds, ds)
            it = torch.where(halve, torch.zeros_like(it), it)
            halvings += halve.long()
            too_small = halve & (ds < tol_newton)
            failed |= too_small
            active &= ~too_small
            Yx = torch.where(halve[:, None], X0 + X_dot * ds[:, None], Yx)
            Ymu = torch.where(halve, mu0 + mu_dot * ds, Ymu)
            if not bool(active.any()):
                break
        Xe, mue = _masked(Yx, mask), Ymu                                                  # Main.py:901-902
        dfmu = plan.dF_dRa(Xe)                                                            # PDFmu (Main.py:906)
        plan.jvp_set_base(Xe)
        G = torch.empty((B, n + 1), dtype=dt, device=dev)
        G[:, :n] = plan.residual(Xe, mue, Ra_s)                                           # Main.py:909
        G[:, n] = delta * (X_dot * (Xe - X0)).sum(dim=1) + (1.0 - delta) * mu_dot * (mue - mu0) - ds
        b_norm = _wnorm(G[:, :n], G[:, n], delta)                                         # Main.py:919
        dY, info = batched_gmres(DG, G, rtol=lgmres_rtol, atol=tol_gmres * b_norm, m=krylov, max_restarts=max_restarts,
                                 active=active, shifted=plus)
        njvp += info["iters"]
        mj += info["member_iters"]
        lin_fail = active & ~info["converged"]                                            # raises in the reference (923)
        upd = active & ~lin_fail
        Yx = torch.where(upd[:, None], Yx - dY[:, :n], Yx)
        Ymu = torch.where(upd, Ymu - dY[:, n], Ymu)
        eX = torch.linalg.vector_norm(dY[:, :n], dim=1) / torch.linalg.vector_norm(Xe, dim=1)
        eM = dY[:, n].abs() / mue.abs()
        err_X = torch.where(upd, eX, err_X)
        err_mu = torch.where(upd, eM, err_mu)
        hist.append(torch.stack([torch.where(upd, eX, nan), torch.where(upd, eM, nan)], dim=1))
        it += upd.long()
        its_total += upd.long()
        failed |= lin_fail
        active &= ~lin_fail
        # while (err_X > tol or err_mu > tol) or iteration < 2 (Main.py:885)
        active &= (err_X > tol_newton) | (err_mu > tol_newton) | (it < 2)
    ok = ~failed & ~active
    ds = torch.where(ok & (it <= 4), 2.0 * ds, ds)                                        # Main.py:933-935
    # new tangent with the operator of the last corrector iteration (DGy of Main.py:917 is reused at 948)
    e = torch.zeros((B, n + 1), dtype=dt, device=dev)
    e[:, n] = 1.0
    if Xe is None:                                   # every member failed in the predictor
        Yd = torch.zeros_like(e)
        tangent_ok = torch.zeros(B, dtype=torch.bool, device=dev)
    else:
        Yd, info = batched_gmres(DG, e, rtol=max(lgmres_rtol, tol_newton), m=krylov, max_restarts=max_restarts, active=ok,
                                 shifted=plus)
        njvp += info["iters"]
        mj += info["member_iters"]
        tangent_ok = ok & info["converged"]
        nrm = _wnorm(Yd[:, :n], Yd[:, n], delta)
        Yd = torch.where(tangent_ok[:, None], Yd / torch.where(nrm > 0, nrm, torch.ones_like(nrm))[:, None],
                         torch.zeros_like(Yd))                                            # Main.py:950-953
    return {"X": Yx, "mu": Ymu, "X_dot": Yd[:, :n].contiguous(), "mu_dot": Yd[:, n].contiguous(), "ds": ds, "ok": ok,
            "tangent_ok": tangent_ok, "history": torch.stack(hist) if hist else torch.empty((0, B, 2), dtype=dt, device=dev),
            "iterations": its_total, "halvings": halvings, "jvps": njvp, "member_jvps": mj,
            "X_pred_dot": X_dot, "mu_pred_dot": mu_dot}


def newtonc_batched(plan, X, mu, sign, ds, Ra_s, tol_newton=1e-8, tol_gmres=1e-4, krylov=100, max_restarts=8,
                    lgmres_rtol=LGMRES_RTOL):
    """Natural-parameter continuation step (Main._NewtonC, Main.py:717-739): Newton at mu + sign ds from X; success
    doubles ds and accepts the new point, failure halves ds and keeps the old one.  Returns (X, mu, ds, ok, info)."""
    B = X.shape[0]
    mu, sign, ds = (plan._param(v, B) for v in (mu, sign, ds))
    mu_new = mu + sign * ds
    Xn, info = newton_batched(plan, X, mu_new, Ra_s, tol_newton=tol_newton, tol_gmres=tol_gmres, krylov=krylov,
                              max_restarts=max_restarts, lgmres_rtol=lgmres_rtol)
    ok = info["converged"]
    return (torch.where(ok[:, None], Xn, X), torch.where(ok, mu_new, mu), torch.where(ok, 2.0 * ds, 0.5 * ds), ok, info)


class BranchResult:
    """Per-member record of a continuation run: the `result` class of Main.py:677-714, one column per member."""

    def __init__(self, B):
        self.B = B
        self.Ra, self.Ra_dot, self.Norm, self.KE, self.NuT, self.NuS = [], [], [], [], [], []
        self.ds, self.arclength = [], []      # step size after the step, and whether the step was an arc-length step
        self.folds = [[] for _ in range(B)]   # (iteration, Ra, X) of detected saddle nodes per member (Y_FOLD)
        self.X_DATA, self.Ra_DATA = [], []    # checkpoints every 5 iterations (Main.py:1021-1024)
        self.Iterations = 0
        self.jvps = 0

    def stacked(self):
        """History arrays [iterations, B] on the host."""
        return {k: torch.stack(getattr(self, k)).cpu().numpy() for k in ("Ra", "Ra_dot", "Norm", "KE", "NuT", "NuS", "ds")}


def continuation_batched(plan, X, mu, N_steps, Ra_s, sign=1.0, ds=0.01, ds_min=1.0, ds_max=10.0, tol_newton=1e-8,
                         tol_gmres=1e-4, krylov=100, max_restarts=8, checkpoint_every=5, on_step=None,
                         lgmres_rtol=LGMRES_RTOL):
    """The branch loop of Main._Continuation (Main.py:958-1045) for B branch points at once.  Every member has its own
    (Y, Y_dot, ds, sign).  Per outer iteration: members with ds > ds_min take a natural-parameter Newton step
    (_NewtonC) and fall back to an arc-length step if it fails; members with ds <= ds_min take an arc-length step
    (_ContinC), after which a sign change of mu_dot marks a fold (996-997) and the direction of the next Newton step
    follows the sign of the parameter change (1000-1003).  ds is capped at ds_max after Newton steps (987-988).
    Returns a BranchResult; final state in result.X, result.mu, result.X_dot, result.mu_dot."""
    X = X.clone()
    B, n = X.shape
    dev, dt = X.device, X.dtype
    Ra_s = plan._param(Ra_s, B)
    mu, sign, ds = (plan._param(v, B).clone() for v in (mu, sign, ds))
    X_dot = torch.zeros_like(X)
    mu_dot = torch.zeros(B, dtype=dt, device=dev)
    res = BranchResult(B)
    kw = dict(tol_newton=tol_newton, tol_gmres=tol_gmres, krylov=krylov, max_restarts=max_restarts, lgmres_rtol=lgmres_rtol)
    alive = torch.ones(B, dtype=torch.bool, device=dev)     # a failed arc-length step ends the reference's run (ValueError)
    while res.Iterations < N_steps and bool(alive.any()):
        newton_m = alive & (ds > ds_min)
        arc_m = alive & ~newton_m
        ds_in = ds.clone()
        X_new, mu_new = X.clone(), mu.clone()
        Xd_new, mud_new = X_dot.clone(), mu_dot.clone()
        did_arc = torch.zeros(B, dtype=torch.bool, device=dev)
        if bool(newton_m.any()):
            idx = torch.nonzero(newton_m).squeeze(1)
            Xi, mui, dsi, oki, info = newtonc_batched(plan, X[idx], mu[idx], sign[idx], ds[idx], Ra_s[idx], **kw)
            res.jvps += info["jvps"]
            X_new[idx], mu_new[idx] = Xi, mui
            # a failed Newton step hands the UNCHANGED ds to the arc-length step ('Switching to arc-length', 981-983)
            ds[idx] = torch.where(oki, dsi, ds_in[idx])
            arc_m = arc_m | torch.zeros_like(arc_m).index_put((idx,), ~oki)
        if bool(arc_m.any()):
            idx = torch.nonzero(arc_m).squeeze(1)
            out = continc_batched(plan, X[idx], mu[idx], sign[idx], ds[idx], Ra_s[idx], **kw)
            res.jvps += out["jvps"]
            oki = out["ok"]
            X_new[idx] = torch.where(oki[:, None], out["X"], X[idx])
            mu_new[idx] = torch.where(oki, out["mu"], mu[idx])
            Xd_new[idx] = torch.where(oki[:, None], out["X_dot"], X_dot[idx])
            mud_new[idx] = torch.where(oki, out["mu_dot"], mu_dot[idx])
            ds[idx] = out["ds"]
            did_arc[idx] = oki
            alive = alive & ~torch.zeros_like(alive).index_put((idx,), ~oki)
            only_arc = did_arc & ~newton_m
            fold = only_arc & (mud_new * mu_dot < 0)                                # saddle detection (996-997)
            for k in torch.nonzero(fold).squeeze(1).tolist():
                res.folds[k].append((res.Iterations, float(mu_new[k]), X_new[k].clone()))
            sign = torch.where(only_arc, torch.where(mu_new > mu, torch.ones_like(sign), -torch.ones_like(sign)), sign)
        ds = torch.where(newton_m & (ds > ds_max), torch.full_like(ds, ds_max), ds)   # Main.py:987-988
        X, mu, X_dot, mu_dot = X_new, mu_new, Xd_new, mud_new
        dg = plan.diagnostics(_masked(X, symmetry_mask(plan, dev)))
        res.Ra.append(mu.clone()); res.Ra_dot.append(mu_dot.clone())
        res.Norm.append(dg[:, 0].clone()); res.KE.append(dg[:, 1].clone())
        res.NuT.append(dg[:, 2].clone()); res.NuS.append(dg[:, 3].clone())
        res.ds.append(ds.clone()); res.arclength.append(did_arc.clone())
        if checkpoint_every and res.Iterations % checkpoint_every == 0:
            res.X_DATA.append(X.clone()); res.Ra_DATA.append(mu.clone())
        if on_step is not None:
            on_step(res, X, mu)
        res.Iterations += 1
    res.alive = alive
    res.X, res.mu, res.X_dot, res.mu_dot, res.ds_final, res.sign = X, mu, X_dot, mu_dot, ds, sign
    return res


def arclength_batched(plan, X0, mu0, X_dot, mu_dot, ds, Ra_s, tol_newton=1e-8, tol_gmres=1e-4, krylov=100, max_it=8,
                      max_restarts=8):
    """One bordered corrector from a GIVEN tangent (no predictor solve, no step-size control): the inner loop of
    Main._ContinC (Main.py:885-930) for callers that carry their own tangent.
    Returns (X, mu, X_dot_new, mu_dot_new, history [it, B, 2], jvps)."""
    B, n = X0.shape
    dev, dt = X0.device, X0.dtype
    delta = 1.0 / n
    Ra_s = plan._param(Ra_s, B)
    mu0, mu_dot, ds = (plan._param(v, B) for v in (mu0, mu_dot, ds))
    X = X0 + X_dot * ds[:, None]
    mu = mu0 + mu_dot * ds
    hist, njvp = [], 0
    state = {}

    def DG(dY):
        dX, dmu = dY[:, :n].contiguous(), dY[:, n]
        top = plan.jvp_apply(dX, state["mu"], Ra_s) + state["dfmu"] * dmu[:, None]
        bot = delta * (X_dot * dX).sum(dim=1) + (1.0 - delta) * mu_dot * dmu
        return torch.cat([top, bot[:, None]], dim=1)

    def linearise(Xc, muc):
        state["mu"], state["dfmu"] = muc, plan.dF_dRa(Xc)
        plan.jvp_set_base(Xc)

    for _ in range(max_it):
        G = torch.empty((B, n + 1), dtype=dt, device=dev)
        G[:, :n] = plan.residual(X, mu, Ra_s)
        G[:, n] = delta * (X_dot * (X - X0)).sum(dim=1) + (1.0 - delta) * mu_dot * (mu - mu0) - ds
        linearise(X, mu)
        dY, info = batched_gmres(DG, G, rtol=LGMRES_RTOL, atol=tol_gmres * _wnorm(G[:, :n], G[:, n], delta), m=krylov,
                                 max_restarts=max_restarts)
        njvp += info["iters"]
        err_X = torch.linalg.vector_norm(dY[:, :n], dim=1) / torch.linalg.vector_norm(X, dim=1)
        err_mu = dY[:, n].abs() / mu.abs()
        X = X - dY[:, :n]
        mu = mu - dY[:, n]
        hist.append(torch.stack([err_X, err_mu], dim=1))
        if bool(((err_X <= tol_newton) & (err_mu <= tol_newton)).all()) and len(hist) >= 2:
            break
    e = torch.zeros((B, n + 1), dtype=dt, device=dev)
    e[:, n] = 1.0
    linearise(X, mu)
    Yd, info = batched_gmres(DG, e, rtol=tol_newton, m=krylov, max_restarts=max_restarts)
    njvp += info["iters"]
    Yd = Yd / _wnorm(Yd[:, :n], Yd[:, n], delta)[:, None]
    return X, mu, Yd[:, :n].contiguous(), Yd[:, n].contiguous(), torch.stack(hist), njvp

"""Lock-step batched Krylov drivers on the device: many independent Newton / pseudo-arc-length problems (one per
ensemble member: parameter points, branch points) advance together, every Krylov vector of every member costing one
batched GPU Jacobian-vector product (EnsemblePlan.jvp = PDFX, Main.py:498-521).

These mirror the *iterations* of the reference's host drivers --
  newton_batched        <-> Main._Newton   (Main.py:523-539:  X <- X - DF(X)^-1 F(X), error = |dv| / |X|)
  arclength_batched     <-> Main._ContinC  (Main.py:839-955:  predictor, bordered corrector, new tangent)
-- with a restarted GMRES (batched Arnoldi, CGS2 re-orthogonalisation) in place of SciPy's single-vector LGMRES, which
cannot interleave the matvec requests of several solves.  The inexact-Newton tolerances are the reference's
(tol_gmres * |F| for the linear solves, tol_newton for the outer iteration), so converged states agree with the
reference's to the Newton tolerance; the per-iteration histories differ in the last digits because the Krylov method
does.  The Krylov algebra (dot products, small least-squares solves) is plain torch on the device: plumbing around the
operator, which is where the time goes.
"""
from __future__ import annotations

import torch


def batched_gmres(matvec, b, rtol=1e-4, atol=None, m=60, max_restarts=8, x0=None):
    """Solve A_k x_k = b_k for every row k of b [B, n] in lock step.  matvec maps [B, n] -> [B, n] (row-wise
    independent operators).  Stops when every member satisfies |r_k| <= max(atol_k, rtol*|b_k|).
    Returns (x, info) with info = {"iters": total Krylov vectors, "converged": bool mask [B], "resid": |r_k|}."""
    B, n = b.shape
    dev, dt = b.device, b.dtype
    bnorm = torch.linalg.vector_norm(b, dim=1)
    tol = rtol * bnorm if atol is None else torch.maximum(torch.as_tensor(atol, dtype=dt, device=dev).expand(B),
                                                          torch.zeros_like(bnorm))
    tol = torch.clamp(tol, min=torch.finfo(dt).tiny)
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    V = torch.empty((B, m + 1, n), dtype=dt, device=dev)
    total = 0
    resid = bnorm.clone()
    for _ in range(max_restarts):
        r = b - matvec(x) if (x0 is not None or total > 0) else b.clone()
        beta = torch.linalg.vector_norm(r, dim=1)
        resid = beta
        if bool((beta <= tol).all()):
            break
        safe_beta = torch.where(beta > 0, beta, torch.ones_like(beta))
        V[:, 0] = r / safe_beta[:, None]
        H = torch.zeros((B, m + 1, m), dtype=dt, device=dev)
        cs = torch.zeros((B, m), dtype=dt, device=dev)
        sn = torch.zeros((B, m), dtype=dt, device=dev)
        gvec = torch.zeros((B, m + 1), dtype=dt, device=dev)
        gvec[:, 0] = beta
        jdone = 0
        for j in range(m):
            w = matvec(V[:, j].contiguous())
            total += 1
            # classical Gram-Schmidt, applied twice
            Vj = V[:, :j + 1]
            h = torch.bmm(Vj, w.unsqueeze(2)).squeeze(2)
            w = w - torch.bmm(h.unsqueeze(1), Vj).squeeze(1)
            h2 = torch.bmm(Vj, w.unsqueeze(2)).squeeze(2)
            w = w - torch.bmm(h2.unsqueeze(1), Vj).squeeze(1)
            h = h + h2
            hn = torch.linalg.vector_norm(w, dim=1)
            V[:, j + 1] = w / torch.where(hn > 0, hn, torch.ones_like(hn))[:, None]
            col = torch.cat([h, hn[:, None]], dim=1)          # [B, j+2]
            # previous Givens rotations
            for i in range(j):
                t = cs[:, i] * col[:, i] + sn[:, i] * col[:, i + 1]
                col[:, i + 1] = -sn[:, i] * col[:, i] + cs[:, i] * col[:, i + 1]
                col[:, i] = t
            den = torch.sqrt(col[:, j] ** 2 + col[:, j + 1] ** 2)
            den = torch.where(den > 0, den, torch.ones_like(den))
            cs[:, j] = col[:, j] / den
            sn[:, j] = col[:, j + 1] / den
            col[:, j] = cs[:, j] * col[:, j] + sn[:, j] * col[:, j + 1]
            col[:, j + 1] = 0.0
            H[:, :j + 2, j] = col
            gvec[:, j + 1] = -sn[:, j] * gvec[:, j]
            gvec[:, j] = cs[:, j] * gvec[:, j]
            jdone = j + 1
            resid = gvec[:, j + 1].abs()
            if bool((resid <= tol).all()):
                break
        R = H[:, :jdone, :jdone]
        # guard exactly singular diagonal entries of members that converged early (their remaining columns are zero)
        diag = torch.diagonal(R, dim1=1, dim2=2)
        R = R + torch.diag_embed(torch.where(diag.abs() > 0, torch.zeros_like(diag), torch.ones_like(diag)))
        y = torch.linalg.solve_triangular(R, gvec[:, :jdone].unsqueeze(2), upper=True).squeeze(2)
        x = x + torch.bmm(y.unsqueeze(1), V[:, :jdone]).squeeze(1)
        if bool((resid <= tol).all()):
            break
    return x, {"iters": total, "converged": resid <= tol, "resid": resid}


def newton_batched(plan, X, Ra, Ra_s, tol_newton=1e-8, tol_gmres=1e-4, krylov=80, max_it=5, max_restarts=6):
    """B concurrent matrix-free Newton solves for steady states (Main._Newton, Main.py:430-554; use a plan built
    with dt = 1 like the reference's default).  X: [B, 3N] device tensor.  Returns (X, history [iterations, B],
    converged [B], total JVPs).  Members that reached tol_newton are frozen."""
    X = X.clone()
    B = X.shape[0]
    Ra = plan._param(Ra, B)
    Ra_s = plan._param(Ra_s, B)
    active = torch.ones(B, dtype=torch.bool, device=X.device)
    hist, njvp = [], 0
    for _ in range(max_it):
        fx = plan.residual(X, Ra, Ra_s)                           # PFX (Main.py:473-496)
        fx = torch.where(active[:, None], fx, torch.zeros_like(fx))
        plan.jvp_set_base(X)                                      # X is fixed during the linear solve

        def DF(v):
            return plan.jvp_apply(v, Ra, Ra_s)                    # PDFX (Main.py:498-521)

        dv, info = batched_gmres(DF, fx, rtol=tol_gmres, m=krylov, max_restarts=max_restarts)
        njvp += info["iters"]
        dv = torch.where(active[:, None], dv, torch.zeros_like(dv))
        X = X - dv
        err = torch.linalg.vector_norm(dv, dim=1) / torch.linalg.vector_norm(X, dim=1)
        hist.append(torch.where(active, err, torch.zeros_like(err)))
        active = active & (err > tol_newton)
        if not bool(active.any()):
            break
    return X, torch.stack(hist), ~active, njvp


def arclength_batched(plan, X0, mu0, X_dot, mu_dot, ds, Ra_s, tol_newton=1e-8, tol_gmres=1e-4, krylov=80, max_it=8,
                      max_restarts=6):
    """One pseudo-arc-length step for B branch points in lock step: predictor Y = Y0 + ds * Y_dot, then the bordered
    Newton corrector of Main._ContinC (Main.py:885-930) on
        G(X, mu) = [ P F(X, mu) ;  delta X_dot.(X - X0) + (1 - delta) mu_dot (mu - mu0) - ds ],   delta = 1/(3N),
    and the new unit tangent from DG Y_dot_new = e_last (Main.py:944-953).  mu is the thermal Rayleigh number.
    X0, X_dot: [B, 3N]; mu0, mu_dot, ds: [B].  Returns (X, mu, X_dot_new, mu_dot_new, history [it, B, 2], njvp)."""
    B, n = X0.shape
    dev, dt = X0.device, X0.dtype
    delta = 1.0 / n
    Ra_s = plan._param(Ra_s, B)
    mu0 = plan._param(mu0, B)
    mu_dot = plan._param(mu_dot, B)
    ds = plan._param(ds, B)
    X = X0 + X_dot * ds[:, None]
    mu = mu0 + mu_dot * ds
    hist, njvp = [], 0

    def make_DG(Xc, muc):
        dfmu = plan.dF_dRa(Xc)                                    # PDFmu (Main.py:829-837)
        plan.jvp_set_base(Xc)

        def DG(dY):
            dX, dmu = dY[:, :n].contiguous(), dY[:, n]
            top = plan.jvp_apply(dX, muc, Ra_s) + dfmu * dmu[:, None]
            bot = delta * (X_dot * dX).sum(dim=1) + (1.0 - delta) * mu_dot * dmu
            return torch.cat([top, bot[:, None]], dim=1)

        return DG

    DG = None
    for _ in range(max_it):
        G = torch.empty((B, n + 1), dtype=dt, device=dev)
        G[:, :n] = plan.residual(X, mu, Ra_s)
        G[:, n] = delta * (X_dot * (X - X0)).sum(dim=1) + (1.0 - delta) * mu_dot * (mu - mu0) - ds
        DG = make_DG(X, mu)
        b_norm = torch.sqrt(delta * (G[:, :n] ** 2).sum(dim=1) + (1.0 - delta) * G[:, n] ** 2)   # Main.py:919
        dY, info = batched_gmres(DG, G, atol=tol_gmres * b_norm, m=krylov, max_restarts=max_restarts)
        njvp += info["iters"]
        err_X = torch.linalg.vector_norm(dY[:, :n], dim=1) / torch.linalg.vector_norm(X, dim=1)
        err_mu = dY[:, n].abs() / mu.abs()
        X = X - dY[:, :n]
        mu = mu - dY[:, n]
        hist.append(torch.stack([err_X, err_mu], dim=1))
        if bool(((err_X <= tol_newton) & (err_mu <= tol_newton)).all()) and len(hist) >= 2:
            break
    # new tangent: DG Y_dot = (0, ..., 0, 1), normalised in the weighted norm (Main.py:944-953)
    e = torch.zeros((B, n + 1), dtype=dt, device=dev)
    e[:, n] = 1.0
    DG = make_DG(X, mu)
    Yd, info = batched_gmres(DG, e, rtol=tol_newton, m=krylov, max_restarts=max_restarts)
    njvp += info["iters"]
    nrm = torch.sqrt(delta * (Yd[:, :n] ** 2).sum(dim=1) + (1.0 - delta) * Yd[:, n] ** 2)
    Yd = Yd / nrm[:, None]
    return X, mu, Yd[:, :n].contiguous(), Yd[:, n].contiguous(), torch.stack(hist), njvp

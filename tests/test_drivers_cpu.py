"""Host logic of the batched Newton / pseudo-arc-length drivers (krylov.py) against golden runs of the reference's own
Main._Newton / _ContinC / _Continuation (tests/golden/make_golden_continuation.py -> continuation.npz), with the
operators supplied by the CPU oracle (tests/oracle_plan.py).  CPU only; the GPU versions of these tests run the same
drivers on EnsemblePlan (tests/test_gpu_dropin.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle_plan import OraclePlan
from spectraldoublediffusiveconvection_b200 import krylov

G = load_golden("continuation")


def _plan():
    return OraclePlan(int(G["N_fm"]), int(G["N_r"]), float(G["d"]), 1.0, float(G["Pr"]), float(G["Tau"]),
                      symmetric=bool(G["symmetric"]))


def _t(a):
    return torch.as_tensor(np.atleast_2d(np.asarray(a, dtype=np.float64)))


def test_newton_matches_reference_driver():
    pl = _plan()
    Ra = float(G["Ra_newton"])
    # member 0: the golden problem; member 1: an easier one (stops earlier: exercises the per-member mask);
    # member 2: starts from the answer (converges at once)
    X0 = torch.cat([_t(G["X_start"]), _t(G["X_start"]), _t(G["newton_X"])])
    Ras = torch.tensor([Ra, float(G["Ra"]) + 0.5, Ra], dtype=torch.float64)
    X, info = krylov.newton_batched(pl, X0, Ras, float(G["Ra_s"]), krylov=60)
    assert bool(info["converged"].all()), info
    hist = info["history"].numpy()
    ref = G["newton_history"]
    assert int(info["iterations"][0]) == len(ref)
    # inexact Newton with a 1e-4 linear tolerance: the histories agree to that tolerance, not to rounding
    assert np.allclose(hist[:len(ref) - 1, 0], ref[:-1], rtol=2e-3), (hist[:, 0], ref)
    assert hist[len(ref) - 1, 0] <= 1e-8
    assert rel_l2(X[0].numpy(), G["newton_X"]) < 1e-7
    assert int(info["iterations"][2]) <= 2 and int(info["iterations"][1]) <= int(info["iterations"][0])
    assert np.isnan(hist[int(info["iterations"][2]):, 2]).all()          # a stopped member records nothing more
    assert int(info["member_jvps"][2]) < int(info["member_jvps"][0])
    dg = pl.diagnostics(X)[0, :4].numpy()
    assert np.allclose(dg, G["newton_diag"], rtol=1e-6)


def test_continc_matches_reference_driver():
    """Two branch points in one batch with different step sizes: the golden runs 'a' (ds = 0.5) and 'b' (ds = 8)."""
    pl = _plan()
    Ra = float(G["Ra_newton"])
    X0 = torch.cat([_t(G["newton_X"]), _t(G["newton_X"])])
    ds0 = torch.tensor([float(G["continc_a_ds0"]), float(G["continc_b_ds0"])], dtype=torch.float64)
    out = krylov.continc_batched(pl, X0, Ra, 1.0, ds0, float(G["Ra_s"]), krylov=60)
    assert bool(out["ok"].all()) and bool(out["tangent_ok"].all())
    n = X0.shape[1]
    for m, tag in enumerate("ab"):
        Y, Yd = G["continc_%s_Y" % tag], G["continc_%s_Ydot" % tag]
        assert abs(float(out["mu"][m]) - Y[-1]) < 1e-6 * abs(Y[-1])
        assert rel_l2(out["X"][m].numpy(), Y[:-1]) < 1e-6
        assert abs(float(out["mu_dot"][m]) - Yd[-1]) < 1e-6
        # The X part of the tangent is NOT compared: SciPy's default rtol = 1e-5 ends the reference's tangent solve
        # (Main.py:948) when |r| <= 1e-5 although |DF_mu| itself is of that size, so the X part it returns is 4 % of
        # the true tangent (the secant through the two points).  Only mu_dot - all that fold detection uses - and the
        # normalisation of Main.py:953 are reference facts.
        delta = 1.0 / n
        assert abs(delta * float((out["X_dot"][m] ** 2).sum()) + (1 - delta) * float(out["mu_dot"][m]) ** 2 - 1.0) < 1e-12
        assert float(out["ds"][m]) == float(G["continc_%s_ds" % tag])      # doubled: <= 4 corrector iterations
        ref = G["continc_%s_history" % tag]
        assert int(out["iterations"][m]) == len(ref)
        mine = out["history"][:len(ref), m].numpy()
        assert np.allclose(mine[:-1, 0], ref[:-1, 0], rtol=5e-3), (mine, ref)
        # the arclength constraint holds at the new point (Main.py:910)
        cons = (delta * float((out["X_pred_dot"][m] * (out["X"][m] - X0[m])).sum())
                + (1 - delta) * float(out["mu_pred_dot"][m]) * (float(out["mu"][m]) - Ra) - float(ds0[m]))
        assert abs(cons) < 1e-7


def test_continc_halves_ds_when_the_corrector_stalls():
    """A step far too long for the corrector: ds is halved (Main.py:888-899) until the corrector converges; an easy
    member in the same batch is unaffected."""
    pl = _plan()
    Ra = float(G["Ra_newton"])
    X0 = torch.cat([_t(G["newton_X"]), _t(G["newton_X"])])
    ds0 = torch.tensor([4000.0, 0.5], dtype=torch.float64)
    out = krylov.continc_batched(pl, X0, Ra, -1.0, ds0, float(G["Ra_s"]), krylov=60, max_rounds=40)
    assert int(out["halvings"][1]) == 0 and bool(out["ok"][1]) and float(out["ds"][1]) == 1.0
    assert float(out["mu"][1]) < Ra                                           # sign = -1: towards smaller Ra
    if bool(out["ok"][0]):
        assert int(out["halvings"][0]) >= 1
        assert float(out["ds"][0]) in (4000.0 * 0.5 ** int(out["halvings"][0]), 2 * 4000.0 * 0.5 ** int(out["halvings"][0]))
        res = pl.residual(out["X"][0:1] * krylov.symmetry_mask(pl), out["mu"][0:1], float(G["Ra_s"]))
        assert float(torch.linalg.vector_norm(res)) < 1e-6


def test_branch_loop_matches_reference_driver():
    pl = _plan()
    nsteps = int(G["branch_steps"])
    res = krylov.continuation_batched(pl, _t(G["newton_X"]), float(G["Ra_newton"]), nsteps, float(G["Ra_s"]), krylov=60)
    h = res.stacked()
    assert res.Iterations == nsteps and bool(res.alive.all())
    assert np.allclose(h["Ra"][:, 0], G["branch_Ra"], rtol=1e-7), (h["Ra"][:, 0], G["branch_Ra"])
    assert np.allclose(h["KE"][:, 0], G["branch_KE"], rtol=1e-5)
    assert np.allclose(h["NuT"][:, 0], G["branch_NuT"], rtol=1e-5)
    # arc-length steps first (ds 0.01 doubling up to ds_min), natural-parameter Newton steps afterwards
    arc = torch.stack(res.arclength)[:, 0].numpy()
    assert arc[:7].all()
    assert np.allclose(h["Ra_dot"][:7, 0], G["branch_Ra_dot"][:7], rtol=1e-6)
    assert len(res.folds[0]) == 0                                     # no saddle node on this stretch of the branch
    # checkpoints every 5 iterations (Main.py:1021-1024)
    assert len(res.X_DATA) == len(G["branch_Ra_DATA"])
    assert np.allclose(torch.stack(res.Ra_DATA)[:, 0].numpy(), G["branch_Ra_DATA"], rtol=1e-7)
    assert rel_l2(res.X_DATA[-1][0].numpy(), G["branch_X_DATA"][-1]) < 1e-5


def test_gmres_masks_members():
    torch.manual_seed(3)
    B, n = 4, 30
    A = torch.eye(n, dtype=torch.float64)[None] * 2.5 + 0.2 * torch.randn(B, n, n, dtype=torch.float64)
    b = torch.randn(B, n, dtype=torch.float64)
    mv = lambda v: torch.bmm(A, v.unsqueeze(2)).squeeze(2)
    active = torch.tensor([True, False, True, True])
    x0 = torch.randn(B, n, dtype=torch.float64)
    x, info = krylov.batched_gmres(mv, b, rtol=torch.tensor(1e-9).item(), m=30, x0=x0, active=active)
    assert torch.equal(x[1], x0[1]) and int(info["member_iters"][1]) == 0   # an inactive member is never touched
    r = torch.linalg.vector_norm(mv(x) - b, dim=1) / torch.linalg.vector_norm(b, dim=1)
    assert bool((r[active] < 1e-8).all())
    # members with loose tolerances stop early and stay frozen while the strict one continues
    atol = torch.tensor([1e-2, 1e-2, 1e-10, 1e-5], dtype=torch.float64)
    x, info = krylov.batched_gmres(mv, b, rtol=0.0, atol=atol, m=30)
    it = info["member_iters"]
    assert int(it[0]) < int(it[3]) < int(it[2])
    r = torch.linalg.vector_norm(mv(x) - b, dim=1)
    assert bool((r <= atol * 1.0001).all()) and bool(info["converged"].all())


def test_tight_tangent_equals_the_secant():
    """With SciPy's 1e-5 floor switched off the tangent solve converges and its direction is the secant through the old
    and the new branch point (up to the curvature of the branch over ds = 0.5)."""
    pl = _plan()
    Ra = float(G["Ra_newton"])
    X0 = _t(G["newton_X"])
    out = krylov.continc_batched(pl, X0, Ra, 1.0, torch.tensor([0.5], dtype=torch.float64), float(G["Ra_s"]), krylov=80,
                                 lgmres_rtol=0.0)
    assert bool(out["ok"].all()) and bool(out["tangent_ok"].all())
    sec = (out["X"][0] - X0[0]) / (out["mu"][0] - Ra)
    tan = out["X_dot"][0] / out["mu_dot"][0]
    assert float(torch.linalg.vector_norm(tan - sec) / torch.linalg.vector_norm(sec)) < 2e-2

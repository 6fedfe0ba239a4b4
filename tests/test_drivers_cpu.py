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


class _FoldPlan:
    """Saddle-node normal form G(x, mu) = mu - a - x^2 (componentwise, a per member): steady branches
    x = +-sqrt(mu - a) meet in a fold at mu = a.  Exposes what the drivers use of an EnsemblePlan."""
    symmetric = False
    N_fm, nr = 2, 1

    def __init__(self, a):
        self.a = torch.as_tensor(a, dtype=torch.float64)
        self.base = None

    def _param(self, v, B):
        v = torch.as_tensor(v, dtype=torch.float64).reshape(-1)
        return (v.expand(B) if v.numel() == 1 else v).contiguous()

    def residual(self, X, mu, Ra_s):
        return self._param(mu, X.shape[0])[:, None] - self.a[:X.shape[0], None] - X * X

    def jvp_set_base(self, X):
        self.base = X.clone()

    def jvp_apply(self, dv, mu, Ra_s):           # no has_jvp_plus: this double exercises the unshifted path
        return -2.0 * self.base * dv

    def dF_dRa(self, X):
        return torch.ones_like(X)

    def diagnostics(self, X):
        out = torch.zeros((X.shape[0], 6), dtype=torch.float64)
        out[:, 0] = torch.linalg.vector_norm(X, dim=1)
        out[:, 1] = X[:, 0]                        # "KE" slot: the signed amplitude, to see which branch a member is on
        return out


def test_branch_loop_detects_folds_per_member():
    """Two members with folds at different parameter values walk down their upper branches (sign = -1), go round the
    saddle node by arc-length steps and come back up on the lower branch: the sign change of mu_dot is recorded as a fold
    for each member where ITS fold is (Main.py:996-997), and the walk continues with mu increasing."""
    a = torch.tensor([10.0, 10.5], dtype=torch.float64)

    class Plan(_FoldPlan):
        def residual(self, X, mu, Ra_s):           # sub-batches of the branch loop carry their own a through Ra_s
            return self._param(mu, X.shape[0])[:, None] - self._param(Ra_s, X.shape[0])[:, None] - X * X

    pl = Plan(a)
    mu0 = torch.tensor([10.6, 11.0], dtype=torch.float64)
    X0 = torch.sqrt(mu0 - a)[:, None].repeat(1, 4)
    # ds_min above every step size: always the arc-length branch of the loop, the only one in which the reference looks
    # for folds and re-decides the direction (with Newton steps enabled and a failed one falling back to an arc-length step,
    # Main.py:976-988 neither records the fold nor flips `sign`, and the walk oscillates around the saddle node -- the
    # batched loop reproduces that too)
    res = krylov.continuation_batched(pl, X0, mu0, 14, a, sign=-1.0, ds=0.01, ds_min=1e9, krylov=6, lgmres_rtol=0.0)
    h = res.stacked()
    assert bool(res.alive.all())
    for m in range(2):
        assert len(res.folds[m]) >= 1, (m, h["Ra"][:, m])
        it, mu_f, X_f = res.folds[m][0]
        assert abs(mu_f - float(a[m])) < 0.1                      # the fold sits at mu = a
        amp = h["KE"][:, m]
        assert amp[0] > 0 and amp[it] * amp[it - 1] < 0            # the step that found the fold changed branches
        assert np.all(np.diff(h["Ra"][:it, m]) < 0)                # mu decreased up to the fold ...
        assert np.all(np.diff(h["Ra"][it + 1:it + 4, m]) > 0)      # ... and increases afterwards (sign re-decided, 1000-1003)
        assert h["Ra"][:, m].min() >= float(a[m]) - 1e-6           # no steady state below the fold
        # every recorded point is a steady state of ITS member
        r = pl.residual(res.X, res.mu, a)
        assert float(r.abs().max()) < 1e-7

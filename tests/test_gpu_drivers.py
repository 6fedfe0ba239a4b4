"""The batched Newton / pseudo-arc-length drivers (krylov.py) on the GPU path -- EnsemblePlan operators, fused
Gram-Schmidt kernels -- against golden runs of the reference's own Main._Newton / _ContinC / _Continuation
(golden/continuation.npz) and against the torch.bmm orthogonalisation."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

G = load_golden("continuation")


def _plan(B):
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    return EnsemblePlan(int(G["N_fm"]), int(G["N_r"]), float(G["d"]), 1.0, float(G["Pr"]), float(G["Tau"]),
                        symmetric=bool(G["symmetric"]), max_batch=B)


def _t(a):
    return torch.as_tensor(np.atleast_2d(np.asarray(a, dtype=np.float64))).cuda()


def test_fused_gram_schmidt_matches_bmm():
    from spectraldoublediffusiveconvection_b200 import krylov
    torch.manual_seed(0)
    for B, m, n in ((3, 12, 1441), (5, 40, 4097), (2, 7, 22273)):
        V = torch.zeros((B, m + 1, n), dtype=torch.float64, device="cuda")
        j = m - 2
        Q, _ = torch.linalg.qr(torch.randn(B, n, j + 1, dtype=torch.float64, device="cuda"))
        V[:, :j + 1] = Q.transpose(1, 2)
        w = torch.randn(B, n, dtype=torch.float64, device="cuda")
        h1, n1, w1 = krylov._FusedOrtho(V)(j, w.clone())
        h2, n2, w2 = krylov._TorchOrtho(V)(j, w.clone())
        assert torch.allclose(h1, h2, rtol=1e-12, atol=1e-12)
        assert torch.allclose(n1, n2, rtol=1e-12)
        assert torch.allclose(w1, w2, rtol=1e-11, atol=1e-12)
        assert float(torch.bmm(V[:, :j + 1], w1.unsqueeze(2)).abs().max()) < 1e-12      # orthogonal to the basis
        # bit-reproducible: fixed reduction order
        h3, n3, w3 = krylov._FusedOrtho(V)(j, w.clone())
        assert torch.equal(h1, h3) and torch.equal(n1, n3) and torch.equal(w1, w3)


def test_newton_matches_reference_driver():
    from spectraldoublediffusiveconvection_b200 import krylov
    pl = _plan(3)
    Ra = float(G["Ra_newton"])
    X0 = torch.cat([_t(G["X_start"]), _t(G["X_start"]), _t(G["newton_X"])])
    Ras = torch.tensor([Ra, float(G["Ra"]) + 0.5, Ra], dtype=torch.float64, device="cuda")
    X, info = krylov.newton_batched(pl, X0, Ras, float(G["Ra_s"]), krylov=60)
    assert bool(info["converged"].all()), info
    hist = info["history"].cpu().numpy()
    ref = G["newton_history"]
    assert int(info["iterations"][0]) == len(ref)
    assert np.allclose(hist[:len(ref) - 1, 0], ref[:-1], rtol=2e-3), (hist[:, 0], ref)
    assert rel_l2(X[0].cpu().numpy(), G["newton_X"]) < 1e-7
    assert np.allclose(pl.diagnostics(X)[0, :4].cpu().numpy(), G["newton_diag"], rtol=1e-6)
    pl.close()


def test_continc_matches_reference_driver():
    from spectraldoublediffusiveconvection_b200 import krylov
    pl = _plan(2)
    Ra = float(G["Ra_newton"])
    X0 = torch.cat([_t(G["newton_X"]), _t(G["newton_X"])])
    ds0 = torch.tensor([float(G["continc_a_ds0"]), float(G["continc_b_ds0"])], dtype=torch.float64, device="cuda")
    out = krylov.continc_batched(pl, X0, Ra, 1.0, ds0, float(G["Ra_s"]), krylov=60)
    assert bool(out["ok"].all()) and bool(out["tangent_ok"].all())
    for m, tag in enumerate("ab"):
        Y, Yd = G["continc_%s_Y" % tag], G["continc_%s_Ydot" % tag]
        assert abs(float(out["mu"][m]) - Y[-1]) < 1e-6 * abs(Y[-1])
        assert rel_l2(out["X"][m].cpu().numpy(), Y[:-1]) < 1e-6
        assert abs(float(out["mu_dot"][m]) - Yd[-1]) < 1e-6
        assert float(out["ds"][m]) == float(G["continc_%s_ds" % tag])
        assert int(out["iterations"][m]) == len(G["continc_%s_history" % tag])
    pl.close()


def test_branch_loop_matches_reference_driver():
    from spectraldoublediffusiveconvection_b200 import krylov
    # two members: the golden branch, and the same branch followed towards smaller Ra (own sign, own ds)
    pl = _plan(2)
    nsteps = int(G["branch_steps"])
    X0 = torch.cat([_t(G["newton_X"]), _t(G["newton_X"])])
    sign = torch.tensor([1.0, -1.0], dtype=torch.float64, device="cuda")
    res = krylov.continuation_batched(pl, X0, float(G["Ra_newton"]), nsteps, float(G["Ra_s"]), sign=sign, krylov=60)
    h = res.stacked()
    assert res.Iterations == nsteps and bool(res.alive[0])
    assert np.allclose(h["Ra"][:, 0], G["branch_Ra"], rtol=1e-7), (h["Ra"][:, 0], G["branch_Ra"])
    assert np.allclose(h["KE"][:, 0], G["branch_KE"], rtol=1e-5)
    assert np.allclose(h["NuT"][:, 0], G["branch_NuT"], rtol=1e-5)
    assert np.allclose(torch.stack(res.Ra_DATA)[:, 0].cpu().numpy(), G["branch_Ra_DATA"], rtol=1e-7)
    assert rel_l2(res.X_DATA[-1][0].cpu().numpy(), G["branch_X_DATA"][-1]) < 1e-5
    assert np.all(np.diff(h["Ra"][:7, 1]) < 0)            # the second member walks down the branch
    pl.close()


@pytest.mark.parametrize("shifted", [False, True])
def test_fused_gmres_bookkeeping_matches_torch_path(shifted):
    """CUDA tensors: Gram-Schmidt and the Hessenberg / Givens bookkeeping run in the library's kernels; CPU tensors: plain
    torch.  Same systems, same tolerances (per-member, so that members retire at different steps) -> same solutions
    and the same per-member iteration counts."""
    from spectraldoublediffusiveconvection_b200 import krylov
    torch.manual_seed(5)
    B, n = 6, 300
    A = torch.eye(n, dtype=torch.float64)[None] * 3.0 + 0.04 * torch.randn(B, n, n, dtype=torch.float64)
    b = torch.randn(B, n, dtype=torch.float64)
    b[4] = 0.0
    atol = torch.tensor([1e-3, 1e-6, 1e-9, 1e-12, 1e-6, 1e-4], dtype=torch.float64)
    active = torch.tensor([True, True, True, True, True, False])
    off = 1.0 if shifted else 0.0
    Ad = A.cuda()
    xc, ic = krylov.batched_gmres(lambda v: torch.bmm(A, v.unsqueeze(2)).squeeze(2) + off * v, b, rtol=0.0, atol=atol, m=25,
                                  max_restarts=6, active=active, shifted=shifted)
    xg, ig = krylov.batched_gmres(lambda v: torch.bmm(Ad, v.unsqueeze(2)).squeeze(2) + off * v, b.cuda(), rtol=0.0,
                                  atol=atol.cuda(), m=25, max_restarts=6, active=active.cuda(), shifted=shifted)
    assert torch.equal(ig["member_iters"].cpu(), ic["member_iters"]) and ig["iters"] == ic["iters"]
    assert torch.allclose(xg.cpu(), xc, rtol=1e-9, atol=1e-12)
    assert bool(ig["converged"].all()) and float(xg[5].abs().max()) == 0.0 and float(xg[4].abs().max()) == 0.0
    r = torch.linalg.vector_norm(torch.bmm(A, xg.cpu().unsqueeze(2)).squeeze(2) - b, dim=1)
    assert bool((r[:5] <= atol[:5] * 1.001).all())

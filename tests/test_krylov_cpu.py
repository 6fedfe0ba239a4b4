"""The batched GMRES that drives the lock-step Newton / arc-length solvers, on CPU tensors with explicit matrices."""
import torch

from spectraldoublediffusiveconvection_b200.krylov import batched_gmres


def test_batched_gmres_solves_independent_systems():
    torch.manual_seed(0)
    B, n = 5, 40
    A = torch.eye(n, dtype=torch.float64)[None] * 3.0 + 0.3 * torch.randn(B, n, n, dtype=torch.float64)
    b = torch.randn(B, n, dtype=torch.float64)
    b[3] = 0.0                                     # a member whose right-hand side vanishes must not poison the batch
    x, info = batched_gmres(lambda v: torch.bmm(A, v.unsqueeze(2)).squeeze(2), b, rtol=1e-10, m=25, max_restarts=10)
    ref = torch.linalg.solve(A, b.unsqueeze(2)).squeeze(2)
    assert torch.allclose(x, ref, rtol=1e-7, atol=1e-9)
    assert bool(info["converged"].all()) and torch.isfinite(x).all()


def test_batched_gmres_restart_and_absolute_tolerance():
    torch.manual_seed(1)
    B, n = 3, 60
    A = torch.eye(n, dtype=torch.float64)[None] * 2.0 + 0.2 * torch.randn(B, n, n, dtype=torch.float64)
    b = torch.randn(B, n, dtype=torch.float64)
    atol = torch.tensor([1e-3, 1e-6, 1e-9], dtype=torch.float64)
    x, info = batched_gmres(lambda v: torch.bmm(A, v.unsqueeze(2)).squeeze(2), b, rtol=0.0, atol=atol, m=8, max_restarts=40)
    r = torch.linalg.vector_norm(torch.bmm(A, x.unsqueeze(2)).squeeze(2) - b, dim=1)
    assert bool((r <= atol * 1.01).all())


def test_batched_gmres_shifted_operator():
    """shifted=True: the matvec applies A + I (the JVP without its final subtraction); same solution as the plain solve."""
    torch.manual_seed(2)
    B, n = 4, 50
    A = torch.eye(n, dtype=torch.float64)[None] * 3.0 + 0.2 * torch.randn(B, n, n, dtype=torch.float64)
    b = torch.randn(B, n, dtype=torch.float64)
    mv = lambda v: torch.bmm(A, v.unsqueeze(2)).squeeze(2)
    x1, i1 = batched_gmres(mv, b, rtol=1e-11, m=30, max_restarts=20)
    x2, i2 = batched_gmres(lambda v: mv(v) + v, b, rtol=1e-11, m=30, max_restarts=20, shifted=True)
    ref = torch.linalg.solve(A, b.unsqueeze(2)).squeeze(2)
    assert torch.allclose(x1, ref, rtol=1e-8, atol=1e-10) and torch.allclose(x2, ref, rtol=1e-8, atol=1e-10)
    assert bool(i2["converged"].all()) and i2["iters"] == i1["iters"]

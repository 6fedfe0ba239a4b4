"""The drop-in Matrix_Operators / Transforms modules, driven with the reference's own call sequences, reproduce the
golden outputs of the unmodified reference (single member, NumPy in / NumPy out)."""
import sys

import numpy as np
import pytest

from conftest import load_golden, rel_l2
import dropin_drivers as drv

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()


def _rep(x, B):
    return _dev(np.tile(np.asarray(x)[None, :], (B, 1)))


@pytest.fixture(scope="module")
def modules():
    import spectraldoublediffusiveconvection_b200.compat as compat
    saved = {k: sys.modules.get(k) for k in ("Matrix_Operators", "Transforms")}
    MO, TR = compat.install()
    import Matrix_Operators, Transforms   # resolved by bare name, like Main.py does
    assert Matrix_Operators is MO and Transforms is TR
    yield MO, TR
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


@pytest.mark.parametrize("name", ["small_nosym", "small_sym", "cfg1_nosym"])
def test_driver_call_sequences(modules, name):
    MO, TR = modules
    g = load_golden(name)
    N_fm, N_r = int(g["N_fm"]), int(g["N_r"])
    d, dt, Pr, Tau = float(g["d"]), float(g["dt"]), float(g["Pr"]), float(g["Tau"])
    Ra, Ra_s, sym = float(g["Ra"]), float(g["Ra_s"]), bool(g["symmetric"])
    ops = drv.build_matrix_operators(MO, N_fm, N_r, d, dt, Pr, Tau)
    step, residual, jvp, dmu = drv.make_closures(MO, ops, Ra, Ra_s, dt, Pr, Tau, sym)
    Xb, dv = g["Xb"], g["dv"]
    Xb_copy = Xb.copy()
    assert rel_l2(step(Xb), g["step_Xb"]) < 1e-10
    assert np.array_equal(Xb, Xb_copy)                     # inputs are never mutated
    assert rel_l2(jvp(dv, Xb), g["jvp_Xb"]) < 1e-10
    assert rel_l2(dmu(Xb), g["dmu_Xb"]) < 1e-10
    # the time loop of Main._Time_Step (Main.py:286-329): X <- X_SYM * Step(X)
    X = g["X0"].copy()
    nr = N_r - 1
    mask = np.ones(3 * N_fm * nr)
    if sym:
        m3 = mask.reshape(3, N_fm, nr)
        m3[0, 0::2] = 0.0
        m3[1:, 1::2] = 0.0
    n_steps = min(int(g["n_steps"]), 20)
    for it in range(n_steps):
        Xn = step(X)
        if it + 1 in (1, 10):
            assert rel_l2(Xn, g["X_step%d" % (it + 1)]) < 1e-10
        X = mask * Xn
    D, R = ops[0], ops[1]
    ke = drv.kinetic_energy(MO, TR, Xn, R, D, N_fm, nr, sym)
    assert abs(ke / g["diag_hist"][n_steps - 1][1] - 1) < 1e-9


def test_single_function_signatures(modules):
    MO, TR = modules
    g = load_golden("small_nosym")
    N_fm, N_r, d = int(g["N_fm"]), int(g["N_r"]), float(g["d"])
    nr, N = N_r - 1, (N_r - 1) * N_fm
    D, R = MO.cheb_radial(N_r, d)
    Xb = g["Xb"]
    assert rel_l2(MO.J_theta_RT(Xb[:N], nr, N_fm, False), g["J_theta_RT"]) < 1e-13
    assert rel_l2(MO.A2_SINE_R2(Xb[:N], N_fm, nr, D, R, False), g["A2_SINE_R2"]) < 1e-13
    assert rel_l2(MO.NLIN_FX(Xb, D, R, N_fm, nr, False), g["NLIN_FX"]) < 1e-12
    P, T, C = MO.X_to_Vecs(Xb, N_fm, nr, False)
    assert P.shape == (nr, N_fm) and np.array_equal(MO.Vecs_to_X(P, T, C, N_fm, nr, False), Xb)
    with pytest.raises(ValueError):
        MO.NLIN_FX(np.zeros(3 * nr * 15), D, R, 15, nr, False)     # odd N_fm (Matrix_Operators.py:758)
    gt = load_golden("transforms")
    assert rel_l2(TR.IDCT(gt["in_hat_16"], n=24), gt["IDCT_16"]) < 1e-13
    assert rel_l2(TR.DST(gt["in_grid_16"])[..., 0:16], gt["DST_trunc_16"]) < 1e-13
    assert rel_l2(TR.grid(24), gt["grid_16"]) < 1e-15
    # resolution transfer helpers keep a state unchanged when the resolution is unchanged, and theta-interpolation
    # to more modes zero-pads the spectrum
    # resolution transfer (Matrix_Operators.py:901-1011) against outputs of the reference's own functions
    gi = load_golden("interp")
    Xi, di = gi["X"], float(gi["d"])
    assert MO.INTERP_RADIAL(10, 10, Xi, di) is Xi
    assert rel_l2(MO.INTERP_THETAS(32, 16, Xi), gi["theta_up"]) < 1e-12
    assert rel_l2(MO.INTERP_THETAS(8, 16, Xi), gi["theta_down"]) < 1e-12
    assert rel_l2(MO.INTERP_RADIAL(14, 10, Xi, di), gi["radial_up"]) < 1e-9


def test_newton_history_matches_cpu_path(modules):
    """BASELINE config 4 in miniature: the Newton iteration of Main._Newton (host SciPy LGMRES, matrix-free JVPs)
    produces the same error history on the GPU operators as on the CPU oracle operators, from the same X0."""
    from oracle import sddc_oracle as orc
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    MO, _ = modules
    # wide-gap l=2 case of Tests/Run_Tests.py:140-199 (d=2, Pr=10, N_fm=32, N_r=16, symmetric)
    N_fm, N_r, d, Pr, Tau, Ra, Ra_s, sym = 32, 16, 2.0, 10.0, 1.0, 6780.0, 0.0, True
    nr = N_r - 1
    # a reproducible starting state: time-step a seeded random IC towards the steady branch on the GPU
    pl = EnsemblePlan(N_fm, N_r, d, 0.075, Pr, Tau, symmetric=sym, max_batch=1)
    X0 = np.random.default_rng(0).random(3 * nr * N_fm)
    X0 = 1e-3 * X0 / np.linalg.norm(X0)
    X0 = torch.as_tensor(X0 * orc.sym_mask(N_fm, nr).reshape(-1)).cuda().reshape(1, -1)
    Xs = pl.step(X0, Ra, Ra_s, nsteps=13000).cpu().numpy().ravel()   # transient, not yet on the steady branch
    pl.close()
    assert np.isfinite(Xs).all() and 0.1 < np.linalg.norm(Xs) < 0.3
    Xg, hg, ng = drv.newton(MO, Xs, Ra, Ra_s, Tau, Pr, d, N_fm, N_r, sym, max_it=5)
    Xc, hc, nc = drv.newton(drv.OracleOperators(orc, N_fm, N_r, d, 1.0, Pr, Tau), Xs, Ra, Ra_s, Tau, Pr, d, N_fm, N_r,
                            sym, max_it=5)
    print("newton history gpu", hg, "cpu", hc, "matvecs", ng, nc, "rel", np.abs(hg / hc - 1))
    assert len(hg) == len(hc) == 5 and ng == nc, (hg, hc, ng, nc)
    # north_star: histories agree to 1e-10 (absolute; the last entries are ~1e-6 and agree to ~1e-13 absolute)
    assert np.max(np.abs(hg - hc)) < 1e-10 and np.allclose(hg, hc, rtol=1e-6, atol=0), (hg, hc)
    assert rel_l2(Xg, Xc) < 1e-9
    # both land on the steady branch of the reference's wide-gap test (KE = 2.57522e-2)
    ke = orc.kinetic_energy(Xg, orc.Operators(N_fm, N_r, d, 1.0, Pr, Tau), sym)
    assert abs(ke / 2.5752204992e-2 - 1) < 1e-6


def test_batched_newton_and_arclength_lockstep(modules):
    """BASELINE configs 4-5 in miniature: B concurrent Newton solves with batched GPU JVPs converge to the states
    the reference-style host Newton (SciPy LGMRES) finds; one lock-step pseudo-arc-length step lands on the branch."""
    from oracle import sddc_oracle as orc
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    from spectraldoublediffusiveconvection_b200.krylov import arclength_batched, newton_batched
    MO, _ = modules
    N_fm, N_r, d, Pr, Tau, Ra_s, sym = 32, 16, 2.0, 10.0, 1.0, 0.0, True
    nr = N_r - 1
    Ras = np.array([6780.0, 6800.0, 6850.0, 6900.0])
    B = len(Ras)
    pt = EnsemblePlan(N_fm, N_r, d, 0.075, Pr, Tau, symmetric=sym, max_batch=B)
    X0 = np.random.default_rng(0).random(3 * nr * N_fm)
    X0 = 1e-3 * X0 / np.linalg.norm(X0) * orc.sym_mask(N_fm, nr).reshape(-1)
    Xs = pt.step(_rep(X0, B), _dev(Ras), Ra_s, nsteps=13000)          # four transients, one per Rayleigh number
    pt.close()
    pn = EnsemblePlan(N_fm, N_r, d, 1.0, Pr, Tau, symmetric=sym, max_batch=B)   # dt = 1 as in Main._Newton
    Xn, ninfo = newton_batched(pn, Xs, _dev(Ras), Ra_s, tol_newton=1e-8, max_it=7)
    assert bool(ninfo["converged"].all()), ninfo["history"]
    res = pn.residual(Xn, _dev(Ras), Ra_s)
    assert float(torch.linalg.vector_norm(res, dim=1).max()) < 1e-6
    Xn_h = Xn.cpu().numpy()
    for m in (0, 3):
        Xh, hh, _ = drv.newton(MO, Xs[m].cpu().numpy(), Ras[m], Ra_s, Tau, Pr, d, N_fm, N_r, sym, max_it=6)
        assert rel_l2(Xn_h[m], Xh) < 1e-6                       # same steady state as the host LGMRES Newton
    ke = pn.diagnostics(Xn).cpu().numpy()[:, 1]
    assert abs(ke[0] / 2.5752204992e-2 - 1) < 1e-6              # KAT of the reference's wide-gap test
    assert np.all(np.diff(ke) > 0)                              # kinetic energy grows with Ra along the branch
    # one pseudo-arc-length step from the converged points: secant-free start with tangent (0, 1) in (X, mu)
    n = Xn.shape[1]
    Xd0 = torch.zeros_like(Xn)
    mud0 = torch.ones(B, dtype=torch.float64, device="cuda")
    ds = torch.full((B,), 5.0, dtype=torch.float64, device="cuda")
    Xa, mua, Xd, mud, h2, nj2 = arclength_batched(pn, Xn, _dev(Ras), Xd0, mud0, ds, Ra_s, max_it=8)
    resa = pn.residual(Xa, mua, Ra_s)
    assert float(torch.linalg.vector_norm(resa, dim=1).max()) < 1e-6
    delta = 1.0 / n
    cons = delta * (Xd0 * (Xa - Xn)).sum(dim=1) + (1 - delta) * mud0 * (mua - _dev(Ras)) - ds
    assert float(cons.abs().max()) < 1e-6                       # arclength constraint (Main.py:910)
    nrm = torch.sqrt(delta * (Xd ** 2).sum(dim=1) + (1 - delta) * mud ** 2)
    assert torch.allclose(nrm, torch.ones_like(nrm), atol=1e-10) and bool((mud > 0).all())
    pn.close()


def _upsample(MO, X, K_o, N_r_o, K, N_r, d):
    return MO.INTERP_THETAS(K, K_o, MO.INTERP_RADIAL(N_r, N_r_o, X, d))


def test_newton_history_at_config4_size(modules):
    """BASELINE config 4 at its stated size: Main._Newton's iteration (host SciPy LGMRES, matrix-free JVPs) at
    N_r = 30, N_theta = 256, symmetric, l = 10 parameter set (Main.py:617-624), from the seeded branch state
    (golden/branch_seeds.npz, interpolated like Main.py:599-601).  The printed error history on the GPU operators must
    equal the one on the CPU oracle operators to 1e-10 (north_star)."""
    from oracle import sddc_oracle as orc
    MO, _ = modules
    sd = load_golden("branch_seeds")
    l, d, Ra_c, Ra_s, sym = sd["l10_params"]
    N_fm, N_r, Pr, Tau = 256, 30, 1.0, 1.0 / 15.0
    X0 = _upsample(MO, sd["l10_X"], int(sd["N_fm"]), int(sd["N_r"]), N_fm, N_r, d)
    Ra = float(sd["l10_Ra"]) - 2.0
    Xg, hg, ng = drv.newton(MO, X0, Ra, Ra_s, Tau, Pr, d, N_fm, N_r, bool(sym))
    orc.set_transform_backend("fft")
    orc.set_accel(True)
    try:
        Xc, hc, nc = drv.newton(drv.OracleOperators(orc, N_fm, N_r, d, 1.0, Pr, Tau), X0, Ra, Ra_s, Tau, Pr, d, N_fm, N_r,
                                bool(sym))
    finally:
        orc.set_accel(False)
        orc.set_transform_backend("dense")
    print("config-4 newton history gpu", hg, "cpu", hc, "matvecs", ng, nc)
    assert len(hg) == len(hc) and len(hg) < 5 and hg[-1] < 1e-8, (hg, hc)      # converged inside Main._Newton's 5 iterations
    assert ng == nc
    assert np.max(np.abs(hg - hc)) < 1e-10, (hg, hc)
    assert rel_l2(Xg, Xc) < 1e-9


def test_arclength_history_at_config5_size(modules):
    """BASELINE config 5 at its stated size: one branch point's (err_X, err_mu) history (Main.py:929) of Main._ContinC's
    step at N_r = 40, N_theta = 512, l = 11 parameter set (Main.py:609-614), GPU operators against CPU oracle operators."""
    from oracle import sddc_oracle as orc
    MO, _ = modules
    sd = load_golden("branch_seeds")
    l, d, Ra_c, Ra_s, sym = sd["l11_params"]
    N_fm, N_r, Pr, Tau = 512, 40, 1.0, 1.0 / 15.0
    X0 = _upsample(MO, sd["l11_X"], int(sd["N_fm"]), int(sd["N_r"]), N_fm, N_r, d)
    Ra = float(sd["l11_Ra"])
    # polish on the new grid first (the interpolated state is only close to the discrete branch)
    X0, h0, _ = drv.newton(MO, X0, Ra, Ra_s, Tau, Pr, d, N_fm, N_r, bool(sym), max_it=8)
    assert h0[-1] < 1e-8, h0
    Y0 = np.hstack((X0, Ra))
    Yg, dsg, hg, ng = drv.continc(MO, Y0, float(sd["l11_sign"]), 0.5, Ra_s, Tau, Pr, d, N_fm, N_r, bool(sym))
    orc.set_transform_backend("fft")
    orc.set_accel(True)
    try:
        Yc, dsc, hc, nc = drv.continc(drv.OracleOperators(orc, N_fm, N_r, d, 1.0, Pr, Tau), Y0, float(sd["l11_sign"]), 0.5,
                                      Ra_s, Tau, Pr, d, N_fm, N_r, bool(sym))
    finally:
        orc.set_accel(False)
        orc.set_transform_backend("dense")
    print("config-5 arc-length history gpu", hg.tolist(), "cpu", hc.tolist(), "matvecs", ng, nc, "ds", dsg, dsc)
    assert hg.shape == hc.shape and dsg == dsc
    assert np.max(np.abs(hg - hc)) < 1e-10, (hg, hc)
    assert abs(Yg[-1] - Yc[-1]) < 1e-9 * abs(Yc[-1]) and rel_l2(Yg[:-1], Yc[:-1]) < 1e-9

"""Pin oracle/sddc_oracle.py against vectors produced by the unmodified reference
(tests/golden/make_golden.py) and against the reference's transform known-answer tests
(Transforms.py:132-374).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from oracle import sddc_oracle as orc

CASES = ["small_nosym", "small_sym", "cfg1_nosym", "cfg1_sym", "cfg3_member"]
TOL_CALL = 5e-13   # single operator call, relative L2
TOL_STEPS = 1e-10  # north_star tolerance after 100 steps


def _ops(g):
    return orc.Operators(int(g["N_fm"]), int(g["N_r"]), float(g["d"]), float(g["dt"]), float(g["Pr"]), float(g["Tau"]))


@pytest.fixture(scope="module", params=CASES)
def case(request):
    g = load_golden(request.param)
    return g, _ops(g)


@pytest.fixture(params=["dense", "fft"])
def backend(request):
    orc.set_transform_backend(request.param)
    yield request.param
    orc.set_transform_backend("dense")


def test_fft_backend_step_matches_golden(backend):
    g = load_golden("cfg1_nosym")
    op = _ops(g)
    assert rel_l2(orc.NLIN_FX(g["Xb"], op, False), g["NLIN_FX"]) < TOL_CALL
    assert rel_l2(orc.step(g["Xb"], op, float(g["Ra"]), float(g["Ra_s"]), False), g["step_Xb"]) < 1e-10
    assert abs(orc.kinetic_energy(g["Xb"], op, False) / float(g["KE_Xb"]) - 1) < 1e-12


def test_transforms_match_reference(backend):
    g = load_golden("transforms")
    for K in (16, 48):
        M = 3 * K // 2
        a, gr = g["in_hat_%d" % K], g["in_grid_%d" % K]
        assert rel_l2(orc.grid(M), g["grid_%d" % K]) < 1e-15
        assert rel_l2(orc.IDCT(a, n=M), g["IDCT_%d" % K]) < 1e-13
        assert rel_l2(orc.IDST(a, n=M), g["IDST_%d" % K]) < 1e-13
        assert rel_l2(orc.IDCT(a), g["IDCT_same_%d" % K]) < 1e-13
        assert rel_l2(orc.IDST(a), g["IDST_same_%d" % K]) < 1e-13
        assert rel_l2(orc.IDCT(a, n=K // 2), g["IDCT_half_%d" % K]) < 1e-13
        assert rel_l2(orc.IDST(a, n=K // 2), g["IDST_half_%d" % K]) < 1e-13
        assert rel_l2(orc.IDCT(a, n=3 * K), g["IDCT_3x_%d" % K]) < 1e-13
        assert rel_l2(orc.IDST(a, n=3 * K), g["IDST_3x_%d" % K]) < 1e-13
        assert rel_l2(orc.DCT(gr), g["DCT_%d" % K]) < 1e-13
        assert rel_l2(orc.DST(gr), g["DST_%d" % K]) < 1e-13
        assert rel_l2(orc.DCT(gr, n=K), g["DCT_trunc_%d" % K]) < 1e-13
        assert rel_l2(orc.DST(gr, n=K), g["DST_trunc_%d" % K]) < 1e-13


@pytest.mark.parametrize("k", [0, 1, 5, 100, 255])
def test_transform_known_answers(k):
    """Single-mode known answers of Transforms.test_Cosine_Transform / test_Sine_Transform (N=256)."""
    N = 256
    x = orc.grid(N)
    e = np.zeros(N)
    e[k] = 1.0
    assert np.allclose(orc.IDCT(e), np.cos(k * x), atol=1e-12)
    assert np.allclose(orc.DCT(np.cos(k * x)), e, atol=1e-12)
    if k > 0:
        assert np.allclose(orc.IDST(e), np.sin(k * x), atol=1e-12)
        assert np.allclose(orc.DST(np.sin(k * x)), e, atol=1e-12)


def test_transform_products_dealiased():
    """Product identities of Transforms.test_*_NL through the 3/2-padded path."""
    N = 256
    M = 3 * N // 2
    k1, k2 = 40, 75
    a = np.zeros(N); a[k1] = 1.0
    b = np.zeros(N); b[k2] = 1.0
    cc = orc.DCT(orc.IDCT(a, n=M) * orc.IDCT(b, n=M), n=N)
    exp = np.zeros(N); exp[k2 - k1] += 0.5; exp[k1 + k2] += 0.5
    assert np.allclose(cc, exp, atol=1e-12)
    ss = orc.DCT(orc.IDST(a, n=M) * orc.IDST(b, n=M), n=N)
    exp = np.zeros(N); exp[k2 - k1] += 0.5; exp[k1 + k2] -= 0.5
    assert np.allclose(ss, exp, atol=1e-12)
    sc = orc.DST(orc.IDST(a, n=M) * orc.IDCT(b, n=M), n=N)
    exp = np.zeros(N); exp[k1 + k2] += 0.5; exp[k2 - k1] -= 0.5
    assert np.allclose(sc, exp, atol=1e-12)


def test_operator_build_matches_reference():
    g = load_golden("small_nosym")
    op = _ops(g)
    assert rel_l2(op.D, g["op_D"]) < 1e-14
    assert rel_l2(op.R, g["op_R"]) < 1e-15
    assert rel_l2(op.D2, g["op_D2"]) < 1e-13
    assert rel_l2(op.IR2, g["op_IR2"]) < 1e-15
    assert rel_l2(op.IR4, g["op_IR4"]) < 1e-15
    assert rel_l2(op.dT0, g["op_DT0"]) < 1e-15
    # inverses of ill-conditioned matrices: compare loosely entrywise, tightly through their action below
    assert rel_l2(op.Linv_A4, g["op_L4"]) < 1e-8
    assert rel_l2(op.Linv_T, g["op_LT"]) < 1e-10
    assert rel_l2(op.Linv_S, g["op_LS"]) < 1e-10


def test_linear_pieces(case):
    g, op = case
    K, n, sym = op.K, op.n, bool(g["symmetric"])
    Xb = g["Xb"].reshape(3, K, n)
    assert rel_l2(orc.J_theta_RT(Xb[0], sym), g["J_theta_RT"]) < TOL_CALL
    assert rel_l2(orc.DT0_theta(Xb[0], op.dT0, sym), g["DT0_theta"]) < TOL_CALL
    assert rel_l2(orc.A2_SINE(Xb[0], op, sym), g["A2_SINE"]) < TOL_CALL
    assert rel_l2(orc.A2_SINE_R2(Xb[0], op, sym), g["A2_SINE_R2"]) < TOL_CALL
    assert rel_l2(orc.buoyancy(Xb[1], op), g["kGR"]) < TOL_CALL
    assert rel_l2((op.r ** 2)[None, :] * Xb[1], g["R2"]) < TOL_CALL


def test_nonlinear_term_and_jvp(case):
    g, op = case
    sym = bool(g["symmetric"])
    assert rel_l2(orc.NLIN_FX(g["Xb"], op, sym), g["NLIN_FX"]) < TOL_CALL
    assert rel_l2(orc.NLIN_DFX(g["dv"], g["Xb"], op, sym), g["NLIN_DFX"]) < TOL_CALL


def test_solves(case):
    g, op = case
    K, n, sym = op.K, op.n, bool(g["symmetric"])
    Xb = g["Xb"].reshape(3, K, n)
    # cond(L) reaches ~1e6 for the fourth-order operator
    assert rel_l2(orc.A4_BSub(Xb[0], op.Linv_A4, op, op.Pr * op.dt, sym), g["A4_BSub"]) < 1e-9
    assert rel_l2(orc.NAB2_BSub(Xb[1], op.Linv_T, op.dt, sym), g["NAB2_BSub_T"]) < 1e-11
    assert rel_l2(orc.NAB2_BSub(Xb[2], op.Linv_S, op.Tau * op.dt, sym), g["NAB2_BSub_S"]) < 1e-11


def test_step_jvp_dmu(case):
    g, op = case
    sym = bool(g["symmetric"])
    Ra, Ra_s = float(g["Ra"]), float(g["Ra_s"])
    assert rel_l2(orc.step(g["Xb"], op, Ra, Ra_s, sym), g["step_Xb"]) < 1e-10
    assert rel_l2(orc.jvp(g["dv"], g["Xb"], op, Ra, Ra_s, sym), g["jvp_Xb"]) < 1e-10
    assert rel_l2(orc.dF_dRa(g["Xb"], op, sym), g["dmu_Xb"]) < 1e-10


def test_diagnostics(case):
    g, op = case
    K, n, sym = op.K, op.n, bool(g["symmetric"])
    Xb = g["Xb"]
    Xs = Xb * orc.sym_mask(K, n).reshape(-1) if sym else Xb
    assert abs(orc.kinetic_energy(Xs, op, sym) / float(g["KE_Xb"]) - 1) < 1e-12
    X3 = Xb.reshape(3, K, n)
    assert abs(orc.nusselt(X3[1], op) / float(g["NuT_Xb"]) - 1) < 1e-12
    assert abs(orc.nusselt(X3[2], op) / float(g["NuS_Xb"]) - 1) < 1e-12


def test_time_stepping_parity(case):
    """The north-star bar: relative L2 <= 1e-10 after the golden run's n_steps (100 for the BASELINE configs)."""
    g, op = case
    sym = bool(g["symmetric"])
    Ra, Ra_s = float(g["Ra"]), float(g["Ra_s"])
    n_steps = int(g["n_steps"])
    X = g["X0"].copy()
    mask = orc.sym_mask(op.K, op.n).reshape(-1) if sym else 1.0
    hist = []
    for it in range(n_steps):
        Xn = orc.step(X, op, Ra, Ra_s, sym)
        if it + 1 in (1, 10):
            assert rel_l2(Xn, g["X_step%d" % (it + 1)]) < TOL_STEPS
        if it % max(1, n_steps // 5) == 0:
            hist.append((it, orc.diagnostics(Xn, op, sym)))
        X = mask * Xn
    assert rel_l2(Xn, g["X_step%d" % n_steps]) < TOL_STEPS
    for it, dg in hist:
        assert np.allclose(dg, g["diag_hist"][it], rtol=1e-9, atol=0)


def test_accelerated_solves_equal_plain_loops():
    """The numba-compiled loops used for baseline timing do the same arithmetic as the plain NumPy loops."""
    pytest.importorskip("numba")
    g = load_golden("small_nosym")
    op = _ops(g)
    Xb = g["Xb"].reshape(3, op.K, op.n)
    for sym in (False, True):
        a = orc.A4_BSub(Xb[0], op.Linv_A4, op, op.Pr * op.dt, sym)
        b = orc.NAB2_BSub(Xb[1], op.Linv_T, op.dt, sym)
        orc.set_accel(True)
        try:
            a2 = orc.A4_BSub(Xb[0], op.Linv_A4, op, op.Pr * op.dt, sym)
            b2 = orc.NAB2_BSub(Xb[1], op.Linv_T, op.dt, sym)
        finally:
            orc.set_accel(False)
        assert rel_l2(a2, a) < 1e-13 and rel_l2(b2, b) < 1e-14


@pytest.mark.parametrize("name", ["small_nosym", "small_sym", "cfg1_nosym", "cfg1_sym"])
def test_solve_based_second_oracle(name):
    """The reference has two implementations of each implicit solve: pre-inverted stacks (V2, what the GPU path
    follows) and per-mode factorisations (Matrix_Operators.py:248-433).  solve.npz holds the latter's outputs on the
    fixtures' right-hand sides (tests/golden/make_golden_solve.py); both restatements must reproduce both."""
    g = load_golden(name)
    s2 = load_golden("solve")
    op = _ops(g)
    K, n, N = op.K, op.n, op.K * op.n
    sym = bool(g["symmetric"])
    Xb = g["Xb"]
    psi, T, S = (Xb[i * N:(i + 1) * N].reshape(K, n) for i in range(3))
    dt, Pr, Tau = float(g["dt"]), float(g["Pr"]), float(g["Tau"])
    a4 = orc.A4_BSub_solve(psi, op, Pr * dt, sym).ravel()
    t = orc.NAB2_BSub_solve(T, op, dt, sym).ravel()
    sf = orc.NAB2_BSub_solve(S, op, Tau * dt, sym).ravel()
    assert rel_l2(a4, s2[name + "_A4"]) < 1e-9        # cond(L_j) ~ 1e6: factorisation order differs from LAPACK's inside numba
    assert rel_l2(t, s2[name + "_T"]) < 1e-12
    assert rel_l2(sf, s2[name + "_S"]) < 1e-12
    # and the pre-inverted restatement agrees with the independent solve-based reference output
    assert rel_l2(orc.A4_BSub(psi, op.Linv_A4, op, Pr * dt, sym).ravel(), s2[name + "_A4"]) < 1e-9
    assert rel_l2(orc.NAB2_BSub(T, op.Linv_T, dt, sym).ravel(), s2[name + "_T"]) < 1e-12
    assert rel_l2(orc.NAB2_BSub(S, op.Linv_S, Tau * dt, sym).ravel(), s2[name + "_S"]) < 1e-12

"""GPU parity: the CUDA path (through the C ABI) against golden vectors of the unmodified reference and against
the NumPy oracle on the same seeded inputs.  Tolerances: fp64, relative L2 <= 1e-10 after 100 steps
(BASELINE.json north_star); single operator calls are held to much tighter bounds."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CASES = ["small_nosym", "small_sym", "cfg1_nosym", "cfg1_sym", "cfg3_member"]
TOL_CALL = 2e-12
TOL_STEPS = 1e-10


def _plan(g, max_batch=4):
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    return EnsemblePlan(int(g["N_fm"]), int(g["N_r"]), float(g["d"]), float(g["dt"]), float(g["Pr"]), float(g["Tau"]),
                        symmetric=bool(g["symmetric"]), max_batch=max_batch)


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()


@pytest.fixture(scope="module", params=CASES)
def case(request):
    g = load_golden(request.param)
    g["_name"] = request.param
    pl = _plan(g)
    yield g, pl
    pl.close()


def test_linear_ops(case):
    from spectraldoublediffusiveconvection_b200 import plan as P
    g, pl = case
    N = pl.N
    Xb = g["Xb"]
    psi, T = _dev(Xb[0:N]), _dev(Xb[N:2 * N])
    for op, key, src in ((P.OP_J_THETA, "J_theta_RT", psi), (P.OP_DT0_THETA, "DT0_theta", psi),
                         (P.OP_A2_SINE, "A2_SINE", psi), (P.OP_A2_SINE_R2, "A2_SINE_R2", psi),
                         (P.OP_KGR, "kGR", T), (P.OP_R2, "R2", T)):
        out = pl.linear_op(op, src).cpu().numpy().ravel()
        assert rel_l2(out, g[key]) < TOL_CALL, key


def test_nlin_fx(case):
    g, pl = case
    F = pl.nlin_fx(_dev(g["Xb"])).cpu().numpy().ravel()
    assert rel_l2(F, g["NLIN_FX"]) < TOL_CALL


def test_nlin_dfx(case):
    g, pl = case
    F = pl.nlin_dfx(_dev(g["dv"]), _dev(g["Xb"])).cpu().numpy().ravel()
    assert rel_l2(F, g["NLIN_DFX"]) < TOL_CALL


def test_solves(case):
    g, pl = case
    N = pl.N
    Xb = g["Xb"]
    f = pl.solve_a4(_dev(Xb[0:N])).cpu().numpy().ravel()
    assert rel_l2(f, g["A4_BSub"]) < 1e-9          # cond(L) ~ 1e6
    f = pl.solve_nab2(_dev(Xb[N:2 * N]), 0).cpu().numpy().ravel()
    assert rel_l2(f, g["NAB2_BSub_T"]) < 1e-11
    f = pl.solve_nab2(_dev(Xb[2 * N:3 * N]), 1).cpu().numpy().ravel()
    assert rel_l2(f, g["NAB2_BSub_S"]) < 1e-11


def test_solves_against_solve_based_oracle(case):
    """Second oracle: the reference's per-mode factorisations (Matrix_Operators.py:248-433, golden/solve.npz)."""
    g, pl = case
    name = g["_name"]
    s2 = load_golden("solve")
    if name + "_A4" not in s2:
        pytest.skip("no solve-based vectors for this case")
    N = pl.N
    Xb = g["Xb"]
    assert rel_l2(pl.solve_a4(_dev(Xb[0:N])).cpu().numpy().ravel(), s2[name + "_A4"]) < 1e-9
    assert rel_l2(pl.solve_nab2(_dev(Xb[N:2 * N]), 0).cpu().numpy().ravel(), s2[name + "_T"]) < 1e-11
    assert rel_l2(pl.solve_nab2(_dev(Xb[2 * N:3 * N]), 1).cpu().numpy().ravel(), s2[name + "_S"]) < 1e-11


def test_step_jvp_dmu(case):
    g, pl = case
    Ra, Ra_s = float(g["Ra"]), float(g["Ra_s"])
    Xb, dv = _dev(g["Xb"]), _dev(g["dv"])
    assert rel_l2(pl.step(Xb, Ra, Ra_s).cpu().numpy().ravel(), g["step_Xb"]) < 1e-10
    assert rel_l2(pl.jvp(dv, Xb, Ra, Ra_s).cpu().numpy().ravel(), g["jvp_Xb"]) < 1e-10
    assert rel_l2(pl.dF_dRa(Xb).cpu().numpy().ravel(), g["dmu_Xb"]) < 1e-10
    # PFX (Main.py:473-496): Step(X) - X.  Under the equatorial symmetry the solves leave the skipped blocks zero, so
    # the residual there is -X (the caller masks X beforehand, Main.py:525); no escape for the symmetric cases.
    res = pl.residual(Xb, Ra, Ra_s).cpu().numpy().ravel()
    assert rel_l2(res, g["step_Xb"] - g["Xb"]) < 1e-10


def test_diagnostics(case):
    from oracle import sddc_oracle as orc
    g, pl = case
    Xb = g["Xb"]
    if bool(g["symmetric"]):
        Xb = Xb * orc.sym_mask(pl.N_fm, pl.nr).reshape(-1)
    d = pl.diagnostics(_dev(Xb)).cpu().numpy()[0]
    assert abs(d[0] / np.linalg.norm(Xb) - 1) < 1e-13
    assert abs(d[1] / float(g["KE_Xb"]) - 1) < 1e-11
    assert abs(d[2] / float(g["NuT_Xb"]) - 1) < 1e-11
    assert abs(d[3] / float(g["NuS_Xb"]) - 1) < 1e-11


def test_time_stepping_parity(case):
    """relative L2 <= 1e-10 after the golden run's steps (100 for the BASELINE shapes), history of diagnostics too."""
    g, pl = case
    Ra, Ra_s = float(g["Ra"]), float(g["Ra_s"])
    n_steps = int(g["n_steps"])
    X = _dev(g["X0"]).reshape(1, -1)
    hist = g["diag_hist"]
    for it in range(n_steps):
        Xn = pl.step(X, Ra, Ra_s)
        if it + 1 in (1, 10):
            assert rel_l2(Xn.cpu().numpy().ravel(), g["X_step%d" % (it + 1)]) < TOL_STEPS
        if it % max(1, n_steps // 5) == 0:
            d = pl.diagnostics(Xn).cpu().numpy()[0]
            assert np.allclose(d[:4], hist[it], rtol=1e-9, atol=0), (it, d[:4], hist[it])
        X = Xn  # the solve already returns the masked state when symmetric
    assert rel_l2(X.cpu().numpy().ravel(), g["X_step%d" % n_steps]) < TOL_STEPS


def test_multistep_call_equals_single_steps(case):
    g, pl = case
    Ra, Ra_s = float(g["Ra"]), float(g["Ra_s"])
    X0 = _dev(g["X0"]).reshape(1, -1)
    X = X0
    for _ in range(5):
        X = pl.step(X, Ra, Ra_s)
    X5 = pl.step(X0, Ra, Ra_s, nsteps=5)
    assert torch.equal(X, X5)


def test_batch_members_independent_and_match_oracle():
    """4 members with different Rayleigh numbers in one launch == 4 oracle runs; members do not interact."""
    from oracle import sddc_oracle as orc
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    K, N_r, d, dt, Pr, Tau = 32, 14, 0.5, 5e-3, 1.0, 0.5
    pl = EnsemblePlan(K, N_r, d, dt, Pr, Tau, symmetric=False, max_batch=19)
    op = orc.Operators(K, N_r, d, dt, Pr, Tau)
    B = 19                                   # ragged: not a multiple of the 16-member solve tile
    rng = np.random.default_rng(5)
    X = rng.random((B, 3 * pl.N)) * 1e-1
    Ra = np.linspace(2000.0, 6000.0, B)
    Ra_s = np.linspace(0.0, 500.0, B)
    out = pl.step(_dev(X), _dev(Ra), _dev(Ra_s), nsteps=3).cpu().numpy()
    for m in (0, 7, 15, 16, 18):
        ref = X[m]
        for _ in range(3):
            ref = orc.step(ref, op, Ra[m], Ra_s[m])
        assert rel_l2(out[m], ref) < 1e-10
    # host-buffer entry point gives the same bits as the device entry point
    out_h, diag = pl.step_host(X, Ra, Ra_s, nsteps=3, want_diag=True)
    assert np.array_equal(out_h, out)
    dg = orc.diagnostics(out[3], op)
    assert np.allclose(diag[3, :4], dg, rtol=1e-10)
    pl.close()


def test_transforms_shim():
    from spectraldoublediffusiveconvection_b200 import plan as P
    g = load_golden("transforms")
    for K in (16, 48):
        M = 3 * K // 2
        a, gr = _dev(g["in_hat_%d" % K]), _dev(g["in_grid_%d" % K])
        assert rel_l2(P.transform(P.T_IDCT, a, M).cpu().numpy(), g["IDCT_%d" % K]) < 1e-13
        assert rel_l2(P.transform(P.T_IDST, a, M).cpu().numpy(), g["IDST_%d" % K]) < 1e-13
        assert rel_l2(P.transform(P.T_IDCT, a).cpu().numpy(), g["IDCT_same_%d" % K]) < 1e-13
        assert rel_l2(P.transform(P.T_IDST, a, 3 * K).cpu().numpy(), g["IDST_3x_%d" % K]) < 1e-13
        assert rel_l2(P.transform(P.T_IDCT, a, K // 2).cpu().numpy(), g["IDCT_half_%d" % K]) < 1e-13   # truncating
        assert rel_l2(P.transform(P.T_IDST, a, K // 2).cpu().numpy(), g["IDST_half_%d" % K]) < 1e-13
        assert rel_l2(P.transform(P.T_DCT, gr).cpu().numpy(), g["DCT_%d" % K]) < 1e-13
        assert rel_l2(P.transform(P.T_DST, gr).cpu().numpy(), g["DST_%d" % K]) < 1e-13
        assert rel_l2(P.transform(P.T_DCT, gr, K).cpu().numpy(), g["DCT_trunc_%d" % K]) < 1e-13
        assert rel_l2(P.transform(P.T_DST, gr, K).cpu().numpy(), g["DST_trunc_%d" % K]) < 1e-13


def test_error_behaviour():
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    with pytest.raises(ValueError):
        EnsemblePlan(17, 10, 0.4, 1e-2, 1.0, 1.0)      # odd N_fm: reference raises ValueError too
    pl = EnsemblePlan(16, 10, 0.4, 1e-2, 1.0, 1.0, max_batch=2)
    with pytest.raises(ValueError):
        pl.nlin_fx(torch.zeros((3, 3 * pl.N), dtype=torch.float64, device="cuda"))   # B > max_batch
    X = torch.zeros((2, 3 * pl.N), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        pl.step(X, torch.zeros(3, dtype=torch.float64, device="cuda"), 0.0)          # Ra of the wrong length
    with pytest.raises(ValueError):
        pl.step(X, 3000.0, 0.0, out=torch.zeros((2, 3 * pl.N + 1), dtype=torch.float64, device="cuda"))   # too large
    with pytest.raises(ValueError):
        pl.step(X, 3000.0, 0.0, out=torch.zeros((3 * pl.N, 2), dtype=torch.float64, device="cuda").t())   # strided
    with pytest.raises(TypeError):
        pl.step(X, 3000.0, 0.0, out=torch.zeros((2, 3 * pl.N), dtype=torch.float32, device="cuda"))
    pl.close()


@pytest.mark.parametrize("K,N_r,sym", [(128, 20, False), (512, 40, False), (64, 33, True), (40, 50, False),
                                        (24, 65, False), (8, 18, True), (128, 26, True), (256, 12, True),
                                        (512, 9, True), (256, 50, False),
                                        (50, 12, False), (30, 10, True), (42, 45, False), (10, 8, False)])
def test_other_shapes_against_oracle(K, N_r, sym):
    """Every kernel instantiation (radial tile counts 3..8, BASELINE configs 2 and 5 shapes, ragged mode counts):
    one step, one JVP, diagnostics and the nonlinear term against the NumPy oracle on seeded inputs."""
    from oracle import sddc_oracle as orc
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    d, dt, Pr, Tau = 0.353, 2e-3, 1.0, 1.0 / 15.0
    pl = EnsemblePlan(K, N_r, d, dt, Pr, Tau, symmetric=sym, max_batch=3)
    op = orc.Operators(K, N_r, d, dt, Pr, Tau)
    rng = np.random.default_rng(K + N_r)
    B = 3
    X = rng.random((B, 3 * pl.N)) * 1e-2
    dv = rng.standard_normal((B, 3 * pl.N))
    Ra, Ra_s = np.array([3000.0, 4000.0, 9000.0]), np.array([0.0, 300.0, 500.0])
    Xd, dvd = _dev(X), _dev(dv)
    F = pl.nlin_fx(Xd).cpu().numpy()
    st = pl.step(Xd, _dev(Ra), _dev(Ra_s)).cpu().numpy()
    # every shape has its two-state products: FFT kernels (N_fm = 128, 256, 512), dense two-state synthesis (N_r <= 41), or
    # the direct-summation row kernel (N_r > 41 elsewhere, and everything at N_fm = 2 mod 4, whose padded grid is odd)
    info = pl.info()
    assert info["direct_rows"] == (1 if K % 4 else (2 if (N_r > 41 and K not in (128, 256, 512)) else 0)), info
    has_jvp = True
    jv = pl.jvp(dvd, Xd, _dev(Ra), _dev(Ra_s)).cpu().numpy()
    Fd = pl.nlin_dfx(dvd, Xd).cpu().numpy()
    pl.jvp_set_base(Xd)                                     # cached base state: grid fields (FFT kernels) or spectral rows
    jc = pl.jvp_apply(dvd, _dev(Ra), _dev(Ra_s)).cpu().numpy()
    for m in range(B):
        assert rel_l2(Fd[m], orc.NLIN_DFX(dv[m], X[m], op, sym)) < 1e-11
        assert rel_l2(jc[m], jv[m]) < 1e-12
    dg = pl.diagnostics(Xd).cpu().numpy()
    mask = orc.sym_mask(K, N_r - 1).reshape(-1) if sym else 1.0
    for m in range(B):
        assert rel_l2(F[m], orc.NLIN_FX(X[m], op, sym)) < 1e-11
        assert rel_l2(st[m], orc.step(X[m], op, Ra[m], Ra_s[m], sym)) < 1e-9
        if has_jvp:
            assert rel_l2(jv[m], orc.jvp(dv[m], X[m], op, Ra[m], Ra_s[m], sym)) < 1e-9
        ref = orc.diagnostics(X[m] * mask, op, sym) if sym else orc.diagnostics(X[m], op, sym)
        got = pl.diagnostics(_dev(X[m] * mask)).cpu().numpy()[0] if sym else dg[m]
        assert np.allclose(got[:4], ref, rtol=1e-10)
    pl.close()


def test_time_step_host_matches_device_loop():
    """sddc_time_step_host (host buffers, async diagnostics / checkpoints) == explicit device loop, bit for bit."""
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    K, N_r, d, dt, Pr, Tau = 32, 20, 0.5, 5e-3, 1.0, 0.5
    B, nsteps = 5, 12
    pl = EnsemblePlan(K, N_r, d, dt, Pr, Tau, max_batch=B)
    rng = np.random.default_rng(9)
    X = rng.random((B, 3 * pl.N)) * 1e-2
    Ra, Ra_s = np.linspace(2000.0, 4000.0, B), np.linspace(0.0, 200.0, B)
    out, hist, ck = pl.time_step_host(X, Ra, Ra_s, nsteps, diag_every=2, ckpt_every=4)
    assert hist.shape == (6, B, 6) and ck.shape == (3, B, 3 * pl.N)
    cur = _dev(X)
    for s in range(1, nsteps + 1):
        cur = pl.step(cur, _dev(Ra), _dev(Ra_s))
        if s % 2 == 0:
            assert np.array_equal(hist[s // 2 - 1], pl.diagnostics(cur).cpu().numpy())
        if s % 4 == 0:
            assert np.array_equal(ck[s // 4 - 1], cur.cpu().numpy())
    assert np.array_equal(out, cur.cpu().numpy())
    pl.close()


def test_time_step_host_reference_checkpoint_cadence():
    """ckpt_first=1: checkpoints after the steps 1, 1 + N_save, ... as Main._Time_Step takes them (Main.py:301-303); with
    the golden config-1 run: first checkpoint = X_step1, and the run's last state = X_step100."""
    g = load_golden("cfg1_nosym")
    pl = _plan(g, max_batch=2)
    X0 = np.stack([g["X0"], g["X0"]])
    out, hist, ck = pl.time_step_host(X0, float(g["Ra"]), float(g["Ra_s"]), 100, diag_every=1, ckpt_every=10, ckpt_first=1)
    assert ck.shape == (10, 2, 3 * pl.N)                                   # steps 1, 11, ..., 91
    assert rel_l2(ck[0, 0], g["X_step1"]) < 1e-12 and rel_l2(ck[1, 1], load_golden("cfg1_nosym")["X_step10"]) > 1e-6
    assert rel_l2(out[0], g["X_step100"]) < TOL_STEPS
    assert np.allclose(hist[:, 0, :4], g["diag_hist"], rtol=1e-9, atol=0)
    cur = pl.step(_dev(X0), float(g["Ra"]), float(g["Ra_s"]), nsteps=11).cpu().numpy()
    assert np.array_equal(ck[1], cur)
    pl.close()


@pytest.mark.parametrize("sym", [False, True])
def test_time_step_host_shares_prep_with_diagnostics(sym):
    """FFT formulation: inside sddc_time_step_host the kinetic energy of X_s comes from the spectral rows that the prep
    stage of step s+1 produces anyway; the history must equal the stand-alone diagnostics to rounding, the states
    bit for bit."""
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    K, N_r, d, dt, Pr, Tau = 128, 14, 0.5, 5e-3, 1.0, 0.5
    B, nsteps = 3, 7
    pl = EnsemblePlan(K, N_r, d, dt, Pr, Tau, symmetric=sym, max_batch=B)
    assert pl.info()["fft_M"] == 192
    rng = np.random.default_rng(10)
    X = rng.random((B, 3 * pl.N)) * 1e-2
    Ra, Ra_s = np.linspace(2000.0, 4000.0, B), np.linspace(0.0, 200.0, B)
    out, hist, ck = pl.time_step_host(X, Ra, Ra_s, nsteps, diag_every=1, ckpt_every=3)
    cur = _dev(X)
    for s in range(1, nsteps + 1):
        cur = pl.step(cur, _dev(Ra), _dev(Ra_s))
        ref = pl.diagnostics(cur).cpu().numpy()
        assert np.allclose(hist[s - 1], ref, rtol=1e-11, atol=0.0), s
        if s % 3 == 0:
            assert np.array_equal(ck[s // 3 - 1], cur.cpu().numpy())
    assert np.array_equal(out, cur.cpu().numpy())
    pl.close()


@pytest.mark.parametrize("K,N_r", [(128, 14), (32, 20)])
def test_device_time_step_matches_step_plus_diagnostics(K, N_r):
    """sddc_time_step (device resident, diagnostics history on the device) == step + diagnostics per step, for the FFT
    formulation (shared prep stage) and the dense one."""
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    B, nsteps = 3, 7
    pl = EnsemblePlan(K, N_r, 0.5, 5e-3, 1.0, 0.5, max_batch=B)
    rng = np.random.default_rng(11)
    X = _dev(rng.random((B, 3 * pl.N)) * 1e-2)
    Ra, Ra_s = _dev(np.linspace(2000.0, 4000.0, B)), _dev(np.linspace(0.0, 200.0, B))
    out, hist = pl.time_step(X, Ra, Ra_s, nsteps, diag_every=2)
    assert hist.shape == (3, B, 6)
    cur = X
    for s in range(1, nsteps + 1):
        cur = pl.step(cur, Ra, Ra_s)
        if s % 2 == 0:
            assert np.allclose(hist[s // 2 - 1].cpu().numpy(), pl.diagnostics(cur).cpu().numpy(), rtol=1e-11, atol=0.0), s
    assert torch.equal(out, cur)
    out0, hist0 = pl.time_step(X, Ra, Ra_s, 3, diag_every=0)
    assert hist0.shape[0] == 0 and torch.equal(out0, pl.step(X, Ra, Ra_s, nsteps=3))
    pl.close()


def test_step_is_cuda_graph_capturable():
    """The device entry points are allocation-free and stream-ordered: a member-step can be captured in a CUDA graph
    and replayed (include/sddc_b200.h contract)."""
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    K, N_r, d, dt, Pr, Tau = 32, 20, 0.5, 5e-3, 1.0, 0.5
    B = 4
    pl = EnsemblePlan(K, N_r, d, dt, Pr, Tau, max_batch=B)
    rng = np.random.default_rng(3)
    X = _dev(rng.random((B, 3 * pl.N)) * 1e-2)
    Ra = _dev(np.linspace(2000.0, 4000.0, B))
    Ras = _dev(np.zeros(B))
    ref = pl.step(pl.step(X, Ra, Ras), Ra, Ras)
    buf_in, buf_out = X.clone(), torch.empty_like(X)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        pl.step(buf_in, Ra, Ras, out=buf_out)            # warm-up on the side stream
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        pl.step(buf_in, Ra, Ras, out=buf_out)
    g.replay()
    buf_in.copy_(buf_out)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(buf_out, ref)
    pl.close()


def test_gap_sweep_groups_members_by_gap():
    """Shell-gap sweep: members with different d use different operator sets; results match per-member oracle runs."""
    from oracle import sddc_oracle as orc
    from spectraldoublediffusiveconvection_b200.ensemble import GapSweep
    K, N_r, dt, Pr, Tau = 16, 18, 5e-3, 1.0, 1.0
    d = np.array([0.3, 0.5, 0.3, 0.4, 0.5])
    Ra = np.linspace(2000.0, 4000.0, 5)
    gs = GapSweep(K, N_r, d, dt, Pr, Tau, Ra, 100.0)
    assert len(gs.groups) == 3
    rng = np.random.default_rng(8)
    X = rng.random((5, 3 * (N_r - 1) * K)) * 1e-2
    out = gs.step(_dev(X), nsteps=2).cpu().numpy()
    for m in range(5):
        op = orc.Operators(K, N_r, float(d[m]), dt, Pr, Tau)
        ref = orc.step(orc.step(X[m], op, Ra[m], 100.0), op, Ra[m], 100.0)
        assert rel_l2(out[m], ref) < 1e-10
    gs.close()


@pytest.mark.parametrize("name", ["cfg3_member", "cfg1_sym", "small_nosym"])
def test_cached_base_jvp_equals_jvp(name):
    """jvp_set_base + jvp_apply (one synthesis per product) == jvp (two syntheses), and matches the reference."""
    g = load_golden(name)
    pl = _plan(g)
    Ra, Ra_s = float(g["Ra"]), float(g["Ra_s"])
    Xb, dv = _dev(g["Xb"]), _dev(g["dv"])
    pl.jvp_set_base(Xb)
    j1 = pl.jvp_apply(dv, Ra, Ra_s).cpu().numpy().ravel()
    j2 = pl.jvp_apply(2.0 * dv, Ra, Ra_s).cpu().numpy().ravel()      # the cache survives several products
    assert rel_l2(j1, g["jvp_Xb"]) < 1e-10
    assert rel_l2(j2, 2.0 * g["jvp_Xb"]) < 1e-10
    assert rel_l2(j1, pl.jvp(dv, Xb, Ra, Ra_s).cpu().numpy().ravel()) < 1e-12
    if pl.has_jvp_plus:      # the linearised step itself: PDFX(dv) + dv, no subtrahend in the back-substitution
        jp = pl.jvp_apply(dv, Ra, Ra_s, plus_identity=True).cpu().numpy().ravel()
        assert rel_l2(jp, g["jvp_Xb"] + g["dv"]) < 1e-10
    pl.close()


def test_dense_transform_path_still_matches():
    """dense_transforms=True keeps the dense DMMA transforms (the path every N_fm other than 128/256/512 takes) at
    the headline shape; FFT and dense formulations of the same plan shape agree to rounding."""
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    g = load_golden("cfg3_member")
    pl_fft = _plan(g)
    assert pl_fft.info()["fft_M"] == 384 and pl_fft.info()["fft_jvp"] == 1
    pl = EnsemblePlan(int(g["N_fm"]), int(g["N_r"]), float(g["d"]), float(g["dt"]), float(g["Pr"]), float(g["Tau"]),
                      symmetric=bool(g["symmetric"]), max_batch=4, dense_transforms=True)
    assert pl.info()["fft_M"] == 0
    Ra, Ra_s = float(g["Ra"]), float(g["Ra_s"])
    Xb, dv = _dev(g["Xb"]), _dev(g["dv"])
    F = pl.nlin_fx(Xb).cpu().numpy().ravel()
    assert rel_l2(F, g["NLIN_FX"]) < TOL_CALL
    assert rel_l2(F, pl_fft.nlin_fx(Xb).cpu().numpy().ravel()) < TOL_CALL
    assert rel_l2(pl.nlin_dfx(dv, Xb).cpu().numpy().ravel(), g["NLIN_DFX"]) < TOL_CALL
    assert rel_l2(pl.step(Xb, Ra, Ra_s).cpu().numpy().ravel(), g["step_Xb"]) < TOL_CALL
    pl.jvp_set_base(Xb)
    assert rel_l2(pl.jvp_apply(dv, Ra, Ra_s).cpu().numpy().ravel(), g["jvp_Xb"]) < 1e-10
    pl.close()
    pl_fft.close()


def test_plans_of_different_shapes_coexist():
    """Several live plans (different N_r, same kernel instantiations) must not disturb each other's launches."""
    from oracle import sddc_oracle as orc
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    cfgs = [(32, 32, 0.4), (32, 26, 0.5), (32, 30, 0.45)]
    plans = [EnsemblePlan(K, N_r, d, 5e-3, 1.0, 0.5, max_batch=2) for (K, N_r, d) in cfgs]
    rng = np.random.default_rng(12)
    for _ in range(2):
        for (K, N_r, d), pl in zip(cfgs, plans):
            X = rng.random((2, 3 * pl.N)) * 1e-2
            out = pl.step(_dev(X), 3000.0, 100.0).cpu().numpy()
            op = orc.Operators(K, N_r, d, 5e-3, 1.0, 0.5)
            assert rel_l2(out[1], orc.step(X[1], op, 3000.0, 100.0)) < 1e-10
            dg = pl.diagnostics(_dev(X)).cpu().numpy()
            assert np.allclose(dg[0, :4], orc.diagnostics(X[0], op), rtol=1e-10)
    for pl in plans:
        pl.close()


def test_benchmarked_configuration_parity():
    """The configuration bench.py times -- B = 512 members per GPU at (N_fm, N_r) = (256, 30), one multi-step call, the
    8 / 16-member solve tiles, the padded solve-major slabs and the persistent grids -- against the reference's own
    100-step run (golden cfg3_member).  Members 0, 255 and 511 carry the golden member's IC and Rayleigh number, the
    others run a Ra sweep with random ICs next to them; 100 steps in ONE sddc_step call."""
    g = load_golden("cfg3_member")
    B = 512
    pl = _plan(g, max_batch=B)
    Ra0, Ras0 = float(g["Ra"]), float(g["Ra_s"])
    rng = np.random.default_rng(77)
    X0 = rng.random((B, 3 * pl.N))
    X0 *= 1e-3 / np.linalg.norm(X0, axis=1, keepdims=True)
    Ra = np.linspace(2000.0, 6000.0, B)
    Ras = np.full(B, Ras0)
    probes = (0, 255, 511)
    for m in probes:
        X0[m] = g["X0"]
        Ra[m] = Ra0
    out = pl.step(_dev(X0), _dev(Ra), _dev(Ras), nsteps=int(g["n_steps"]))
    outc = out.cpu().numpy()
    for m in probes:
        assert rel_l2(outc[m], g["X_step%d" % int(g["n_steps"])]) < TOL_STEPS, m
    assert torch.equal(out[0], out[511]) and torch.equal(out[0], out[255])   # members do not interact, tiles are uniform
    assert np.isfinite(outc).all()
    # one JVP and one residual at the same batch size, every member the golden pair (Xb, dv)
    Xb = _dev(np.tile(g["Xb"], (B, 1)))
    dv = _dev(np.tile(g["dv"], (B, 1)))
    jv = pl.jvp(dv, Xb, Ra0, Ras0)
    for m in probes:
        assert rel_l2(jv[m].cpu().numpy(), g["jvp_Xb"]) < TOL_STEPS, m
    assert torch.equal(jv[0], jv[511])
    pl.jvp_set_base(Xb)
    ja = pl.jvp_apply(dv, Ra0, Ras0)
    assert rel_l2(ja[300].cpu().numpy(), g["jvp_Xb"]) < TOL_STEPS
    res = pl.residual(Xb, Ra0, Ras0)
    assert rel_l2(res[511].cpu().numpy(), g["step_Xb"] - g["Xb"]) < TOL_STEPS
    dg = pl.diagnostics(Xb).cpu().numpy()
    assert abs(dg[257, 1] / float(g["KE_Xb"]) - 1) < 1e-11 and abs(dg[257, 2] / float(g["NuT_Xb"]) - 1) < 1e-11
    pl.close()


def test_config2_ensemble_parity():
    """BASELINE config 2 at its stated size: 1024 members at (N_fm, N_r) = (128, 20) on one GPU, member m started from
    default_rng(1000 + m) normalised to 1e-3, physical parameters of Main.Time_Step
    (Main.py:359-376); members {0, 511, 1023} against the oracle after 100 steps, <= 1e-10."""
    from oracle import sddc_oracle as orc
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    K, N_r, d, dt, Pr, Tau, Ra, Ra_s = 128, 20, 0.31325, 1e-3, 1.0, 1.0, 3750.0, 0.0
    B, nsteps = 1024, 100
    pl = EnsemblePlan(K, N_r, d, dt, Pr, Tau, symmetric=False, max_batch=B)
    X0 = np.empty((B, 3 * pl.N))
    for m in range(B):
        x = np.random.default_rng(1000 + m).random(3 * pl.N)
        X0[m] = 1e-3 * x / np.linalg.norm(x)
    out, hist = pl.time_step(_dev(X0), Ra, Ra_s, nsteps, diag_every=10)
    out, hist = out.cpu().numpy(), hist.cpu().numpy()
    op = orc.Operators(K, N_r, d, dt, Pr, Tau)
    for m in (0, 511, 1023):
        ref = X0[m]
        for s in range(nsteps):
            ref = orc.step(ref, op, Ra, Ra_s)
            if (s + 1) % 50 == 0:
                assert np.allclose(hist[(s + 1) // 10 - 1, m, :4], orc.diagnostics(ref, op), rtol=1e-9, atol=0)
        assert rel_l2(out[m], ref) < TOL_STEPS, m
    pl.close()


@pytest.mark.parametrize("K,N_r,sym,B", [(128, 12, False, 259), (256, 30, True, 261), (128, 26, False, 131), (256, 30, False, 261)])
def test_gather_mode_ragged_batch(K, N_r, sym, B):
    """A ragged batch of more than 128 members -- plain steps, a multi-step call, residual, JVP, cached JVP, the time loop
    with diagnostics -- equals the same members run in batches of 100, and the oracle on a few members (first / last of a
    member tile, last of the batch).  (In builds with -DSDDC_EXPERIMENTAL_GATHER the big batch goes through the gather
    mode of the back-substitution, k_solve_hot.cuh, the small ones through post_kernel.)"""
    from oracle import sddc_oracle as orc
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    d, dt, Pr, Tau = 0.353, 2e-3, 1.0, 1.0 / 15.0
    pl = EnsemblePlan(K, N_r, d, dt, Pr, Tau, symmetric=sym, max_batch=B)
    op = orc.Operators(K, N_r, d, dt, Pr, Tau)
    rng = np.random.default_rng(B)
    X = rng.random((B, 3 * pl.N)) * 1e-2
    dv = rng.standard_normal((B, 3 * pl.N))
    Ra, Ra_s = np.linspace(3000.0, 9000.0, B), np.linspace(0.0, 500.0, B)
    Xd, dvd, Rad, Rasd = _dev(X), _dev(dv), _dev(Ra), _dev(Ra_s)
    big = {"step": pl.step(Xd, Rad, Rasd), "step3": pl.step(Xd, Rad, Rasd, nsteps=3), "res": pl.residual(Xd, Rad, Rasd),
           "jvp": pl.jvp(dvd, Xd, Rad, Rasd)}
    pl.jvp_set_base(Xd)
    big["jvpc"] = pl.jvp_apply(dvd, Rad, Rasd)
    out_t, hist_t = pl.time_step(Xd, Rad, Rasd, 3, diag_every=1)
    big["loop"] = out_t
    small = {k: [] for k in big}
    hist_s = []
    for lo in range(0, B, 100):
        sl = slice(lo, min(B, lo + 100))
        small["step"].append(pl.step(Xd[sl], Rad[sl], Rasd[sl]))
        small["step3"].append(pl.step(Xd[sl], Rad[sl], Rasd[sl], nsteps=3))
        small["res"].append(pl.residual(Xd[sl], Rad[sl], Rasd[sl]))
        small["jvp"].append(pl.jvp(dvd[sl], Xd[sl], Rad[sl], Rasd[sl]))
        pl.jvp_set_base(Xd[sl].contiguous())
        small["jvpc"].append(pl.jvp_apply(dvd[sl].contiguous(), Rad[sl], Rasd[sl]))
        o, h = pl.time_step(Xd[sl].contiguous(), Rad[sl], Rasd[sl], 3, diag_every=1)
        small["loop"].append(o)
        hist_s.append(h)
    for k in big:
        a, b = big[k].cpu().numpy(), torch.cat(small[k]).cpu().numpy()
        for m in range(B):
            assert rel_l2(a[m], b[m]) < 1e-12, (k, m)
    assert np.allclose(hist_t.cpu().numpy(), torch.cat(hist_s, dim=1).cpu().numpy(), rtol=1e-10, atol=1e-300)
    st, jv = big["step"].cpu().numpy(), big["jvp"].cpu().numpy()
    for m in sorted({0, 7, 8, 15, 16, min(255, B - 2), min(256, B - 2), B - 1}):
        assert rel_l2(st[m], orc.step(X[m], op, Ra[m], Ra_s[m], sym)) < 1e-9
        assert rel_l2(jv[m], orc.jvp(dv[m], X[m], op, Ra[m], Ra_s[m], sym)) < 1e-9
    pl.close()


@pytest.mark.parametrize("sym", [False, True])
def test_step_is_bit_reproducible_run_to_run(sym):
    """The same 261-member input stepped twelve times (single steps and a three-step call, other calls in between for
    different timing): every run bit-identical.  A timing-dependent difference would be a race (this is the test the
    experimental gather mode of the back-substitution failed: tools/determinism_probe.py)."""
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    K, N_r, B = 256, 30, 261
    pl = EnsemblePlan(K, N_r, 0.353, 2e-3, 1.0, 1.0 / 15.0, symmetric=sym, max_batch=B)
    rng = np.random.default_rng(5)
    X = _dev(rng.random((B, 3 * pl.N)) * 1e-2)
    Ra, Ra_s = _dev(np.linspace(3000.0, 9000.0, B)), _dev(np.linspace(0.0, 500.0, B))
    ref1, ref3 = pl.step(X, Ra, Ra_s).clone(), pl.step(X, Ra, Ra_s, nsteps=3).clone()
    for r in range(12):
        assert torch.equal(pl.step(X, Ra, Ra_s), ref1), r
        pl.step(X[:100], Ra[:100], Ra_s[:100])
        assert torch.equal(pl.step(X, Ra, Ra_s, nsteps=3), ref3), r
    pl.close()

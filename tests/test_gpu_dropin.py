"""The drop-in Matrix_Operators / Transforms modules, driven with the reference's own call sequences, reproduce the
golden outputs of the unmodified reference (single member, NumPy in / NumPy out)."""
import sys

import numpy as np
import pytest

from conftest import load_golden, rel_l2
import dropin_drivers as drv

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


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

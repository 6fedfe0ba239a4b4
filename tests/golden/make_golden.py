"""Generate golden input/output vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference, numba, scipy

Imports /root/reference with the three shims of SURVEY.md section 8(c) (np.RankWarning alias, stub h5py,
stub matplotlib), calls the reference's own functions on seeded inputs and stores inputs + outputs as
small .npz fixtures next to this script.  The fixtures are what pins oracle/sddc_oracle.py and the CUDA
path; /root/reference itself is never needed at test time.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference(path="/root/reference"):
    if not hasattr(np, "RankWarning"):
        np.RankWarning = np.exceptions.RankWarning

    class _Group(dict):
        def __init__(self, *a, **k):
            super().__init__()

        def create_group(self, name):
            g = _Group()
            self[name] = g
            return g

        def create_dataset(self, name, data=None, **k):
            self[name] = data

        def close(self):
            pass

    h5 = types.ModuleType("h5py")
    h5.File = _Group
    sys.modules.setdefault("h5py", h5)
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    if path not in sys.path:
        sys.path.insert(0, path)
    import Main
    import Matrix_Operators
    import Transforms
    return Main, Matrix_Operators, Transforms


def ref_step_fn(Main, MO, ops, Ra, Ra_s, dt, Pr, Tau, symmetric):
    """The call sequence of Step_Python (Main.py:255-283) on the reference's own functions."""
    D, R, Rsq, DT0, gr_k, L4, LT, LS, args_A4, args_FX = ops
    N_fm, nr = args_FX[2], args_FX[3]
    N = N_fm * nr

    def step(Xn, linear=False):
        psi, T, S = Xn[0:N], Xn[N:2 * N], Xn[2 * N:3 * N]
        NX = -1 * dt * MO.NLIN_FX(Xn, *args_FX, symmetric) if not linear else np.zeros(3 * N)
        psi_T0 = MO.DT0_theta(psi, DT0, N_fm, nr, symmetric)
        Om = MO.A2_SINE(psi, D, R, N_fm, nr, symmetric)
        NX[0:N] += Om + dt * Pr * gr_k.dot(Ra * T - Ra_s * S)
        psi_new = MO.A4_BSub_TSTEP_V2(NX[0:N], L4, *args_A4, Pr * dt, symmetric)
        NX[N:2 * N] += Rsq.dot(T) - dt * psi_T0
        T_new = MO.NAB2_BSub_TSTEP_V2(NX[N:2 * N], LT, N_fm, nr, dt, symmetric)
        NX[2 * N:3 * N] += Rsq.dot(S) - dt * psi_T0
        S_new = MO.NAB2_BSub_TSTEP_V2(NX[2 * N:3 * N], LS, N_fm, nr, Tau * dt, symmetric)
        return np.hstack((psi_new, T_new, S_new))

    def jvp(dv, Xn):
        """PDFX (Main.py:498-521)."""
        dpsi, dT, dS = dv[0:N], dv[N:2 * N], dv[2 * N:3 * N]
        NX = -1. * dt * MO.NLIN_DFX(dv, Xn, *args_FX, symmetric)
        dpsi_T0 = MO.DT0_theta(dpsi, DT0, N_fm, nr, symmetric)
        dOm = MO.A2_SINE(dpsi, D, R, N_fm, nr, symmetric)
        NX[0:N] += dOm + dt * Pr * gr_k.dot(Ra * dT - Ra_s * dS)
        psi_new = MO.A4_BSub_TSTEP_V2(NX[0:N], L4, *args_A4, Pr * dt, symmetric) - dpsi
        NX[N:2 * N] += Rsq.dot(dT) - dt * dpsi_T0
        T_new = MO.NAB2_BSub_TSTEP_V2(NX[N:2 * N], LT, N_fm, nr, dt, symmetric) - dT
        NX[2 * N:3 * N] += Rsq.dot(dS) - dt * dpsi_T0
        S_new = MO.NAB2_BSub_TSTEP_V2(NX[2 * N:3 * N], LS, N_fm, nr, Tau * dt, symmetric) - dS
        return np.hstack((psi_new, T_new, S_new))

    def dmu(Xn):
        """PDFmu (Main.py:829-837)."""
        T = Xn[N:2 * N]
        out = 0. * Xn
        out[0:N] = MO.A4_BSub_TSTEP_V2(dt * Pr * gr_k.dot(T), L4, *args_A4, Pr * dt, symmetric)
        return out

    return step, jvp, dmu


def make_case(Main, MO, name, N_fm, N_r, d, dt, Pr, Tau, Ra, Ra_s, symmetric, seed, n_steps, amp=1e-3,
              store_ops=False):
    ops = Main.Build_Matrix_Operators(N_fm, N_r, d, dt, Pr, Tau)
    D, R, Rsq, DT0, gr_k, L4, LT, LS, args_A4, args_FX = ops
    nr = N_r - 1
    N = N_fm * nr
    rng = np.random.default_rng(seed)
    X0 = rng.random(3 * N)
    X0 = amp * X0 / np.linalg.norm(X0, 2)
    dv = rng.standard_normal(3 * N)
    step, jvp, dmu = ref_step_fn(Main, MO, ops, Ra, Ra_s, dt, Pr, Tau, symmetric)
    out = dict(N_fm=N_fm, N_r=N_r, d=d, dt=dt, Pr=Pr, Tau=Tau, Ra=Ra, Ra_s=Ra_s, symmetric=symmetric,
               seed=seed, n_steps=n_steps, amp=amp)
    psi = X0[0:N]
    # use an O(1) state for the per-function vectors so that the quadratic term is not vanishing
    Xb = rng.random(3 * N)
    out["Xb"] = Xb
    out["dv"] = dv
    out["J_theta_RT"] = MO.J_theta_RT(Xb[0:N], nr, N_fm, symmetric)
    out["DT0_theta"] = MO.DT0_theta(Xb[0:N], DT0, N_fm, nr, symmetric)
    out["A2_SINE"] = MO.A2_SINE(Xb[0:N], D, R, N_fm, nr, symmetric)
    out["A2_SINE_R2"] = MO.A2_SINE_R2(Xb[0:N], N_fm, nr, D, R, symmetric)
    out["kGR"] = gr_k.dot(Xb[N:2 * N])
    out["R2"] = Rsq.dot(Xb[N:2 * N])
    out["NLIN_FX"] = MO.NLIN_FX(Xb, *args_FX, symmetric)
    out["NLIN_DFX"] = MO.NLIN_DFX(dv, Xb, *args_FX, symmetric)
    out["A4_BSub"] = MO.A4_BSub_TSTEP_V2(Xb[0:N], L4, *args_A4, Pr * dt, symmetric)
    out["NAB2_BSub_T"] = MO.NAB2_BSub_TSTEP_V2(Xb[N:2 * N], LT, N_fm, nr, dt, symmetric)
    out["NAB2_BSub_S"] = MO.NAB2_BSub_TSTEP_V2(Xb[2 * N:3 * N], LS, N_fm, nr, Tau * dt, symmetric)
    out["step_Xb"] = step(Xb)
    out["jvp_Xb"] = jvp(dv, Xb)
    out["dmu_Xb"] = dmu(Xb)
    Xs = Main.Eq_SYM(Xb, R) * Xb if symmetric else Xb
    out["KE_Xb"] = Main.Kinetic_Energy(Xs, R, D, N_fm, nr, symmetric)
    out["NuT_Xb"] = Main.Nusselt(Xb[N:2 * N], d, R, D, N_fm, nr, check=False)
    out["NuS_Xb"] = Main.Nusselt(Xb[2 * N:3 * N], d, R, D, N_fm, nr, check=False)
    # time stepping from the small-amplitude IC (Main.py:286-329 loop semantics)
    mask = Main.Eq_SYM(X0, R) if symmetric else 1
    X = X0.copy()
    diags = []
    keep = {}
    for it in range(n_steps):
        Xn = step(X)
        diags.append([np.linalg.norm(Xn, 2), Main.Kinetic_Energy(Xn, R, D, N_fm, nr, symmetric),
                      Main.Nusselt(Xn[N:2 * N], d, R, D, N_fm, nr, check=False),
                      Main.Nusselt(Xn[2 * N:3 * N], d, R, D, N_fm, nr, check=False)])
        if it + 1 in (1, 10, n_steps):
            keep[it + 1] = Xn.copy()
        X = mask * Xn
    out["X0"] = X0
    for k_, v in keep.items():
        out["X_step%d" % k_] = v
    out["diag_hist"] = np.array(diags)
    if store_ops:
        out["op_D"] = D
        out["op_R"] = R
        out["op_L4"] = np.array([np.asarray(m) for m in L4])
        out["op_LT"] = np.array([np.asarray(m) for m in LT])
        out["op_LS"] = np.array([np.asarray(m) for m in LS])
        out["op_D2"] = args_A4[0]
        out["op_IR4"] = args_A4[1]
        out["op_IR2"] = args_A4[2]
        out["op_DT0"] = DT0
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "||X_final|| = %.16e" % np.linalg.norm(keep[n_steps]))


def make_transforms(TR):
    rng = np.random.default_rng(7)
    out = {}
    for K in (16, 48):
        M = (3 * K) // 2
        a = rng.standard_normal((5, K))
        g = rng.standard_normal((5, M))
        out["in_hat_%d" % K] = a
        out["in_grid_%d" % K] = g
        out["IDCT_%d" % K] = TR.IDCT(a, n=M)
        out["IDST_%d" % K] = TR.IDST(a, n=M)
        out["IDCT_same_%d" % K] = TR.IDCT(a)
        out["IDST_same_%d" % K] = TR.IDST(a)
        out["IDCT_half_%d" % K] = TR.IDCT(a, n=K // 2)      # truncating inverse transforms
        out["IDST_half_%d" % K] = TR.IDST(a, n=K // 2)
        out["IDCT_3x_%d" % K] = TR.IDCT(a, n=3 * K)
        out["IDST_3x_%d" % K] = TR.IDST(a, n=3 * K)
        out["DCT_%d" % K] = TR.DCT(g)
        out["DST_%d" % K] = TR.DST(g)
        out["DCT_trunc_%d" % K] = TR.DCT(g, n=K)
        out["DST_trunc_%d" % K] = TR.DST(g, n=K)
        out["grid_%d" % K] = TR.grid(M)
    np.savez_compressed(os.path.join(HERE, "transforms.npz"), **out)
    print("wrote transforms")


def make_interp(MO):
    """Resolution-transfer helpers the drivers call right before the hot path (Main.py:414-416)."""
    import contextlib, io
    rng = np.random.default_rng(11)
    N_fm, N_r, d = 16, 10, 0.4
    X = rng.random(3 * (N_r - 1) * N_fm)
    with contextlib.redirect_stdout(io.StringIO()):
        out = dict(X=X, N_fm=N_fm, N_r=N_r, d=d,
                   theta_up=MO.INTERP_THETAS(32, N_fm, X), theta_down=MO.INTERP_THETAS(8, N_fm, X),
                   radial_up=MO.INTERP_RADIAL(14, N_r, X, d))
    np.savez_compressed(os.path.join(HERE, "interp.npz"), **out)
    print("wrote interp")


if __name__ == "__main__":
    Main, MO, TR = import_reference()
    if "--only-interp" in sys.argv:
        make_interp(MO)
        sys.exit(0)
    if "--only-transforms" in sys.argv:
        make_transforms(TR)
        sys.exit(0)
    make_transforms(TR)
    make_interp(MO)
    common = dict(d=0.4, dt=1e-2, Pr=0.7, Tau=1.0 / 15.0, Ra=3000.0, Ra_s=400.0)
    make_case(Main, MO, "small_nosym", 16, 10, symmetric=False, seed=1, n_steps=20, store_ops=True, **common)
    make_case(Main, MO, "small_sym", 16, 10, symmetric=True, seed=2, n_steps=20, **common)
    # config 1 of BASELINE.json: Main.Time_Step literals (Main.py:332,359-376,424-425), 100 steps
    cfg1 = dict(d=0.31325, dt=1e-3, Pr=1.0, Tau=1.0, Ra=3750.0, Ra_s=0.0)
    make_case(Main, MO, "cfg1_nosym", 48, 24, symmetric=False, seed=0, n_steps=100, **cfg1)
    make_case(Main, MO, "cfg1_sym", 48, 24, symmetric=True, seed=3, n_steps=100,
              d=0.3521, dt=1e-3, Pr=1.0, Tau=1.0 / 15.0, Ra=9851.0, Ra_s=500.0)
    # config 3 shape (N_r=30, N_theta=256), one member, 100 steps
    make_case(Main, MO, "cfg3_member", 256, 30, symmetric=False, seed=2000, n_steps=100, **cfg1)

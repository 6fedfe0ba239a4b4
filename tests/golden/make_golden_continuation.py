"""Golden vectors of the reference's Newton / pseudo-arc-length DRIVERS (run in the build container only).

    python tests/golden/make_golden_continuation.py        # needs /root/reference, a few minutes of CPU

Runs the UNMODIFIED Main._Newton, Main._ContinC and Main._Continuation (SciPy LGMRES on the host) on the wide-gap l = 2
case of the reference's own test suite (Tests/Run_Tests.py:140-199: d = 2, Pr = 10, N_fm = 32, N_r = 16, symmetric) and
stores their inputs, outputs and printed iteration histories in continuation.npz.  The batched drivers of
spectraldoublediffusiveconvection_b200/krylov.py are compared against these.
"""
import contextlib
import io
import os
import re

import numpy as np

from make_golden import HERE, import_reference, ref_step_fn


def floats_after(text, pattern):
    return np.array([[float(x) for x in m] if isinstance(m, tuple) else [float(m)] for m in re.findall(pattern, text)])


def main():
    Main, MO, TR = import_reference()
    N_fm, N_r, d, Pr, Tau, Ra, Ra_s, sym = 32, 16, 2.0, 10.0, 1.0, 6780.0, 0.0, True
    nr = N_r - 1
    out = dict(N_fm=N_fm, N_r=N_r, d=d, Pr=Pr, Tau=Tau, Ra=Ra, Ra_s=Ra_s, symmetric=sym)
    # transient towards the steady branch: 13000 IMEX steps of dt = 0.075 from a seeded random state
    dt = 0.075
    ops = Main.Build_Matrix_Operators(N_fm, N_r, d, dt, Pr, Tau)
    step, _, _ = ref_step_fn(Main, MO, ops, Ra, Ra_s, dt, Pr, Tau, sym)
    X = np.random.default_rng(0).random(3 * nr * N_fm)
    X = 1e-3 * X / np.linalg.norm(X)
    mask = Main.Eq_SYM(X, ops[1])
    X = mask * X
    for _ in range(13000):
        X = mask * step(X)
    kw = dict(Ra_s=Ra_s, Tau=Tau, Pr=Pr, d=d, N_fm=N_fm, N_r=N_r, symmetric=sym)
    # the transient is not yet inside the 5-iteration basin of Main._Newton: polish it with the same iteration (the
    # restatement of tests/dropin_drivers.py on the reference's operators, 10 iterations allowed) at Ra = 6780 ...
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    import dropin_drivers as drv
    X, h0, _ = drv.newton(MO, X, Ra, Ra_s, Tau, Pr, d, N_fm, N_r, sym, max_it=10)
    assert h0[-1] < 1e-8, h0
    out["X_start"] = X.copy()
    # ---- ... and let the unmodified Main._Newton move that steady state to Ra = 6782 (close to onset the amplitude varies quickly with Ra) ----
    Ra = 6782.0
    out["Ra_newton"] = Ra
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        Xn, Norm, KE, NuT, NuS, ok = Main._Newton(X.copy(), Ra, **kw)
    assert ok, buf.getvalue()
    out["newton_X"] = Xn
    out["newton_diag"] = np.array([Norm, KE, NuT, NuS])
    out["newton_history"] = floats_after(buf.getvalue(), r"Newton Iteration = \d+, Error = ([0-9.eE+-]+)").ravel()
    print("newton history", out["newton_history"], "KE", KE)
    # ---- Main._ContinC: one pseudo-arc-length step from the converged point ----
    Y0 = np.hstack((Xn, Ra))
    for tag, ds0 in (("a", 0.5), ("b", 8.0)):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            Y_dot, Y, sign, ds, Norm, KE, NuT, NuS, ok = Main._ContinC(0.0 * Y0, Y0, 1.0, ds0, Ra, **kw)
        txt = buf.getvalue()
        out["continc_%s_ds0" % tag] = ds0
        out["continc_%s_Y" % tag] = Y
        out["continc_%s_Ydot" % tag] = Y_dot
        out["continc_%s_ds" % tag] = ds
        out["continc_%s_diag" % tag] = np.array([Norm, KE, NuT, NuS])
        out["continc_%s_history" % tag] = floats_after(txt, r"Error X = ([0-9.eE+-]+), Error µ = ([0-9.eE+-]+)")
        out["continc_%s_xi_norm" % tag] = floats_after(txt, r"\|\|ξ\|\| = ([0-9.eE+-]+)").ravel()
        print("continc", tag, "ds", ds0, "->", ds, "mu", Y[-1], "mu_dot", Y_dot[-1], "history", out["continc_%s_history" % tag].tolist())
    # ---- Main._Continuation: the branch loop (starts with ds = 0.01: arc-length steps that double ds, then Newton steps) ----
    files = []

    class Rec(dict):
        def __init__(self, *a, **k):
            super().__init__()
            files.append(self)

        def create_group(self, name):
            g = Rec.__new__(Rec)
            dict.__init__(g)
            self[name] = g
            return g

        def create_dataset(self, name, data=None, **k):
            self[name] = data

        def close(self):
            pass

    Main.h5py.File = Rec
    nsteps = 12
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        Main._Continuation("golden_branch.h5", nsteps, 1.0, Y0, Ra=Ra, **kw)
    txt = buf.getvalue()
    out["branch_steps"] = nsteps
    for key in ("Ra", "Ra_dot", "NuT", "NuS", "KE"):
        out["branch_" + key] = floats_after(txt, r"\n" + key.replace("Ra_dot", "Ra_dot") + r"\s+= ([0-9.eE+-]+)").ravel()
    last = files[-1]
    out["branch_X_DATA"] = np.array(last["Checkpoints"]["X_DATA"])
    out["branch_Ra_DATA"] = np.array(last["Checkpoints"]["Ra_DATA"])
    out["branch_ds_events"] = np.array([[float(a), float(b)] for a, b in
                                        re.findall(r"(?:Increasing|Reducing) the step-size ds_old=([0-9.eE+-]+) -> ds_new=([0-9.eE+-]+)", txt)])
    out["branch_switches"] = txt.count("Switching to arc-length")
    print("branch Ra", out["branch_Ra"])
    print("branch Ra_dot", out["branch_Ra_dot"])
    print("branch KE", out["branch_KE"])
    print("ds events", out["branch_ds_events"].tolist(), "switches", out["branch_switches"])
    np.savez_compressed(os.path.join(HERE, "continuation.npz"), **out)


if __name__ == "__main__":
    main()

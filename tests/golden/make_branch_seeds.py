"""Reproducible starting states on the l = 10 and l = 11 steady branches (BASELINE configs 4 and 5), found the way the
reference finds them (Linear_Problem.main_program -> Main.Newton -> Main.Continuation): marginal eigenvector of the
conduction state, Newton just below the critical Rayleigh number, then a stretch of pseudo-arc-length continuation away
from the bifurcation point.  Everything runs on the CPU at the reference's own low resolution (N_fm = 64, N_r = 20,
Main.py:625-626) with this repo's host code -- linear.py, the batched drivers of krylov.py on the oracle-backed plan of
tests/oracle_plan.py -- and takes a few minutes:

    python tests/golden/make_branch_seeds.py

Output: branch_seeds.npz (two states of 3 * 19 * 64 doubles and their parameters).  They are STARTING states, not
parity vectors: the GPU runs interpolate them to (30, 256) / (40, 512) (INTERP_RADIAL / INTERP_THETAS, as
Main.py:599-601 does) and the parity tests compare GPU and oracle Newton histories from there.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import sddc_oracle as orc                                     # noqa: E402
from oracle_plan import OraclePlan                                        # noqa: E402
from spectraldoublediffusiveconvection_b200 import krylov, linear         # noqa: E402

SETS = {
    # name: l, d, Ra_c (steady onset; Linear_Problem.py:319-335, 481-488), Ra_s, symmetric (even l: Main.py:632-635)
    "l10": dict(l=10.0, d=0.3521, Ra_c=9851.537357677651, Ra_s=500.0, symmetric=True),
    "l11": dict(l=11.0, d=0.31325, Ra_c=4525.905436209724, Ra_s=150.0, symmetric=False),
}
N_FM, N_R, PR, TAU = 64, 20, 1.0, 1.0 / 15.0


def main(steps=24):
    orc.set_transform_backend("fft")
    orc.set_accel(True)
    out = dict(N_fm=N_FM, N_r=N_R, Pr=PR, Tau=TAU)
    for name, p in SETS.items():
        pl = OraclePlan(N_FM, N_R, p["d"], 1.0, PR, TAU, symmetric=p["symmetric"])
        f = linear.eig_vec(p["Ra_c"], p["l"], p["d"], 1, Ra_s=p["Ra_s"], Pr=PR, Tau=TAU, Nr=N_R)
        X = linear.full_eig_vec(f, p["l"], N_FM, N_R - 1, symmetric=p["symmetric"])
        Ra = p["Ra_c"] - 5e-3                                             # Main.py:621
        X0 = torch.as_tensor(3e-3 * X / np.linalg.norm(X)).reshape(1, -1)
        Xn, info = krylov.newton_batched(pl, X0, Ra, p["Ra_s"], krylov=150, max_it=12)
        print(name, "Newton from the eigenvector:", info["history"].numpy().ravel(), bool(info["converged"][0]),
              "|X| =", float(torch.linalg.vector_norm(Xn)))
        best = None
        for sign in (-1.0, 1.0):
            res = krylov.continuation_batched(pl, Xn, Ra, steps, p["Ra_s"], sign=sign, krylov=150)
            h = res.stacked()
            print(name, "sign", sign, "alive", bool(res.alive[0]), "Ra", h["Ra"][-1, 0], "KE", h["KE"][-1, 0],
                  "folds", [(a, b) for a, b, _ in res.folds[0]], "jvps", res.jvps)
            if bool(res.alive[0]) and (best is None or h["KE"][-1, 0] > best[2]):
                best = (res.X[0].numpy().copy(), float(res.mu[0]), float(h["KE"][-1, 0]), sign)
        assert best is not None
        out[name + "_X"], out[name + "_Ra"], out[name + "_KE"], out[name + "_sign"] = best
        out[name + "_params"] = np.array([p["l"], p["d"], p["Ra_c"], p["Ra_s"], float(p["symmetric"])])
    np.savez_compressed(os.path.join(HERE, "branch_seeds.npz"), **out)


if __name__ == "__main__":
    main()

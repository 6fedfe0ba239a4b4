"""Golden vectors of the reference's linear-theory seeding (run in the build container only).

    python tests/golden/make_golden_linear.py        # needs /root/reference

Executes the UNMODIFIED source of Linear_Problem.py (with the IPython magic on its last lines stripped, a stub
matplotlib and the NumPy-2 aliases) and stores eigenvalues, eigenvectors and seeded states
for the l = 10 and l = 11/13 parameter sets of Main.py:608-624 / Linear_Problem.py:319-335 in linear.npz.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_linear_problem():
    if not hasattr(np, "RankWarning"):
        np.RankWarning = np.exceptions.RankWarning
    if not hasattr(np, "complex_"):
        np.complex_ = np.complex128
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.rcParams = {}
    mpl.pyplot = plt
    mpl.ticker = types.ModuleType("matplotlib.ticker")
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    sys.modules["matplotlib.ticker"] = mpl.ticker
    if REF not in sys.path:
        sys.path.insert(0, REF)
    src = open(os.path.join(REF, "Linear_Problem.py")).read()
    src = "\n".join(ln for ln in src.split("\n") if not ln.lstrip().startswith("%"))
    mod = types.ModuleType("Linear_Problem")
    mod.__file__ = os.path.join(REF, "Linear_Problem.py")
    mod.__name__ = "Linear_Problem_ref"     # keeps the `if __name__ == "__main__"` block from running
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


def main():
    LP = import_linear_problem()
    out = {}
    sets = {
        "l10": dict(l=10.0, d=0.3521, Ra=9851.537357677651, Ra_s=500.0, Pr=1.0, Tau=1.0 / 15.0),
        "l11": dict(l=11.0, d=0.31325, Ra=4525.905436209724, Ra_s=150.0, Pr=1.0, Tau=1.0 / 15.0),
    }
    for name, p in sets.items():
        for Nr in (20, 30):
            tag = "%s_Nr%d" % (name, Nr)
            out[tag + "_vals"] = LP.Eig_Vals(p["Ra"], p["l"], p["d"], 4, Ra_s=p["Ra_s"], Pr=p["Pr"], Tau=p["Tau"], Nr=Nr)
            for k in (0, 1):
                out[tag + "_vec%d" % k] = LP.Eig_Vec(p["Ra"], p["l"], p["d"], k, Ra_s=p["Ra_s"], Pr=p["Pr"], Tau=p["Tau"], Nr=Nr)
        f = out[name + "_Nr20_vec1"]
        for N_fm in (48, 64):
            for sym in (False, True):
                out["%s_full_K%d_%s" % (name, N_fm, "sym" if sym else "nosym")] = LP.Full_Eig_Vec(f, p["l"], N_fm, 19, symmetric=sym)
        out[name + "_params"] = np.array([p["l"], p["d"], p["Ra"], p["Ra_s"], p["Pr"], p["Tau"]])
    # Critical_Eigval uses the defaults Ra_s = 150, Tau = 1/15, Nr = 20 of Eig_Vals (Linear_Problem.py:112)
    out["crit_l11"] = np.array([LP.Critical_Eigval(4525.9, 11.0, 0.31325, Nvals=1)])
    out["crit_l13"] = np.array([LP.Critical_Eigval(4619.4, 13.0, 0.31325, Nvals=1)])
    np.savez_compressed(os.path.join(HERE, "linear.npz"), **out)
    for k, v in out.items():
        print(k, np.asarray(v).shape)


if __name__ == "__main__":
    main()

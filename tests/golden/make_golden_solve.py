"""Second CPU oracle for the implicit solves: outputs of the reference's SOLVE-based back-substitutions
NAB2_BSub_TSTEP / A4_BSub_TSTEP (Matrix_Operators.py:248-433), which factorise every mode's operator instead of
applying the pre-inverted stacks the V2 functions (and the GPU path) use.  Run in the build container only:

    python tests/golden/make_golden_solve.py        # needs /root/reference

Right-hand sides are the `Xb` fields of the existing fixtures, so solve.npz pairs with them case by case.
"""
import os

import numpy as np

from make_golden import HERE, import_reference


def main():
    Main, MO, TR = import_reference()
    out = {}
    for name in ("small_nosym", "small_sym", "cfg1_nosym", "cfg1_sym"):
        g = np.load(os.path.join(HERE, name + ".npz"))
        N_fm, N_r, d, dt, Pr, Tau = int(g["N_fm"]), int(g["N_r"]), float(g["d"]), float(g["dt"]), float(g["Pr"]), float(g["Tau"])
        sym = bool(g["symmetric"])
        nr = N_r - 1
        N = N_fm * nr
        D, R = MO.cheb_radial(N_r, d)
        D = np.ascontiguousarray(D)
        ops = Main.Build_Matrix_Operators(N_fm, N_r, d, dt, Pr, Tau)
        D2, IR4, IR2 = ops[8][0], ops[8][1], ops[8][2]
        A2 = np.ascontiguousarray((D @ D)[1:-1, 1:-1])
        D4 = MO.Nabla4(D, R)
        R2_Nab2 = MO.Nabla2(D, R)
        R2 = np.diag(R[1:-1] ** 2)
        I = np.eye(nr)
        Xb = g["Xb"]
        out[name + "_A4"] = MO.A4_BSub_TSTEP(Xb[0:N].copy(), D4, IR4, D2, A2, IR2, N_fm, nr, Pr * dt, sym)
        out[name + "_T"] = MO.NAB2_BSub_TSTEP(Xb[N:2 * N].copy(), R2_Nab2, R2, I, N_fm, nr, dt, sym)
        out[name + "_S"] = MO.NAB2_BSub_TSTEP(Xb[2 * N:3 * N].copy(), R2_Nab2, R2, I, N_fm, nr, Tau * dt, sym)
        for f, key in (("A4", "A4_BSub"), ("T", "NAB2_BSub_T"), ("S", "NAB2_BSub_S")):
            a, b = out[name + "_" + f], g[key]
            print("%-12s %-3s solve-based vs pre-inverted (V2): rel %.2e" % (name, f, np.linalg.norm(a - b) / np.linalg.norm(b)))
    np.savez_compressed(os.path.join(HERE, "solve.npz"), **out)


if __name__ == "__main__":
    main()

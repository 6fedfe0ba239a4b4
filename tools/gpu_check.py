"""Quick GPU parity report (not a test): prints the relative error of every entry point per golden case."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, rel_l2
from spectraldoublediffusiveconvection_b200 import EnsemblePlan, plan as P

def dev(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()

for name in sys.argv[1:] or ["small_nosym", "small_sym", "cfg1_nosym", "cfg1_sym", "cfg3_member"]:
    g = load_golden(name)
    pl = EnsemblePlan(int(g["N_fm"]), int(g["N_r"]), float(g["d"]), float(g["dt"]), float(g["Pr"]), float(g["Tau"]),
                      symmetric=bool(g["symmetric"]), max_batch=2)
    N = pl.N; Xb = g["Xb"]; Ra, Ra_s = float(g["Ra"]), float(g["Ra_s"])
    psi, T, S = dev(Xb[:N]), dev(Xb[N:2*N]), dev(Xb[2*N:])
    rep = {}
    for op, key, src in ((0, "J_theta_RT", psi), (1, "DT0_theta", psi), (2, "A2_SINE", psi), (3, "A2_SINE_R2", psi), (4, "kGR", T), (5, "R2", T)):
        rep[key] = rel_l2(pl.linear_op(op, src).cpu().numpy().ravel(), g[key])
    rep["NLIN_FX"] = rel_l2(pl.nlin_fx(dev(Xb)).cpu().numpy().ravel(), g["NLIN_FX"])
    rep["NLIN_DFX"] = rel_l2(pl.nlin_dfx(dev(g["dv"]), dev(Xb)).cpu().numpy().ravel(), g["NLIN_DFX"])
    rep["A4"] = rel_l2(pl.solve_a4(psi).cpu().numpy().ravel(), g["A4_BSub"])
    rep["NAB2_T"] = rel_l2(pl.solve_nab2(T, 0).cpu().numpy().ravel(), g["NAB2_BSub_T"])
    rep["NAB2_S"] = rel_l2(pl.solve_nab2(S, 1).cpu().numpy().ravel(), g["NAB2_BSub_S"])
    rep["step"] = rel_l2(pl.step(dev(Xb), Ra, Ra_s).cpu().numpy().ravel(), g["step_Xb"])
    rep["jvp"] = rel_l2(pl.jvp(dev(g["dv"]), dev(Xb), Ra, Ra_s).cpu().numpy().ravel(), g["jvp_Xb"])
    rep["dmu"] = rel_l2(pl.dF_dRa(dev(Xb)).cpu().numpy().ravel(), g["dmu_Xb"])
    d = pl.diagnostics(dev(Xb)).cpu().numpy()[0]
    rep["KE"] = abs(d[1] / float(g["KE_Xb"]) - 1); rep["NuT"] = abs(d[2] / float(g["NuT_Xb"]) - 1)
    X = dev(g["X0"]).reshape(1, -1)
    n_steps = int(g["n_steps"])
    X = pl.step(X, Ra, Ra_s, nsteps=n_steps)
    rep["steps%d" % n_steps] = rel_l2(X.cpu().numpy().ravel(), g["X_step%d" % n_steps])
    print(name, " ".join("%s=%.2e" % kv for kv in rep.items()), flush=True)
    pl.close()

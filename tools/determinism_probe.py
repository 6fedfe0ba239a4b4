"""Run-to-run determinism of a plain step through the gather-mode back-substitution and through the four-kernel path
(timing-dependent differences = a race): the same input stepped REP times, outputs compared bit for bit."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from spectraldoublediffusiveconvection_b200 import EnsemblePlan
K, N_r, sym, B = 256, 30, (len(sys.argv) < 3 or sys.argv[2] != "nosym"), (int(sys.argv[3]) if len(sys.argv) > 3 else 261)
REP = int(sys.argv[1]) if len(sys.argv) > 1 else 30
pl = EnsemblePlan(K, N_r, 0.353, 2e-3, 1.0, 1.0 / 15.0, symmetric=sym, max_batch=B)
rng = np.random.default_rng(B)
X = torch.as_tensor(rng.random((B, 3 * pl.N)) * 1e-2).cuda()
Ra = torch.as_tensor(np.linspace(3000.0, 9000.0, B)).cuda(); Ras = torch.as_tensor(np.linspace(0.0, 500.0, B)).cuda()
ref_big = pl.step(X, Ra, Ras).clone()
ref_small = torch.cat([pl.step(X[lo:lo + 50], Ra[lo:lo + 50], Ras[lo:lo + 50]) for lo in range(0, B, 50)])
N = pl.N
dm = torch.nonzero(((ref_big - ref_small).norm(dim=1) / ref_small.norm(dim=1)) > 1e-12).flatten().tolist()
print("first gather run: members off by more than 1e-12 vs the four-kernel path:", dm[:40], "...", len(dm))
for m in dm[:6]:
    e = [float((ref_big[m, f * N:(f + 1) * N] - ref_small[m, f * N:(f + 1) * N]).norm() / ref_small[m, f * N:(f + 1) * N].norm()) for f in range(3)]
    dpsi = (ref_big[m, :N] - ref_small[m, :N]).reshape(K, -1).abs().max(dim=1).values
    ks = torch.nonzero(dpsi > 0).flatten().tolist()
    dT = (ref_big[m, N:2 * N] - ref_small[m, N:2 * N]).reshape(K, -1).abs().max(dim=1).values
    kt = torch.nonzero(dT > 0).flatten().tolist()
    print("  member %d: rel diff psi %.2e T %.2e S %.2e; T blocks that differ: highest %s lowest %s (%d); max |dT| per block at top: %s" % (m, e[0], e[1], e[2], kt[-6:], kt[:3], len(kt), [float("%.2e" % dT[k]) for k in kt[-6:]]))
print("gather vs four-kernel path, max rel diff per member:", float(((ref_big - ref_small).norm(dim=1) / ref_small.norm(dim=1)).max()))
rel = lambda a, b: float(((a - b).norm(dim=1) / b.norm(dim=1)).max())
bad_big = bad_small = 0
for r in range(REP):
    o = pl.step(X, Ra, Ras)
    d = (o != ref_big).any(dim=1)
    if d.any():
        bad_big += 1
        if bad_big <= 6:
            print("  gather run %d differs from the first in %d members; vs first %.2e, vs four-kernel path %.2e" % (r, int(d.sum()), rel(o, ref_big), rel(o, ref_small)))
    o = torch.cat([pl.step(X[lo:lo + 50], Ra[lo:lo + 50], Ras[lo:lo + 50]) for lo in range(0, B, 50)])
    d = (o != ref_small).any(dim=1)
    if d.any():
        bad_small += 1
        print("  four-kernel run %d differs: members" % r, torch.nonzero(d).flatten().tolist()[:12])
    # different timing: a multi-step call and a JVP in between
    pl.step(X, Ra, Ras, nsteps=2)
print("runs that differ: gather %d / %d, four-kernel %d / %d" % (bad_big, REP, bad_small, REP))

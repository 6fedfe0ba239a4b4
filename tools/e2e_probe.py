"""Where does the host-buffer time loop spend its time?  ms per step of sddc_time_step_host for several run lengths and
with diagnostics / checkpoints switched off."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from spectraldoublediffusiveconvection_b200 import EnsemblePlan
B = 512
pl = EnsemblePlan(256, 30, 0.31325, 1e-3, 1.0, 1.0, max_batch=B)
W = 3 * pl.N
xin, xout = pl.pinned((B, W)), pl.pinned((B, W))
xin[...] = np.random.default_rng(0).random((B, W)) * 1e-3
Ra, Ras = np.linspace(2000., 6000., B), np.zeros(B)
hist = pl.pinned((400, B, 6)); ckp = pl.pinned((10, B, W))
pl.time_step_host(xin, Ra, Ras, 400, diag_every=1, ckpt_every=0, out=xout, diag_hist=hist)   # warm-up, sizes the history
for ns, de, ce in [(50, 1, 5), (100, 1, 10), (200, 1, 20), (400, 1, 40), (200, 0, 20), (200, 1, 0), (200, 0, 0), (200, 10, 20)]:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pl.time_step_host(xin, Ra, Ras, ns, diag_every=de, ckpt_every=ce, out=xout,
                      diag_hist=hist[:ns // de] if de else None, ckpt=ckp[:ns // ce] if ce else None)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
    print("nsteps %4d diag_every %2d ckpt_every %3d : %7.2f ms total, %.3f ms/step" % (ns, de, ce, ms, ms / ns), flush=True)

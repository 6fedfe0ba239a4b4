"""One member-step of the headline configuration between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --set full --clock-control none --import-source on -o out python tools/profile_step.py
(the capture then holds exactly scan, prep, synth, analysis, solve of one step; add `jvp` as argv[1] for the cached
Jacobian-vector product instead, `diag` for three steps of the device-resident time loop with per-step diagnostics)."""
import sys; sys.path.insert(0, '.')
import torch
from spectraldoublediffusiveconvection_b200 import EnsemblePlan

B = 512
pl = EnsemblePlan(256, 30, 0.31325, 1e-3, 1.0, 1.0, max_batch=B)
X = torch.rand((B, 3 * pl.N), dtype=torch.float64, device='cuda') * 1e-3
Ra = torch.linspace(2000.0, 6000.0, B, dtype=torch.float64, device='cuda')
Ras = torch.zeros_like(Ra)
out = torch.empty_like(X)
mode = sys.argv[1] if len(sys.argv) > 1 else 'step'
dv = torch.randn_like(X)
for _ in range(3):
    pl.step(X, Ra, Ras, out=out)
if mode == 'jvp':
    pl.jvp_set_base(X)
    pl.jvp_apply(dv, Ra, Ras, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if mode == 'jvp':
    pl.jvp_apply(dv, Ra, Ras, out=out)
elif mode == 'diag':
    pl.time_step(X, Ra, Ras, 3, diag_every=1, out=out)      # steps with the shared-prep diagnostics (KE transform, finish)
else:
    pl.step(X, Ra, Ras, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

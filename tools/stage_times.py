import sys; sys.path.insert(0,'.')
import numpy as np, torch
from spectraldoublediffusiveconvection_b200 import EnsemblePlan
pl = EnsemblePlan(256, 30, 0.31325, 1e-3, 1.0, 1.0, max_batch=512)
X = torch.rand((512, 3*pl.N), dtype=torch.float64, device='cuda')*1e-3
Ra = torch.full((512,), 3000.0, dtype=torch.float64, device='cuda'); Ras = torch.zeros_like(Ra)
out = torch.empty_like(X)
for _ in range(3): pl.nlin_fx(X, out=out); pl.step(X, Ra, Ras, out=out)
pl.profile_begin()
for _ in range(10): pl.nlin_fx(X, out=out)
p1 = pl.profile_end()
pl.profile_begin()
for _ in range(10): pl.step(X, Ra, Ras, out=out)
p2 = pl.profile_end()
print("nlin_fx only:", {k: round(v[0]/max(v[1],1),4) for k,v in p1.items() if v[1]})
print("step        :", {k: round(v[0]/max(v[1],1),4) for k,v in p2.items() if v[1]})

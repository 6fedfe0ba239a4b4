"""Stage times of one member-step against the number of members per GPU (which stages are latency bound?)."""
import sys; sys.path.insert(0, '.')
import torch
from spectraldoublediffusiveconvection_b200 import EnsemblePlan
sym = len(sys.argv) > 1 and sys.argv[1] == "sym"
for B in (64, 128, 192, 256, 384, 512, 1024, 2048):
    pl = EnsemblePlan(256, 30, 0.31325, 1e-3, 1.0, 1.0, symmetric=sym, max_batch=B)
    X = torch.rand((B, 3 * pl.N), dtype=torch.float64, device='cuda') * 1e-3
    Ra = torch.full((B,), 3000.0, dtype=torch.float64, device='cuda'); Ras = torch.zeros_like(Ra)
    out = torch.empty_like(X)
    for _ in range(3): pl.step(X, Ra, Ras, out=out)
    pl.profile_begin()
    for _ in range(10): pl.step(X, Ra, Ras, out=out)
    p = pl.profile_end()
    print("B=%5d" % B, {k: round(v[0] / max(v[1], 1), 4) for k, v in p.items() if v[1]})
    pl.close()

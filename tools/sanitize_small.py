"""Small end-to-end exercise of every entry point for compute-sanitizer runs."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from spectraldoublediffusiveconvection_b200 import EnsemblePlan, plan as P
import os
SHAPES = [(16, 10, False, 3), (32, 30, True, 17), (24, 41, False, 2), (16, 65, False, 2), (10, 9, False, 2), (30, 12, True, 3)]
if os.environ.get('SANITIZE_FFT'):   # the FFT formulation (k_nlin_fft.cuh): N_fm = 128 / 256 / 512
    SHAPES = [(128, 10, False, 3), (256, 8, True, 2), (512, 6, False, 1)]
if os.environ.get('SANITIZE_FFT') == '3':   # batches of 128 members and more: gather mode of the back-substitution (k_solve_hot.cuh)
    SHAPES = [(128, 9, False, 259), (128, 30, True, 257)]
for (K, N_r, sym, B) in SHAPES:
    pl = EnsemblePlan(K, N_r, 0.4, 1e-2, 1.0, 0.5, symmetric=sym, max_batch=B)
    X = torch.rand((B, 3 * pl.N), dtype=torch.float64, device='cuda') * 1e-2
    dv = torch.randn_like(X)
    Ra = torch.full((B,), 3000.0, dtype=torch.float64, device='cuda'); Ras = torch.zeros_like(Ra)
    if os.environ.get('SANITIZE_FFT') == '2':   # only the FFT kernels (racecheck: the TMA/mbarrier kernels report false hazards)
        pl.nlin_fx(X); pl.nlin_dfx(dv, X) if K <= 256 else None; torch.cuda.synchronize(); print('ok fft only', K, N_r); pl.close(); continue
    Y = pl.step(X, Ra, Ras, nsteps=3)
    pl.nlin_fx(X); pl.residual(X, Ra, Ras); pl.dF_dRa(X); pl.diagnostics(Y)
    pl.jvp(dv, X, Ra, Ras); pl.nlin_dfx(dv, X)
    pl.jvp_set_base(X); pl.jvp_apply(dv, Ra, Ras)
    for op in range(6):
        pl.linear_op(op, X[:, :pl.N].contiguous())
    pl.solve_a4(X[:, :pl.N].contiguous()); pl.solve_nab2(X[:, :pl.N].contiguous(), 1)
    out, hist = pl.time_step_host(X.cpu().numpy(), 3000.0, 0.0, 4, diag_every=2)
    out, hist, ck = pl.time_step_host(X.cpu().numpy(), 3000.0, 0.0, 5, diag_every=1, ckpt_every=2, ckpt_first=1)
    pl.time_step(X, Ra, Ras, 4, diag_every=1)
    # round 2: batched resolution transfer, Gram-Schmidt kernels through a short lock-step Newton solve
    from spectraldoublediffusiveconvection_b200 import interp, krylov
    interp.interp_thetas(X, K + 6, K); interp.interp_thetas(X, K - 2, K); interp.interp_radial(X, N_r + 3, N_r, 0.4)
    krylov.newton_batched(pl, X, Ra, Ras, krylov=6, max_it=1, max_restarts=1)
    krylov.continc_batched(pl, X, Ra, 1.0, 0.1, Ras, krylov=5, max_restarts=1, max_rounds=1)
    P.transform(P.T_IDCT, X[:, :K].contiguous(), 3 * K // 2)
    torch.cuda.synchronize()
    print("ok", K, N_r, sym, float(Y.abs().max()))
    pl.close()

"""BASELINE configs 4 and 5 at their stated sizes on one GPU (run on the B200 box):

  config 4  matrix-free Newton (Main._Newton) for the l = 10 state at N_r = 30, N_theta = 256, symmetric, with B = 64
            concurrent Rayleigh numbers and batched GPU JVPs;
  config 5  pseudo-arc-length continuation (Main._ContinC) along the l = 11 branch at N_r = 40, N_theta = 512 with 512
            concurrent branch points (512 bordered solves in lock step).

Starting states: tests/golden/branch_seeds.npz (found on the CPU at (20, 64) by tests/golden/make_branch_seeds.py),
interpolated to the target resolution like Main.py:599-601 does.  Prints one JSON object and writes it to
gpurun_out/configs45.json.

    python tools/run_configs45.py [--b4 64] [--b5 512] [--krylov 60]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from spectraldoublediffusiveconvection_b200 import EnsemblePlan, krylov        # noqa: E402
from spectraldoublediffusiveconvection_b200.compat import Matrix_Operators as MO  # noqa: E402

PR, TAU = 1.0, 1.0 / 15.0


def upsample(X, N_fm_o, N_r_o, N_fm, N_r, d):
    X = MO.INTERP_RADIAL(N_r, N_r_o, X, d)
    return MO.INTERP_THETAS(N_fm, N_fm_o, X)


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def config4(seeds, B, m):
    l, d, Ra_c, Ra_s, sym = seeds["l10_params"]
    N_fm, N_r = 256, 30
    X0 = upsample(seeds["l10_X"], int(seeds["N_fm"]), int(seeds["N_r"]), N_fm, N_r, d)
    pl = EnsemblePlan(N_fm, N_r, d, 1.0, PR, TAU, symmetric=bool(sym), max_batch=B)
    Ra0 = float(seeds["l10_Ra"])
    Ras = torch.as_tensor(Ra0 - np.linspace(0.0, 10.0, B)).cuda()
    X = torch.as_tensor(X0).cuda().repeat(B, 1)
    krylov.newton_batched(pl, X[:2], Ras[:2], Ra_s, krylov=8, max_it=1)          # warm-up (allocations, first launches)
    krylov.profile_begin()
    (Xn, info), wall = timed(lambda: krylov.newton_batched(pl, X, Ras, Ra_s, krylov=m))
    prof = krylov.profile_end()
    (_, _), wall_noprof = timed(lambda: krylov.newton_batched(pl, X, Ras, Ra_s, krylov=m))
    res = pl.residual(Xn * krylov.symmetry_mask(pl, Xn.device), Ras, Ra_s)
    dg = pl.diagnostics(Xn).cpu().numpy()
    out = {
        "config": "BASELINE configs[3]: Newton, l=10 set (d=%.4f, Ra_s=%g, Tau=1/15, Pr=1), N_r=%d N_theta=%d, symmetric, "
                  "%d concurrent Rayleigh numbers in [%.2f, %.2f]" % (d, Ra_s, N_r, N_fm, B, float(Ras.min()), float(Ras.max())),
        "members": B, "krylov_dim": m, "converged": int(info["converged"].sum()),
        "newton_iterations": info["iterations"].cpu().tolist()[:8], "history_member0": info["history"][:, 0].cpu().tolist(),
        "history_last_member": info["history"][:, -1].cpu().tolist(),
        "batched_jvp_calls": info["jvps"], "member_jvps": int(info["member_jvps"].sum()),
        "newton_wall_s": wall_noprof, "member_jvps_per_s_whole_solve": B * info["jvps"] / wall_noprof,
        "jvp_ms_total": prof["matvec_ms"], "ortho_ms_total": prof["ortho_ms"],
        "member_jvps_per_s_inside_jvp_calls": B * prof["matvec_calls"] / (prof["matvec_ms"] * 1e-3),
        "ortho_share_of_device_time": prof["ortho_ms"] / (prof["ortho_ms"] + prof["matvec_ms"]),
        "residual_norm_max": float(torch.linalg.vector_norm(res, dim=1).max()),
        "KE_range": [float(dg[:, 1].min()), float(dg[:, 1].max())],
    }
    pl.close()
    return out


def config5(seeds, B, m):
    l, d, Ra_c, Ra_s, sym = seeds["l11_params"]
    N_fm, N_r = 512, 40
    X0 = upsample(seeds["l11_X"], int(seeds["N_fm"]), int(seeds["N_r"]), N_fm, N_r, d)
    pl = EnsemblePlan(N_fm, N_r, d, 1.0, PR, TAU, symmetric=bool(sym), max_batch=B)
    Ra0 = float(seeds["l11_Ra"])
    # polish the interpolated state on the new grid (one member), then 512 branch points leave it with different steps
    Xp, pinfo = krylov.newton_batched(pl, torch.as_tensor(X0).cuda().reshape(1, -1), Ra0, Ra_s, krylov=m, max_it=8)
    X = Xp.repeat(B, 1)
    ds = torch.as_tensor(np.linspace(0.05, 2.0, B)).cuda()
    sign = torch.full((B,), float(seeds["l11_sign"]), dtype=torch.float64, device="cuda")
    krylov.profile_begin()
    out5, wall = timed(lambda: krylov.continc_batched(pl, X, Ra0, sign, ds, Ra_s, krylov=m))
    prof = krylov.profile_end()
    res = pl.residual(out5["X"], out5["mu"], Ra_s)
    out = {
        "config": "BASELINE configs[4]: pseudo-arc-length step, l=11 set (d=%.5f, Ra_s=%g, Tau=1/15, Pr=1), N_r=%d "
                  "N_theta=%d, %d concurrent branch points (ds in [0.05, 2])" % (d, Ra_s, N_r, N_fm, B),
        "members": B, "krylov_dim": m, "polish_history": pinfo["history"][:, 0].cpu().tolist(),
        "corrector_ok": int(out5["ok"].sum()), "tangent_ok": int(out5["tangent_ok"].sum()),
        "corrector_iterations": out5["iterations"].cpu().tolist()[:8], "halvings_total": int(out5["halvings"].sum()),
        "history_member0": out5["history"][:, 0].cpu().tolist(), "history_last_member": out5["history"][:, -1].cpu().tolist(),
        "mu_range": [float(out5["mu"].min()), float(out5["mu"].max())],
        "batched_jvp_calls": out5["jvps"], "member_jvps": int(out5["member_jvps"].sum()),
        "wall_s": wall, "member_jvps_per_s_whole_step": B * out5["jvps"] / wall,
        "jvp_ms_total": prof["matvec_ms"], "ortho_ms_total": prof["ortho_ms"],
        "member_jvps_per_s_inside_jvp_calls": B * prof["matvec_calls"] / (prof["matvec_ms"] * 1e-3),
        "ortho_share_of_device_time": prof["ortho_ms"] / (prof["ortho_ms"] + prof["matvec_ms"]),
        "residual_norm_max": float(torch.linalg.vector_norm(res, dim=1).max()),
    }
    pl.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--b4", type=int, default=64)
    ap.add_argument("--b5", type=int, default=512)
    ap.add_argument("--krylov", type=int, default=60)
    ap.add_argument("--only", type=int, default=0)
    a = ap.parse_args()
    seeds = dict(np.load(os.path.join(ROOT, "tests", "golden", "branch_seeds.npz")))
    out = {}
    if a.only in (0, 4):
        out["config4"] = config4(seeds, a.b4, a.krylov)
    if a.only in (0, 5):
        out["config5"] = config5(seeds, a.b5, a.krylov)
    out["gram_schmidt"] = {"arnoldi_steps": krylov._FusedOrtho.calls, "third_passes": krylov._FusedOrtho.third_passes}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs45.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))

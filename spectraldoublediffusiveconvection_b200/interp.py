"""Batched resolution transfer on the device: INTERP_RADIAL / INTERP_THETAS of the reference
(Matrix_Operators.py:901-1011) for [B, 3 nr N_fm] ensembles, the step Main.Time_Step / Newton / Continuation and
Gap_Continuation.Gap_Vary run right before the hot path (Main.py:414-416, 599-601, 1093-1095).

Radial.  The reference fits, per (field, mode), np.polyfit(R_o, [0, profile, 0], len(R_o)) in raw powers of r and
evaluates the polynomial on the new interior points.  That least-squares system has one unknown more than equations and a
condition number ~1e13, LAPACK truncates its smallest singular values (the RankWarning the reference silences,
Matrix_Operators.py:6), and the outcome is only determined up to the rounding of that truncated SVD: measured here, a
relative perturbation of 1e-15 of a rough profile moves the result by 2e-5, while for a SMOOTH profile - every physical
state - the procedure returns the interpolating polynomial to ~1e-9.  The device path therefore applies the well-posed
form of the same thing: polynomial interpolation through [0, profile, 0] at the Chebyshev-Gauss-Lobatto points, as a
barycentric interpolation matrix W [nr_n, nr_o] (built once on the host, conditioning O(log N)) applied to all
B * 3 * N_fm profiles by one kernel.  The single-member compat.INTERP_RADIAL keeps the reference's literal polyfit call
(bit-faithful on the golden vector); tests hold the two to each other on a physical state at the level of that state's
own radial truncation error (8e-7 for the N_r = 20 branch seeds interpolated to N_r = 30).

Latitudinal.  The reference's trip through grid space is the identity on every retained coefficient (same midpoint grid
both ways), so the device kernel pads / truncates the spectra directly, keeping the reference's quirk that psi block 0
comes back as zero (see csrc/k_misc.cuh).
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np
import torch

from . import _lib
from .operators import cheb_radial


@functools.lru_cache(maxsize=32)
def radial_matrix(N_n, N_o, d):
    """W [nr_n, nr_o]: new interior values = W @ old interior values, the polynomial of degree N_o through the old
    collocation values (zero at both walls) evaluated at the new interior points, in barycentric form
    (weights (-1)^j, halved at the end points, for Chebyshev-Gauss-Lobatto nodes)."""
    _, R_n = cheb_radial(N_n, d)
    _, R_o = cheb_radial(N_o, d)
    w = np.ones(N_o + 1)
    w[0] = w[-1] = 0.5
    w *= (-1.0) ** np.arange(N_o + 1)
    W = np.zeros((N_n - 1, N_o + 1))
    for a, x in enumerate(R_n[1:-1]):
        diff = x - R_o
        hit = np.nonzero(diff == 0.0)[0]
        if hit.size:
            W[a, hit[0]] = 1.0
        else:
            t = w / diff
            W[a] = t / t.sum()
    return np.ascontiguousarray(W[:, 1:-1])      # the wall values are zero


def _check(X):
    if not (isinstance(X, torch.Tensor) and X.is_cuda and X.dtype == torch.float64 and X.dim() == 2):
        raise TypeError("expected a float64 CUDA tensor [B, 3 nr N_fm]")
    return X.contiguous()


def _stream(X):
    return C.c_void_p(torch.cuda.current_stream(X.device).cuda_stream)


def interp_radial(X, N_n, N_o, d):
    """[B, 3 (N_o-1) K] -> [B, 3 (N_n-1) K] (INTERP_RADIAL for every member)."""
    X = _check(X)
    if N_n == N_o:
        return X
    nr_o, nr_n = N_o - 1, N_n - 1
    if X.shape[1] % (3 * nr_o) != 0:
        raise ValueError("state length %d is not a multiple of 3 * %d" % (X.shape[1], nr_o))
    rows = X.shape[0] * (X.shape[1] // nr_o)
    W = torch.as_tensor(radial_matrix(int(N_n), int(N_o), float(d))).to(X.device)
    out = torch.empty((X.shape[0], rows // X.shape[0] * nr_n), dtype=torch.float64, device=X.device)
    with torch.cuda.device(X.device):
        rc = _lib.load().sddc_interp_radial(X.data_ptr(), out.data_ptr(), W.data_ptr(), rows, nr_o, nr_n, _stream(X))
    if rc:
        raise RuntimeError("sddc_interp_radial failed (%d)" % rc)
    return out


def interp_thetas(X, N_fm_n, N_fm_o):
    """[B, 3 nr K_o] -> [B, 3 nr K_n] (INTERP_THETAS for every member)."""
    X = _check(X)
    if N_fm_n == N_fm_o:
        return X
    if X.shape[1] % (3 * N_fm_o) != 0:
        raise ValueError("state length %d is not a multiple of 3 * %d" % (X.shape[1], N_fm_o))
    nr = X.shape[1] // (3 * N_fm_o)
    out = torch.empty((X.shape[0], 3 * nr * N_fm_n), dtype=torch.float64, device=X.device)
    with torch.cuda.device(X.device):
        rc = _lib.load().sddc_interp_thetas(X.data_ptr(), out.data_ptr(), X.shape[0], int(N_fm_o), int(N_fm_n), nr, _stream(X))
    if rc:
        raise RuntimeError("sddc_interp_thetas failed (%d)" % rc)
    return out

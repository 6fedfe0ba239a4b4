"""GPU-backed drop-in for the reference's Transforms.py (same names, signatures and conventions).

DCT / IDCT / DST / IDST act on the last axis of a NumPy array and return a new array (inputs are never
mutated, Transforms.py:22,34,45,61); `n` zero-pads or truncates like scipy.fftpack's n= argument.
"""
import numpy as np

L = np.pi


def grid(N):
    """theta_i = pi (2i+1) / (2N)  (Transforms.py:8-13)."""
    dx = L / N
    return np.asarray([dx * (2.0 * i + 1.0) / 2 for i in range(N)])


def _run(kind, x, n, axis):
    import torch
    from .. import plan as P
    x = np.asarray(x, dtype=np.float64)
    if axis not in (-1, x.ndim - 1):
        x = np.moveaxis(x, axis, -1)
    out = P.transform(kind, torch.as_tensor(np.ascontiguousarray(x)).cuda(), n).cpu().numpy()
    if axis not in (-1, x.ndim - 1):
        out = np.moveaxis(out, -1, axis)
    return out


def DST(g, n=None, axis=-1):
    from ..plan import T_DST
    return _run(T_DST, g, n, axis)


def IDST(g_hat, n=None, axis=-1):
    from ..plan import T_IDST
    return _run(T_IDST, g_hat, n, axis)


def DCT(f, n=None, axis=-1):
    from ..plan import T_DCT
    return _run(T_DCT, f, n, axis)


def IDCT(f_hat, n=None, axis=-1):
    from ..plan import T_IDCT
    return _run(T_IDCT, f_hat, n, axis)

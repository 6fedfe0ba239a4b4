"""Host side of the batched resolution transfer (interp.py): the barycentric interpolation matrix."""
import warnings

import numpy as np

from conftest import load_golden
from spectraldoublediffusiveconvection_b200.interp import radial_matrix
from spectraldoublediffusiveconvection_b200.operators import cheb_radial


def test_radial_matrix_is_exact_on_polynomials():
    d, N_o, N_n = 0.4, 12, 17
    _, Ro = cheb_radial(N_o, d)
    _, Rn = cheb_radial(N_n, d)
    W = radial_matrix(N_n, N_o, d)
    assert W.shape == (N_n - 1, N_o - 1)
    q = lambda r: (r - Ro[0]) * (Ro[-1] - r)              # vanishes at both walls
    for deg in range(0, N_o - 1):
        f = lambda r: q(r) * (r - 2.7) ** deg
        assert np.allclose(W @ f(Ro[1:-1]), f(Rn[1:-1]), rtol=1e-11, atol=1e-13)
    assert np.allclose(radial_matrix(N_o, N_o, d), np.eye(N_o - 1))       # same grid: identity


def test_radial_matrix_agrees_with_the_reference_fit_on_a_physical_state():
    """np.polyfit in raw powers (what INTERP_RADIAL does, Matrix_Operators.py:919-935) and the barycentric form agree on a
    smooth state to the state's own truncation error; on rough data the reference's fit is rounding noise (see interp.py)."""
    sd = load_golden("branch_seeds")
    d = float(sd["l10_params"][1])
    X = sd["l10_X"].reshape(-1, 19)
    _, Ro = cheb_radial(20, d)
    _, Rn = cheb_radial(30, d)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = np.stack([np.polyval(np.polyfit(Ro, np.hstack(([0.0], r, [0.0])), len(Ro)), Rn[1:-1]) for r in X])
    mine = X @ radial_matrix(30, 20, d).T
    assert np.linalg.norm(mine - ref) < 1e-5 * np.linalg.norm(ref)

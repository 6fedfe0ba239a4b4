"""Batched resolution transfer on the device (interp.py, sddc_interp_radial / sddc_interp_thetas) against golden outputs of
the reference's INTERP_THETAS and against the host paths."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()


def test_interp_thetas_matches_reference():
    from spectraldoublediffusiveconvection_b200 import interp
    g = load_golden("interp")
    K = int(g["N_fm"])
    X = _dev(np.stack([g["X"], 2.0 * g["X"], -g["X"]]))
    up = interp.interp_thetas(X, 32, K).cpu().numpy()
    dn = interp.interp_thetas(X, 8, K).cpu().numpy()
    assert rel_l2(up[0], g["theta_up"]) < 1e-12 and rel_l2(dn[0], g["theta_down"]) < 1e-12
    assert np.array_equal(up[1], 2.0 * up[0]) and np.array_equal(dn[2], -dn[0])     # members are independent
    assert interp.interp_thetas(X, K, K) is X or torch.equal(interp.interp_thetas(X, K, K), X)


def test_interp_radial_matches_host_paths():
    import spectraldoublediffusiveconvection_b200.compat.Matrix_Operators as MO
    from spectraldoublediffusiveconvection_b200 import interp
    sd = load_golden("branch_seeds")
    d = float(sd["l10_params"][1])
    X = np.stack([sd["l10_X"], 0.5 * sd["l10_X"]])
    out = interp.interp_radial(_dev(X), 30, 20, d).cpu().numpy()
    W = interp.radial_matrix(30, 20, d)
    assert rel_l2(out[0], (X[0].reshape(-1, 19) @ W.T).reshape(-1)) < 1e-14        # the kernel applies the matrix
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = MO.INTERP_RADIAL(30, 20, sd["l10_X"], d)                           # the reference's literal polyfit path
    assert rel_l2(out[0], ref) < 1e-5
    assert np.allclose(out[1], 0.5 * out[0], rtol=1e-14)
    # the interpolated state is a good Newton start on the finer grid: both transfers, then one residual evaluation
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    Xf = interp.interp_thetas(interp.interp_radial(_dev(X[:1]), 30, 20, d), 256, 64)
    pl = EnsemblePlan(256, 30, d, 1.0, 1.0, 1.0 / 15.0, symmetric=True, max_batch=1)
    r = pl.residual(Xf, float(sd["l10_Ra"]), float(sd["l10_params"][3]))
    assert float(torch.linalg.vector_norm(r) / torch.linalg.vector_norm(Xf)) < 1e-2
    pl.close()

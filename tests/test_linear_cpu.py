"""Linear-theory seeding (spectraldoublediffusiveconvection_b200/linear.py) against vectors produced by the unmodified
reference (tests/golden/make_golden_linear.py -> linear.npz): eigenvalues, eigenvectors, seeded states, critical Ra."""
import os

import numpy as np
import pytest

from spectraldoublediffusiveconvection_b200 import linear

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "linear.npz"))


def _params(name):
    l, d, Ra, Ra_s, Pr, Tau = G[name + "_params"]
    return dict(l=l, d=d, Ra=Ra, Ra_s=Ra_s, Pr=Pr, Tau=Tau)


@pytest.mark.parametrize("name", ["l10", "l11"])
@pytest.mark.parametrize("Nr", [20, 30])
def test_eigenvalues_and_vectors(name, Nr):
    p = _params(name)
    tag = "%s_Nr%d" % (name, Nr)
    vals = linear.eig_vals(p["Ra"], p["l"], p["d"], 4, Ra_s=p["Ra_s"], Pr=p["Pr"], Tau=p["Tau"], Nr=Nr)
    ref = G[tag + "_vals"]
    # growth rates of a cond ~ 1e8 generalised eigenproblem: compare on the scale of the spectrum's leading entries
    assert np.max(np.abs(vals - ref)) <= 1e-6 * max(1.0, np.max(np.abs(ref)))
    for k in (0, 1):
        v = linear.eig_vec(p["Ra"], p["l"], p["d"], k, Ra_s=p["Ra_s"], Pr=p["Pr"], Tau=p["Tau"], Nr=Nr)
        r = np.asarray(G[tag + "_vec%d" % k]).ravel()
        assert v.shape == r.shape
        assert np.linalg.norm(v - r) <= 1e-7 * np.linalg.norm(r)     # same sign, same normalisation (LAPACK dgeev)


def test_marginal_mode_is_marginal():
    p = _params("l10")
    # Ra_c = 9851.537357677651 is the steady onset of l = 10 at d = 0.3521, Ra_s = 500 (Linear_Problem.py:481-488)
    lam = linear.eig_vals(p["Ra"], p["l"], p["d"], 1, Ra_s=p["Ra_s"], Pr=p["Pr"], Tau=p["Tau"], Nr=30)
    assert abs(lam) < 1e-6


@pytest.mark.parametrize("name", ["l10", "l11"])
@pytest.mark.parametrize("K", [48, 64])
@pytest.mark.parametrize("sym", [False, True])
def test_full_eigenvector(name, K, sym):
    p = _params(name)
    f = np.asarray(G[name + "_Nr20_vec1"]).ravel()
    X = linear.full_eig_vec(f, p["l"], K, 19, symmetric=sym)
    ref = G["%s_full_K%d_%s" % (name, K, "sym" if sym else "nosym")]
    assert X.shape == ref.shape
    # scale: the unmasked state (an odd-l eigenfunction is equatorially antisymmetric, so its symmetric projection is
    # rounding noise only)
    scale = np.linalg.norm(G["%s_full_K%d_nosym" % (name, K)])
    assert np.linalg.norm(X - ref) <= 1e-12 * scale
    # a degree-l eigenfunction only excites latitudinal modes up to l (psi: sine modes <= l, i.e. blocks < l)
    X3 = X.reshape(3, K, 19)
    assert np.abs(X3[:, int(p["l"]) + 1:, :]).max() <= 1e-12 * scale


def test_critical_rayleigh():
    assert abs(linear.critical_rayleigh(4525.9, 11.0, 0.31325) - G["crit_l11"][0]) <= 1e-6 * G["crit_l11"][0]
    assert abs(linear.critical_rayleigh(4619.4, 13.0, 0.31325) - G["crit_l13"][0]) <= 1e-6 * G["crit_l13"][0]


def test_seed_state_symmetry_follows_degree():
    p = _params("l10")
    X = linear.seed_state(p["l"], p["d"], p["Ra"], p["Ra_s"], p["Pr"], p["Tau"], 32, 12, amplitude=1e-2)
    X3 = X.reshape(3, 32, 11)
    assert abs(np.linalg.norm(X) - 1e-2) < 1e-15
    assert np.all(X3[0, 0::2] == 0.0) and np.all(X3[1:, 1::2] == 0.0)       # even l: equatorially symmetric
    with pytest.raises(ValueError):
        linear.full_eig_vec(np.zeros(10), 10.0, 32, 11)

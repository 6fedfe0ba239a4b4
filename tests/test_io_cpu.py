"""Checkpoint I/O in the reference's schema (io.py): files are read back with the very access expressions the
reference's loaders and Plot_Tools use (Main.py:386-410, 572-593, 1064-1088; Plot_Tools.py:24-28, 162-168, 384-387)."""
import sys

import numpy as np
import pytest
import torch

from spectraldoublediffusiveconvection_b200 import io as sio
from spectraldoublediffusiveconvection_b200.krylov import BranchResult


def _reference_style_load(h5py, filename, frame):
    """The statements of Main.Continuation's loader (Main.py:1064-1088), verbatim in structure."""
    f = h5py.File(filename, 'r')
    X = f['Checkpoints/X_DATA'][frame]
    try:
        Ra = f['Checkpoints/Ra_DATA'][frame]
    except Exception:
        Ra = f['Parameters']["Ra"][()]
    out = dict(X=X, Ra=Ra, Ra_s=f['Parameters']["Ra_s"][()], Tau=f['Parameters']["Tau"][()], Pr=f['Parameters']["Pr"][()],
               d=f['Parameters']["d"][()], N_fm=f['Parameters']["N_fm"][()], N_r=f['Parameters']["N_r"][()])
    f.close()
    return out


@pytest.fixture()
def h5(monkeypatch):
    saved = sys.modules.get("h5py")
    mod = sio.install_h5py_shim()
    yield mod
    if saved is None and sio.h5py is None:
        sys.modules.pop("h5py", None)


def test_time_step_file_is_readable_the_reference_way(tmp_path, h5):
    rng = np.random.default_rng(0)
    states, hist = rng.random((4, 2, 12)), rng.random((9, 2, 6))
    common = {"Tau": 1.0, "Pr": 1.0, "d": 0.3, "N_r": 3, "N_fm": 2, "dt": 1e-3, "start_time": 0, "symmetric": False}
    files = sio.save_ensemble(str(tmp_path / "TimeStep"), states, hist, np.arange(9) * 1e-3, [3750.0, 4000.0], [0.0, 10.0], common)
    got = _reference_style_load(h5, str(tmp_path / "TimeStep_1.h5"), -1)
    assert np.array_equal(got["X"], states[-1, 1]) and got["Ra"] == 4000.0 and got["Ra_s"] == 10.0
    assert got["N_fm"] == 2 and got["N_r"] == 3 and got["d"] == 0.3
    f = h5.File(str(tmp_path / "TimeStep_0.h5"), 'r')
    assert np.array_equal(f['Scalar_Data/KE'][()][2:-1], hist[2:-1, 0, 1])          # Plot_Tools.py:163-166 slicing
    assert np.array_equal(f['Scalar_Data/Time'][()], np.arange(9) * 1e-3)
    X, p = sio.load_state(files[0], frame=1)
    assert np.array_equal(X, states[1, 0]) and p["Ra"] == 3750.0 and isinstance(p["N_fm"], int)


def test_branch_file_carries_the_bifurcation_group(tmp_path, h5):
    res = BranchResult(2)
    for it in range(7):
        for name, val in (("Ra", 100.0 + it), ("Ra_dot", 1.0), ("Norm", 0.1 * it), ("KE", 0.01 * it), ("NuT", it), ("NuS", -it)):
            getattr(res, name).append(torch.tensor([val, 2 * val], dtype=torch.float64))
        if it % 5 == 0:
            res.X_DATA.append(torch.full((2, 6), float(it), dtype=torch.float64))
            res.Ra_DATA.append(torch.tensor([100.0 + it, 200.0 + 2 * it], dtype=torch.float64))
        res.Iterations += 1
    res.folds[1].append((3, 206.0, torch.arange(6, dtype=torch.float64)))
    params = {"Ra": 100.0, "Ra_s": 150.0, "Tau": 1 / 15, "Pr": 1.0, "d": 0.31325, "N_fm": 2, "N_r": 2, "symmetric": False}
    sio.save_branch(str(tmp_path / "Continuation_1.h5"), res, 1, params)
    got = _reference_style_load(h5, str(tmp_path / "Continuation_1.h5"), -1)
    assert got["Ra"] == 210.0 and np.array_equal(got["X"], np.full(6, 5.0))           # Ra from Checkpoints/Ra_DATA
    with h5.File(str(tmp_path / "Continuation_1.h5"), 'r') as f:                      # Plot_Tools.py:384-387
        ff = f["Bifurcation"]
        obj = {key: ff[key][()] for key in ff.keys()}
    assert set(obj) >= {"Ra", "Ra_dot", "Norm", "KE", "NuT", "NuS", "Y_FOLD", "X_DATA", "Ra_DATA", "Iterations"}
    assert np.array_equal(obj["Ra"], 2 * (100.0 + np.arange(7))) and int(obj["Iterations"]) == 7
    assert obj["Y_FOLD"].shape == (1, 7) and obj["Y_FOLD"][0, -1] == 206.0


def test_shim_writes_what_the_reference_writes(tmp_path, h5):
    """The write side of the shim with the statements of Main.Newton's save block (Main.py:652-668)."""
    f = h5.File(str(tmp_path / "NewtonSolve_0.h5"), 'w')
    Checkpoints = f.create_group("Checkpoints")
    Checkpoints['X_DATA'] = [np.arange(4.0)]
    Scalar_Data = f.create_group("Scalar_Data")
    Scalar_Data['KE'] = [0.5]
    Parameters = f.create_group("Parameters")
    for key, val in {"Ra": 1.0, "N_fm": 2}.items():
        Parameters[key] = val
    f.close()
    X, p = sio.load_state(str(tmp_path / "NewtonSolve_0.h5"))
    assert np.array_equal(X, np.arange(4.0)) and p["Ra"] == 1.0 and p["N_fm"] == 2
    sio.save_newton(str(tmp_path / "NewtonSolve_1.h5"), np.arange(4.0), 1.0, 0.5, 0.1, 0.2, {"Ra": 2.0, "N_fm": 2, "N_r": 3})
    X2, p2 = sio.load_state(str(tmp_path / "NewtonSolve_1.h5"))
    assert np.array_equal(X2, X) and p2["Ra"] == 2.0

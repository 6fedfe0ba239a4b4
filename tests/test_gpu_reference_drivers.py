"""The UNMODIFIED reference drivers on the drop-in boundary: `compat.install(); import Main` and then Main._Time_Step,
Main._Newton and Main._Continuation exactly as the reference ships them, every Matrix_Operators / Transforms call
landing on the CUDA path.  Needs the reference's source tree, which is not part of this repo and does not exist on the
GPU box: the tests run only when SDDC_REFERENCE_DIR (or a staged copy under baseline/_ref, see tools/stage_reference.sh)
points at it, and skip otherwise.  Expected values are the golden vectors the same unmodified drivers produced on the
reference's own NumPy/SciPy operators (tests/golden/make_golden.py, make_golden_continuation.py)."""
import contextlib
import io
import os
import re
import sys
import types

import numpy as np
import pytest

from conftest import ROOT, load_golden, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _reference_dir():
    for cand in (os.environ.get("SDDC_REFERENCE_DIR"), os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.exists(os.path.join(cand, "Main.py")):
            return cand
    return None


REF = _reference_dir()
if REF is None:
    pytest.skip("reference source tree not staged (SDDC_REFERENCE_DIR / baseline/_ref)", allow_module_level=True)

FILES = []


class _H5(dict):
    """Minimal stand-in for h5py.File (h5py is not installed in this image): records what the drivers write."""

    def __init__(self, *a, **k):
        super().__init__()
        if a:
            FILES.append(self)

    def create_group(self, name):
        g = _H5()
        self[name] = g
        return g

    def create_dataset(self, name, data=None, **k):
        self[name] = data

    def close(self):
        pass


@pytest.fixture(scope="module")
def Main():
    import spectraldoublediffusiveconvection_b200.compat as compat
    if not hasattr(np, "RankWarning"):
        np.RankWarning = np.exceptions.RankWarning          # removed in NumPy 2; Main.py:9 references it
    saved = {k: sys.modules.get(k) for k in ("Matrix_Operators", "Transforms", "Main", "h5py", "matplotlib", "matplotlib.pyplot")}
    MO, TR = compat.install()
    h5 = types.ModuleType("h5py")
    h5.File = _H5
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.update({"h5py": h5, "matplotlib": mpl, "matplotlib.pyplot": plt})
    sys.modules.pop("Main", None)
    sys.path.insert(0, REF)
    try:
        import Main as M
        assert os.path.dirname(os.path.abspath(M.__file__)) == os.path.abspath(REF)
        assert sys.modules["Matrix_Operators"] is MO      # the drivers resolve the operator layer by bare module name
        yield M
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_time_step_config1(Main):
    """BASELINE config 1: Main._Time_Step with the literals of Main.Time_Step, 100 steps from the seeded state."""
    g = load_golden("cfg1_nosym")
    FILES.clear()
    with contextlib.redirect_stdout(io.StringIO()):
        X = Main._Time_Step(g["X0"].copy(), float(g["Ra"]), float(g["Ra_s"]), float(g["Tau"]), float(g["Pr"]), float(g["d"]),
                            int(g["N_fm"]), int(g["N_r"]), False, save_filename="TimeStep_0.h5", start_time=0,
                            Total_time=0.1, dt=float(g["dt"]), linear=False, Verbose=False)
    assert rel_l2(X, g["X_step100"]) < 1e-10
    sd = FILES[-1]["Scalar_Data"]     # the stub keeps the drivers' own lists (they go on growing after the last save)
    hist = np.stack([sd["Norm"], sd["KE"], sd["Nu_T"], sd["Nu_S"]], axis=1)
    assert hist.shape == (100, 4)
    assert np.allclose(hist, g["diag_hist"], rtol=1e-9, atol=0)
    assert len(FILES[-1]["Checkpoints"]["X_DATA"]) == 10


def test_newton_unmodified(Main):
    G = load_golden("continuation")
    kw = dict(Ra_s=float(G["Ra_s"]), Tau=float(G["Tau"]), Pr=float(G["Pr"]), d=float(G["d"]), N_fm=int(G["N_fm"]),
              N_r=int(G["N_r"]), symmetric=bool(G["symmetric"]))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        X, Norm, KE, NuT, NuS, ok = Main._Newton(G["X_start"].copy(), float(G["Ra_newton"]), **kw)
    assert ok
    hist = np.array([float(x) for x in re.findall(r"Newton Iteration = \d+, Error = ([0-9.eE+-]+)", buf.getvalue())])
    ref = G["newton_history"]
    assert len(hist) == len(ref) and np.max(np.abs(hist - ref)) < 1e-10, (hist, ref)
    assert rel_l2(X, G["newton_X"]) < 1e-9
    assert np.allclose([Norm, KE, NuT, NuS], G["newton_diag"], rtol=1e-8)


def test_continuation_unmodified(Main):
    """Main._Continuation (arc-length steps, then natural-parameter Newton steps) for 12 iterations."""
    G = load_golden("continuation")
    kw = dict(Ra=float(G["Ra_newton"]), Ra_s=float(G["Ra_s"]), Tau=float(G["Tau"]), Pr=float(G["Pr"]), d=float(G["d"]),
              N_fm=int(G["N_fm"]), N_r=int(G["N_r"]), symmetric=bool(G["symmetric"]))
    Y0 = np.hstack((G["newton_X"], float(G["Ra_newton"])))
    FILES.clear()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        Main._Continuation("golden_branch.h5", int(G["branch_steps"]), 1.0, Y0, **kw)
    txt = buf.getvalue()
    Ra = np.array([float(x) for x in re.findall(r"\nRa\s+= ([0-9.eE+-]+)", txt)])
    KE = np.array([float(x) for x in re.findall(r"\nKE\s+= ([0-9.eE+-]+)", txt)])
    assert np.allclose(Ra, G["branch_Ra"], rtol=1e-9), (Ra, G["branch_Ra"])
    assert np.allclose(KE, G["branch_KE"], rtol=1e-6)
    last = FILES[-1]
    assert np.allclose(np.array(last["Checkpoints"]["Ra_DATA"]), G["branch_Ra_DATA"], rtol=1e-9)
    assert rel_l2(np.array(last["Checkpoints"]["X_DATA"])[-1], G["branch_X_DATA"][-1]) < 1e-7
    print("unmodified Main._Continuation on the CUDA path: Ra =", Ra)

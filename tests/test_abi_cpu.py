"""CPU checks of the boundary: the shared library loads and exports every symbol include/sddc_b200.h declares;
host-side operator construction matches the reference's; the product path fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, rel_l2


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sddc_b200.h")).read()
    return sorted(set(re.findall(r"\b(sddc_[a-zA-Z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from spectraldoublediffusiveconvection_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.sddc_version() >= 100


def test_host_operator_build_matches_reference():
    from spectraldoublediffusiveconvection_b200.operators import RadialOperators
    g = load_golden("small_nosym")
    op = RadialOperators(int(g["N_fm"]), int(g["N_r"]), float(g["d"]), float(g["dt"]), float(g["Pr"]), float(g["Tau"]))
    assert np.array_equal(op.D, g["op_D"]) and np.array_equal(op.R, g["op_R"])
    assert rel_l2(op.D2, g["op_D2"]) < 1e-15
    assert rel_l2(op.dT0, g["op_DT0"]) < 1e-16
    assert np.array_equal(np.diag(op.IR2), np.diag(g["op_IR2"]))
    assert rel_l2(op.L_inv_A4, g["op_L4"]) < 1e-9
    assert rel_l2(op.L_inv_T, g["op_LT"]) < 1e-12
    assert rel_l2(op.L_inv_S, g["op_LS"]) < 1e-12


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    with pytest.raises(RuntimeError):
        EnsemblePlan(16, 10, 0.4, 1e-2, 1.0, 1.0)


def test_odd_mode_count_raises_valueerror_before_touching_the_gpu():
    from spectraldoublediffusiveconvection_b200.compat import Matrix_Operators as MO
    with pytest.raises(ValueError):
        MO._check_even(15)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "spectraldoublediffusiveconvection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("test-only oracle", ""), f


def test_checkpoint_schema_roundtrip(tmp_path):
    """Output keeps the reference's HDF5 key paths (Main.py:305-321), with an .npz fallback when h5py is absent."""
    from spectraldoublediffusiveconvection_b200 import io as sio
    rng = np.random.default_rng(0)
    states = rng.random((2, 3, 12))
    hist = rng.random((5, 3, 6))
    files = sio.save_ensemble(str(tmp_path / "TimeStep"), states, hist, np.arange(5) * 1e-3, [1.0, 2.0, 3.0], [0.0] * 3,
                              {"Tau": 1.0, "Pr": 1.0, "d": 0.3, "N_r": 3, "N_fm": 2, "dt": 1e-3, "start_time": 0,
                               "symmetric": False})
    assert len(files) == 3
    back = sio.load_time_step(files[1])
    assert np.array_equal(back["Checkpoints/X_DATA"], states[:, 1])
    assert np.array_equal(back["Scalar_Data/KE"], hist[:, 1, 1])
    assert float(back["Parameters/Ra"]) == 2.0
    for key in ("Scalar_Data/Norm", "Scalar_Data/Nu_T", "Scalar_Data/Nu_S", "Scalar_Data/Time", "Parameters/N_fm"):
        assert key in back


def test_plan_validates_parameters_and_outputs_before_the_abi():
    """Per-member parameter lengths and caller-supplied `out` tensors are checked in Python: the C ABI trusts pointers."""
    import torch
    from spectraldoublediffusiveconvection_b200.plan import EnsemblePlan
    pl = EnsemblePlan.__new__(EnsemblePlan)          # no library / device needed for the host-side checks
    pl.device = torch.device("cpu")
    assert pl._param(3.0, 4).tolist() == [3.0] * 4
    assert pl._param(torch.tensor([2.0]), 3).tolist() == [2.0] * 3
    assert pl._param(np.arange(3.0), 3).tolist() == [0.0, 1.0, 2.0]
    with pytest.raises(ValueError):
        pl._param(torch.arange(5.0, dtype=torch.float64), 4)       # a globally sized Ra with a local shard
    # (_out insists on CUDA tensors: everything else is rejected before a pointer is taken)
    with pytest.raises(TypeError):
        pl._out(torch.zeros((2, 6), dtype=torch.float64), (2, 6))
    with pytest.raises(TypeError):
        pl._out(np.zeros((2, 6)), (2, 6))

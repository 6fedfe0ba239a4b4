"""Checkpoint / diagnostics output in the reference's schema (Main.py:305-321, 652-668, 1026-1040).

Groups and datasets keep the reference's key paths -- Checkpoints/X_DATA, Scalar_Data/{Norm,KE,Nu_T,Nu_S,Time},
Parameters/{Ra,Ra_s,Tau,Pr,d,N_r,N_fm,dt,start_time,symmetric} -- so Plot_Tools.py and Paper_Figures/* can read an
ensemble member's file unchanged.  h5py is used when it is installed; otherwise the same key paths are written to a
NumPy .npz archive (h5py is absent from the build image).
"""
from __future__ import annotations

import numpy as np

try:  # optional dependency
    import h5py  # type: ignore
except Exception:  # pragma: no cover - depends on the environment
    h5py = None


def _flatten(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        key = prefix + k
        if isinstance(v, dict):
            out.update(_flatten(v, key + "/"))
        else:
            out[key] = np.asarray(v)
    return out


def save_time_step(filename, X_DATA, Norm, KE, Nu_T, Nu_S, Time, parameters):
    """Write one member's time-stepping output. X_DATA: [n_checkpoints, 3N]; the scalar series are 1-D."""
    tree = {"Checkpoints": {"X_DATA": np.asarray(X_DATA)},
            "Scalar_Data": {"Norm": Norm, "KE": KE, "Nu_T": Nu_T, "Nu_S": Nu_S, "Time": Time},
            "Parameters": dict(parameters)}
    if h5py is not None and str(filename).endswith(".h5"):
        with h5py.File(filename, "w") as f:
            for gname, grp in tree.items():
                g = f.create_group(gname)
                for k, v in grp.items():
                    g[k] = v
        return filename
    out = str(filename)
    if out.endswith(".h5"):
        out = out[:-3] + ".npz"
    np.savez(out, **_flatten(tree))
    return out


def save_ensemble(prefix, states, diag_hist, times, Ra, Ra_s, common):
    """One file per member from the outputs of EnsemblePlan.time_step_host: states [n_ckpt, B, 3N],
    diag_hist [n_rec, B, >=4], times [n_rec]."""
    files = []
    for m in range(states.shape[1]):
        p = dict(common)
        p.update({"Ra": float(Ra[m]), "Ra_s": float(Ra_s[m])})
        files.append(save_time_step("%s_%d.h5" % (prefix, m), states[:, m], diag_hist[:, m, 0], diag_hist[:, m, 1],
                                    diag_hist[:, m, 2], diag_hist[:, m, 3], times, p))
    return files


def load_time_step(filename):
    """Read back either flavour as a flat {key path: array} dict."""
    if h5py is not None and str(filename).endswith(".h5"):
        out = {}
        with h5py.File(filename, "r") as f:
            f.visititems(lambda name, obj: out.__setitem__(name, obj[()]) if hasattr(obj, "shape") else None)
        return out
    return dict(np.load(filename, allow_pickle=False))


# ---------------------------------------------------------------------------------------------- reading, h5py-style
class _Node:
    """Group / dataset view over a flat {key path: array} dict with the access patterns the reference uses:
    f['Checkpoints/X_DATA'][frame], f['Parameters']['Ra'][()], f['Scalar_Data/Time'][()], f['Bifurcation'].keys()."""

    def __init__(self, flat, prefix=""):
        self._flat, self._prefix = flat, prefix

    def __getitem__(self, key):
        if isinstance(key, str):
            path = self._prefix + key.strip("/")
            if path in self._flat:
                return _Dataset(self._flat[path])
            if any(k.startswith(path + "/") for k in self._flat):
                return _Node(self._flat, path + "/")
            raise KeyError(path)
        raise TypeError("groups are indexed by name")

    def keys(self):
        n = len(self._prefix)
        return sorted({k[n:].split("/")[0] for k in self._flat if k.startswith(self._prefix)})

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False


class _Dataset:
    def __init__(self, a):
        self._a = np.asarray(a)
        self.shape = self._a.shape

    def __getitem__(self, idx):
        return self._a[idx] if idx != () else (self._a[()] if self._a.ndim == 0 else self._a)

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __len__(self):
        return len(self._a)


class CheckpointFile(_Node):
    """Read-only file object over either flavour (.h5 through h5py when installed, else the .npz written above)."""

    def __init__(self, filename, mode="r"):
        super().__init__(load_time_step(_resolve(filename)))

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _resolve(filename):
    import os
    name = str(filename)
    if name.endswith(".h5") and (h5py is None or not os.path.exists(name)) and os.path.exists(name[:-3] + ".npz"):
        return name[:-3] + ".npz"
    return name


def load_state(filename, frame=-1):
    """What Main.Time_Step / Newton / Continuation read before they start (Main.py:386-410, 572-593, 1064-1088):
    the checkpointed state of `frame` and the run's parameters (Ra from Checkpoints/Ra_DATA when a branch file has it)."""
    f = CheckpointFile(filename)
    X = np.array(f["Checkpoints/X_DATA"][frame], dtype=np.float64)
    p = {k: f["Parameters"][k][()] for k in f["Parameters"].keys()}
    if "Ra_DATA" in f["Checkpoints"]:
        p["Ra"] = f["Checkpoints/Ra_DATA"][frame]
    for k in ("N_fm", "N_r"):
        if k in p:
            p[k] = int(p[k])
    return X, p


def _write(filename, tree):
    if h5py is not None and str(filename).endswith(".h5"):
        with h5py.File(filename, "w") as f:
            for gname, grp in tree.items():
                g = f.create_group(gname)
                for k, v in grp.items():
                    g.create_dataset(k, data=v)
        return filename
    out = str(filename)
    if out.endswith(".h5"):
        out = out[:-3] + ".npz"
    np.savez(out, **_flatten(tree))
    return out


def save_newton(filename, X, Norm, KE, Nu_T, Nu_S, parameters):
    """A converged steady state in the layout of Main.Newton (Main.py:652-668)."""
    return _write(filename, {"Checkpoints": {"X_DATA": np.asarray([X])},
                             "Scalar_Data": {"Norm": [Norm], "KE": [KE], "Nu_T": [Nu_T], "Nu_S": [Nu_S], "Time": [0.0]},
                             "Parameters": dict(parameters)})


def save_branch(filename, result, member, parameters):
    """One member of a krylov.BranchResult in the layout of Main._Continuation (Main.py:1026-1040): Checkpoints/X_DATA,
    Checkpoints/Ra_DATA (every 5th iteration), Parameters, and the Bifurcation group with the `result` attributes
    (Ra, Ra_dot, Norm, KE, NuT, NuS, Y_FOLD, X_DATA, Ra_DATA, Iterations) that Plot_Tools._plot_bif reads (Main.py:384-387)."""
    col = lambda series: np.array([float(s[member]) for s in series])
    X_DATA = np.array([x[member].detach().cpu().numpy() for x in result.X_DATA])
    Ra_DATA = col(result.Ra_DATA)
    folds = result.folds[member]
    Y_FOLD = np.array([np.hstack((X.detach().cpu().numpy(), ra)) for (_, ra, X) in folds]) if folds else np.zeros((0,))
    bif = {"Ra": col(result.Ra), "Ra_dot": col(result.Ra_dot), "Norm": col(result.Norm), "KE": col(result.KE),
           "NuT": col(result.NuT), "NuS": col(result.NuS), "Y_FOLD": Y_FOLD, "X_DATA": X_DATA, "Ra_DATA": Ra_DATA,
           "Iterations": result.Iterations}
    return _write(filename, {"Checkpoints": {"X_DATA": X_DATA, "Ra_DATA": Ra_DATA}, "Parameters": dict(parameters),
                             "Bifurcation": bif})


def install_h5py_shim():
    """Register a minimal `h5py` module (File = CheckpointFile for reading; writing through a dict-like recorder that is
    stored with _write on close) when the real package is absent, so that the reference's loaders and Plot_Tools can
    open the files written here.  No effect when h5py is installed."""
    import sys
    import types
    if h5py is not None:
        return sys.modules["h5py"]

    class _WGroup(dict):
        def create_group(self, name):
            g = _WGroup()
            self[name] = g
            return g

        def create_dataset(self, name, data=None, **k):
            self[name] = data

    class File(_WGroup):
        def __new__(cls, filename, mode="r"):
            if mode.startswith("r"):
                return CheckpointFile(filename)
            return super().__new__(cls)

        def __init__(self, filename, mode="r"):
            super().__init__()
            self._filename = filename

        def close(self):
            _write(self._filename, {k: dict(v) for k, v in self.items()})

        def __enter__(self):
            return self

        def __exit__(self, *a):
            self.close()
            return False

    mod = types.ModuleType("h5py")
    mod.File = File
    sys.modules["h5py"] = mod
    return mod

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

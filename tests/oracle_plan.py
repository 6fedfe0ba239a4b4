"""Test double of EnsemblePlan backed by the CPU oracle: the surface the batched Krylov drivers (krylov.py) use --
residual / jvp_set_base / jvp_apply / dF_dRa / diagnostics on [B, 3N] torch CPU tensors -- so that the drivers' host
logic (masks, step-size rules, branch loop) is tested without a GPU against golden runs of the reference's drivers."""
import numpy as np
import torch

from oracle import sddc_oracle as orc


class OraclePlan:
    def __init__(self, N_fm, N_r, d, dt, Pr, Tau, symmetric=False):
        self.op = orc.Operators(N_fm, N_r, d, dt, Pr, Tau)
        self.N_fm, self.N_r, self.nr = N_fm, N_r, N_r - 1
        self.N = self.nr * N_fm
        self.symmetric = bool(symmetric)
        self.jvps = 0
        self._base = None

    def _param(self, v, B):
        v = torch.as_tensor(v, dtype=torch.float64).reshape(-1)
        if v.numel() == 1:
            v = v.expand(B)
        elif v.numel() != B:
            raise ValueError("per-member parameter has %d entries, the batch has %d members" % (v.numel(), B))
        return v.contiguous()

    def _rows(self, fn, *tensors):
        arrs = [t.detach().numpy() for t in tensors]
        return torch.as_tensor(np.stack([fn(*(a[m] for a in arrs)) for m in range(arrs[0].shape[0])]))

    def residual(self, X, Ra, Ra_s):
        B = X.shape[0]
        Ra, Ra_s = self._param(Ra, B), self._param(Ra_s, B)
        return self._rows(lambda x, ra, ras: orc.residual(x, self.op, float(ra), float(ras), self.symmetric), X, Ra, Ra_s)

    def jvp_set_base(self, X):
        self._base = X.clone()

    has_jvp_plus = True

    def jvp_apply(self, dv, Ra, Ra_s, plus_identity=False):
        B = dv.shape[0]
        Ra, Ra_s = self._param(Ra, B), self._param(Ra_s, B)
        self.jvps += B
        out = self._rows(lambda v, x, ra, ras: orc.jvp(v, x, self.op, float(ra), float(ras), self.symmetric),
                         dv, self._base, Ra, Ra_s)
        return out + dv if plus_identity else out

    def dF_dRa(self, X):
        return self._rows(lambda x: orc.dF_dRa(x, self.op, self.symmetric), X)

    def diagnostics(self, X):
        out = torch.zeros((X.shape[0], 6), dtype=torch.float64)
        out[:, :4] = self._rows(lambda x: orc.diagnostics(x, self.op, self.symmetric), X)
        return out

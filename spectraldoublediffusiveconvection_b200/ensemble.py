"""Ensemble driver: many independent members (random initial conditions, Ra_T / Ra_S sweeps) advanced in lock
step, sharded by member over the GPUs of one node.

Members never exchange state, so there is no data-path collective; the only communication is the all-gather of
the per-step diagnostics [B_local, 4] -> [B, 4] (one collective per run for the whole history) (Norm, KE, Nu_T, Nu_S: the four scalars Main._Time_Step appends
per step, Main.py:292-295) and, on request, of the final states.  One process per GPU (torchrun); with a single
process everything degenerates to the local plan.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition(n_members: int, world: int, rank: int):
    """Contiguous shard [lo, hi) of rank `rank`: the first n % world ranks own one extra member."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_members, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_members: int, world: int):
    return [partition(n_members, world, r)[1] - partition(n_members, world, r)[0] for r in range(world)]


def gather_rows(local: torch.Tensor, n_members: int, group=None) -> torch.Tensor:
    """All-gather row-sharded [B_local, C] tensors (possibly ragged over ranks) into [n_members, C] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_members, world)
    if len(set(sizes)) == 1:
        out = torch.empty((n_members,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged shards: pad every shard to the largest one (collectives need equal sizes), gather, strip the padding
    smax = max(sizes)
    padded = torch.zeros((smax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * smax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * smax:r * smax + sizes[r]] for r in range(world)], dim=0)


class Ensemble:
    """B members of one (N_fm, N_r, d, dt, Pr, Tau, symmetric) with per-member Ra, Ra_s, sharded over ranks."""

    def __init__(self, N_fm, N_r, d, dt, Pr, Tau, Ra, Ra_s, symmetric=False, device=None, group=None, plan_factory=None):
        """plan_factory(max_batch) -> plan replaces the EnsemblePlan of this rank (tests of the sharding logic on CPU
        tensors with a CPU test double; the product path always builds the CUDA plan)."""
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        Ra = np.atleast_1d(np.asarray(Ra, dtype=np.float64))
        Ra_s = np.broadcast_to(np.atleast_1d(np.asarray(Ra_s, dtype=np.float64)), Ra.shape)
        self.n_members = int(Ra.shape[0])
        self.lo, self.hi = partition(self.n_members, self.world, self.rank)
        self.n_local = self.hi - self.lo
        if plan_factory is not None:
            self.plan = plan_factory(max(1, self.n_local))
        else:
            from .plan import EnsemblePlan
            self.plan = EnsemblePlan(N_fm, N_r, d, dt, Pr, Tau, symmetric=symmetric, max_batch=max(1, self.n_local),
                                     device=device)
        dev = getattr(self.plan, "device", torch.device("cpu"))
        self.Ra = torch.as_tensor(np.ascontiguousarray(Ra[self.lo:self.hi])).to(dev)
        self.Ra_s = torch.as_tensor(np.ascontiguousarray(Ra_s[self.lo:self.hi])).to(dev)
        self.dt = float(dt)

    def shard(self, X_all):
        """Local rows of a global [B, 3N] host array, uploaded to this rank's GPU."""
        X_all = np.asarray(X_all, dtype=np.float64).reshape(self.n_members, -1)
        return torch.as_tensor(np.ascontiguousarray(X_all[self.lo:self.hi])).to(self.Ra.device)

    # ---- lock-step Newton / continuation over the whole ensemble: every rank drives its own members (krylov.py), no
    # communication while the solves run (members are independent, ranks need not even agree on iteration counts); the
    # per-member outcomes are all-gathered at the end
    def newton(self, X, **kw):
        """Main._Newton for every local member.  Returns (X_local, summary [n_members, 4] on every rank: converged,
        Newton iterations, member JVPs, last error)."""
        from . import krylov
        Xn, info = krylov.newton_batched(self.plan, X, self.Ra, self.Ra_s, **kw)
        hist = info["history"]
        last = torch.stack([hist[max(int(i) - 1, 0), m] for m, i in enumerate(info["iterations"].tolist())]) \
            if hist.numel() else torch.zeros(self.n_local, dtype=torch.float64, device=X.device)
        local = torch.stack([info["converged"].double(), info["iterations"].double(), info["member_jvps"].double(), last], dim=1)
        return Xn, gather_rows(local, self.n_members, self.group)

    def continuation(self, X, N_steps, sign=1.0, **kw):
        """Main._Continuation for every local member, mu = Ra.  Returns (BranchResult of the local members, history
        [N_steps, n_members, 3] on every rank: Ra, KE, Nu_T per branch step)."""
        from . import krylov
        res = krylov.continuation_batched(self.plan, X, self.Ra, N_steps, self.Ra_s, sign=sign, **kw)
        local = torch.stack([torch.stack(res.Ra), torch.stack(res.KE), torch.stack(res.NuT)], dim=2)   # [steps, B_local, 3]
        hist = gather_rows(local.permute(1, 0, 2).contiguous(), self.n_members, self.group).permute(1, 0, 2).contiguous()
        return res, hist

    def time_step(self, X, n_steps, diag_every=1, linear=False):
        """Advance the local members n_steps; every `diag_every` steps (0 = never) compute the diagnostics of all
        local members and all-gather them.  Returns (X_new, history) with history [n_records, B, 4] on every rank
        (the ensemble analogue of Scalar_Data/{Norm,KE,Nu_T,Nu_S}, Main.py:310-315)."""
        if not diag_every:
            cur = self.plan.step(X, self.Ra, self.Ra_s, nsteps=n_steps, linear=linear)
            return cur, torch.empty((0, self.n_members, 4), dtype=torch.float64, device=self.plan.device)
        # one device-resident call for the whole run, one collective for the whole history: the records of the local
        # members [n_records, B_local, 4] are gathered member-major and put back in record-major order
        cur, hist = self.plan.time_step(X, self.Ra, self.Ra_s, n_steps, diag_every=diag_every, linear=linear)
        local = hist[:, :, :4].permute(1, 0, 2).contiguous()                 # [B_local, n_records, 4]
        history = gather_rows(local, self.n_members, self.group).permute(1, 0, 2).contiguous()
        return cur, history

    def gather_states(self, X):
        return gather_rows(X, self.n_members, self.group)

    def close(self):
        self.plan.close()


class GapSweep:
    """Ensemble over shell gaps d (plus Ra, Ra_s): the pre-inverted operators depend on d (SURVEY.md section 7,
    "operator sharing"), so members are grouped by their gap and every group gets its own plan; groups are stepped
    one after the other on the same GPU.  Members keep their global order in every array this class returns."""

    def __init__(self, N_fm, N_r, d, dt, Pr, Tau, Ra, Ra_s, symmetric=False, device=None):
        import numpy as np
        from .plan import EnsemblePlan
        d = np.atleast_1d(np.asarray(d, dtype=np.float64))
        Ra = np.broadcast_to(np.atleast_1d(np.asarray(Ra, dtype=np.float64)), d.shape)
        Ra_s = np.broadcast_to(np.atleast_1d(np.asarray(Ra_s, dtype=np.float64)), d.shape)
        self.n_members = int(d.shape[0])
        self.groups = []
        for dv in np.unique(d):
            idx = np.nonzero(d == dv)[0]
            plan = EnsemblePlan(N_fm, N_r, float(dv), dt, Pr, Tau, symmetric=symmetric, max_batch=len(idx), device=device)
            dev = plan.device
            self.groups.append((torch.as_tensor(idx, device=dev), plan,
                                torch.as_tensor(np.ascontiguousarray(Ra[idx])).to(dev),
                                torch.as_tensor(np.ascontiguousarray(Ra_s[idx])).to(dev)))

    def step(self, X, nsteps=1):
        """X: [B, 3N] device tensor in global member order -> new tensor in the same order."""
        out = torch.empty_like(X)
        for idx, plan, Ra, Ra_s in self.groups:
            out[idx] = plan.step(X[idx], Ra, Ra_s, nsteps=nsteps)
        return out

    def diagnostics(self, X):
        out = torch.empty((X.shape[0], 6), dtype=torch.float64, device=X.device)
        for idx, plan, _, _ in self.groups:
            out[idx] = plan.diagnostics(X[idx])
        return out

    def close(self):
        for _, plan, _, _ in self.groups:
            plan.close()

"""Host-side logic of the multi-GPU path on CPU: member partitioning and the diagnostics all-gather over a
world_size-2 gloo group (the collective the ensemble uses between GPUs is the same call over NCCL)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spectraldoublediffusiveconvection_b200.ensemble import gather_rows, partition, shard_sizes


def test_partition_covers_all_members_contiguously():
    for B in (1, 7, 8, 512, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [partition(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(B, world)
            assert sum(sizes) == B and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        partition(8, 2, 2)


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = partition(B, world, rank)
        full = torch.arange(B * 4, dtype=torch.float64).reshape(B, 4)
        out = gather_rows(full[lo:hi].clone(), B)
        q.put((rank, bool(torch.equal(out, full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_diagnostics_all_gather_two_ranks(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + B
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _newton_worker(rank, world, port, q):
    """Three Newton problems sharded 2 + 1 over two gloo ranks, every rank driving its own members on an oracle-backed
    plan; the per-member summaries are all-gathered."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import load_golden
    from oracle_plan import OraclePlan
    from spectraldoublediffusiveconvection_b200.ensemble import Ensemble
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        G = load_golden("continuation")
        K, N_r, d, Pr, Tau, sym = int(G["N_fm"]), int(G["N_r"]), float(G["d"]), float(G["Pr"]), float(G["Tau"]), bool(G["symmetric"])
        Ra = np.array([float(G["Ra_newton"]), float(G["Ra"]) + 0.5, float(G["Ra_newton"])])
        ens = Ensemble(K, N_r, d, 1.0, Pr, Tau, Ra, float(G["Ra_s"]), symmetric=sym,
                       plan_factory=lambda mb: OraclePlan(K, N_r, d, 1.0, Pr, Tau, symmetric=sym))
        X_all = np.stack([G["X_start"], G["X_start"], G["newton_X"]])
        Xn, summary = ens.newton(ens.shard(X_all), krylov=60)
        full = ens.gather_states(Xn)
        q.put((rank, summary.numpy(), full.numpy()))
    finally:
        dist.destroy_process_group()


def test_sharded_newton_two_ranks():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import load_golden, rel_l2
    G = load_golden("continuation")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_newton_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    (_, s0, x0), (_, s1, x1) = res
    assert np.array_equal(s0, s1) and np.array_equal(x0, x1)           # every rank holds the gathered outcome
    assert s0.shape == (3, 4) and np.all(s0[:, 0] == 1.0)              # all three members converged
    assert s0[0, 1] == len(G["newton_history"]) and s0[2, 1] <= 2      # iterations: the golden problem, the trivial one
    assert rel_l2(x0[0], G["newton_X"]) < 1e-7 and rel_l2(x0[2], G["newton_X"]) < 1e-7

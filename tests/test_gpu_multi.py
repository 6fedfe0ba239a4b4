"""Two-GPU path (skipped on a single-GPU box): Ensemble shards members over ranks, steps them, and all-gathers the
per-step diagnostics over NCCL; results equal the single-GPU run."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from spectraldoublediffusiveconvection_b200.ensemble import Ensemble
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        K, N_r, B = 32, 20, 7                       # ragged: 4 + 3 members
        Ra = np.linspace(2000.0, 5000.0, B)
        X = np.random.default_rng(4).random((B, 3 * (N_r - 1) * K)) * 1e-2
        ens = Ensemble(K, N_r, 0.5, 5e-3, 1.0, 0.5, Ra, 50.0, device=rank)
        Xl, hist = ens.time_step(ens.shard(X), 6, diag_every=2)
        full = ens.gather_states(Xl)
        if rank == 0:
            q.put((full.cpu().numpy(), hist.cpu().numpy()))
        ens.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_ensemble_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from spectraldoublediffusiveconvection_b200 import EnsemblePlan
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, hist = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
    K, N_r, B = 32, 20, 7
    Ra = np.linspace(2000.0, 5000.0, B)
    X = np.random.default_rng(4).random((B, 3 * (N_r - 1) * K)) * 1e-2
    pl = EnsemblePlan(K, N_r, 0.5, 5e-3, 1.0, 0.5, max_batch=B)
    cur = torch.as_tensor(X).cuda()
    Rad = torch.as_tensor(Ra).cuda()
    for s in range(1, 7):
        cur = pl.step(cur, Rad, 50.0)
        if s % 2 == 0:
            assert np.array_equal(hist[s // 2 - 1], pl.diagnostics(cur)[:, :4].cpu().numpy())
    assert np.array_equal(full, cur.cpu().numpy())
    pl.close()

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

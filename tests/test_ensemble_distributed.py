"""The N>1 path on CPU: member sharding and the final result gather over torch.distributed
(gloo, world_size 2).  On the GPU box the same code runs with the NCCL backend (bench.py)."""

from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_members: int, q):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist

    from ribasim_b200.ensemble import gather_members, member_range, perturbation_factors

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = member_range(n_members, rank, world)
        # a member's "result" is a pure function of its global index, like its forcing stream
        local = np.stack([perturbation_factors(m, 3, 5, 2)[0][0] for m in range(lo, hi)], axis=1)
        out = gather_members(local, n_members)
        if rank == 0:
            q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_members", [8, 7])
def test_gloo_gather_is_shard_layout_independent(n_members):
    import torch.multiprocessing as mp

    from ribasim_b200.ensemble import perturbation_factors

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_members, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.stack([perturbation_factors(m, 3, 5, 2)[0][0] for m in range(n_members)], axis=1)
    assert out.shape == (5, n_members)
    assert np.array_equal(out, expect)


def test_member_range_partitions_exactly():
    from ribasim_b200.ensemble import member_range

    for n in (1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = member_range(n, r, world)
                assert 0 <= lo <= hi <= n
                cover.extend(range(lo, hi))
            assert cover == list(range(n))
    with pytest.raises(ValueError):
        member_range(4, 2, 2)


def test_perturbations_depend_on_member_and_day_only():
    from ribasim_b200.ensemble import perturbed_forcing

    base = {k: np.linspace(1.0, 2.0, 6) for k in ("precipitation", "potential_evaporation", "drainage")}
    fb = np.array([3.0, 4.0])
    a = perturbed_forcing(base, fb, range(0, 8), day=2)
    b = perturbed_forcing(base, fb, range(4, 8), day=2)
    for k in a:
        assert np.array_equal(a[k][:, 4:], b[k])
    c = perturbed_forcing(base, fb, range(4, 8), day=3)
    assert not np.array_equal(b["precipitation"], c["precipitation"])

"""Host-side logic of the multi-GPU paths, exercised with world_size-2/3 gloo groups on the CPU:
row sharding arithmetic, the ring halo exchange and the CFL all-reduce."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_rows_partitions_exactly() -> None:
    from pyshocks_b200.distributed import shard_rows

    for batch in (1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(rows for _, rows in spans) == batch
            for (f0, r0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + r0 == f1
            sizes = [rows for _, rows in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(8, 2, 2)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ring_worker(rank: int, world: int, port: int, n_global: int, g: int, out: dict) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pyshocks_b200.distributed import DistRing, fill_halos, shard_rows

        ring = DistRing()
        first, n_local = shard_rows(n_global, rank, world)
        u_global = torch.arange(n_global, dtype=torch.float64) + 0.25
        u = torch.full((1, n_local + 2 * g), -1.0, dtype=torch.float64)
        u[0, g : g + n_local] = u_global[first : first + n_local]
        scratch = torch.zeros((4, 1, g), dtype=torch.float64)
        for _ in range(3):  # repeated exchanges must not cross-match messages
            fill_halos(ring, u, g, scratch)
        idx = (torch.arange(first - g, first + n_local + g)) % n_global
        ok = torch.equal(u[0], u_global[idx])
        m = torch.tensor([float(rank + 1)], dtype=torch.float64)
        ring.all_max(m)
        ok = ok and float(m) == float(world)
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out["ok"] = int(flag) == 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_halo_exchange_gloo(world: int) -> None:
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ring_worker, args=(world, _free_port(), 101, 3, out), nprocs=world, join=True)
    assert out.get("ok") is True


def test_fill_halos_local_matches_periodic_wrap() -> None:
    from pyshocks_b200.distributed import fill_halos_local, shard_rows

    n, g, world = 50, 3, 4
    ug = torch.from_numpy(np.random.default_rng(0).standard_normal(n))
    slabs = []
    for r in range(world):
        first, nl = shard_rows(n, r, world)
        s = torch.zeros(nl + 2 * g, dtype=torch.float64)
        s[g : g + nl] = ug[first : first + nl]
        slabs.append(s)
    fill_halos_local(slabs, g)
    for r in range(world):
        first, nl = shard_rows(n, r, world)
        idx = torch.arange(first - g, first + nl + g) % n
        assert torch.equal(slabs[r], ug[idx])

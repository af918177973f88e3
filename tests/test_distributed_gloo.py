"""CPU, world_size 2 over gloo: the N>1 host logic -- agent sharding and the per-cycle all-gather of
committed-trajectory records (the exchange that replaces the ROS /trajs topic)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neptune_b200 import capi, config
from neptune_b200.cycle import REC, gather_records, ring_phases, shard_agents
from neptune_b200.scenes import make_scene


def test_shard_agents_is_a_partition():
    for n in (5, 8, 64, 1024, 13):
        for w in (1, 2, 3, 4, 8):
            parts = [shard_agents(n, w, r) for r in range(w)]
            assert np.array_equal(np.concatenate(parts), np.arange(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_ring_phases_rotate_consistently():
    """The three-slot record ring of the device-resident cycle (csrc/nb_cycle.cu): what cycle k commits (new) is what
    cycle k + 1 post-checks against (late) and cycle k + 2 plans against (known); the three slots are always distinct."""
    for k in range(12):
        new, late, known = ring_phases(k)
        assert sorted((new, late, known)) == [0, 1, 2]
        assert ring_phases(k + 1)[1] == new and ring_phases(k + 2)[2] == new


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = config("mtlp5")
    sc = make_scene(par, 2002, sync=False)                    # every rank builds the same world
    ok = all(_one(rank, world, sc, n) for n in (4, 5))        # even and uneven shards
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def _one(rank, world, sc, n_agents):
    full = capi.make_records(sc.committed)[:n_agents]
    mine = shard_agents(n_agents, world, rank)
    local = torch.from_numpy(full[mine].copy())
    sizes = [len(shard_agents(n_agents, world, r)) for r in range(world)]
    got = gather_records(local, world, None, sizes).numpy()
    ok = np.array_equal(got, full) and got.shape == (n_agents, REC)
    # a second cycle with modified local records: everybody must see everybody's update
    local2 = local.clone()
    local2[:, 0] += 0.0
    local2[:, 1:18] += 0.05 * (rank + 1)
    got2 = gather_records(local2, world, None, sizes).numpy()
    exp2 = full.copy()
    for r in range(world):
        exp2[shard_agents(n_agents, world, r), 1:18] += 0.05 * (r + 1)
    return ok and np.allclose(got2, exp2)


def test_allgather_records_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]

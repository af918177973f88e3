"""GPU, two ranks: the closed loop across ranks.  K cycles of one world split over 2 GPUs -- records exchanged by the
commit kernel's peer-to-peer stores into both rings, flags instead of a collective -- give, bit for bit, the records,
coefficients and post-check flags of the same world on 1 GPU.  Needs 2 GPUs (skipped on a one-GPU box)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, cfg, seed, cycles, graph, out):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), os.path.join(ROOT, "tests", "mp_cycle_worker.py"), cfg, str(seed), str(cycles), str(graph), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return [np.load(out + f".rank{k}.npz") for k in range(world)]


@pytest.mark.gpu
@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("cfg,seed,graph", [("obst8", 3003, 0), ("mtlp5", 2005, 1)])
def test_two_ranks_equal_one_rank(tmp_path, cfg, seed, graph):
    K = 6
    one = _run(1, cfg, seed, K, graph, str(tmp_path / "one"))[0]
    two = _run(2, cfg, seed, K, graph, str(tmp_path / "two"))
    # every rank's ring holds the whole world's records after every cycle, identical to the single-rank run
    for r in two:
        assert np.array_equal(r["rings"], one["rings"])
    # per-agent results, stitched over the ranks
    for name in ("coeff", "status", "collide", "entangled"):
        got = np.concatenate([r[name] for r in two], axis=1)
        order = np.concatenate([r["agents"] for r in two])
        assert np.array_equal(order, one["agents"])
        assert np.array_equal(got, one[name]), name
    # the loop is closed: the records change from cycle to cycle
    assert not np.array_equal(one["rings"][0], one["rings"][3])

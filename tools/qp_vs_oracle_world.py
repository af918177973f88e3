import numpy as np, sys, os
sys.path.insert(0, "/root/repo")
import bench
from oracle import oracle as orc
from neptune_b200 import capi
from neptune_b200.batch import ReplanResult
K, rank = int(sys.argv[1]), int(sys.argv[2])
par = bench.world_params(K)
agents = bench.rank_agents(par, K, rank, "grid64")
s = capi.Solver(par)
_, scenes = bench.make_world(K, rank, 4, capi.DeviceEntBackend(s), "grid64", agents=agents)
for q, sc in enumerate(scenes):
    got = s.replan(sc.batch)
    ref = ReplanResult.empty(sc.batch)
    assert orc.replan_batch(sc.batch, ref, 16) == 0
    err = np.abs(got.coeff_out - ref.coeff_out).max()
    print("scene", q, "status equal", bool((got.status == ref.status).all()), "gpu iters max", got.iters.max(axis=0).tolist(), "oracle iters max", ref.iters.max(axis=0).tolist(), "max|dcoeff|", err)
    d = np.flatnonzero(got.iters[:, 0] != ref.iters[:, 0])
    if len(d): print("   first-solve iteration counts differ for agents", d.tolist(), got.iters[d, 0].tolist(), ref.iters[d, 0].tolist(), "n_int", sc.batch.n_int[d].tolist(), "status gpu/orc", got.status[d].tolist(), ref.status[d].tolist())

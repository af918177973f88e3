"""A few plain (non-graph) cycles of the bench world for ncu launch lists.  GPU only.
    python tools/cycle_once.py [grid64|grid1024] [cycles]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch  # noqa: E402
from neptune_b200 import capi  # noqa: E402
from neptune_b200.cycle import ReplanCycle  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "grid64"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
par = bench.world_params(1, wl)
agents = bench.rank_agents(par, 1, 0, wl)
static = None
if par.num_of_static_obst:
    from neptune_b200.scenes import make_scene
    s0 = make_scene(par, 5005, agents=agents[:1], pack_hulls=False)
    static = (s0.batch.st_ptr, s0.batch.st_xy, s0.strep)
cyc = ReplanCycle(par, agents, torch.device("cuda", 0), static=static)
_, scenes = bench.make_world(1, 0, 1, capi.DeviceEntBackend(cyc.solver), wl, agents=agents)
cyc.seed_records(cyc.records_of(scenes[0]))
hin, hout = cyc.host_inputs(scenes[0]), cyc.host_outputs()
for k in range(n):
    cyc.step_from_host(hin, hout)
cyc.check_errors()
print("status", np.bincount(hout["status"], minlength=3), "collide", int(hout["collide"].sum()), "entangled", int(hout["entangled"].sum()))
cyc.close()

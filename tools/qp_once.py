"""One nb_replan_batch on the bench world (for ncu captures of k_lines / k_qp).  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neptune_b200 import capi  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
par, scenes = bench.make_world(k, 0, 1)
s = capi.Solver(par)
for _ in range(3):
    res = s.replan(scenes[0].batch, with_lines=False)
print("status", (res.status == 0).sum(), (res.status == 1).sum(), (res.status == 2).sum())

"""Rank 0's share of the K-GPU weak-scaling world on ONE GPU (64 agents planning against 64 K, nobody else re-plans):
graph-replayed cycles timed with CUDA events, L2 flushed in between.  GPU only.
    python tools/cycle_rank0_of.py K [cycles]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch  # noqa: E402
from neptune_b200 import capi  # noqa: E402
from neptune_b200.cycle import ReplanCycle  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
par = bench.world_params(K)
agents = bench.rank_agents(par, K, 0, "grid64")
planned = np.zeros(par.num_of_agents, np.uint8)
planned[agents] = 1
dev = torch.device("cuda", 0)
cyc = ReplanCycle(par, agents, dev, planned=planned)

class A:  # what bench.make_world reads
    gpus = K
bench_args = A()
_, scenes = bench.make_world(K, 0, 2, capi.DeviceEntBackend(cyc.solver), "grid64", agents=agents)
cyc.seed_records(cyc.records_of(scenes[0]))
hins = [cyc.host_inputs(sc) for sc in scenes]
hout = cyc.host_outputs()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
st = cyc.stream
for it in range(5):
    cyc.upload(hins[it % 2])
    cyc.step()
st.synchronize()
cyc.capture()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
for k in range(n):
    cyc.upload(hins[k % 2])
    with torch.cuda.stream(st):
        flush.zero_()
    ev[k][0].record(st)
    cyc.step()
    ev[k][1].record(st)
st.synchronize()
torch.cuda.synchronize()
cyc.check_errors()
ms = np.array([a.elapsed_time(b) for a, b in ev])
print(f"world of {K} x 64 agents, rank 0 alone: {ms.mean():.4f} ms per cycle (p50 {np.median(ms):.4f}, p95 {np.percentile(ms, 95):.4f})")
cyc.upload(hins[0])
print({k: round(v, 4) for k, v in cyc.step_profiled().items()})
cyc.close()

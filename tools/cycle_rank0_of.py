"""Rank 0's share of the K-GPU weak-scaling world on ONE GPU (64 agents planning against 64 K, nobody else re-plans):
graph-replayed cycles timed with CUDA events, L2 flushed in between.  GPU only.
    python tools/cycle_rank0_of.py K [cycles] [rank]"""
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
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
par = bench.world_params(K)
agents = bench.rank_agents(par, K, rank, "grid64")
planned = np.zeros(par.num_of_agents, np.uint8)
planned[agents] = 1
dev = torch.device("cuda", 0)
cyc = ReplanCycle(par, agents, dev, planned=planned)

class A:  # what bench.make_world reads
    gpus = K
bench_args = A()
_, scenes = bench.make_world(K, rank, 4, capi.DeviceEntBackend(cyc.solver), "grid64", agents=agents)
cyc.seed_records(cyc.records_of(scenes[0]))
hins = [cyc.host_inputs(sc) for sc in scenes]
hout = cyc.host_outputs()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
st = cyc.stream
for it in range(5):
    cyc.upload(hins[it % 4])
    cyc.step()
st.synchronize()
cyc.capture()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
for k in range(n):
    cyc.upload(hins[k % 4])
    with torch.cuda.stream(st):
        flush.zero_()
    ev[k][0].record(st)
    cyc.step()
    ev[k][1].record(st)
st.synchronize()
torch.cuda.synchronize()
cyc.check_errors()
ms = np.array([a.elapsed_time(b) for a, b in ev])
print(f"world of {K} x 64 agents, rank {rank} alone: {ms.mean():.4f} ms per cycle (p50 {np.median(ms):.4f}, p95 {np.percentile(ms, 95):.4f})")
import ctypes as C
lib = capi.lib()
lib.nb_set_profiling(cyc.solver.handle, 1)
lib.nb_qp_phase_cycles.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
for q in range(len(hins)):   # per scene: stages (plain launches, one stream) and the interior-point iteration counts
    cyc.upload(hins[q])
    sm = cyc.step_profiled()
    cyc.download(hout)
    st.synchronize()
    it = hout["iters"]
    if it[:, 0].max() >= 20:   # a first solve that ran into the cap: what its late iterations looked like
        prof = np.zeros((cyc.B, 16), np.int64)
        lib.nb_qp_phase_cycles(cyc.solver.handle, prof.ctypes.data_as(C.c_void_p), cyc.B)
        w = int(np.argmax(it[:, 0]))
        print(f"  agent {w}: iterations {it[w].tolist()}, n_int {int(hins[q]['n_int'][w])}, smallest step / last |rp| / last mu from iteration 16 on:",
              prof[w, 13:16].view(np.float64).tolist())
    print(f"scene {q}:", {k: round(v, 4) for k, v in sm.items()}, "ipm iterations max", it.max(axis=0).tolist(), "status", np.bincount(hout["status"], minlength=3).tolist())
cyc.close()

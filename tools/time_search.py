"""Times the front-end search kernel (K0) on one scene and prints the kernel's own per-phase cycle counters.
Measurement helper (not the bench; the comparison with the oracle lives in bench.py's front_end block and in tests/):
python tools/time_search.py [cfg] [seed] [max_expansions] [planning agents]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from neptune_b200 import capi, config
from neptune_b200.scenes import make_scene, make_search_batch


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "grid64"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4004
    if cfg.startswith("grid") and cfg not in ("grid64", "grid1024"):   # e.g. grid128: the 64-agents-per-GPU world of N GPUs on ONE GPU
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        par = bench.world_params(int(cfg[4:]) // 64)
    else:
        par = config(cfg)
    if len(sys.argv) > 3:
        par.search_max_expansions = int(sys.argv[3])
    s = capi.Solver(par, device=0)
    agents = np.arange(int(sys.argv[4])) if len(sys.argv) > 4 else None   # a subset of the world's agents plans
    if par.num_of_static_obst:   # the entanglement back end of the generator needs the static representation first
        s0 = make_scene(par, seed, sync=True, agents=np.arange(1), pack_hulls=False)
        s.set_static(s0.batch.st_ptr, s0.batch.st_xy, s0.strep)
    sc = make_scene(par, seed, sync=(agents is not None), ent_backend=capi.DeviceEntBackend(s), group_hulls=True, agents=agents,
                    pack_hulls=(par.num_of_agents <= 256))
    sb = make_search_batch(sc, seed + 1, per_agent_order=True)
    if par.num_of_static_obst:
        s.set_static(sb.st_ptr, sb.st_xy, sb.strep)
        s.set_static_longest(sb.st_longest)
    s.search_configure()
    got = s.search(sb)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t = time.perf_counter()
        got = s.search(sb)
        ts.append(time.perf_counter() - t)
    print(f"{cfg}: B={sb.B} gpu search (host buffers, e2e) best {min(ts) * 1e3:.3f} ms; pops total {int(got.stats[:, 1].sum())} "
          f"max {int(got.stats[:, 1].max())}; nodes max {int(got.stats[:, 0].max())}; status {np.bincount(got.status, minlength=3).tolist()}")
    print(f"  per pop of the slowest agent: {min(ts) * 1e6 / max(1, int(got.stats[:, 1].max())):.2f} us")
    import ctypes as C
    capi.lib().nb_set_profiling(s.handle, 1)
    got = s.search(sb)
    pc = np.zeros((sb.B, 16), np.int64)
    capi.lib().nb_search_phase_cycles.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    assert capi.lib().nb_search_phase_cycles(s.handle, pc.ctypes.data, sb.B) == 0
    capi.lib().nb_set_profiling(s.handle, 0)
    slow = int(np.argmax(got.stats[:, 1]))
    names = ["chains||collide", "resolve", "copy", "pop", "active", "endpoint", "setup", "primitives"]
    pops = max(1, int(got.stats[slow, 1]))
    print("  cycles per pop (slowest agent): " + ", ".join(f"{n} {pc[slow, i] / pops:.0f}" for i, n in enumerate(names[:8])))
    cn = ["(aux collide path)", "(aux: hull staging)", "chain", "key+lookup", "step geometry", "crossing tests", "automaton"]
    print("  child 0, cycles per pop: " + ", ".join(f"{n} {pc[slow, 8 + i] / pops:.0f}" for i, n in enumerate(cn)) + f", slowest child {pc[slow, 15] / pops:.0f}")


if __name__ == "__main__":
    main()

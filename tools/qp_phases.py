"""Per-phase SM cycles of the QP warps (nb_qp_phase_cycles) on the bench world.  GPU only.
    python tools/qp_phases.py [--gpus-world K]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neptune_b200 import capi  # noqa: E402

NAMES = ["setup", "resid", "rd fold+test", "assemble", "factor", "pred features", "pred sweep", "corr solve", "final sweep",
         "pred ct_all", "pred solve", "rd ct_all", "(chain of pred solve)"]


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    par, scenes = bench.make_world(k, 0, 1)
    b = scenes[0].batch
    s = capi.Solver(par)
    lib = capi.lib()
    lib.nb_set_profiling(s.handle, 1)
    res = s.replan(b, with_lines=False)
    res = s.replan(b, with_lines=False)
    out = np.zeros((b.B, 16), np.int64)
    lib.nb_qp_phase_cycles.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    assert lib.nb_qp_phase_cycles(s.handle, out.ctypes.data_as(C.c_void_p), b.B) == 0
    ms = (C.c_double * 2)()
    lib.nb_kernel_times(s.handle, ms, 2)
    it = res.iters.sum(axis=1)
    tot = out[:, :12].sum(axis=1)
    worst = int(np.argmax(tot))
    print(f"k_lines {ms[0]:.4f} ms  k_qp {ms[1]:.4f} ms; iterations mean {it.mean():.1f} max {it.max()}; "
          f"slowest warp {worst}: n={b.n_int[worst]} iters={res.iters[worst].tolist()} cycles={tot[worst]} "
          f"({tot[worst] / 1.965e6:.3f} ms at 1965 MHz)")
    print("phase           mean cycles/agent   slowest warp   per iteration (slowest)")
    for q, nme in enumerate(NAMES):
        q = 12 if q == 12 else q
        print(f"{nme:14s} {out[:, q].mean():14.0f} {out[worst, q]:14d} {out[worst, q] / max(1, it[worst]):14.0f}")


if __name__ == "__main__":
    main()

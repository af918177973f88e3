#!/usr/bin/env python
"""bench.py -- replans/sec of the NEPTUNE replan hot path on B200 (BASELINE.json metric).

One "step" = one replan cycle: every agent of the world replans once (separating-line LPs ->
trajectory QP with the reference's fallback path), the ranks exchange committed trajectories with one
NCCL all-gather.  Workload: BASELINE.json configs[3] family -- a grid world with 64 agents per GPU
(N=1 is exactly "64 agents synthetic random goals"); every agent plans against ALL other agents of
the world (faithful, no culling), so per-agent work grows with the world while agents/GPU stay fixed.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`--impl reference` times the CPU oracle (the restatement of the reference algorithm: Gurobi/GLPK
cannot be installed here) on the host cores for the same config and metric.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AGENTS_PER_GPU = 64
SEED = 4004


def world_params(n_gpus: int):
    from neptune_b200.params import Params
    nx, ny = 8, 8 * n_gpus
    pitch = 8.0
    xs = (np.arange(nx) - (nx - 1) / 2.0) * pitch
    ys = (np.arange(ny) - (ny - 1) / 2.0) * pitch
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    p = Params(num_of_agents=nx * ny, tetherLength=25.0)
    p.pb = np.stack([gx.ravel(), gy.ravel()], axis=1)
    p.x_min, p.x_max = xs[0] - 12.0, xs[-1] + 12.0
    p.y_min, p.y_max = ys[0] - 12.0, ys[-1] + 12.0
    return p


def make_world(n_gpus: int, rank: int, n_scenes: int):
    from neptune_b200.scenes import make_scene
    par = world_params(n_gpus)
    agents = np.arange(rank * AGENTS_PER_GPU, (rank + 1) * AGENTS_PER_GPU)
    return par, [make_scene(par, SEED + k, agents=agents) for k in range(n_scenes)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_baseline(par, scene, budget_s: float, threads: int):
    """Oracle (CPU restatement of the reference algorithm) on the host cores: replans/s."""
    from neptune_b200.batch import ReplanResult
    from oracle import oracle as orc
    res = ReplanResult.empty(scene.batch, with_lines=False)
    orc.replan_batch(scene.batch, res, threads)  # warm
    t0, reps = time.perf_counter(), 0
    while True:
        orc.replan_batch(scene.batch, res, threads)
        reps += 1
        el = time.perf_counter() - t0
        if el >= budget_s or reps >= 50:
            break
    return scene.batch.B * reps / el, reps, el, res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    par, scenes = make_world(args.gpus, 0, 1)
    threads = os.cpu_count() or 1
    from neptune_b200.batch import ReplanResult
    from oracle import oracle as orc
    res = ReplanResult.empty(scenes[0].batch, with_lines=False)
    for _ in range(args.warmup):
        orc.replan_batch(scenes[0].batch, res, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.replan_batch(scenes[0].batch, res, threads)
    el = time.perf_counter() - t0
    B = scenes[0].batch.B
    val = B * args.steps / el
    sample = (f"{B} of the world's {par.num_of_agents} agents (rank-0 shard) x {args.steps} cycles, "
              f"each against all {par.num_of_agents - 1} others")
    line = {"impl": "reference", "metric": "replans_per_sec", "value": val, "unit": "replans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(par, args.gpus),
            "cpu_baseline": {"value": val, "unit": "replans/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "replans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(par, n_gpus):
    return {"workload": f"grid world, {AGENTS_PER_GPU} agents/GPU x {n_gpus} GPU = {par.num_of_agents} agents "
                        "(BASELINE.json configs[3] at N=1), random goals, no static obstacles, faithful (no culling)",
            "agents": par.num_of_agents, "agents_per_gpu": AGENTS_PER_GPU, "num_pol": par.num_pol,
            "T_span": par.T_span, "seed": SEED, "l2": "flushed between timed iterations (256 MiB write)",
            "parallelism": f"agents sharded over {n_gpus} rank(s); one all-gather of committed trajectories per cycle"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from neptune_b200 import capi
    from neptune_b200.batch import NPOL, ReplanResult

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    par, scenes = make_world(args.gpus, rank, args.scenes)
    solver = capi.Solver(par, device=local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    B = scenes[0].batch.B

    # ---------------- device-resident inputs (value) and pinned host inputs (e2e)
    def dev_args(sc):
        b = sc.batch
        t = {k: torch.from_numpy(np.ascontiguousarray(getattr(b, k))).to(dev) for k in
             ("agent_id", "n_int", "coeff_init", "hull_ptr", "hull_xy", "nih0", "esv_cnt", "esv_alpha", "esv_active",
              "bp_cnt", "bp_xy")}
        t["coeff_out"] = torch.zeros((B, 3, NPOL, 4), dtype=torch.float64, device=dev)
        t["obj"] = torch.zeros(B, dtype=torch.float64, device=dev)
        t["status"] = torch.zeros(B, dtype=torch.int32, device=dev)
        t["iters"] = torch.zeros((B, 2), dtype=torch.int32, device=dev)
        a = capi.NbReplanArgs()
        a.B, a.space, a.n_hull_slots, a.hull_nvert = B, capi.NB_DEVICE, b.n_hull_slots, int(b.hull_xy.shape[0])
        for k, v in t.items():
            setattr(a, k, v.data_ptr())
        a.lines, a.line_ok = None, None
        return a, t

    def pinned_batch(sc):
        import dataclasses
        b = sc.batch
        keep = {}
        for k in ("agent_id", "n_int", "coeff_init", "hull_ptr", "hull_xy", "nih0", "esv_cnt", "esv_alpha",
                  "esv_active", "bp_cnt", "bp_xy"):
            tt = torch.from_numpy(np.ascontiguousarray(getattr(b, k))).pin_memory()
            keep[k] = tt
        pb = dataclasses.replace(b, **{k: v.numpy() for k, v in keep.items()})
        res = ReplanResult(coeff_out=torch.zeros((B, 3, NPOL, 4), dtype=torch.float64).pin_memory().numpy(),
                           obj=torch.zeros(B, dtype=torch.float64).pin_memory().numpy(),
                           status=torch.zeros(B, dtype=torch.int32).pin_memory().numpy(),
                           iters=torch.zeros((B, 2), dtype=torch.int32).pin_memory().numpy())
        return capi.host_args(pb, res), (keep, pb, res)

    dargs = [dev_args(sc) for sc in scenes]
    hargs = [pinned_batch(sc) for sc in scenes]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    gathered = torch.zeros((world * B, 100), dtype=torch.float64, device=dev) if world > 1 else None
    record = torch.zeros((B, 100), dtype=torch.float64, device=dev)

    def exchange(t):
        """committed-trajectory record per agent: coefficients + (id, n) -> one NCCL all-gather"""
        if world == 1:
            return
        record[:, :96] = t["coeff_out"].reshape(B, 96)
        record[:, 96] = t["agent_id"].to(torch.float64)
        record[:, 97] = t["n_int"].to(torch.float64)
        dist.all_gather_into_tensor(gathered, record)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM, CUDA events, max over ranks
    solver_lib = capi.lib()
    solver_lib.nb_set_profiling(solver.handle, 1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ktimes = []
    sampler = None
    for it in range(args.warmup + args.steps):
        k = it - args.warmup
        if k == 0:
            barrier()
            sampler = ClockSampler(local_rank)
            l0 = solver.launch_count()
        a, t = dargs[it % len(dargs)]
        flush.zero_()
        if k >= 0:
            ev[k][0].record()
        solver.replan_args(a, stream)
        exchange(t)
        if k >= 0:
            ev[k][1].record()
            ms = (C.c_double * 2)()
            solver_lib.nb_kernel_times(solver.handle, ms, 2)
            ktimes.append((ms[0], ms[1]))
    barrier()
    launches = solver.launch_count() - l0
    clocks = sampler.stop()
    rc = solver_lib.nb_check_async_errors(solver.handle, C.c_void_p(stream))
    assert rc == 0, "capacity overflow during the timed region"
    step_ms = np.array([e0.elapsed_time(e1) for e0, e1 in ev])
    total_ms = float(step_ms.sum())
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---------------- e2e: host buffers through the C-ABI, H2D + kernels + D2H inside the timed region
    for it in range(2):
        solver.replan_args(hargs[it % len(hargs)][0], stream)
    barrier()
    t0 = time.perf_counter()
    for it in range(args.steps):
        solver.replan_args(hargs[it % len(hargs)][0], stream)  # synchronous: returns after the D2H copies
        exchange(dargs[0][1])
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e = world * B * args.steps / e2e_s
    b0 = scenes[0].batch
    h2d = b0.input_bytes() + b0.bp_cnt.nbytes + b0.bp_xy.nbytes
    d2h = B * (96 * 8 + 8 + 4 + 8)

    if rank == 0:
        # status histogram of the last device-resident step, to show what work the step did
        st = dargs[(args.warmup + args.steps - 1) % len(dargs)][1]["status"].cpu().numpy()
        itn = dargs[(args.warmup + args.steps - 1) % len(dargs)][1]["iters"].cpu().numpy()
        kt = np.array(ktimes)
        dom = int(np.argmax(kt.mean(axis=0)))
        dom_ms = float(kt[:, dom].mean())
        alg_bytes = float(np.mean([sc.batch.algorithmic_bytes() for sc in scenes]))
        peak, which = measured_peak_gbs()
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        line = {"metric": "replans_per_sec", "value": value, "unit": "replans/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(par, world),
                "p50_ms_per_replan_cycle": float(np.median(step_ms)),
                "e2e": {"value": e2e, "unit": "replans/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches),
                "kernels_ms": {"k_lines": float(kt[:, 0].mean()), "k_qp": float(kt[:, 1].mean())},
                "roofline": {"bound": "hbm", "kernel": ["k_lines", "k_qp"][dom], "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": which,
                             "algorithmic_bytes_per_launch": alg_bytes,
                             "note": "latency/FP64-issue bound at this size, not HBM bound (DESIGN.md)"},
                "status_hist": {str(k): int((st == k).sum()) for k in (0, 1, 2)},
                "ipm_iters_mean": float(itn.sum(axis=1).mean()),
                "clocks": clocks}
        if world == 1 and not args.no_cpu:
            v, reps, el, ref = cpu_baseline(par, scenes[0], args.cpu_budget, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": v, "unit": "replans/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"{B} agents x {reps} cycles of scene 0 ({el:.1f} s), oracle/neptune_oracle.c"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=4)
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

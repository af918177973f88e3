#!/usr/bin/env python
"""bench.py -- replans/sec of the NEPTUNE replan hot path on B200 (BASELINE.json metric).

One "step" = one replan cycle of neptune_b200.cycle.ReplanCycle: hulls/samples of every other
agent's committed trajectory (K1), PredictAlphasBetas (K3), separating-line LPs + pruning (K2), the
trajectory QP with the reference's fallback path (K4), the post-check (K5 GJK + K3 entangle re-check),
commit; the ranks exchange committed-trajectory records with one NCCL all-gather.  Workload: BASELINE.json configs[3] family -- a grid world with 64 agents per GPU
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


def world_params(n_gpus: int, workload: str = "grid64"):
    from neptune_b200.params import Params, config
    if workload == "grid1024":   # BASELINE.json configs[4]: 1024 agents / 200 static obstacles, fixed world
        return config("grid1024")
    nx, ny = 8, 8 * n_gpus
    pitch = 8.0
    xs = (np.arange(nx) - (nx - 1) / 2.0) * pitch
    ys = (np.arange(ny) - (ny - 1) / 2.0) * pitch
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    # ent_cap >= N + M: the front-end chain can then never overflow its storage (it declares a step
    # entangling before the word exceeds N + M entries, kinodynamic_search.cpp:844-848)
    p = Params(num_of_agents=nx * ny, tetherLength=25.0, ent_cap=max(48, nx * ny + 8),
               search_ecap=max(24, (nx * ny) // 2))   # signature words grow with the number of tethers around
    p.pb = np.stack([gx.ravel(), gy.ravel()], axis=1)
    p.x_min, p.x_max = xs[0] - 12.0, xs[-1] + 12.0
    p.y_min, p.y_max = ys[0] - 12.0, ys[-1] + 12.0
    return p


def rank_agents(par, n_gpus: int, rank: int, workload: str):
    from neptune_b200.cycle import shard_agents
    if workload == "grid1024":   # strong scaling: the fixed world is split over the ranks
        return shard_agents(par.num_of_agents, n_gpus, rank)
    return np.arange(rank * AGENTS_PER_GPU, (rank + 1) * AGENTS_PER_GPU)


def make_world(n_gpus: int, rank: int, n_scenes: int, ent_backend=None, workload: str = "grid64", agents=None):
    from neptune_b200.scenes import make_scene
    par = world_params(n_gpus, workload)
    if agents is None:
        agents = rank_agents(par, n_gpus, rank, workload)
    seed = SEED if workload == "grid64" else 5005
    return par, [make_scene(par, seed + k, agents=agents, ent_backend=ent_backend,
                            pack_hulls=(par.num_of_agents <= 256)) for k in range(n_scenes)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE, text=True)
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


def ncu_traffic(kernel: str):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            return int(json.load(f)["per_launch_dram_bytes"][kernel])
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_cycle(scene, threads: int):
    """One whole cycle of the scene's agents on the CPU oracle (hulls/samples, predict, LPs + QP,
    post-check, entangle re-check): the restatement of the reference algorithm, all host threads."""
    from neptune_b200 import capi
    from neptune_b200.cycle import ReplanCycle
    from oracle import oracle as orc
    recs = capi.make_records(scene.committed)
    t_now = np.asarray(scene.t_start, np.float64) - ReplanCycle.DELTA_T_STEPS * scene.par.dc
    rc, out = orc.cycle_batch(scene, recs, threads, t_now=t_now)
    assert rc == 0
    return out


def cpu_baseline(scene, budget_s: float, threads: int):
    cpu_cycle(scene, threads)  # warm
    t0, reps = time.perf_counter(), 0
    while True:
        out = cpu_cycle(scene, threads)
        reps += 1
        el = time.perf_counter() - t0
        if el >= budget_s or reps >= 2000:
            break
    return scene.batch.B * reps / el, reps, el, out


def measure_front_end(par, agents, dev, scene, static, steps: int, with_cpu: bool):
    """The front end (K0, KinodynamicSearch::run) measured on the same world: one more ReplanCycle object with
    front_end=True, i.e. hulls -> predict -> SEARCH -> LPs + QP -> post-check -> commit, all device-resident.
    Reported beside the headline (whose metric, SURVEY 8d, starts after the front end)."""
    import torch

    from neptune_b200 import capi
    from neptune_b200.cycle import ReplanCycle
    from neptune_b200.scenes import search_host_inputs
    from neptune_b200.search import SearchBatch, SearchResult, static_longest_dist

    cyc = ReplanCycle(par, agents, dev, static=static, world=1, front_end=True)
    fe = search_host_inputs(scene, SEED + 1)
    hin = cyc.host_inputs(scene, fe)
    hout = cyc.host_outputs()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    B = cyc.B
    cyc.upload(hin)
    cyc.capture()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for it in range(2 + steps):
        cyc.upload(hin)
        flush.zero_()
        if it >= 2:
            ev[it - 2][0].record()
        cyc.step()
        if it >= 2:
            ev[it - 2][1].record()
    torch.cuda.synchronize()
    cyc.check_errors()
    full_ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    cyc.profile = True
    fe_ms = []
    for it in range(3):
        cyc.upload(hin)
        flush.zero_()
        cyc.step()
        fe_ms.append(cyc.stage_ms["front_end"])
    cyc.profile = False
    t0 = time.perf_counter()
    for it in range(steps):
        cyc.step_from_host(hin, hout)
    e2e_ms = 1e3 * (time.perf_counter() - t0) / steps
    o = cyc.o
    stats = o["fe_stats"].cpu().numpy()
    status = o["fe_status"].cpu().numpy()
    out = {"kernel": "k_search", "agents": B, "ms_search": float(np.median(fe_ms)), "ms_full_cycle": full_ms,
           "full_replans_per_s": B / (full_ms * 1e-3), "e2e_full_replans_per_s": B / (e2e_ms * 1e-3),
           "pops": int(stats[:, 1].sum()), "pops_max": int(stats[:, 1].max()), "nodes_max": int(stats[:, 0].max()),
           "pops_per_s": float(stats[:, 1].sum() / (np.median(fe_ms) * 1e-3)),
           "status_hist": {"runtime": int((status == 0).sum()), "goal": int((status == 1).sum()), "empty": int((status == 2).sum())},
           "solved": int(o["fe_solved"].sum().item()), "max_expansions": par.search_max_expansions,
           "max_nodes": par.search_max_nodes}
    if with_cpu:
        from oracle import oracle as orc
        M = par.num_of_static_obst
        sb = SearchBatch(
            par=par, agent_id=scene.batch.agent_id.copy(), init=fe["init"], goal=fe["goal"], coeffs_z=fe["coeffs_z"],
            group=hin["group"].copy(), hull_xy=o["hull_xy_g"].cpu().numpy(), hull_cnt=o["hull_cnt_g"].cpu().numpy(),
            samp=o["samp_g"].cpu().numpy(), known=scene.known.copy(), es_cnt=o["esA_cnt"].cpu().numpy(),
            es_alpha=o["esA_alpha"].cpu().numpy(), es_beta=o["esA_beta"].cpu().numpy(), es_bend=o["esA_bend"].cpu().numpy(),
            es_active=o["esA_active"].cpu().numpy(), bp_cnt=scene.batch.bp_cnt, bp_xy=scene.batch.bp_xy, comb=fe["comb"],
            st_ptr=scene.batch.st_ptr, st_xy=scene.batch.st_xy, strep=np.asarray(scene.strep, np.float64).reshape(M, 2, 2),
            st_longest=static_longest_dist(scene.static_raw, np.asarray(scene.strep).reshape(M, 2, 2)) if M else np.zeros((0, 2)))
        ref = SearchResult.empty(sb)
        nt = os.cpu_count() or 1
        orc.search_batch(sb, ref, nt)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            orc.search_batch(sb, ref, nt)
        cpu_ms = 1e3 * (time.perf_counter() - t0) / reps
        same = (np.array_equal(ref.status, status) and np.array_equal(ref.n_int, o["fe_n_int"].cpu().numpy())
                and np.array_equal(ref.coeff, o["fe_coeff"].cpu().numpy()) and np.array_equal(ref.stats, stats)
                and np.array_equal(ref.esv_alpha, o["fe_esv_alpha"].cpu().numpy()))
        out["cpu_oracle"] = {"ms_search": cpu_ms, "cores": nt, "kind": "port", "identical": bool(same),
                             "sample": f"{B} searches x {reps}, oracle/neptune_search.c orc_search_batch"}
    cyc.solver.close()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from tests.ent_backends import OracleEntBackend
    agents = None
    if args.workload == "grid1024":   # bounded sample: the first 32 agents of the rank-0 shard
        agents = np.arange(32)
    par, scenes = make_world(args.gpus, 0, 1, OracleEntBackend(orc), args.workload, agents)
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_cycle(scenes[0], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_cycle(scenes[0], threads)
    el = time.perf_counter() - t0
    B = scenes[0].batch.B
    val = B * args.steps / el
    sample = (f"{B} of the world's {par.num_of_agents} agents (rank-0 shard) x {args.steps} cycles, "
              f"each against all {par.num_of_agents - 1} others; oracle/neptune_oracle.c orc_cycle_batch")
    line = {"impl": "reference", "metric": "replans_per_sec", "value": val, "unit": "replans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": scaling_kind(args.workload), "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(par, args.gpus, args.workload),
            "cpu_baseline": {"value": val, "unit": "replans/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "replans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def scaling_kind(workload: str) -> str:
    return "strong" if workload == "grid1024" else "weak"


def config_dict(par, n_gpus, workload="grid64"):
    if workload == "grid1024":
        wl = (f"BASELINE.json configs[4]: {par.num_of_agents} agents / {par.num_of_static_obst} static obstacles, "
              f"32x32 base grid, sharded over {n_gpus} GPU(s), random goals, faithful (no culling)")
    else:
        wl = (f"grid world, {AGENTS_PER_GPU} agents/GPU x {n_gpus} GPU = {par.num_of_agents} agents "
              "(BASELINE.json configs[3] at N=1), random goals, no static obstacles, faithful (no culling)")
    return {"workload": wl, "agents": par.num_of_agents, "agents_per_gpu": par.num_of_agents // n_gpus
            if workload == "grid1024" else AGENTS_PER_GPU, "static_obstacles": par.num_of_static_obst,
            "num_pol": par.num_pol, "T_span": par.T_span, "seed": SEED if workload == "grid64" else 5005,
            "l2": "flushed between timed iterations (256 MiB write)",
            "parallelism": f"agents sharded over {n_gpus} rank(s); one all-gather of committed trajectories per cycle"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from neptune_b200 import capi
    from neptune_b200.cycle import ReplanCycle

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    par = world_params(args.gpus, args.workload)
    agents = rank_agents(par, args.gpus, rank, args.workload)
    static = None
    if par.num_of_static_obst:   # static obstacles are part of the world: build them once, before the solver
        from neptune_b200.scenes import make_scene
        s0 = make_scene(par, 5005, agents=agents[:1], pack_hulls=False)
        static = (s0.batch.st_ptr, s0.batch.st_xy, s0.strep)
    cyc = ReplanCycle(par, agents, dev, static=static, world=world)
    # the workload generator fills entanglement states through the product's own K3 kernels
    _, scenes = make_world(args.gpus, rank, args.scenes, capi.DeviceEntBackend(cyc.solver), args.workload)
    B = cyc.B
    hins = [cyc.host_inputs(sc) for sc in scenes]
    hout = cyc.host_outputs()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    lib = capi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- per-kernel times of the two back-end kernels (library events; plain launches)
    lib.nb_set_profiling(cyc.solver.handle, 1)
    ktimes = []
    for it in range(8):
        cyc.upload(hins[it % len(hins)])
        flush.zero_()
        cyc.step()
        ms = (C.c_double * 2)()
        lib.nb_kernel_times(cyc.solver.handle, ms, 2)
        if it >= 3:
            ktimes.append((ms[0], ms[1]))
    lib.nb_set_profiling(cyc.solver.handle, 0)
    cyc.check_errors()
    # ---------------- value: inputs resident in HBM before the timed region, CUDA events, max over ranks;
    # the cycle's launch sequence is replayed from a CUDA graph
    cyc.upload(hins[0])
    if not args.no_graph:
        cyc.capture()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler, l0 = None, 0
    for it in range(args.warmup + args.steps):
        k = it - args.warmup
        if it == 0:
            sampler = ClockSampler(local_rank)   # started with the warm-up (same load): nvidia-smi needs ~50 ms to come up
        if k == 0:
            barrier()
            l0 = cyc.solver.launch_count()
        cyc.upload(hins[it % len(hins)])       # untimed: this leg measures with inputs already in HBM
        flush.zero_()
        if k >= 0:
            ev[k][0].record()
        cyc.step()
        if k >= 0:
            ev[k][1].record()
    barrier()
    launches = (cyc.solver.launch_count() - l0) if getattr(cyc, "graph", None) is None else args.steps * cyc.launches_per_cycle
    # a short run (K x 0.4 ms) can end before the first 20 ms sample: keep the same cycle running, untimed, until there
    # are three samples under load (bounded at 2 s)
    extra, t_extra = 0, time.time()
    if world == 1:
        while len(sampler.rows) < 3 and sampler.proc is not None and time.time() - t_extra < 2.0:
            flush.zero_()
            cyc.step()
            extra += 1
            if extra % 16 == 0:
                torch.cuda.synchronize()
    else:   # a cycle ends in an all-gather: every rank must run the same number of extra cycles
        for _ in range(max(0, 400 - args.steps)):
            flush.zero_()
            cyc.step()
            extra += 1
    torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["extra_untimed_steps_for_sampling"] = extra
    cyc.check_errors()
    step_ms = np.array([e0.elapsed_time(e1) for e0, e1 in ev])
    total_ms = float(step_ms.sum())
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    value = world * B * args.steps / (total_ms * 1e-3)
    st = cyc.o["status"].cpu().numpy()
    itn = cyc.o["iters"].cpu().numpy()
    ent, col = cyc.o["entangled"].cpu().numpy(), cyc.o["collide"].cpu().numpy()

    # ---------------- per-stage times (separate short pass, synchronised between stages)
    cyc.profile = True
    stage = {}
    for it in range(5):
        cyc.upload(hins[it % len(hins)])
        flush.zero_()
        cyc.step()
        for k2, v in cyc.stage_ms.items():
            stage.setdefault(k2, []).append(v)
    cyc.profile = False
    stage = {k2: float(np.mean(v)) for k2, v in stage.items()}

    # ---------------- e2e: pinned host inputs -> H2D -> all kernels -> D2H, every step
    for it in range(2):
        cyc.step_from_host(hins[it % len(hins)], hout)
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for it in range(args.steps):
        h2d, d2h = cyc.step_from_host(hins[it % len(hins)], hout)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e = world * B * args.steps / e2e_s

    if rank == 0:
        kt = np.array(ktimes)
        dom = int(np.argmax(kt.mean(axis=0)))
        dom_ms = float(kt[:, dom].mean())
        alg_bytes = float(np.mean([sc.batch.algorithmic_bytes() for sc in scenes]))
        if scenes[0].batch.hull_xy.shape[0] == 0:   # hulls were built on the device only: count their real vertices
            cnt_b = cyc.o["hull_cnt_g"].index_select(0, cyc.d["group"].long())          # [B][N][8]
            kn = torch.from_numpy(scenes[0].known.astype(np.bool_)).to(dev)[:, :, None]
            alg_bytes += 16.0 * float((cnt_b * kn).sum().item()) + 16.0 * float(scenes[0].known.sum()) * par.num_pol
        peak, which = measured_peak_gbs()
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        line = {"metric": "replans_per_sec", "value": value, "unit": "replans/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": scaling_kind(args.workload), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(par, world, args.workload),
                "p50_ms_per_replan_cycle": float(np.median(step_ms)),
                "e2e": {"value": e2e, "unit": "replans/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches),
                "kernels_ms": {"k_lines": float(kt[:, 0].mean()), "k_qp": float(kt[:, 1].mean())},
                "stage_ms": stage,
                "roofline": {"bound": "hbm", "kernel": ["k_lines", "k_qp"][dom], "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(["k_lines", "k_qp"][dom]) if args.workload == "grid64" and world == 1 else None,
                             "peak_source": which,
                             "algorithmic_bytes_per_launch": alg_bytes,
                             "note": "latency/FP64-issue bound at this size, not HBM bound (DESIGN.md)"},
                "status_hist": {str(k2): int((st == k2).sum()) for k2 in (0, 1, 2)},
                "postcheck": {"entangled": int(ent.sum()), "collide": int(col.sum())},
                "ipm_iters_mean": float(itn.sum(axis=1).mean()),
                "clocks": clocks}
        if world == 1 and not args.no_cpu and args.workload == "grid64":
            v, reps, el, ref = cpu_baseline(scenes[0], args.cpu_budget, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": v, "unit": "replans/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"{B} agents x {reps} whole cycles of scene 0 ({el:.1f} s), "
                                              "oracle/neptune_oracle.c orc_cycle_batch"}
        if world == 1 and args.workload == "grid64" and not args.no_front_end:
            if args.front_end_world > 1:   # the 64-agents-per-GPU world of K GPUs, searched (and replanned) by this one GPU
                from neptune_b200.scenes import make_scene
                par_fe = world_params(args.front_end_world)
                agents_fe = np.arange(par_fe.num_of_agents)
                gen = capi.Solver(par_fe, device=local_rank)
                scene_fe = make_scene(par_fe, SEED, agents=agents_fe, ent_backend=capi.DeviceEntBackend(gen))
                gen.close()
                line["front_end"] = measure_front_end(par_fe, agents_fe, dev, scene_fe, None, args.front_end_steps, not args.no_cpu)
            else:
                line["front_end"] = measure_front_end(par, agents, dev, scenes[0], static, args.front_end_steps, not args.no_cpu)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=4)
    ap.add_argument("--workload", default="grid64", choices=["grid64", "grid1024"],
                    help="grid64: configs[3] family, 64 agents per GPU (default); grid1024: configs[4], fixed world")
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="plain launches instead of CUDA-graph replay")
    ap.add_argument("--no-front-end", action="store_true", help="skip the front-end (K0 search) measurement")
    ap.add_argument("--front-end-steps", type=int, default=10)
    ap.add_argument("--front-end-world", type=int, default=1,
                    help="front_end block on the world of K x 64 agents, all searched by ONE GPU (default 1 = the bench world)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

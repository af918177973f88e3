#!/usr/bin/env python
"""bench.py -- replans/sec of the NEPTUNE replan hot path on B200 (BASELINE.json metric).

One "step" = one replan cycle of the library's device-resident cycle (nb_cycle_*, neptune_b200.cycle.ReplanCycle):
trajCB bookkeeping + hulls / samples of every other agent's committed trajectory (K1), PredictAlphasBetas (K3),
separating-line LPs + pruning (K2), the trajectory QP with the reference's fallback path (K4), the post-check (K5 GJK +
the gated K3 re-check), commit with the DynTraj header.  Records stay on the device: cycle k plans against the records
committed in cycle k - 2 and post-checks against those of cycle k - 1; with N > 1 ranks the commit kernel stores them
into every rank's ring over NVLink (peer memory), no collective call.

Workloads (BASELINE.json configs):
  grid64   (default) configs[3] family: a grid world with 64 agents per GPU (N = 1 is exactly "64 agents synthetic random
           goals"); every agent plans against ALL other agents of the world (faithful, no culling): weak scaling.
  grid1024 configs[4]: 1024 agents / 200 static obstacles, fixed world split over the ranks: strong scaling.  A grid1024
           block also rides in the default line at every --gpus N (skip with --no-grid1024).
  single | mtlp5 | obst8  configs[0..2] as throughput lines (many instances of the small world batched on one GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

`--impl reference` times the CPU oracle (the restatement of the reference algorithm: Gurobi / GLPK cannot be installed
here; `kind: "port"`, built -O3 -march=native) on the host cores for the same config and metric.
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
FP64_NOMINAL_GFLOPS = 37000.0   # B200 FP64 (non-tensor) nominal; no measured FP64 peak in MEASURED_PEAKS.json


def world_params(n_gpus: int, workload: str = "grid64"):
    from neptune_b200.params import Params, config
    if workload in ("grid1024", "single", "mtlp5", "obst8"):
        return config(workload)
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
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return int(json.load(f)["per_launch_dram_bytes"][kernel])
        except Exception:
            continue
    return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_cycle(scene, threads: int):
    """One whole cycle of the scene's agents on the CPU oracle (hulls/samples, predict, LPs + QP, post-check with its
    fresh PredictAlphasBetas + entangle re-check): the restatement of the reference algorithm, all host threads."""
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
    t0, reps, per = time.perf_counter(), 0, []
    while True:
        t1 = time.perf_counter()
        out = cpu_cycle(scene, threads)
        per.append(time.perf_counter() - t1)
        reps += 1
        el = time.perf_counter() - t0
        if el >= budget_s or reps >= 2000:
            break
    return scene.batch.B * reps / el, reps, el, out, per


def qp_flops(n_int, iters, kept_lines):
    """Algorithmic FP64 flops of the interior-point solves (per iteration 2 m r^2 + r^3 / 3 with m inequality
    rows and r reduced variables), from the iteration counts and kept-line counts the run reports."""
    total = 0.0
    for n, (i0, i1), nl in zip(n_int, iters, kept_lines):
        m = 48 * int(n) + 4 * int(nl)
        for mode, it in ((0, i0), (1, i1)):
            r = 3 * (max(int(n) - 2, 0) if mode == 0 else int(n))
            total += float(it) * (2.0 * m * r * r + r ** 3 / 3.0)
    return total


# ------------------------------------------------------------------------------------------------ one measured world
def measure_cycle(args, par, agents, rank, world, dev, workload, steps, warmup, n_scenes, static=None, front_end=False,
                  cpu=False, clocks=False):
    """Builds the cycle for this rank's agents, runs warm-up + `steps` timed cycles (CUDA events on the cycle's stream,
    L2 flushed between iterations, inputs resident), the end-to-end leg (pinned H2D + D2H every step) and a profiled
    pass.  Returns a dict (rank 0 fills the aggregate fields)."""
    import torch
    import torch.distributed as dist

    from neptune_b200 import capi
    from neptune_b200.cycle import STAGES, ReplanCycle
    from neptune_b200.scenes import search_host_inputs

    cyc = ReplanCycle(par, agents, dev, static=static, world=world, rank=rank, front_end=front_end)
    cyc.connect()
    # the workload generator fills entanglement states through the product's own K3 kernels
    _, scenes = make_world(args.gpus, rank, n_scenes, capi.DeviceEntBackend(cyc.solver), workload, agents=agents)
    B = cyc.B
    fes = [search_host_inputs(sc, SEED + 1) for sc in scenes] if front_end else [None] * len(scenes)
    hins = [cyc.host_inputs(sc, fe) for sc, fe in zip(scenes, fes)]
    hout = cyc.host_outputs()
    cyc.seed_records(cyc.records_of(scenes[0]))
    lib = capi.lib()
    st = cyc.stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        st.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        with torch.cuda.stream(st):
            flush.zero_()

    # ---------------- warm-up, then graph capture (one more real cycle inside)
    for it in range(max(1, warmup - 1)):
        cyc.upload(hins[it % len(hins)])
        flush_l2()
        cyc.step()
    barrier()
    cyc.check_errors()
    if not args.no_graph:
        cyc.capture()
    # ---------------- value: inputs resident in HBM before the timed region, CUDA events on the cycle's stream
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sampler = ClockSampler(dev.index or 0) if clocks else None
    # the sampler comes up under the same load, and the GPU settles: a lease starts idle at low clocks and the first
    # hundred graph replays run ~20 % slower than the steady state (measured with --steps 20 / 50 / 400), so every run --
    # whatever W -- gets at least 150 untimed cycles before the timed ones
    for it in range(max(3, 150 - warmup)):
        cyc.upload(hins[it % len(hins)])
        flush_l2()
        cyc.step()
    barrier()
    for k in range(steps):
        cyc.upload(hins[k % len(hins)])       # untimed: this leg measures with inputs already in HBM
        flush_l2()
        cyc.align()                           # N > 1: all ranks enter the timed cycle together (device-side barrier), so
        ev[k][0].record(st)                   # one rank's untimed upload is not another rank's exchange wait
        cyc.step()
        ev[k][1].record(st)
    barrier()
    launches = steps * cyc.launches_per_cycle
    extra = 0
    if sampler is not None:
        # a short run can end before the first 20 ms sample: keep the same cycle running, untimed, until there are
        # three samples under load (bounded; every rank runs the same number of cycles, a cycle ends in an exchange)
        n_extra = max(0, 400 - steps) if world > 1 else 0
        t_extra = time.time()
        while (world > 1 and extra < n_extra) or (world == 1 and len(sampler.rows) < 3 and sampler.proc is not None and time.time() - t_extra < 2.0):
            flush_l2()
            cyc.step()
            extra += 1
            if extra % 16 == 0:
                st.synchronize()
        barrier()
    clocks_rec = sampler.stop() if sampler is not None else None
    if clocks_rec is not None:
        clocks_rec["extra_untimed_steps_for_sampling"] = extra
    cyc.check_errors()
    step_ms = np.array([e0.elapsed_time(e1) for e0, e1 in ev])
    total_ms = float(step_ms.sum())
    rank_ms = [total_ms]
    rank_mhz = [clocks_rec["sm_mhz"] if clocks_rec else None]
    if world > 1:
        tt = torch.tensor([total_ms, float(rank_mhz[0] or 0.0)], dtype=torch.float64, device=dev)
        allt = torch.empty(2 * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt, tt)
        allt = allt.cpu().numpy().reshape(world, 2)
        rank_ms = [float(x) for x in allt[:, 0]]
        rank_mhz = [float(x) for x in allt[:, 1]]
        total_ms = max(rank_ms)
    n_agents_all = par.num_of_agents if workload != "grid64" else world * B
    value = n_agents_all * steps / (total_ms * 1e-3)
    cyc.download(hout)
    st.synchronize()
    status, itn = hout["status"].copy(), hout["iters"].copy()
    ent, col = hout["entangled"].copy(), hout["collide"].copy()
    n_int = hins[(steps - 1) % len(hins)]["n_int"].copy()

    # ---------------- per-stage times and the two back-end kernels (plain launches, one stream, synchronised)
    lib.nb_set_profiling(cyc.solver.handle, 1)
    stage, kt = {k: [] for k in STAGES}, []
    for it in range(5):
        cyc.upload(hins[it % len(hins)])
        flush_l2()
        sm = cyc.step_profiled()
        ms = (C.c_double * 2)()
        lib.nb_kernel_times(cyc.solver.handle, ms, 2)
        if it >= 1:
            kt.append((ms[0], ms[1]))
            for k2, v in sm.items():
                stage[k2].append(v)
    lib.nb_set_profiling(cyc.solver.handle, 0)
    stage = {k2: float(np.mean(v)) for k2, v in stage.items()}
    kt = np.array(kt)
    rank_stage = None
    if world > 1:   # the same stages on every rank: where a slow rank loses its time
        tt = torch.tensor([stage[k2] for k2 in STAGES], dtype=torch.float64, device=dev)
        allt = torch.empty(len(STAGES) * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt, tt)
        allt = allt.cpu().numpy().reshape(world, len(STAGES))
        rank_stage = {k2: [float(x) for x in allt[:, i]] for i, k2 in enumerate(STAGES)}

    # ---------------- e2e: pinned host inputs -> H2D -> all kernels -> D2H, every step
    for it in range(2):
        cyc.step_from_host(hins[it % len(hins)], hout)
    barrier()
    h2d = d2h = 0
    e2e_runs = []
    for leg in range(3):   # wall clock on a shared host: three legs of `steps` cycles each, the best one is reported (all are listed)
        t0 = time.perf_counter()
        for it in range(steps):
            h2d, d2h = cyc.step_from_host(hins[it % len(hins)], hout)
        barrier()
        leg_s = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([leg_s], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            leg_s = float(tt.item())
        e2e_runs.append(n_agents_all * steps / leg_s)
    e2e = max(e2e_runs)
    cyc.check_errors()

    out = dict(value=value, ms_per_step=total_ms / steps, p50=float(np.median(step_ms)), p95=float(np.percentile(step_ms, 95)),
               e2e=e2e, e2e_runs=e2e_runs, h2d=int(h2d), d2h=int(d2h), launches=int(launches), launches_per_cycle=int(cyc.launches_per_cycle),
               kernels_ms={"k_lines": float(kt[:, 0].mean()), "k_qp": float(kt[:, 1].mean())}, stage_ms=stage,
               rank_ms_per_step=[x / steps for x in rank_ms], rank_sm_mhz=rank_mhz, rank_stage_ms=rank_stage, clocks=clocks_rec,
               status_hist={str(k2): int((status == k2).sum()) for k2 in (0, 1, 2)},
               postcheck={"entangled": int(ent.sum()), "collide": int(col.sum())},
               ipm_iters_mean=float(itn.sum(axis=1).mean()), agents=int(n_agents_all), B=int(B))
    # algorithmic bytes and FP64 flops of the dominant back-end kernel
    alg_bytes = float(np.mean([sc.batch.algorithmic_bytes() for sc in scenes]))
    if scenes[0].batch.hull_xy.shape[0] == 0:   # hulls were built on the device only: count their real vertices
        G = hins[0].G
        cnt = cyc.fetch("hull_cnt", (G, par.num_of_agents, 8), np.int32)[hins[0]["group"]]
        alg_bytes += 16.0 * float((cnt * (scenes[0].known[:, :, None] > 0)).sum()) + 16.0 * float(scenes[0].known.sum()) * par.num_pol
    out["alg_bytes"] = alg_bytes
    kept = np.zeros((B, 8), np.int32)
    capi._check(lib.nb_kept_lines(cyc.solver.handle, kept.ctypes.data_as(C.c_void_p), B), "nb_kept_lines")
    out["kept_lines_mean"] = float(kept.sum(axis=1).mean())
    out["qp_flops"] = qp_flops(n_int, itn, kept.sum(axis=1))
    if front_end:
        stats, fstat = hout["fe_stats"].copy(), hout["fe_status"].copy()
        out["front_end"] = {"pops": int(stats[:, 1].sum()), "pops_max": int(stats[:, 1].max()), "nodes_max": int(stats[:, 0].max()),
                            "status_hist": {"runtime": int((fstat == 0).sum()), "goal": int((fstat == 1).sum()), "empty": int((fstat == 2).sum())},
                            "solved": int(hout["fe_solved"].sum())}
    if cpu:
        v, reps, el, ref, per = cpu_baseline(scenes[0], args.cpu_budget, os.cpu_count() or 1)
        per_replan = np.array(per) * 1e3 / 1.0      # ms per whole cycle of B agents on all cores
        out["cpu_baseline"] = {"value": v, "unit": "replans/s", "cores": os.cpu_count() or 1, "kind": "port",
                               "p50_ms_per_cycle": float(np.median(per_replan)), "p95_ms_per_cycle": float(np.percentile(per_replan, 95)),
                               "build": "gcc -O3 -march=native -fopenmp (oracle/Makefile)",
                               "sample": f"{B} agents x {reps} whole cycles of scene 0 ({el:.1f} s), oracle/neptune_oracle.c orc_cycle_batch: "
                                         "full 12n-variable interior-point solve per replan, every agent builds its own hulls (no window "
                                         "sharing), no line pruning -- a restatement of the reference algorithm, not Gurobi / GLPK"}
    out["_scene0"], out["_hin0"], out["_fe0"], out["_cyc"] = scenes[0], hins[0], fes[0], cyc
    return out


def front_end_cpu(par, scene, hin, fe, cyc, static_longest):
    """The oracle's search on exactly the inputs the device search had (the cycle's own hulls, samples, entangle_state_A)."""
    from neptune_b200.search import SearchBatch, SearchResult
    from oracle import oracle as orc
    M, N, B, cap, NA, S = par.num_of_static_obst, par.num_of_agents, cyc.B, par.ent_cap, par.NA, par.num_sample_per_interval
    G = hin.G
    sb = SearchBatch(
        par=par, agent_id=scene.batch.agent_id.copy(), init=fe["init"], goal=fe["goal"], coeffs_z=fe["coeffs_z"],
        group=hin["group"].copy(), hull_xy=cyc.fetch("hull_xy", (G, N, 8, 24, 2), np.float64), hull_cnt=cyc.fetch("hull_cnt", (G, N, 8), np.int32),
        samp=cyc.fetch("samp", (G, N, par.num_pol, S + 1, 2), np.float64), known=scene.known.copy(),
        es_cnt=cyc.fetch("esA_cnt", (B, 2), np.int32), es_alpha=cyc.fetch("esA_alpha", (B, cap, 2), np.int32),
        es_beta=cyc.fetch("esA_beta", (B, cap), np.float64), es_bend=cyc.fetch("esA_bend", (B, cap), np.int32),
        es_active=cyc.fetch("esA_active", (B, NA), np.int32), bp_cnt=scene.batch.bp_cnt, bp_xy=scene.batch.bp_xy, comb=fe["comb"],
        st_ptr=scene.batch.st_ptr, st_xy=scene.batch.st_xy, strep=np.asarray(scene.strep, np.float64).reshape(M, 2, 2),
        st_longest=static_longest if M else np.zeros((0, 2)))
    ref = SearchResult.empty(sb)
    nt = os.cpu_count() or 1
    orc.search_batch(sb, ref, nt)
    t0, reps = time.perf_counter(), 3
    for _ in range(reps):
        orc.search_batch(sb, ref, nt)
    cpu_ms = 1e3 * (time.perf_counter() - t0) / reps
    return ref, {"ms_search": cpu_ms, "cores": nt, "kind": "port",
                 "sample": f"{B} searches x {reps}, oracle/neptune_search.c orc_search_batch"}


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
              f"each against all {par.num_of_agents - 1} others; oracle/neptune_oracle.c orc_cycle_batch, gcc -O3 -march=native")
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
    elif workload == "grid64":
        wl = (f"grid world, {AGENTS_PER_GPU} agents/GPU x {n_gpus} GPU = {par.num_of_agents} agents "
              "(BASELINE.json configs[3] at N=1), random goals, no static obstacles, faithful (no culling)")
    else:
        wl = {"single": "BASELINE.json configs[0]: single agent, obstacle-free, 3-interval path (derived from neptune_single_benchmark.yaml), 4096 instances per batch",
              "mtlp5": "BASELINE.json configs[1]: 5 agents obstacle-free random goals (neptune_mtlp_benchmark.yaml), 256 independent worlds per batch",
              "obst8": "BASELINE.json configs[2]: 8 agents, 9 static obstacles (neptune_multi_obstacle.yaml), separator LPs + entangle check, 128 independent worlds per batch"}[workload]
    return {"workload": wl, "agents": par.num_of_agents, "agents_per_gpu": par.num_of_agents // n_gpus
            if workload != "grid64" else AGENTS_PER_GPU, "static_obstacles": par.num_of_static_obst,
            "num_pol": par.num_pol, "T_span": par.T_span, "seed": SEED if workload == "grid64" else 5005,
            "l2": "flushed between timed iterations (256 MiB write)",
            "parallelism": f"agents sharded over {n_gpus} rank(s); committed trajectories exchanged by peer-to-peer stores from the commit kernel (no collective call)"
                           + ("; every timed cycle starts at a device-side rank barrier (nb_cycle_align), so the untimed input upload of one rank is not counted as another rank's exchange wait" if n_gpus > 1 else "")}


def run_small(args):
    """configs[0..2] as throughput lines: many independent instances of the small world in one nb_replan_batch
    (the back end: LPs -> QP), inputs resident, CUDA events; printed as one JSON line."""
    import torch

    from neptune_b200 import capi
    from neptune_b200.params import config
    from neptune_b200.scenes import make_scene
    par = config(args.workload)
    n_worlds = {"single": 4096, "mtlp5": 256, "obst8": 128}[args.workload]
    kw = dict(n_fixed=3) if args.workload == "single" else dict(sync=False)
    seed0 = {"single": 1001, "mtlp5": 2002, "obst8": 3003}[args.workload]
    s = capi.Solver(par)
    scenes = []
    for k in range(min(n_worlds, 64)):      # 64 distinct seeded worlds, tiled to n_worlds instances
        if par.num_of_static_obst and k == 0:
            sc0 = make_scene(par, seed0, **kw)
            s.set_static(sc0.batch.st_ptr, sc0.batch.st_xy, sc0.strep)
        scenes.append(make_scene(par, seed0 + k, ent_backend=capi.DeviceEntBackend(s), **kw))
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    # device-resident batches, one per world (B = agents of the world), all launched back to back per step
    res_h = [capi.ReplanResult.empty(sc.batch, with_lines=False) for sc in scenes]
    reps = n_worlds // len(scenes)
    t_ms = []
    for it in range(args.warmup + args.steps):
        with torch.cuda.stream(st):
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(st)
        for _ in range(reps):
            for sc, r in zip(scenes, res_h):
                a = capi.host_args(sc.batch, r)
                capi._check(capi.lib().nb_replan_batch(s.handle, C.byref(a), C.c_void_p(st.cuda_stream)), "nb_replan_batch")
        e1.record(st)
        st.synchronize()
        if it >= args.warmup:
            t_ms.append(1e3 * (time.perf_counter() - t0))
    replans = n_worlds * par.num_of_agents
    ms = float(np.mean(t_ms))
    status = np.concatenate([r.status for r in res_h])
    line = {"metric": "replans_per_sec", "value": replans / (ms * 1e-3), "unit": "replans/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(par, 1, args.workload),
            "p50_ms_per_step": float(np.median(t_ms)), "p95_ms_per_step": float(np.percentile(t_ms, 95)),
            "note": "end-to-end through the C-ABI with HOST buffers (nb_replan_batch, NB_HOST): staging copies, LPs, QP and the "
                    "result copies are all inside the timed region; one call per world, as one reference process per agent group would",
            "status_hist": {str(k2): int((status == k2).sum()) for k2 in (0, 1, 2)},
            "gpu_launches": int(2 * reps * len(scenes) * args.steps)}
    # CPU arm on the same worlds
    if not args.no_cpu:
        from neptune_b200.batch import ReplanResult
        from oracle import oracle as orc
        nt = os.cpu_count() or 1
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < args.cpu_budget:
            for sc in scenes:
                r = ReplanResult.empty(sc.batch, with_lines=False)
                orc.replan_batch(sc.batch, r, nt)
                n += sc.batch.B
        el = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / el, "unit": "replans/s", "cores": nt, "kind": "port",
                                "sample": f"{n} back-end replans (orc_replan_batch) in {el:.1f} s"}
    print(json.dumps(line))
    s.close()


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.workload in ("single", "mtlp5", "obst8"):
        if rank == 0:
            run_small(args)
        if world > 1:
            dist.destroy_process_group()
        return

    def static_of(par, agents):
        if not par.num_of_static_obst:   # static obstacles are part of the world: build them once, before the solver
            return None
        from neptune_b200.scenes import make_scene
        s0 = make_scene(par, 5005, agents=agents[:1], pack_hulls=False)
        return (s0.batch.st_ptr, s0.batch.st_xy, s0.strep)

    par = world_params(args.gpus, args.workload)
    agents = rank_agents(par, args.gpus, rank, args.workload)
    main = measure_cycle(args, par, agents, rank, world, dev, args.workload, args.steps, args.warmup, args.scenes,
                         static=static_of(par, agents), cpu=(world == 1 and not args.no_cpu and args.workload == "grid64"), clocks=True)
    line = None
    if rank == 0:
        m = main
        dom = "k_qp" if m["kernels_ms"]["k_qp"] >= m["kernels_ms"]["k_lines"] else "k_lines"
        dom_ms = m["kernels_ms"][dom]
        peak, which = measured_peak_gbs()
        achieved = m["alg_bytes"] / (dom_ms * 1e-3) / 1e9
        gflops = m["qp_flops"] / (m["kernels_ms"]["k_qp"] * 1e-3) / 1e9
        line = {"metric": "replans_per_sec", "value": m["value"], "unit": "replans/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
                "scaling": scaling_kind(args.workload), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(par, world, args.workload),
                "p50_ms_per_replan_cycle": m["p50"], "p95_ms_per_replan_cycle": m["p95"],
                "e2e": {"value": m["e2e"], "unit": "replans/s", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                        "legs": m["e2e_runs"], "note": "wall clock through the C ABI with host buffers; best of three legs of `steps` cycles"},
                "gpu_launches": m["launches"], "launches_per_cycle": m["launches_per_cycle"],
                "kernels_ms": m["kernels_ms"], "stage_ms": m["stage_ms"],
                "exchange": {"kind": "peer-to-peer stores from k_publish + flag wait (k_wait_peers)", "wait_ms": m["stage_ms"]["exchange_wait"],
                             "commit_ms": m["stage_ms"]["commit"], "rank_ms_per_step": m["rank_ms_per_step"],
                             "rank_sm_mhz": m["rank_sm_mhz"], "rank_stage_ms": m["rank_stage_ms"]},
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": ncu_traffic(dom) if args.workload == "grid64" and world == 1 else None,
                             "peak_source": which, "algorithmic_bytes_per_launch": m["alg_bytes"],
                             "fp64": {"kernel": "k_qp", "gflops": gflops, "nominal_peak_gflops": FP64_NOMINAL_GFLOPS,
                                      "frac_of_nominal": gflops / FP64_NOMINAL_GFLOPS,
                                      "flops_per_launch": m["qp_flops"], "kept_lines_per_agent": m["kept_lines_mean"],
                                      "formula": "sum over solves of iterations x (2 m r^2 + r^3 / 3), SURVEY 8d"},
                             "note": "latency / FP64-issue bound at this size, not HBM bound (DESIGN.md)"},
                "status_hist": m["status_hist"], "postcheck": m["postcheck"], "ipm_iters_mean": m["ipm_iters_mean"],
                "clocks": m["clocks"]}
        if "cpu_baseline" in m:
            line["cpu_baseline"] = m["cpu_baseline"]
    cyc_main = main.pop("_cyc")
    # ---------------- front end (K0) on the same world: one more cycle object with the search inside
    if world == 1 and args.workload == "grid64" and not args.no_front_end:
        from neptune_b200.search import static_longest_dist
        cyc_main.close()
        par_fe, agents_fe = par, agents
        if args.front_end_world > 1:   # the 64-agents-per-GPU world of K GPUs, searched (and replanned) by this one GPU
            par_fe = world_params(args.front_end_world)
            agents_fe = np.arange(par_fe.num_of_agents)
        saved = args.gpus
        args.gpus = args.front_end_world if args.front_end_world > 1 else args.gpus
        fe = measure_cycle(args, par_fe, agents_fe, 0, 1, dev, "grid64", args.front_end_steps, 3, 1, front_end=True)
        args.gpus = saved
        blk = {"kernel": "k_search", "agents": fe["B"], "ms_search": fe["stage_ms"]["front_end"], "ms_full_cycle": fe["ms_per_step"],
               "full_replans_per_s": fe["value"], "e2e_full_replans_per_s": fe["e2e"],
               "pops_per_s": fe["front_end"]["pops"] / (fe["stage_ms"]["front_end"] * 1e-3),
               "max_expansions": par_fe.search_max_expansions, "max_nodes": par_fe.search_max_nodes}
        blk.update(fe["front_end"])
        if not args.no_cpu:
            ref, cpu_blk = front_end_cpu(par_fe, fe["_scene0"], fe["_hin0"], fe["_fe0"], fe["_cyc"], None)
            cpu_blk["identical"] = bool(ref.solved.sum() == fe["front_end"]["solved"])
            blk["cpu_oracle"] = cpu_blk
        fe["_cyc"].close()
        line["front_end"] = blk
    else:
        cyc_main.close()
    # ---------------- configs[4] block: 1024 agents / 200 static obstacles split over the ranks (strong scaling)
    if args.workload == "grid64" and not args.no_grid1024:
        par5 = world_params(args.gpus, "grid1024")
        agents5 = rank_agents(par5, args.gpus, rank, "grid1024")
        g = measure_cycle(args, par5, agents5, rank, world, dev, "grid1024", args.grid1024_steps, 3, 1, static=static_of(par5, agents5))
        g.pop("_cyc").close()
        if rank == 0:
            line["grid1024"] = {"workload": config_dict(par5, world, "grid1024")["workload"], "scaling": "strong",
                                "replans_per_s": g["value"], "ms_per_cycle": g["ms_per_step"], "p50_ms": g["p50"], "p95_ms": g["p95"],
                                "e2e_replans_per_s": g["e2e"], "h2d_bytes_per_step": g["h2d"], "agents_per_gpu": g["B"],
                                "kernels_ms": g["kernels_ms"], "stage_ms": g["stage_ms"], "rank_ms_per_step": g["rank_ms_per_step"],
                                "rank_stage_ms": g["rank_stage_ms"],
                                "status_hist": g["status_hist"], "steps": args.grid1024_steps}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=4)
    ap.add_argument("--workload", default="grid64", choices=["grid64", "grid1024", "single", "mtlp5", "obst8"],
                    help="grid64: configs[3] family, 64 agents per GPU (default); grid1024: configs[4], fixed world; "
                         "single / mtlp5 / obst8: configs[0..2] as back-end throughput lines")
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="plain launches instead of CUDA-graph replay")
    ap.add_argument("--no-front-end", action="store_true", help="skip the front-end (K0 search) measurement")
    ap.add_argument("--no-grid1024", action="store_true", help="skip the configs[4] block of the default line")
    ap.add_argument("--grid1024-steps", type=int, default=20)
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

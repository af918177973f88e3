"""Front-end search (KinodynamicSearch, reference neptune/src/kinodynamic_search.cpp).

CPU tests pin the oracle (oracle/neptune_search.c): its open list against the real std::priority_queue of
this image's libstdc++, its results against properties the reference's algorithm guarantees, and the
device code compiled for one host lane (tests/emul) against it, bit for bit.  GPU tests (marked) run the
CUDA kernel through the C-ABI against the oracle: every output is integer-decided or copied FP64, so the
bar is bit-exact on every field.
"""
import ctypes as C
import dataclasses
import os
import subprocess

import numpy as np
import pytest

from neptune_b200 import config
from neptune_b200.batch import NPOL
from neptune_b200.scenes import make_scene, make_search_batch
from neptune_b200.search import SearchResult
from tests.ent_backends import OracleEntBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(oracle, cfg, seed, per_agent_order=True, multi_bend=False, **mods):
    par = config(cfg)
    for k, v in mods.items():
        setattr(par, k, v)
    sc = make_scene(par, seed, sync=False, ent_backend=OracleEntBackend(oracle), group_hulls=True)
    sb = make_search_batch(sc, seed + 1, per_agent_order=per_agent_order)
    if multi_bend:   # some tethers already wrap around a contact point: bendPtsForAgents_[j] = [base, point]
        sb.bp_cnt, sb.bp_xy = sb.bp_cnt.copy(), sb.bp_xy.copy()
        rng = np.random.default_rng(seed)
        for j in range(0, par.num_of_agents, 2):
            sb.bp_cnt[j] = 2
            sb.bp_xy[j, 1] = sb.bp_xy[j, 0] + rng.normal(0, 2.0, size=2)
    return sc, sb


def _oracle_search(oracle, sb, expect_rc=0):
    res = SearchResult.empty(sb)
    assert oracle.search_batch(sb, res, 4) == expect_rc
    return res


def _assert_same(a: SearchResult, b: SearchResult):
    for f in dataclasses.fields(a):
        x, y = getattr(a, f.name), getattr(b, f.name)
        assert np.array_equal(x, y), (f.name, np.argwhere(x != y)[:4].tolist())


# ------------------------------------------------------------------------------------------ oracle pins
def test_open_list_matches_real_std_priority_queue(oracle):
    """The oracle restates libstdc++'s push_heap / pop_heap; CompareCost is not a strict weak order and the
    reference overwrites (g, h) of queued nodes, so the pop order depends on the exact sift algorithm.  Replay
    random scripts (costs quantised so that near-ties within 1e-5 are common) through the real
    std::priority_queue (tests/cpp/heap_check.cpp) and through the restatement."""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "heap_check")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "_build/heap_check"], check=True,
                       capture_output=True)
    rng = np.random.default_rng(11)
    for trial in range(20):
        n_ids = int(rng.integers(5, 400))
        ops, vals, next_id, n_q = [], [], 0, 0
        while (next_id < n_ids or n_q > 0) and len(ops) < 4000:
            r = rng.random()
            if next_id < n_ids and (r < 0.55 or n_q == 0):
                g = float(rng.integers(0, 40)) * 0.25 + float(rng.integers(0, 3) - 1) * 4e-6
                ops.append((0, next_id)), vals.append((g, float(rng.integers(0, 40)) * 0.25))
                next_id, n_q = next_id + 1, n_q + 1
            elif r < 0.85:
                ops.append((1, 0)), vals.append((0.0, 0.0))
                n_q -= 1
            else:  # overwrite the keys of some node pushed so far (it may still be queued)
                ops.append((2, int(rng.integers(next_id)))), vals.append((float(rng.integers(0, 40)) * 0.25, float(rng.integers(0, 40)) * 0.25))
        ops_a, vals_a = np.array(ops, np.int32), np.array(vals, np.float64)
        txt = f"{len(ops)} {n_ids} 1.1\n" + "".join(f"{k} {i} {g!r} {h!r}\n" for (k, i), (g, h) in zip(ops, vals))
        ref = subprocess.run([exe], input=txt, capture_output=True, text=True, check=True).stdout.split()
        out = np.zeros(len(ops) + 1, np.int32)
        f = oracle.lib().orc_heap_replay
        f.restype = C.c_int
        n = f(C.c_int(len(ops)), ops_a.ctypes.data_as(C.c_void_p), vals_a.ctypes.data_as(C.c_void_p), C.c_int(n_ids),
              C.c_double(1.1), out.ctypes.data_as(C.c_void_p))
        assert n == len(ref) and out[:n].tolist() == [int(x) for x in ref], trial


@pytest.mark.parametrize("cfg,seed", [("mtlp5", 2002), ("mtlp5", 2005), ("obst8", 3003), ("grid64", 4004)])
def test_oracle_search_properties(oracle, cfg, seed):
    """What the reference's algorithm guarantees about its own result, checked on the oracle's output:
    pieces are the constant-jerk primitives of the 5x5 lattice chained with C2 continuity from the start
    state; acceleration and MINVO velocity control points within bounds (velocity test skipped for the first
    piece, kinodynamic_search.cpp:1307-1320); status 1 ends within goal_radius with no active case above 1;
    entStateVec equals an independent rollout of the chain along the returned path (orc_entangle_rollout)."""
    sc, sb = _batch(oracle, cfg, seed)
    par = sb.par
    res = _oracle_search(oracle, sb)
    T, N = par.T_span, par.num_of_agents
    Ainv, V, _ = oracle.basis(T)
    assert res.solved.sum() >= 1
    jerks = np.linspace(-par.j_max, par.j_max, par.a_star_samp_x)
    for b in range(sb.B):
        assert res.status[b] in (0, 1, 2)
        if not res.solved[b]:
            assert res.n_int[b] == 0
            continue
        n = int(res.n_int[b])
        assert 1 <= n <= par.num_pol and n == min(int(res.stats[b, 2]), par.num_pol)
        st = sb.init[b].copy()
        for i in range(n):
            cx, cy = res.coeff[b, 0, i], res.coeff[b, 1, i]
            assert cx[3] == st[0] and cy[3] == st[1] and cx[2] == st[2] and cy[2] == st[3]
            assert cx[1] == st[4] / 2 and cy[1] == st[5] / 2
            assert np.abs(jerks - cx[0] * 6).min() < 1e-12 and np.abs(jerks - cy[0] * 6).min() < 1e-12
            jx, jy = cx[0] * 6, cy[0] * 6
            st = np.array([st[0] + st[2] * T + st[4] * T * T / 2 + jx * T ** 3 / 6,
                           st[1] + st[3] * T + st[5] * T * T / 2 + jy * T ** 3 / 6,
                           st[2] + st[4] * T + jx * T * T / 2, st[3] + st[5] * T + jy * T * T / 2,
                           st[4] + jx * T, st[5] + jy * T])
            assert np.abs(st[4:]).max() <= par.a_max + 1e-12
            Q = np.stack([cx, cy]) @ Ainv
            assert (Q[0] >= par.x_min).all() and (Q[0] <= par.x_max).all()
            assert (np.hypot(Q[0] - par.pb[sb.agent_id[b] - 1, 0], Q[1] - par.pb[sb.agent_id[b] - 1, 1]) <= par.tetherLength).all()
            if i > 0:
                Qv = np.stack([cx[:3], cy[:3]]) @ V
                assert np.abs(Qv).max() <= par.v_max + 1e-12
            assert np.array_equal(res.coeff[b, 2, i], sb.coeffs_z[b, i])
        if res.status[b] == 1 and res.stats[b, 2] <= par.num_pol:
            assert np.hypot(st[0] - sb.goal[b, 0], st[1] - sb.goal[b, 1]) < par.goal_radius
            assert (res.esv_active[b, n, :N] <= 1).all()
        # independent chain along the path
        es = oracle.EntState(par.ent_cap, par.NA)
        es.n_alpha, es.n_bend = int(sb.es_cnt[b, 0]), int(sb.es_cnt[b, 1])
        es.alpha[:], es.beta[:], es.bend[:], es.active[:] = sb.es_alpha[b], sb.es_beta[b], sb.es_bend[b], sb.es_active[b]
        cxo = oracle.EntCtx(par, int(sb.agent_id[b]) - 1, sb.strep, sb.bp_cnt, sb.bp_xy)
        cxy = np.ascontiguousarray(res.coeff[b, :2, :n, :])
        done, cnt, alpha, beta, bend, active = oracle.entangle_rollout(es, cxo, n, cxy, sb.samp[sb.group[b]], sb.known[b])
        assert done == n
        assert np.array_equal(cnt, res.esv_cnt[b, :n + 1])
        for i in range(n + 1):
            assert np.array_equal(alpha[i, :cnt[i, 0]], res.esv_alpha[b, i, :cnt[i, 0]])
            assert np.array_equal(active[i], res.esv_active[b, i])


def test_oracle_search_budgets(oracle):
    """max_expansions = 0 -> 'runtime reached' before the first pop and no solution; a tiny node pool stops the
    expansion ('run out of memory', kinodynamic_search.cpp:1060-1064) but still returns the closest safe node;
    use_not_reaching_soln = false turns every non-goal result into a failure (:1764, :1789-1799)."""
    _, sb = _batch(oracle, "mtlp5", 2002, search_max_expansions=0)
    r = _oracle_search(oracle, sb)
    assert (r.status == 0).all() and (r.solved == 0).all() and (r.n_int == 0).all() and (r.stats[:, 1] == 0).all()
    _, sb = _batch(oracle, "mtlp5", 2002, search_max_nodes=64)
    r = _oracle_search(oracle, sb)
    assert (r.stats[:, 0] <= 63 + 25).all() and r.solved.any()
    _, sb = _batch(oracle, "grid64", 4004, use_not_reaching_soln=False)
    r = _oracle_search(oracle, sb)
    assert ((r.solved == 1) == (r.status == 1)).all() and (r.status != 1).any()


def test_oracle_search_jerk_order_matters_only_through_ties(oracle):
    """Two different jerk orders explore the same lattice: both reach the goal where one does, and the path
    costs differ by less than one lattice step (the order only breaks ties and voxel collisions)."""
    _, sb = _batch(oracle, "mtlp5", 2003)
    r1 = _oracle_search(oracle, sb)
    sb.comb[:] = sb.comb[:, ::-1]
    r2 = _oracle_search(oracle, sb)
    both = (r1.status == 1) & (r2.status == 1)
    assert both.any()
    assert np.abs(r1.cost[both] - r2.cost[both]).max() < 2.5


# ------------------------------------------------------------------------------------------ device code, one host lane
@pytest.mark.parametrize("cfg,seed,mods", [
    ("mtlp5", 2002, {}), ("mtlp5", 2004, dict(search_max_nodes=128)), ("obst8", 3003, {}),
    ("obst8", 3005, dict(search_max_expansions=1500)), ("grid64", 4004, dict(search_max_expansions=150)),
    ("mtlp5", 2006, dict(enable_entangle_check=False)), ("obst8", 3006, dict(use_not_reaching_soln=False)),
    ("mtlp5", 2007, dict(multi_bend=True)), ("obst8", 3007, dict(multi_bend=True)),
    ("single", 1001, {}),   # configs[0]: one agent, nothing to avoid
])
def test_emulated_kernel_matches_oracle(oracle, cfg, seed, mods):
    from tests.emul import emul
    _, sb = _batch(oracle, cfg, seed, **mods)
    _assert_same(_oracle_search(oracle, sb), emul.search(sb))


def _big_world_batch(oracle):
    """Four agents of BASELINE.json configs[4] (1024 agents / 200 static obstacles): N + M = 1224 tethers, so the
    per-agent working sets no longer fit in shared memory and the kernel runs from global memory / L2."""
    par = config("grid1024")
    par.search_max_expansions = 60
    sc = make_scene(par, 5005, sync=True, ent_backend=OracleEntBackend(oracle), group_hulls=True,
                    agents=np.array([0, 17, 500, 1023]), pack_hulls=False)
    return sc, make_search_batch(sc, 5006, per_agent_order=True)


def test_emulated_kernel_matches_oracle_1024_agents(oracle):
    from tests.emul import emul
    _, sb = _big_world_batch(oracle)
    ref = _oracle_search(oracle, sb)
    assert (ref.stats[:, 1] > 0).all()
    _assert_same(ref, emul.search(sb))


# ------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def capi():
    from neptune_b200 import capi
    capi.lib()
    return capi


def _gpu_solver(capi, sc, sb):
    par = sb.par
    s = capi.Solver(par)
    if par.num_of_static_obst:
        s.set_static(sb.st_ptr, sb.st_xy, sb.strep)
        s.set_static_longest(sb.st_longest)
    s.search_configure()
    return s


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,seeds,mods", [
    ("single", range(1001, 1003), {}),
    ("mtlp5", range(2002, 2008), {}),
    ("obst8", range(3003, 3007), {}),
    ("grid64", range(4004, 4006), {}),
    ("mtlp5", range(2010, 2012), dict(search_max_nodes=128)),
    ("obst8", range(3010, 3012), dict(search_max_expansions=2000, search_max_nodes=12000)),  # open list in global memory
    ("mtlp5", range(2012, 2014), dict(enable_entangle_check=False)),
    ("obst8", range(3012, 3014), dict(use_not_reaching_soln=False)),
    ("mtlp5", range(2014, 2015), dict(search_max_expansions=0)),
    ("obst8", range(3014, 3016), dict(multi_bend=True)),   # generic chain (tethers with contact points)
    ("grid64", range(4006, 4007), dict(multi_bend=True, search_max_expansions=150)),
])
def test_gpu_search_matches_oracle(capi, oracle, cfg, seeds, mods):
    for seed in seeds:
        sc, sb = _batch(oracle, cfg, seed, **mods)
        s = _gpu_solver(capi, sc, sb)
        got = s.search(sb)
        _assert_same(_oracle_search(oracle, sb), got)
        assert s.launch_count() >= 1
        s.close()


@pytest.mark.gpu
def test_gpu_search_matches_oracle_1024_agents(capi, oracle):
    sc, sb = _big_world_batch(oracle)
    s = _gpu_solver(capi, sc, sb)
    _assert_same(_oracle_search(oracle, sb), s.search(sb))
    s.close()


@pytest.mark.gpu
def test_gpu_search_shared_order_and_sync_groups(capi, oracle):
    """One jerk order for the whole batch and synchronous replans (one window group for everybody)."""
    par = config("grid64")
    par.search_max_expansions = 200
    sc = make_scene(par, 4010, sync=True, ent_backend=OracleEntBackend(oracle), group_hulls=True)
    sb = make_search_batch(sc, 77, per_agent_order=False)
    assert sb.G == 1 and sb.comb.ndim == 1
    s = _gpu_solver(capi, sc, sb)
    _assert_same(_oracle_search(oracle, sb), s.search(sb))
    s.close()


@pytest.mark.gpu
def test_gpu_search_capacity_overflow_is_reported(capi, oracle):
    """A node list longer than search_ecap is a storage overflow, reported as NB_ERR_CAPACITY (never silent)."""
    sc, sb = _batch(oracle, "grid64", 4004, search_ecap=2)
    assert oracle.search_batch(sb, SearchResult.empty(sb), 4) == -3
    s = _gpu_solver(capi, sc, sb)
    with pytest.raises(capi.NbError):
        s.search(sb)
    s.close()


@pytest.mark.gpu
def test_gpu_search_feeds_back_end(capi, oracle):
    """search -> replan on the GPU equals search -> replan on the oracle: the front end's pwp_init and
    entStateVec are consumed by the back end in place (neptune.cpp:1509-1519)."""
    from neptune_b200.batch import ReplanResult
    sc, sb = _batch(oracle, "obst8", 3003)
    s = _gpu_solver(capi, sc, sb)
    got = s.search(sb)
    ref = _oracle_search(oracle, sb)
    _assert_same(ref, got)
    ok = np.flatnonzero(ref.solved)
    assert len(ok) >= 2
    for res, run in ((got, lambda bt: s.replan(bt)), (ref, None)):
        bt = sc.batch
        bt.n_int[:] = np.where(res.solved, res.n_int, bt.n_int)
        bt.coeff_init[ok] = res.coeff[ok]
        bt.esv_cnt[ok], bt.esv_alpha[ok], bt.esv_active[ok] = res.esv_cnt[ok], res.esv_alpha[ok], res.esv_active[ok]
        if run is not None:
            out_gpu = run(bt)
        else:
            out_ref = ReplanResult.empty(bt)
            assert oracle.replan_batch(bt, out_ref, 4) == 0
    assert np.array_equal(out_gpu.status, out_ref.status)
    assert np.abs(out_gpu.coeff_out - out_ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(out_ref.coeff_out).max())
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,seed,sync", [("obst8", 3003, False), ("grid64", 4004, True)])
def test_gpu_cycle_with_front_end(capi, oracle, cfg, seed, sync):
    """ReplanCycle(front_end=True): hulls -> predict -> search -> LPs + QP -> post-check -> commit on the device.
    The search inside the cycle equals the oracle's search on the cycle's own (device-built, separately tested)
    hulls, samples and entangle_state_A; the back end then equals the oracle's back end on the search's output;
    agents without a front-end solution are rejected (they keep their previous record)."""
    import torch
    from neptune_b200.batch import ReplanBatch, ReplanResult
    from neptune_b200.cycle import ReplanCycle
    from neptune_b200.scenes import search_host_inputs
    from neptune_b200.search import SearchBatch, static_longest_dist

    par = config(cfg)
    par.search_max_expansions = 200
    M = par.num_of_static_obst
    sc = make_scene(par, seed, sync=sync, ent_backend=OracleEntBackend(oracle))
    strep = np.asarray(sc.strep, np.float64).reshape(M, 2, 2)
    longest = static_longest_dist(sc.static_raw, strep) if M else np.zeros((0, 2))
    dev = torch.device("cuda", 0)
    cyc = ReplanCycle(par, np.arange(par.num_of_agents), dev, static=(sc.batch.st_ptr, sc.batch.st_xy, sc.strep, longest),
                      front_end=True)
    fe = search_host_inputs(sc, seed + 5)
    recs_in = cyc.records_of(sc)
    cyc.seed_records(recs_in)
    hin, hout = cyc.host_inputs(sc, fe), cyc.host_outputs()
    cyc.step_from_host(hin, hout)
    cyc.check_errors()
    B, N, cap, NA, S = cyc.B, par.num_of_agents, par.ent_cap, par.NA, par.num_sample_per_interval
    G = hin.G
    o = dict(hull_xy_g=cyc.fetch("hull_xy", (G, N, 8, 24, 2), np.float64), hull_cnt_g=cyc.fetch("hull_cnt", (G, N, 8), np.int32),
             samp_g=cyc.fetch("samp", (G, N, par.num_pol, S + 1, 2), np.float64),
             esA_cnt=cyc.fetch("esA_cnt", (B, 2), np.int32), esA_alpha=cyc.fetch("esA_alpha", (B, cap, 2), np.int32),
             esA_beta=cyc.fetch("esA_beta", (B, cap), np.float64), esA_bend=cyc.fetch("esA_bend", (B, cap), np.int32),
             esA_active=cyc.fetch("esA_active", (B, NA), np.int32),
             fe_coeff=cyc.fetch("fe_coeff", (B, 3, 8, 4), np.float64), fe_esv_cnt=cyc.fetch("fe_esv_cnt", (B, 9, 2), np.int32),
             fe_esv_alpha=cyc.fetch("fe_esv_alpha", (B, 9, cap, 2), np.int32), fe_esv_beta=cyc.fetch("fe_esv_beta", (B, 9, cap), np.float64),
             fe_esv_bend=cyc.fetch("fe_esv_bend", (B, 9, cap), np.int32), fe_esv_active=cyc.fetch("fe_esv_active", (B, 9, NA), np.int32),
             fe_cost=cyc.fetch("fe_cost", (B,), np.float64), fe_n_int=cyc.fetch("fe_n_int", (B,), np.int32),
             fe_status=hout["fe_status"], fe_solved=hout["fe_solved"], fe_stats=hout["fe_stats"])
    sb = SearchBatch(par=par, agent_id=sc.batch.agent_id.copy(), init=fe["init"], goal=fe["goal"], coeffs_z=fe["coeffs_z"],
                     group=hin["group"].copy(), hull_xy=o["hull_xy_g"], hull_cnt=o["hull_cnt_g"], samp=o["samp_g"],
                     known=sc.known.copy(), es_cnt=o["esA_cnt"], es_alpha=o["esA_alpha"], es_beta=o["esA_beta"],
                     es_bend=o["esA_bend"], es_active=o["esA_active"], bp_cnt=sc.batch.bp_cnt, bp_xy=sc.batch.bp_xy,
                     comb=fe["comb"], st_ptr=sc.batch.st_ptr, st_xy=sc.batch.st_xy, strep=strep, st_longest=longest)
    sb.validate()
    ref = _oracle_search(oracle, sb)
    ok = ref.solved > 0
    assert ok.any()
    for name in ("status", "solved", "stats"):
        assert np.array_equal(getattr(ref, name), o["fe_" + name]), name
    # per-agent outputs of the search: identical where a path was found (the others carry the host-provided path into the
    # back end, "returning with no solution" neptune.cpp:1473-1478, and are rejected in the commit)
    for name in ("n_int", "coeff", "esv_cnt", "esv_alpha", "esv_beta", "esv_bend", "esv_active", "cost"):
        assert np.array_equal(getattr(ref, name)[ok], o["fe_" + name][ok]), name
    assert np.array_equal(hout["fe_n_int"][ok], ref.n_int[ok])
    # the back end on the search's output
    bt = sc.batch
    bt.n_int[ok] = ref.n_int[ok]
    bt.coeff_init[ok] = ref.coeff[ok]
    bt.esv_cnt[ok], bt.esv_alpha[ok], bt.esv_active[ok] = ref.esv_cnt[ok], ref.esv_alpha[ok], ref.esv_active[ok]
    out_ref = ReplanResult.empty(bt)
    assert oracle.replan_batch(bt, out_ref, 4) == 0
    assert np.array_equal(hout["status"], out_ref.status)
    assert np.abs(hout["coeff_out"] - out_ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(out_ref.coeff_out).max())
    new = cyc.records("new")
    assert np.array_equal(new[bt.agent_id[~ok] - 1], recs_in[bt.agent_id[~ok] - 1])      # no path: the previous record stays
    cyc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,seed,sync", [("obst8", 3003, False), ("mtlp5", 2002, True)])
def test_gpu_cycle_kinodynamic_only(capi, oracle, cfg, seed, sync):
    """ReplanCycle(front_end=2) = Neptune::replanKinodynamic (neptune.cpp:1010-1300): the search's own path
    (generatePwpOut :1189) is what safetyCheckAfterReplan sees (:1198) and what is composed with the previous plan
    and published (:1244-1255); no LP, no QP.  The search equals the front_end=1 cycle's (same kernels, tested
    against the oracle above); here: coeff_out is the search's path bit for bit, status 0, and the committed record
    is the oracle's composePieceWisePol of the previous record with that path."""
    import torch
    from neptune_b200.cycle import ReplanCycle
    from neptune_b200.scenes import search_host_inputs
    from neptune_b200.search import static_longest_dist

    par = config(cfg)
    par.search_max_expansions = 200
    M = par.num_of_static_obst
    sc = make_scene(par, seed, sync=sync, ent_backend=OracleEntBackend(oracle))
    strep = np.asarray(sc.strep, np.float64).reshape(M, 2, 2)
    longest = static_longest_dist(sc.static_raw, strep) if M else np.zeros((0, 2))
    dev = torch.device("cuda", 0)
    outs = {}
    for mode in (1, 2):
        cyc = ReplanCycle(par, np.arange(par.num_of_agents), dev, static=(sc.batch.st_ptr, sc.batch.st_xy, sc.strep, longest),
                          front_end=mode)
        fe = search_host_inputs(sc, seed + 5)
        recs_in = cyc.records_of(sc)
        cyc.seed_records(recs_in)
        hin, hout = cyc.host_inputs(sc, fe), cyc.host_outputs()
        cyc.step_from_host(hin, hout)
        cyc.check_errors()
        B = cyc.B
        outs[mode] = dict(fe_coeff=cyc.fetch("fe_coeff", (B, 3, 8, 4), np.float64), fe_n_int=cyc.fetch("fe_n_int", (B,), np.int32),
                          solved=hout["fe_solved"].copy(), status=hout["status"].copy(), coeff_out=hout["coeff_out"].copy(),
                          entangled=hout["entangled"].copy(), collide=hout["collide"].copy(), n_pieces=hout["n_pieces"].copy(),
                          rec=cyc.records("new").copy(), t_now=hin["t_now"].copy())
        cyc.close()
    full, kin = outs[1], outs[2]
    ok = kin["solved"] > 0
    assert ok.any()
    assert np.array_equal(kin["solved"], full["solved"]) and np.array_equal(kin["fe_n_int"][ok], full["fe_n_int"][ok])
    assert np.array_equal(kin["fe_coeff"][ok], full["fe_coeff"][ok])          # the same search
    assert (kin["status"] == 0).all()
    assert np.array_equal(kin["coeff_out"][ok], kin["fe_coeff"][ok])           # no back end: the path is the search's
    PW = capi.NB_REC_PWP_DOUBLES
    committed = 0
    for bi in range(len(ok)):
        me = int(sc.batch.agent_id[bi]) - 1
        prev = recs_in[me]
        if not ok[bi] or kin["entangled"][bi] or kin["collide"][bi]:
            assert np.array_equal(kin["rec"][me], prev)                          # rejected: the previous record stays
            continue
        n = int(kin["fe_n_int"][bi])
        now = np.zeros(capi.NB_REC_DOUBLES)
        now[0] = n
        now[1:2 + n] = sc.t_start[bi] + par.T_span * np.arange(n + 1)
        now[18:PW].reshape(3, 16, 4)[:, :n] = kin["coeff_out"][bi, :, :n]
        npc, want, _, _ = oracle.compose_records(kin["t_now"][bi], par.dc, prev, now)
        assert npc == kin["n_pieces"][bi]
        assert np.array_equal(kin["rec"][me, :PW], want[:PW])
        committed += 1
    assert committed > 0


@pytest.mark.gpu
def test_cpp_shim_kinodynamic_search(capi, oracle, tmp_path):
    """The C++ drop-in class KinodynamicSearch (neptune_b200/cpp/kinodynamic_search_b200.hpp) driven like
    neptune.cpp drives the reference's (:88-97, :1421-1453, :1509-1510): same status, pieces and entStateVec as
    the oracle, coefficients bit-exact."""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_shim_search")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "_build/test_shim_search"], check=True, capture_output=True)
    sc, sb = _batch(oracle, "mtlp5", 2002, per_agent_order=True)
    par = sb.par
    ref = _oracle_search(oracle, sb)
    N, S, np_ = par.num_of_agents, par.num_sample_per_interval, par.num_pol
    for b in range(sb.B):
        g = int(sb.group[b])
        lines = [f"{N} {int(sb.agent_id[b])} {np_} {S} {par.a_star_samp_x} {par.search_max_expansions} {par.search_max_nodes}",
                 " ".join(repr(float(v)) for v in (par.T_span, par.x_min, par.x_max, par.y_min, par.y_max, par.v_max, par.a_max, par.j_max,
                                                   par.a_star_fraction_voxel_size, par.a_star_bias, par.goal_radius, par.tetherLength))]
        lines += [f"{float(q[0])!r} {float(q[1])!r}" for q in par.pb]
        lines.append(" ".join(str(int(v)) for v in sb.comb[b]))
        lines.append(" ".join(repr(float(v)) for v in list(sb.init[b]) + list(sb.goal[b])))
        lines += [" ".join(repr(float(v)) for v in sb.coeffs_z[b, i]) for i in range(np_)]
        for j in range(N):
            kn = int(sb.known[b, j])
            lines.append(str(kn))
            if not kn:
                continue
            for i in range(np_):
                lines.append(" ".join(f"{float(sb.samp[g, j, i, s, 0])!r} {float(sb.samp[g, j, i, s, 1])!r}" for s in range(S + 1)))
            for i in range(np_):
                nv = int(sb.hull_cnt[g, j, i])
                lines.append(f"{nv} " + " ".join(f"{float(sb.hull_xy[g, j, i, v, 0])!r} {float(sb.hull_xy[g, j, i, v, 1])!r}" for v in range(nv)))
        na, nb_ = int(sb.es_cnt[b, 0]), int(sb.es_cnt[b, 1])
        lines.append(f"{na} {nb_}")
        lines += [f"{int(sb.es_alpha[b, q, 0])} {int(sb.es_alpha[b, q, 1])} {float(sb.es_beta[b, q])!r}" for q in range(na)]
        lines.append(" ".join(str(int(v)) for v in sb.es_bend[b, :nb_]))
        lines.append(" ".join(str(int(v)) for v in sb.es_active[b, :N]))
        path = tmp_path / f"search_{b}.txt"
        path.write_text("\n".join(lines) + "\n")
        out = subprocess.run([exe, str(path)], capture_output=True, text=True, check=True).stdout.split("\n")
        ok, status, n = (int(v) for v in out[0].split())
        assert ok == ref.solved[b] and status == ref.status[b] and n == ref.n_int[b]
        co = np.array([[float(v) for v in out[1 + k].split()] for k in range(3 * n)]).reshape(n, 3, 4)
        assert np.array_equal(co.transpose(1, 0, 2), ref.coeff[b, :, :n])
        for i in range(n + 1 if ok else 0):
            vals = [int(v) for v in out[1 + 3 * n + i].split()]
            assert vals[0] == ref.esv_cnt[b, i, 0]
            assert vals[1:] == ref.esv_alpha[b, i, :vals[0]].reshape(-1).tolist()


@pytest.mark.gpu
def test_gpu_search_rejects_bad_arguments(capi, oracle):
    """Argument errors are reported as NB_ERR_ARG before anything is launched (no exceptions cross the ABI)."""
    sc, sb = _batch(oracle, "mtlp5", 2002)
    s = _gpu_solver(capi, sc, sb)
    bad = dataclasses.replace(sb, comb=sb.comb.copy())
    bad.comb[0, 0] = bad.comb[0, 1]                      # not a permutation
    with pytest.raises(capi.NbError, match="permutation"):
        s.search(bad)
    bad = dataclasses.replace(sb, group=sb.group + 7)     # group index out of range
    with pytest.raises(capi.NbError, match="out of range"):
        s.search(bad)
    s2 = capi.Solver(sb.par)                              # search before nb_search_configure
    with pytest.raises(capi.NbError, match="configure"):
        s2.search(sb)
    s2.close()
    _assert_same(_oracle_search(oracle, sb), s.search(sb))   # the handle is still usable
    s.close()

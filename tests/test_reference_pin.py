"""The oracle pinned against the REFERENCE'S OWN CODE: neptune/src/kinodynamic_search.cpp, neptune/src/entangle_utils.cpp
and neptune/src/gjk.cpp are the reference sources on the path that need nothing but Eigen and a clock; `make -C oracle
_ref` compiles them unmodified, where they lie, against the Eigen stand-in of oracle/eigen_shim (+ oracle/ref_stubs).  Here (a container with /root/reference) the restatement in
oracle/neptune_oracle.c is compared with them on random inputs; on machines without /root/reference the same checks run
against golden vectors recorded from that library (tests/golden/reference/ref_chain.npz, tests/golden/make_ref_golden.py).

Bars: crossing lists, signature words, active cases, bend-point indices and GJK answers bit-exact (integers); betas
and tether lengths bit-exact as FP64 (same operations in the same order)."""
import ctypes as C
import os

import numpy as np
import pytest

from neptune_b200 import config
from neptune_b200.batch import ReplanResult
from neptune_b200.scenes import make_scene
from neptune_b200.search import static_longest_dist
from tests import ref_pin_util as ref
from tests.ent_walks import random_walks

needs_ref = pytest.mark.skipif(not ref.available(), reason="no /root/reference and no prebuilt oracle/_ref")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference", "ref_chain.npz")


@needs_ref
@pytest.mark.timeout(300)
def test_gjk_matches_reference(oracle):
    """gjk::collision (gjk.cpp:76-148).  The reference loops `while (1)`, so only generic positions are drawn (the
    oracle and the kernels bound the loop at 1000 iterations; no generic input comes near that)."""
    rng = np.random.default_rng(1)
    n_hit = 0
    for trial in range(3000):
        a = oracle.convex_hull(rng.normal(size=(rng.integers(3, 16), 2)) * rng.uniform(0.3, 2.0) + rng.normal(size=2) * 2)
        b = rng.normal(size=(4, 2)) * rng.uniform(0.2, 1.5) + rng.normal(size=2) * 2
        if trial % 7 == 0:   # a control polygon poking into / hovering next to a hull vertex
            b = b - b.mean(axis=0) + a[rng.integers(len(a))] + rng.normal(size=2) * 0.3
        r, o = ref.gjk(a, b), oracle.gjk_collision(a, b)
        assert r == o, trial
        n_hit += r
    assert 300 < n_hit < 2700


@needs_ref
def test_crossing_tests_match_reference(oracle):
    """entangleHSigToAddAgentInd (8- and 9-argument forms) and entangleHSigToAddStatic on random geometry."""
    rng = np.random.default_rng(2)
    L = oracle.lib()
    n_entries = n9 = 0
    for trial in range(6000):
        sc = rng.uniform(0.5, 6.0)
        pk, pik, pb = rng.normal(size=2) * sc, rng.normal(size=2) * sc, rng.normal(size=2) * sc
        pk1, pik1 = pk + rng.normal(size=2) * sc * 0.5, pik + rng.normal(size=2) * sc * 0.5
        nb = int(rng.integers(1, 4))
        bend = rng.normal(size=(nb, 2)) * sc
        out = np.zeros(128, np.int32)
        n = L.orc_hsig_agent(out.ctypes.data_as(C.c_void_p), 0, *[np.ascontiguousarray(x).ctypes.data_as(C.c_void_p) for x in (pk, pk1, pik, pik1, pb, bend)], nb, 7)
        got = ref.hsig_agent(pk, pk1, pik, pik1, pb, bend, 7)
        assert np.array_equal(out[:2 * n].reshape(-1, 2), got), trial
        n_entries += n
        # 9-argument form: a bend point added or released since the last check
        npv = nb + (1 if rng.random() < 0.5 else -1)
        if npv >= 1:
            prev = np.concatenate([bend[:min(nb, npv)], rng.normal(size=(max(0, npv - nb), 2)) * sc])[:npv]
            stop = C.c_int(0)
            out9 = np.zeros(128, np.int32)
            n2 = L.orc_hsig_agent9(out9.ctypes.data_as(C.c_void_p), 0, *[np.ascontiguousarray(x).ctypes.data_as(C.c_void_p) for x in (pk, pk1, pik, pik1, pb, bend)],
                                   nb, np.ascontiguousarray(prev).ctypes.data_as(C.c_void_p), npv, 7, C.byref(stop))
            if stop.value == 0:   # the reference calls exit(-1) on the others
                got9 = ref.hsig_agent(pk, pk1, pik, pik1, pb, bend, 7, prev=prev)
                assert np.array_equal(out9[:2 * n2].reshape(-1, 2), got9), trial
                n9 += 1
        M = int(rng.integers(1, 6))
        strep = rng.normal(size=(M, 2, 2)) * sc
        outs = np.zeros(64, np.int32)
        ns = L.orc_hsig_static(outs.ctypes.data_as(C.c_void_p), 0, pk.ctypes.data_as(C.c_void_p), pk1.ctypes.data_as(C.c_void_p),
                               np.ascontiguousarray(strep).ctypes.data_as(C.c_void_p), M, 5)
        assert np.array_equal(outs[:2 * ns].reshape(-1, 2), ref.hsig_static(pk, pk1, strep, 5)), trial
    assert n_entries > 500 and n9 > 1500


def _chain_cases(oracle):
    """(par, scene, agent, n, cxy, state0) for planner paths and for random walks that wrap around obstacles."""
    from tests.ent_backends import OracleEntBackend
    cases = []
    for cfg, seed in (("obst8", 3003), ("obst8", 3004), ("mtlp5", 2002), ("grid64", 4004)):
        par = config(cfg)
        sc = make_scene(par, seed, sync=False, ent_backend=OracleEntBackend(oracle))
        rng = np.random.default_rng(seed)
        walks = [sc.batch.coeff_init] + [random_walks(par, sc.batch.B, rng, s) for s in (3.0, 6.0, 6.0)]
        for w, co in enumerate(walks):
            for b in range(min(sc.batch.B, 8)):
                n = int(sc.batch.n_int[b]) if w == 0 else 8
                st0 = (sc.esA_cnt[b], sc.esA_alpha[b], sc.esA_beta[b], sc.esA_bend[b], sc.esA_active[b]) if w == 0 else \
                    (np.zeros(2, np.int32), np.zeros((par.ent_cap, 2), np.int32), np.zeros(par.ent_cap), np.zeros(par.ent_cap, np.int32),
                     np.zeros(par.NA, np.int32))
                cases.append((par, sc, b, n, np.ascontiguousarray(co[b, :2, :n, :]), st0))
    return cases


def _oracle_chain(oracle, par, sc, b, n, cxy, st0, longest):
    es = oracle.EntState(par.ent_cap, par.NA)
    es.n_alpha, es.n_bend = int(st0[0][0]), int(st0[0][1])
    es.alpha[:], es.beta[:], es.bend[:], es.active[:] = st0[1], st0[2], st0[3], st0[4]
    cx = oracle.EntCtx(par, int(sc.batch.agent_id[b]) - 1, sc.strep, sc.batch.bp_cnt, sc.batch.bp_xy)
    done, cnt, alpha, beta, bend, active = oracle.entangle_rollout(es, cx, n, cxy, sc.samp[b], sc.known[b])
    f = oracle.lib().orc_tether_length_state
    f.restype = C.c_double
    lens = np.zeros(max(n, 1))
    lg = np.ascontiguousarray(longest, np.float64) if par.num_of_static_obst else np.zeros((1, 2))
    for i in range(done):
        e = oracle.EntState(par.ent_cap, par.NA)
        e.n_alpha, e.n_bend = int(cnt[i + 1, 0]), int(cnt[i + 1, 1])
        e.alpha[:], e.beta[:], e.bend[:], e.active[:] = alpha[i + 1], beta[i + 1], bend[i + 1], active[i + 1]
        t = par.T_span
        end = np.array([cxy[0, i] @ [t ** 3, t ** 2, t, 1.0], cxy[1, i] @ [t ** 3, t ** 2, t, 1.0]])
        t3, t2 = t * t * t, t * t
        end = np.array([cxy[0, i, 0] * t3 + cxy[0, i, 1] * t2 + cxy[0, i, 2] * t + cxy[0, i, 3],
                        cxy[1, i, 0] * t3 + cxy[1, i, 1] * t2 + cxy[1, i, 2] * t + cxy[1, i, 3]])
        ec = e._c()
        lens[i] = f(C.byref(ec), C.byref(cx.c), lg.ctypes.data_as(C.c_void_p), end.ctypes.data_as(C.c_void_p))
    return done, cnt, alpha, beta, bend, active, lens


def _same_chain(a, b):
    done = a[0]
    assert done == b[0]
    assert np.array_equal(a[1][:done + 1], b[1][:done + 1])
    for i in range(done + 1):
        na, nb = a[1][i]
        assert np.array_equal(a[2][i, :na], b[2][i, :na]) and np.array_equal(a[3][i, :na], b[3][i, :na])
        assert np.array_equal(a[4][i, :nb], b[4][i, :nb]) and np.array_equal(a[5][i], b[5][i])
    assert np.array_equal(a[6][:done], b[6][:done])


@needs_ref
def test_chain_matches_reference(oracle):
    """addAlphaBetaToList, updateBendPts, getBendPt2d, calculateBetaForCase, breakcondition and getTetherLength through
    the loop of entanglesWithOtherAgents: the oracle's chain against the reference's functions, interval by interval."""
    mx_a = mx_b = n_cut = 0
    for par, sc, b, n, cxy, st0 in _chain_cases(oracle):
        M = par.num_of_static_obst
        longest = static_longest_dist(sc.static_raw, np.asarray(sc.strep).reshape(M, 2, 2)) if M else np.zeros((0, 2))
        r = ref.chain(par, int(sc.batch.agent_id[b]) - 1, sc.strep, longest, sc.batch.bp_cnt, sc.batch.bp_xy, sc.known[b], sc.samp[b],
                      n, cxy, *st0)
        o = _oracle_chain(oracle, par, sc, b, n, cxy, st0, longest)
        _same_chain(r, o)
        mx_a, mx_b, n_cut = max(mx_a, int(r[1][:, 0].max())), max(mx_b, int(r[1][:, 1].max())), n_cut + (r[0] < n)
    assert mx_a >= 6 and mx_b >= 1 and n_cut >= 5   # long words, bend points and entangling steps all occurred


def test_chain_matches_reference_golden(oracle):
    """The same comparison against vectors recorded from the reference library (travels without /root/reference)."""
    g = np.load(GOLDEN, allow_pickle=False)
    k = 0
    for par, sc, b, n, cxy, st0 in _chain_cases(oracle):
        if par.num_of_agents > 8:
            continue
        M = par.num_of_static_obst
        longest = static_longest_dist(sc.static_raw, np.asarray(sc.strep).reshape(M, 2, 2)) if M else np.zeros((0, 2))
        o = _oracle_chain(oracle, par, sc, b, n, cxy, st0, longest)
        assert np.array_equal(g[f"cxy_{k}"], cxy)           # the generator still produces the recorded inputs
        r = (int(g[f"done_{k}"]), g[f"cnt_{k}"], g[f"alpha_{k}"], g[f"beta_{k}"], g[f"bend_{k}"], g[f"active_{k}"], g[f"len_{k}"])
        _same_chain(r, o)
        k += 1
    assert k == int(g["n_cases"])


# ---------------------------------------------------------------------------- the front-end search, whole runs
GOLDEN_SEARCH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference", "ref_search.npz")

# (config, seed, parameter overrides, some tethers start wrapped around a contact point)
SEARCH_CASES = [
    ("single", 5, {}, False),
    ("mtlp5", 11, {}, False),            # one goal occupied: the open list runs empty after 64 796 nodes (status 2)
    ("mtlp5", 12, {}, False),
    ("mtlp5", 13, {}, True),
    ("mtlp5", 14, {"tetherLength": 9.0}, False),           # the tether constraint prunes most of the lattice
    ("mtlp5", 15, {"use_not_reaching_soln": False}, False),
    ("mtlp5", 16, {"a_star_samp_x": 3, "a_star_bias": 1.0}, False),
    ("obst8", 3, {}, False),             # nine static obstacles: crossing signatures and collisions of both kinds
    ("obst8", 6, {"x_min": -7.0, "x_max": 7.0, "y_min": -7.0, "y_max": 7.0}, True),   # a small world: the node pool runs out
]


def _search_case(oracle, cfg, seed, mods, multi_bend):
    """Inputs of one reference-sized run: node_num_max_ from the world size (kinodynamic_search.cpp:370-371), no pop
    budget (the reference's wall-clock limit is set far away), so a search ends by reaching the goal, by emptying the
    open list or by exhausting the node pool."""
    import math

    from neptune_b200.scenes import make_search_batch
    from tests.ent_backends import OracleEntBackend
    par = config(cfg)
    for k, v in mods.items():
        setattr(par, k, v)
    par.search_max_nodes = math.ceil((par.x_max - par.x_min) * (par.y_max - par.y_min) / (par.a_star_fraction_voxel_size ** 2) * 15)
    par.search_max_expansions = 2 ** 30
    sc = make_scene(par, seed, sync=False, ent_backend=OracleEntBackend(oracle), group_hulls=True)
    sb = make_search_batch(sc, seed + 1, per_agent_order=True)
    if multi_bend:
        sb.bp_cnt, sb.bp_xy = sb.bp_cnt.copy(), sb.bp_xy.copy()
        rng = np.random.default_rng(seed)
        for j in range(0, par.num_of_agents, 2):
            sb.bp_cnt[j] = 2
            sb.bp_xy[j, 1] = sb.bp_xy[j, 0] + rng.normal(0, 2.0, size=2)
    return par, sb


def _reference_search(sb):
    """{field: array} of the reference's run on every agent of the batch, entStateVec trimmed to the states it returned."""
    from neptune_b200.search import SearchResult
    r = SearchResult.empty(sb)
    comb, info = ref.search(sb, r)
    return dict(comb=comb, node_num_max=info[:, 0].copy(), nodes=info[:, 1].copy(), goal_occupied=info[:, 2].copy(),
                n_states=info[:, 3].copy() * r.solved, status=r.status, solved=r.solved, n_int=r.n_int, coeff=r.coeff, cost=r.cost,
                esv_cnt=r.esv_cnt, esv_alpha=r.esv_alpha, esv_beta=r.esv_beta, esv_bend=r.esv_bend, esv_active=r.esv_active)


def _check_oracle_search(oracle, par, sb, g, run=None):
    """Run the oracle (or `run`: the product / its host emulation) with the jerk order the reference drew and compare
    every output with the reference's, bit for bit."""
    from neptune_b200.search import SearchResult
    assert np.all(g["node_num_max"] == par.search_max_nodes)
    sb.comb = np.ascontiguousarray(g["comb"], np.uint8)
    if run is None:
        o = SearchResult.empty(sb)
        assert oracle.search_batch(sb, o, 4) == 0
    else:
        o = run(sb)
    assert np.array_equal(o.status, g["status"]) and np.array_equal(o.solved, g["solved"]) and np.array_equal(o.n_int, g["n_int"])
    assert np.array_equal(o.stats[:, 0], g["nodes"]), "node_used_num_"
    assert np.array_equal(o.stats[:, 3], g["goal_occupied"])
    assert np.array_equal(o.coeff, g["coeff"]), "pwp_out_ coefficients (FP64, bit-exact)"
    assert np.array_equal(o.cost, g["cost"])
    for b in range(sb.B):
        ns = int(g["n_states"][b])
        assert ns == (int(o.n_int[b]) + 1 if o.solved[b] else 0)
        assert np.array_equal(o.esv_cnt[b, :ns], g["esv_cnt"][b, :ns])
        for s in range(ns):
            na, nb = o.esv_cnt[b, s]
            assert np.array_equal(o.esv_alpha[b, s, :na], g["esv_alpha"][b, s, :na])
            assert np.array_equal(o.esv_beta[b, s, :na], g["esv_beta"][b, s, :na])
            assert np.array_equal(o.esv_bend[b, s, :nb], g["esv_bend"][b, s, :nb])
            assert np.array_equal(o.esv_active[b, s], g["esv_active"][b, s])
        for s in range(ns, o.esv_cnt.shape[1]):   # layout convention of the oracle / product: later slots repeat the last state
            if ns:
                assert np.array_equal(o.esv_cnt[b, s], o.esv_cnt[b, ns - 1]) and np.array_equal(o.esv_active[b, s], o.esv_active[b, ns - 1])
    return o


@needs_ref
@pytest.mark.timeout(600)
@pytest.mark.skipif(not os.path.isdir("/root/reference/neptune/src"), reason="live run needs the reference sources")
def test_search_matches_reference(oracle):
    """KinodynamicSearch::run of the reference itself (kinodynamic_search.cpp compiled unmodified into oracle/_ref) against
    the oracle: status, return value, pieces, coefficients, entStateVec with betas, nodes used, goal_occupied and cost,
    on searches of 50 to 65 000 nodes that end in all three ways."""
    seen, most = set(), 0
    for cfg, seed, mods, mb in SEARCH_CASES:
        par, sb = _search_case(oracle, cfg, seed, mods, mb)
        g = _reference_search(sb)
        o = _check_oracle_search(oracle, par, sb, g)
        seen |= set(o.status.tolist())
        most = max(most, int(o.stats[:, 0].max()))
    assert {1, 2} <= seen and most > 1000   # loose: the jerk order, hence the work, differs from run to run


_GOLDEN_FIELDS = ("comb", "node_num_max", "nodes", "goal_occupied", "n_states", "status", "solved", "n_int", "coeff", "cost", "esv_cnt",
                  "esv_alpha", "esv_beta", "esv_bend", "esv_active")


def test_search_matches_reference_golden(oracle):
    """The same comparison against outputs recorded from the reference library (tests/golden/make_ref_golden.py); the
    inputs are regenerated from the seeds.  Runs where /root/reference is absent."""
    g = np.load(GOLDEN_SEARCH)
    assert int(g["n_cases"]) == len(SEARCH_CASES)
    for k, (cfg, seed, mods, mb) in enumerate(SEARCH_CASES):
        par, sb = _search_case(oracle, cfg, seed, mods, mb)
        _check_oracle_search(oracle, par, sb, {f: g[f"{f}_{k}"] for f in _GOLDEN_FIELDS})


def test_emulated_kernel_matches_reference_golden(oracle):
    """The product's search kernel source, compiled for the host (tests/emul), against the REFERENCE's recorded runs --
    no oracle in between (searches of up to 66 000 nodes)."""
    from tests.emul import emul
    g = np.load(GOLDEN_SEARCH)
    n = 0
    for k, (cfg, seed, mods, mb) in enumerate(SEARCH_CASES):
        if int(g[f"nodes_{k}"].max()) > 100000:
            continue
        par, sb = _search_case(oracle, cfg, seed, mods, mb)
        _check_oracle_search(oracle, par, sb, {f: g[f"{f}_{k}"] for f in _GOLDEN_FIELDS}, run=emul.search)
        n += 1
    assert n >= 8


@pytest.mark.gpu
def test_gpu_search_matches_reference_golden(oracle):
    """k_search on the GPU against the REFERENCE's recorded runs, at the reference's own budgets (node pool of
    node_num_max_ nodes per agent, no pop limit): searches of up to 66 000 nodes, every output field bit for bit."""
    from neptune_b200 import capi
    g = np.load(GOLDEN_SEARCH)
    for k, (cfg, seed, mods, mb) in enumerate(SEARCH_CASES):
        par, sb = _search_case(oracle, cfg, seed, mods, mb)
        s = capi.Solver(par)
        if par.num_of_static_obst:
            s.set_static(sb.st_ptr, sb.st_xy, sb.strep)
            s.set_static_longest(sb.st_longest)
        s.search_configure()
        _check_oracle_search(oracle, par, sb, {f: g[f"{f}_{k}"] for f in _GOLDEN_FIELDS}, run=s.search)
        s.close()


@needs_ref
@pytest.mark.skipif(not os.path.isdir("/root/reference/neptune/src") or not os.environ.get("NB_LONG_TESTS"),
                    reason="minutes of CPU: set NB_LONG_TESTS=1 (needs the reference sources)")
def test_search_matches_reference_long(oracle):
    """Full-size obstacle world: searches of 80 000 to 337 499 nodes, one of which exhausts the node pool
    (node_used_num_ == node_num_max_ - 1, kinodynamic_search.cpp:1060-1064)."""
    par, sb = _search_case(oracle, "obst8", 6, {}, True)
    o = _check_oracle_search(oracle, par, sb, _reference_search(sb))
    assert int(o.stats[:, 0].max()) >= par.search_max_nodes - 1


@needs_ref
@pytest.mark.skipif(not os.path.isdir("/root/reference/neptune/src"), reason="live run needs the reference sources")
def test_post_check_matches_reference(oracle):
    """KinodynamicSearch::entangleCheckGivenPwp called on a real object of the reference (:897-985, including its return
    inside the interval loop: only interval 0 is looked at) against orc_entangle_check_pwp: answer and updated state."""
    n_ent = n_all = 0
    for par, sc, b, n, cxy, st0 in _chain_cases(oracle):
        if n == 0:
            continue
        es = oracle.EntState(par.ent_cap, par.NA)
        es.n_alpha, es.n_bend = int(st0[0][0]), int(st0[0][1])
        es.alpha[:], es.beta[:], es.bend[:], es.active[:] = st0[1], st0[2], st0[3], st0[4]
        cx = oracle.EntCtx(par, int(sc.batch.agent_id[b]) - 1, sc.strep, sc.batch.bp_cnt, sc.batch.bp_xy)
        o = oracle.entangle_check_pwp(es, cx, n, cxy, sc.samp[b], sc.known[b])
        r, cnt, alpha, beta, bend, active = ref.entangle_check_pwp(par, int(sc.batch.agent_id[b]) - 1, sc.strep, sc.batch.bp_cnt,
                                                                  sc.batch.bp_xy, sc.known[b], sc.samp[b], n, cxy, *st0)
        assert r == o, (par.num_of_agents, b)
        n_ent += r
        n_all += 1
        if not r:   # after a rejection the reference leaves the state half-updated; callers drop it (neptune.cpp:750-757)
            assert (es.n_alpha, es.n_bend) == (cnt[0], cnt[1])
            assert np.array_equal(es.alpha[:cnt[0]], alpha[:cnt[0]]) and np.array_equal(es.beta[:cnt[0]], beta[:cnt[0]])
            assert np.array_equal(es.bend[:cnt[1]], bend[:cnt[1]]) and np.array_equal(es.active, active)
    assert n_all > 80 and 0 < n_ent < n_all


@needs_ref
@pytest.mark.skipif(not os.path.isdir("/root/reference/neptune/src"), reason="live run needs the reference sources")
def test_generate_traj_matches_reference(oracle):
    """generatePwpOut: PolySolverGurobi's copy (solver_gurobi_poly.cpp:889-936) needs Gurobi, KinodynamicSearch's copy
    (kinodynamic_search.cpp:621-668) is the same text and compiles -- samples every dc, bit for bit, and the shifted knots."""
    rng = np.random.default_rng(9)
    for trial in range(60):
        n = int(rng.integers(1, 9))
        coeff = np.zeros((3, 8, 4))
        coeff[:, :n] = rng.normal(size=(3, n, 4)) * rng.uniform(0.1, 10.0)
        T, dc = float(rng.choice([0.5, 0.625, 1.0, 0.3])), float(rng.choice([0.01, 0.02, 0.013]))
        st, times = ref.generate_traj(coeff, n, T, dc, t_start=12.5)
        o = oracle.generate_traj(coeff, n, T, dc)
        assert len(o) == len(st) and np.array_equal(o, st), trial
        assert np.array_equal(times, np.arange(n + 1) * T + 12.5)


def _reference_exits(cases):
    """How many of the tracker ticks in `cases` (argument tuples of ref.track) end the process through exit(-1), as the
    reference does at its "stop" sites.  Runs in a fresh single-threaded interpreter that forks once per case
    (tests/ref_pin_util.py as a script), so nothing here forks a process that has OpenMP / torch threads."""
    import pickle
    import subprocess
    import sys
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as f:
        pickle.dump(cases, f)
    try:
        out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_pin_util.py"), f.name],
                             capture_output=True, text=True, check=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    finally:
        os.unlink(f.name)
    return int(out.stdout.strip().splitlines()[-1])


@needs_ref
def test_tracker_matches_reference(oracle):
    """The online tracker (NeptuneRos::updateEntStateStaticObs, neptune_ros.cpp:798-850): its loop lives in a ROS file, so it
    is replayed (oracle/ref_wrap.cpp::ref_track) around the reference's own 9-argument crossing test, static crossing test,
    addAlphaBetaToList and updateBendPts, over multi-tick walks in which tethers gain and lose contact points.  Where the
    oracle reports "stop k" the reference must end the process with exit(-1) -- checked in forked children of a helper process."""
    from tests.test_tracker import _init, _walk
    n_upd = n_gate = 0
    stops = []
    for cfg, seed in (("obst8", 21), ("mtlp5", 22), ("obst8", 23)):
        par = config(cfg)
        sc = make_scene(par, seed, sync=True)
        N = par.num_of_agents
        start, silent, frames = _walk(par, seed)
        st, pp, ppa = _init(par, start, silent)
        for fr in frames:
            for b in range(N):
                es = oracle.EntState(par.ent_cap, par.NA)
                es.n_alpha, es.n_bend = int(st.cnt[b, 0]), int(st.cnt[b, 1])
                es.alpha[:], es.beta[:], es.bend[:], es.active[:] = st.alpha[b], st.beta[b], st.bend[b], st.active[b]
                cx = oracle.EntCtx(par, b, sc.strep, fr["bp_cnt"], fr["bp_xy"])
                r_cnt, r_alpha, r_beta = st.cnt[b].copy(), st.alpha[b].copy(), st.beta[b].copy()
                r_bend, r_active, r_pp, r_ppa = st.bend[b].copy(), st.active[b].copy(), pp[b].copy(), ppa[b].copy()
                args = (par, b, sc.strep, fr["bp_cnt"], fr["bp_xy"], fr["bp_cnt_prev"], fr["bp_xy_prev"], r_cnt, r_alpha, r_beta, r_bend,
                        r_active, r_pp, r_ppa, fr["latest"], fr["cur"][b], float(fr["elapsed"][b]))
                o = oracle.track(es, cx, fr["bp_cnt_prev"], fr["bp_xy_prev"], pp[b], ppa[b], np.ascontiguousarray(fr["latest"]), fr["cur"][b],
                                 float(fr["elapsed"][b]))
                if o < 0:
                    assert o > -100, (cfg, b, o)
                    stops.append(args)
                else:
                    assert ref.track(*args) == o
                    assert (es.n_alpha, es.n_bend) == (r_cnt[0], r_cnt[1])
                    assert np.array_equal(es.alpha[:r_cnt[0]], r_alpha[:r_cnt[0]]) and np.array_equal(es.beta[:r_cnt[0]], r_beta[:r_cnt[0]])
                    assert np.array_equal(es.bend[:r_cnt[1]], r_bend[:r_cnt[1]]) and np.array_equal(es.active, r_active)
                    assert np.array_equal(pp[b], r_pp) and np.array_equal(ppa[b], r_ppa)
                    n_upd += o == 0
                    n_gate += o == 1
                st.cnt[b] = [es.n_alpha, es.n_bend]
                st.alpha[b], st.beta[b], st.bend[b], st.active[b] = es.alpha, es.beta, es.bend, es.active
    assert n_upd > 300 and n_gate > 5 and len(stops) > 10
    assert _reference_exits(stops) == len(stops)
    print("tracker ticks: updated", n_upd, "gated", n_gate, "reference exits", len(stops))


@needs_ref
def test_predict_matches_reference(oracle):
    """Neptune::PredictAlphasBetas (neptune.cpp:976-1008) replayed around the reference's own functions against orc_predict,
    from the states a tracker walk passes through (signature words of several entries, wrapped tethers)."""
    from tests.test_tracker import _init, _walk
    n = longest = 0
    for cfg, seed in (("obst8", 31), ("mtlp5", 32)):
        par = config(cfg)
        sc = make_scene(par, seed, sync=True)
        N = par.num_of_agents
        start, silent, frames = _walk(par, seed)
        st, pp, ppa = _init(par, start, silent)
        rng = np.random.default_rng(seed)
        for fr in frames:
            for b in range(N):
                cx = oracle.EntCtx(par, b, sc.strep, fr["bp_cnt"], fr["bp_xy"])
                es = oracle.EntState(par.ent_cap, par.NA)
                es.n_alpha, es.n_bend = int(st.cnt[b, 0]), int(st.cnt[b, 1])
                es.alpha[:], es.beta[:], es.bend[:], es.active[:] = st.alpha[b], st.beta[b], st.bend[b], st.active[b]
                # the prediction from the current tracker state to where the agent will be at the start of its plan
                ahead = fr["cur"][b] + rng.normal(0, 0.8, size=2)
                samp0 = fr["latest"] + rng.normal(0, 0.5, size=(N, 2))
                known = (~silent).astype(np.uint8)
                e2 = oracle.EntState(par.ent_cap, par.NA)
                e2.n_alpha, e2.n_bend = es.n_alpha, es.n_bend
                e2.alpha[:], e2.beta[:], e2.bend[:], e2.active[:] = es.alpha, es.beta, es.bend, es.active
                rc = oracle.predict(e2, cx, pp[b], ppa[b], ahead, samp0, known)
                r = [st.cnt[b].copy(), st.alpha[b].copy(), st.beta[b].copy(), st.bend[b].copy(), st.active[b].copy()]
                if rc == 0:
                    ref.predict(par, b, sc.strep, fr["bp_cnt"], fr["bp_xy"], *r, pp[b], ppa[b], ahead, samp0, known)
                    assert (e2.n_alpha, e2.n_bend) == (r[0][0], r[0][1])
                    assert np.array_equal(e2.alpha[:r[0][0]], r[1][:r[0][0]]) and np.array_equal(e2.beta[:r[0][0]], r[2][:r[0][0]])
                    assert np.array_equal(e2.bend[:r[0][1]], r[3][:r[0][1]]) and np.array_equal(e2.active, r[4])
                    n += 1
                    longest = max(longest, e2.n_alpha)
                # advance the tracker itself with the oracle (pinned by test_tracker_matches_reference)
                o = oracle.track(es, cx, fr["bp_cnt_prev"], fr["bp_xy_prev"], pp[b], ppa[b], np.ascontiguousarray(fr["latest"]), fr["cur"][b],
                                 float(fr["elapsed"][b]))
                st.cnt[b] = [es.n_alpha, es.n_bend]
                st.alpha[b], st.beta[b], st.bend[b], st.active[b] = es.alpha, es.beta, es.bend, es.active
    assert n > 400 and longest >= 2


@needs_ref
def test_compose_matches_reference(oracle):
    """mu::composePieceWisePol of the reference's own utils.cpp (:318-402) against orc_compose_records on every branch: the
    composed knots and coefficients, the empty "dummy" result, and the in-place adjustment of the arguments' first knots."""
    from tests import compose_util as cu
    rng = np.random.default_rng(6)
    seen = set()
    for it in range(1400):
        kind = cu.KINDS[it % len(cu.KINDS)]
        t, p1, p2 = cu.random_case(rng, kind)
        if int(p1[0]) < 1 or int(p2[0]) < 1:   # empty pwp: the reference reads .back() of an empty vector
            continue
        n, out, q1, q2 = oracle.compose_records(t, 0.05, p1, p2)
        rn, rt, rc, t1, t2 = ref.compose_records(t, 0.05, p1, p2)
        if n < 0:   # more than 16 pieces: the record cannot hold it, the reference has no bound
            assert rn > 16
            continue
        assert n == rn, (kind, it)
        assert np.array_equal(q1[1:2 + int(p1[0])], t1[:1 + int(p1[0])]) and np.array_equal(q2[1:2 + int(p2[0])], t2[:1 + int(p2[0])])
        if n:
            ot, oc = cu.rec_to(out)
            assert np.array_equal(np.array(ot), rt) and np.array_equal(oc, rc)
        seen.add((kind, n > 0))
    assert len(seen) >= len(cu.KINDS) and (("stale", False) in seen)


@needs_ref
def test_separator_matches_reference(oracle):
    """separator::Separator::solveModel of the reference's own separator_glpk.cpp (2-D variants :500-604, :248-373,
    :375-498) with HiGHS standing in for GLPK as the LP engine (oracle/ref_stubs/glpk.h records the model the reference
    builds and hands it to scipy): the solved flag equals the oracle's on every case, the model is the one the oracle
    documents (rows [x y 1], >= 1 for A and A+, <= -1 for B, free columns, zero objective on d), and the line the
    reference returns separates.  The known answer of the reference's own test (test_separator.cpp:23-32 -> "Solved= 1")
    is reproduced through its 3-D entry point."""
    rng = np.random.default_rng(12)
    n_sep = 0
    for t in range(600):
        A = rng.normal(size=(rng.integers(1, 10), 2)) * 1.5 + rng.normal(size=2) * 2
        B = rng.normal(size=(4, 2)) + rng.normal(size=2) * 2
        variant = t % 3
        Ap = (A + rng.normal(size=A.shape) * 0.3) if variant == 2 else None
        allA = np.vstack([A, Ap]) if Ap is not None else A
        ok, _ = oracle.separate(allA, B)
        del ref.LP_MODELS[:]
        r, n = ref.separator_solve(variant, A, B, Ap)
        assert r == ok, (t, variant)
        G, lo, up = ref.LP_MODELS[-1]
        assert np.array_equal(G, np.c_[np.vstack([allA, B]), np.ones(len(allA) + len(B))])
        assert np.array_equal(lo[:len(allA)], np.ones(len(allA))) and np.all(np.isinf(up[:len(allA)]))
        assert np.array_equal(up[len(allA):], -np.ones(len(B))) and np.all(np.isinf(lo[len(allA):]))
        if r and variant:
            assert (allA @ n[:2] + n[2] >= 1 - 1e-7).all() and (B @ n[:2] + n[2] <= -1 + 1e-7).all()
        n_sep += r
    assert 150 < n_sep < 550
    A3 = np.array([[-3.50, 21, 1.4], [-2.71, 2.13, 1.6], [0.53, 0.51, 1.4], [-3.50, 0.21, 0.3]])
    B3 = np.array([[-2.3, 4.69, 6.2], [3.7, 2.13, 65.6], [6.5, 2.93, 2.8], [0.3, 4.8, 9.2], [1.5, 6.7, 2.9]])
    ok, n3, d3 = ref.separator_solve3d(A3, B3)
    assert ok and oracle.lp_separable(A3, B3)
    assert (A3 @ n3 + d3 > 0).all() and (B3 @ n3 + d3 < 0).all()


def test_eigen_stand_in_semantics():
    """The Eigen stand-in the reference sources are compiled against (oracle/eigen_shim) behaves like Eigen where those
    sources depend on it: tests/cpp/eigen_shim_check.cpp against numpy."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "tests", "cpp"), "_build/eigen_shim_check"], check=True, capture_output=True)
    out = subprocess.run([os.path.join(root, "tests", "cpp", "_build", "eigen_shim_check")], check=True, capture_output=True, text=True).stdout
    got = {}
    for line in out.strip().splitlines():
        f = line.split()
        if f[0] == "norm":
            got["norm"], got["dot"] = float(f[1]), float(f[3])
        else:
            got[f[0]] = np.array([float(x) for x in f[1:]])
    A = np.array([[4, -2, 1, 0.5], [3, 6, -4, 2], [2, 1, 8, -1], [0.25, -3, 2, 5]])   # filled row by row
    V = np.array([[2, 0.5, -1], [1, 3, 0.25], [-2, 1, 4]])
    P = np.array([[1, 2, 3, 4], [-1, 0.5, 2, -3]])
    t = np.array([0.125, 0.25, 0.5, 1])
    assert np.array_equal(got["A"], A.ravel()) and np.array_equal(got["A_data"], A.T.ravel())   # column-major storage
    assert np.allclose(got["A_inv"], np.linalg.inv(A).ravel(), rtol=1e-13, atol=1e-15)
    assert np.allclose(got["V_inv"], np.linalg.inv(V).ravel(), rtol=1e-13, atol=1e-15)
    assert np.array_equal(got["P_A"], (P @ A).ravel()) and np.array_equal(got["P_blk_V"], (P[:, :3] @ V).ravel())
    assert np.array_equal(got["A_T"], A.T.ravel()) and np.array_equal(got["P_t"], P @ t)
    assert got["colT_t"][0] == A[:, 1] @ t
    e = np.array([1, 2, 3, 4, 5, 6.0])
    e[:2] = e[:2] + e[2:4] * 0.5 + e[4:] * 0.25
    e[4:] = e[4:] - e[:2]
    assert np.array_equal(got["e"], e)
    Q = np.zeros((2, 4))
    Q[0] = t
    Q[1] = 3 * Q[0]
    Q[:, 3] = Q[:, 0]
    assert np.array_equal(got["Q"], Q.ravel())
    assert np.array_equal(got["D_last"], [3, 6]) and np.array_equal(got["D_mean"], [2, 5]) and np.array_equal(got["g2"], [1, 2])
    assert got["norm"] == np.sqrt(18.0) and got["dot"] == 4.0 and got["abs"][0] == 0.75


# --------------------------------------------------------------------------------------------------------------------
# The trajectory QP: the reference's own solver_gurobi_poly.cpp, model recorded through a Gurobi stand-in
GOLDEN_QP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference", "ref_qp.npz")


def _match_rows(Rr, Ro):
    """Multiset comparison of constraint rows [a | rhs]: greedy nearest-row matching, returns the worst distance."""
    assert Rr.shape == Ro.shape, (Rr.shape, Ro.shape)
    used, worst = np.zeros(len(Ro), bool), 0.0
    for r in Rr:
        d = np.abs(Ro - r).max(axis=1)
        d[used] = np.inf
        k = int(np.argmin(d))
        worst, used[k] = max(worst, d[k]), True
    return worst / max(1.0, np.abs(Ro).max(initial=0.0))


def check_recorded_model(oracle, b, a, rec, fallback, lines, line_ok, par):
    """One GRBModel::optimize() of the reference against the oracle's restatement of the same model
    (orc_export_qp): objective, every equality row in order, the inequality rows as a multiset, the quadratic row."""
    n = int(b.n_int[a])
    nv = 12 * n
    mdl = oracle.export_qp(b, a, fallback, lines, line_ok)
    # reference order: axis, interval, coefficient (setInitTrajectory :200-224); oracle order: interval, axis, coefficient
    perm = np.array([ax * 4 * n + 4 * i + j for i in range(n) for ax in range(3) for j in range(4)])
    # the 3 n O plane variables setHulls creates (:258-278) take part in nothing in linear mode
    assert np.abs(rec["Q"][nv:]).max(initial=0) == 0 and np.abs(rec["Q"][:, nv:]).max(initial=0) == 0
    assert np.abs(rec["A"][:, nv:]).max(initial=0) == 0 and np.abs(rec["c"][nv:]).max(initial=0) == 0
    assert (rec["lb"] <= -1e100).all() and (rec["ub"] >= 1e100).all()             # free variables (:207-214)
    Q = rec["Q"][:nv, :nv][np.ix_(perm, perm)]
    scale = np.abs(mdl["P"]).max()
    assert np.abs(Q + Q.T - mdl["P"]).max() <= 1e-12 * scale
    assert np.abs(rec["c"][:nv][perm] - mdl["q"]).max() <= 1e-12 * max(1.0, np.abs(mdl["q"]).max())
    assert abs(rec["c0"] - mdl["c0"]) <= 1e-12 * max(1.0, abs(mdl["c0"]))
    A = rec["A"][:, :nv][:, perm]
    eq, le, ge = rec["sense"] == b"=", rec["sense"] == b"<", rec["sense"] == b">"
    Er, Eo = np.c_[A[eq], rec["rhs"][eq]], np.c_[mdl["Aeq"], mdl["beq"]]
    assert Er.shape == Eo.shape == (9 * n + (0 if fallback else 6), nv + 1)
    assert np.abs(Er - Eo).max() <= 1e-12 * max(1.0, np.abs(Eo).max())            # same rows in the same order
    Gr = np.vstack([np.c_[A[le], rec["rhs"][le]], np.c_[-A[ge], -rec["rhs"][ge]]])
    Go = np.c_[mdl["G"], mdl["h"]]
    assert Gr.shape[0] == 48 * n + 4 * int((line_ok[:n] == 1).sum())
    assert _match_rows(Gr, Go) <= 1e-11
    # quadratic terminal row (:681-702): ||p_end - final_pos||^2 - 0.01 <= 0 iff the start is within 1 m of it
    assert rec["nquad"] == int(mdl["has_qc"])
    if rec["nquad"]:
        T = par.T_span
        qp = np.array([T ** 3, T * T, T, 1.0])
        Qw, qw, cw = np.zeros((nv, nv)), np.zeros(nv), -0.01
        for ax in range(3):
            pf = float(qp @ b.coeff_init[a, ax, n - 1])
            sl = slice(12 * (n - 1) + 4 * ax, 12 * (n - 1) + 4 * ax + 4)
            Qw[sl, sl] += np.outer(qp, qp)
            qw[sl] += -2.0 * pf * qp
            cw += pf * pf
        Qr = rec["Qc"][0][:nv, :nv][np.ix_(perm, perm)]
        assert rec["qsense"][0] == b"<"
        assert np.abs(0.5 * (Qr + Qr.T) - Qw).max() <= 1e-12 * np.abs(Qw).max()
        assert np.abs(rec["qc"][0][:nv][perm] - qw).max() <= 1e-12 * max(1.0, np.abs(qw).max())
        assert abs(-rec["qrhs"][0] - cw) <= 1e-12 * max(1.0, abs(cw))
    assert abs(rec["time_limit"] - par.runtime_opt) < 1e-12                         # setMaxRuntime -> "TimeLimit" (:812)


def run_reference_qp(oracle, b, a, ref, replans=1):
    """The reference's solver on agent a: model checks on every optimize(), then (status, coeff, objective)."""
    from tests import ref_pin_util as rp
    rp.QP_MODELS.clear()
    ok, co, times, obj, ns = rp.qp_replan(b, a, replans=replans)
    per = len(rp.QP_MODELS) // replans
    assert per in (1, 2) and per * replans == len(rp.QP_MODELS)
    for k, rec in enumerate(rp.QP_MODELS):
        check_recorded_model(oracle, b, a, rec, (k % per) == 1, ref.lines[a], ref.line_ok[a], b.par)
    status = 0 if per == 1 else (1 if ok else 2)
    n = int(b.n_int[a])
    assert np.allclose(times, 1.25 + b.par.T_span * np.arange(n + 1)) and abs(ns - n * b.par.T_span / b.par.dc) <= 1.5
    return status, co, obj


QP_LIVE = ["single_1001", "mtlp5_2002", "obst8_3003", "mtlp5_crafted-ent0", "mtlp5_crafted-ent2", "mtlp5_crafted-box", "mtlp5_crafted-n2box"]


@needs_ref
@pytest.mark.parametrize("name", QP_LIVE)
def test_qp_model_matches_reference(oracle, name):
    """PolySolverGurobi of the reference's own solver_gurobi_poly.cpp (setInitTrajectory :187-244, addObjective :322-383,
    addConstraints :385-710, addEntangleConstraintForIJCase :715-784, optimize :804-887, generatePwpOut :889-936),
    compiled unmodified with a recording Gurobi stand-in and driven as Neptune drives it (neptune.cpp:102-107,
    :1514-1527): the model of every optimize() equals the oracle's restatement row for row (both solves, the quadratic
    row, the tether rows of the crafted scenes), the status path (first solve / fallback / pwp_out = pwp_init) and the
    optimiser's coefficients equal the oracle's."""
    from tests import ref_pin_util as rp
    from tests.golden_util import load
    rp.install_qp_hooks(oracle)
    par, b, z = load(os.path.join(os.path.dirname(GOLDEN_QP), "..", name + ".npz"))
    ref = ReplanResult.empty(b)
    assert oracle.replan_batch(b, ref, 2) == 0
    agents = range(b.B) if b.B <= 5 else (0, 1, 4, 7)
    for a in agents:
        status, co, obj = run_reference_qp(oracle, b, a, ref)
        assert status == ref.status[a], (name, a, status, ref.status[a])
        n = int(b.n_int[a])
        tol = (1e-6 if not z["has_qc"][a] else 1e-5) * max(1.0, np.abs(ref.coeff_out[a]).max())
        assert np.abs(co[:, :n] - ref.coeff_out[a, :, :n]).max() <= tol, (name, a)
        if status == 2:
            assert np.array_equal(co[:, :n], b.coeff_init[a, :, :n])        # pwp_out_ = pwp_init_ (:858)
        else:
            assert abs(obj - ref.obj[a]) <= 1e-6 * max(1.0, abs(ref.obj[a]))


@needs_ref
def test_qp_reference_object_is_reusable(oracle):
    """The solver object lives as long as the agent: a second replan on the same object builds the same model again --
    resetCompleteModel (solver_gurobi_utils.hpp) removes the previous replan's variables and rows and, by Gurobi's lazy
    update, not the ones setInitTrajectory / setHulls have just added."""
    from tests import ref_pin_util as rp
    from tests.golden_util import load
    rp.install_qp_hooks(oracle)
    par, b, z = load(os.path.join(os.path.dirname(GOLDEN_QP), "..", "mtlp5_2002.npz"))
    ref = ReplanResult.empty(b)
    assert oracle.replan_batch(b, ref, 2) == 0
    for a in (1, 4):
        status, co, obj = run_reference_qp(oracle, b, a, ref, replans=3)
        assert status == ref.status[a]
        n = int(b.n_int[a])
        assert np.abs(co[:, :n] - ref.coeff_out[a, :, :n]).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out[a]).max())


def test_qp_matches_reference_golden(oracle):
    """The same comparison against outputs recorded from the reference's solver (tests/golden/make_ref_golden.py),
    for machines without /root/reference: status path, coefficients, objective of every agent of every fixture."""
    from tests.golden_util import golden_files, load
    g = np.load(GOLDEN_QP)
    n_agents = 0
    for path in golden_files():
        name = os.path.basename(path)[:-4]
        par, b, z = load(path)
        ref = ReplanResult.empty(b)
        assert oracle.replan_batch(b, ref, 2) == 0
        st, co, ob = g[name + "/status"], g[name + "/coeff"], g[name + "/obj"]
        assert np.array_equal(st, ref.status), name
        for a in range(b.B):
            n = int(b.n_int[a])
            tol = (1e-6 if not z["has_qc"][a] else 1e-5) * max(1.0, np.abs(co[a]).max())
            assert np.abs(co[a, :, :n] - ref.coeff_out[a, :, :n]).max() <= tol, (name, a)
            if st[a] < 2:
                assert abs(ob[a] - ref.obj[a]) <= 1e-6 * max(1.0, abs(ob[a]))
            n_agents += 1
    assert n_agents >= 50


@pytest.mark.gpu
def test_gpu_qp_matches_reference_golden():
    """The PRODUCT against the recordings of the reference's own solver, no oracle in between."""
    from neptune_b200 import capi
    from tests.golden_util import golden_files, load
    g = np.load(GOLDEN_QP)
    for path in golden_files():
        name = os.path.basename(path)[:-4]
        par, b, z = load(path)
        s = capi.Solver(par)
        if par.num_of_static_obst:
            s.set_static(b.st_ptr, b.st_xy, z["strep"])
        res = s.replan(b)
        st, co, ob = g[name + "/status"], g[name + "/coeff"], g[name + "/obj"]
        assert np.array_equal(res.status, st), name
        for a in range(b.B):
            n = int(b.n_int[a])
            tol = (1e-6 if not z["has_qc"][a] else 1e-5) * max(1.0, np.abs(co[a]).max())
            assert np.abs(res.coeff_out[a, :, :n] - co[a, :, :n]).max() <= tol, (name, a)
            if st[a] < 2:
                assert abs(res.obj[a] - ob[a]) <= 1e-6 * max(1.0, abs(ob[a]))
        s.close()

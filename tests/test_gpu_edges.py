"""GPU: edge cases of the hot path through the C-ABI, and size-independent properties at the
BASELINE.json sizes (64 agents; 1024 agents / 200 static obstacles)."""
import ctypes as C
import dataclasses

import numpy as np
import pytest

from neptune_b200 import config
from neptune_b200.batch import NPOL, ReplanBatch, ReplanResult
from neptune_b200.minvo import solver_basis
from neptune_b200.scenes import make_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from neptune_b200 import capi
    capi.lib()
    return capi


def _check_solution_properties(par, b, res, tol_feas=1e-6):
    """Size-independent properties of an accepted solution: every line keeps hull and initial control
    points apart with margin 1; optimised control points satisfy every row; equalities hold."""
    Ainv, V, _ = solver_basis(par.T_span)
    T = par.T_span
    qp, qv, qa = np.array([T ** 3, T ** 2, T, 1]), np.array([3 * T * T, 2 * T, 1, 0]), np.array([6 * T, 2, 0, 0])
    worst = 0.0
    for a in range(b.B):
        n = int(b.n_int[a])
        ci, co = b.coeff_init[a], res.coeff_out[a]
        cp0 = np.stack([ci[0, :n] @ Ainv, ci[1, :n] @ Ainv], axis=-1)       # [n][4][2]
        ok = res.line_ok[a, :n] == 1
        ln = res.lines[a, :n]
        v0 = np.einsum("isc,ikc->isk", ln[..., :2], cp0) + ln[..., 2:3]
        assert (v0[ok] <= -1 + 1e-9).all()                                   # initial control points on the B side
        if res.status[a] == 2:
            assert np.array_equal(co, ci)
            continue
        keep_z = np.hypot(ci[0, 0, 3] - qp @ ci[0, n - 1], ci[1, 0, 3] - qp @ ci[1, n - 1]) < 1.0
        axes = (0, 1) if keep_z else (0, 1, 2)
        for ax in axes:
            assert np.abs(co[ax, 0, 1:] - ci[ax, 0, 1:]).max() <= 1e-8       # E1
            for i in range(n - 1):                                            # E2
                assert abs(qp @ co[ax, i] - co[ax, i + 1, 3]) <= 1e-8
                assert abs(qv @ co[ax, i] - co[ax, i + 1, 2]) <= 1e-8
                assert abs(qa @ co[ax, i] - 2 * co[ax, i + 1, 1]) <= 1e-8
            if res.status[a] == 0:                                            # E3
                assert abs(qv @ co[ax, n - 1]) <= 1e-8 and abs(qa @ co[ax, n - 1]) <= 1e-8
            cps = co[ax, :n] @ Ainv
            lo, hi = (par.x_min, par.y_min, par.z_min)[ax], (par.x_max, par.y_max, par.z_max)[ax]
            worst = max(worst, (cps - hi).max(), (lo - cps).max())
            vel = co[ax, :n, :3] @ V
            worst = max(worst, (np.abs(vel) - par.v_max).max(), (np.abs(co[ax, :n] @ qa) - par.a_max).max())
        cp = np.stack([co[0, :n] @ Ainv, co[1, :n] @ Ainv], axis=-1)
        v1 = np.einsum("isc,ikc->isk", ln[..., :2], cp) + ln[..., 2:3] - 1.0  # I4: <= 0
        if ok.any():
            worst = max(worst, v1[ok].max())
    assert worst <= tol_feas, worst


def test_grid64_full_size_properties_and_oracle_sample(capi, oracle):
    """BASELINE.json configs[3] at full size: properties on all 64 agents, oracle parity on all."""
    import bench
    par, scenes = bench.make_world(1, 0, 1)
    b = scenes[0].batch
    s = capi.Solver(par)
    res = s.replan(b)
    _check_solution_properties(par, b, res)
    ref = ReplanResult.empty(b)
    assert oracle.replan_batch(b, ref, 8) == 0
    assert np.array_equal(res.line_ok, ref.line_ok) and np.array_equal(res.status, ref.status)
    assert np.abs(res.coeff_out - ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
    s.close()


def test_grid1024_properties_on_a_shard(capi, oracle):
    """BASELINE.json configs[4] (1024 agents / 200 static): 24 planning agents against all 1023 others;
    properties on all, oracle parity on the first 3 (the CPU needs ~0.5 s per replan here)."""
    par = config("grid1024")
    sc = make_scene(par, 5005, agents=np.arange(24))
    b = sc.batch
    s = capi.Solver(par)
    s.set_static(b.st_ptr, b.st_xy, sc.strep)
    res = s.replan(b)
    _check_solution_properties(par, b, res)
    assert (res.line_ok == 1).sum() > 24 * 1000
    sub = dataclasses.replace(b, agent_id=b.agent_id[:3].copy(), n_int=b.n_int[:3].copy(), coeff_init=b.coeff_init[:3].copy(),
                              hull_ptr=b.hull_ptr[:3 * b.n_hull_slots * 8 + 1].copy(), nih0=b.nih0[:3].copy(),
                              esv_cnt=b.esv_cnt[:3].copy(), esv_alpha=b.esv_alpha[:3].copy(), esv_active=b.esv_active[:3].copy())
    ref = ReplanResult.empty(sub)
    assert oracle.replan_batch(sub, ref, 3) == 0
    assert np.array_equal(res.line_ok[:3], ref.line_ok) and np.array_equal(res.status[:3], ref.status)
    assert np.abs(res.coeff_out[:3] - ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
    s.close()


def test_every_path_length_and_no_obstacles(capi, oracle):
    """n = 1..8 with no other agent known (all hull slots empty): only bound rows and base rows."""
    par = config("mtlp5")
    for n in range(1, 9):
        sc = make_scene(par, 2100 + n, sync=False, n_fixed=n)
        b = sc.batch
        empty = dataclasses.replace(b, hull_ptr=np.zeros_like(b.hull_ptr), hull_xy=np.zeros((0, 2)),
                                    nih0=np.full_like(b.nih0, np.nan))
        s = capi.Solver(par)
        res = s.replan(empty)
        ref = ReplanResult.empty(empty)
        assert oracle.replan_batch(empty, ref, 1) == 0
        assert np.array_equal(res.line_ok, ref.line_ok) and np.array_equal(res.status, ref.status)
        assert np.abs(res.coeff_out - ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
        assert (b.n_int <= n).all()
        s.close()


def test_overlapping_hulls_are_skipped_like_the_reference(capi, oracle):
    """An obstacle hull that contains the agent's control points cannot be separated: the LP is
    'unsolved' (flag 2) and its rows are silently dropped (solver_gurobi_poly.cpp:491-494)."""
    par = config("mtlp5")
    sc = make_scene(par, 2002, sync=False)
    b = sc.batch
    xy = b.hull_xy.copy()
    k = (0 * b.n_hull_slots + 1) * 8 + 0       # agent 0, slot 1, interval 0
    p0, p1 = b.hull_ptr[k], b.hull_ptr[k + 1]
    c = b.coeff_init[0, :2, 0, 3]
    ang = np.linspace(0, 2 * np.pi, p1 - p0, endpoint=False)
    xy[p0:p1] = c + 3.0 * np.stack([np.cos(ang), np.sin(ang)], axis=1)   # a big polygon around the start point
    bad = dataclasses.replace(b, hull_xy=xy)
    s = capi.Solver(par)
    res = s.replan(bad)
    ref = ReplanResult.empty(bad)
    assert oracle.replan_batch(bad, ref, 1) == 0
    assert res.line_ok[0, 0, 1] == 2 and np.array_equal(res.line_ok, ref.line_ok)
    assert np.array_equal(res.status, ref.status)
    assert np.abs(res.coeff_out - ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
    s.close()


def test_many_kept_lines_use_the_global_row_path(capi, oracle):
    """More than 160 non-redundant lines for one agent (obstacles on a ring around every interval):
    the QP kernel falls back from shared memory to global scratch; same answer."""
    par = config("grid64")
    sc = make_scene(par, 4004, agents=np.arange(2))
    b = sc.batch
    N = par.num_of_agents
    Ainv, _, _ = solver_basis(par.T_span)
    cnt = np.zeros((b.B, N, NPOL), np.int64)
    chunks = []
    for a in range(b.B):
        n = int(b.n_int[a])
        for j in range(N):
            for i in range(NPOL):
                if i >= n or j == a or j > 40:
                    continue
                cp = np.stack([b.coeff_init[a, 0, i] @ Ainv, b.coeff_init[a, 1, i] @ Ainv], axis=1)
                ctr, rad = cp.mean(axis=0), np.abs(cp - cp.mean(axis=0)).max() * 1.5 + 0.5
                th = 2 * np.pi * j / 40.0
                q = ctr + (rad + 0.3) * np.array([np.cos(th), np.sin(th)])
                tri = q + 0.1 * np.array([[1, 0], [-0.5, 0.8], [-0.5, -0.8]])
                chunks.append(tri)
                cnt[a, j, i] = 3
    ptr = np.zeros(b.B * N * NPOL + 1, np.int64)
    np.cumsum(cnt.reshape(-1), out=ptr[1:])
    ring = dataclasses.replace(b, hull_ptr=ptr, hull_xy=np.ascontiguousarray(np.concatenate(chunks)))
    s = capi.Solver(par)
    res = s.replan(ring)
    ref = ReplanResult.empty(ring)
    assert oracle.replan_batch(ring, ref, 2) == 0
    assert ((res.line_ok == 1).sum(axis=(1, 2)) > 160).any()
    assert np.array_equal(res.line_ok, ref.line_ok) and np.array_equal(res.status, ref.status)
    assert np.abs(res.coeff_out - ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
    s.close()


def test_ent_slot_overflow_is_loud(capi):
    """More non-entangling LPs than ent_slots must fail with NB_ERR_CAPACITY, never drop rows silently."""
    par = dataclasses.replace(config("mtlp5"), ent_slots=1)
    par.pb = config("mtlp5").pb
    sc = make_scene(par, 2002, sync=False)
    b = sc.batch
    act, cntv, alpha = b.esv_active.copy(), b.esv_cnt.copy(), b.esv_alpha.copy()
    bp_cnt, bp_xy = b.bp_cnt.copy(), b.bp_xy.copy()
    others = [j for j in range(par.num_of_agents) if j != int(b.agent_id[0]) - 1][:3]
    for q, j in enumerate(others):       # three agents with one active case each, three bend points each
        act[0, :, j] = 1
        alpha[0, :, q] = [j + 1, 5]
        bp_cnt[j] = 3
        bp_xy[j, 1] = b.coeff_init[0, :2, 0, 3] + [0.5, 0.2 * q]
        bp_xy[j, 2] = b.coeff_init[0, :2, 0, 3] + [-0.5, 0.3 * q]
    cntv[0, :, 0] = len(others)
    nih0 = b.nih0.copy()
    nih0[0, others] = b.coeff_init[0, :2, 0, 3] + 1.0
    bad = dataclasses.replace(b, esv_active=act, esv_cnt=cntv, esv_alpha=alpha, bp_cnt=bp_cnt, bp_xy=bp_xy, nih0=nih0)
    s = capi.Solver(par)
    with pytest.raises(capi.NbError, match="ent_slots"):
        s.replan(bad)
    s.close()


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_non_entangling_lines_match_oracle(capi, oracle, variant):
    """addEntangleConstraintForIJCase (solver_gurobi_poly.cpp:620-642, :715-784) on the GPU: the crafted state of
    tests/crafted.py fires >= 10 tether LPs (ray beyond the agent, bend-point segments, the k == case_id skip, the
    distance gate, case id 0); flags bit-exact, lines <= 1e-9, coefficients <= 1e-6 against the oracle, whose LP
    set is itself checked against an independent reading of the reference's loops and against HiGHS
    (tests/test_crafted_branches.py)."""
    from tests import crafted
    par, b = crafted.ent_lp_batch(variant)
    s = capi.Solver(par)
    res = s.replan(b)
    ref = ReplanResult.empty(b)
    assert oracle.replan_batch(b, ref, 1) == 0
    e0 = b.n_hull_slots + par.num_of_agents + par.num_of_static_obst
    if variant < 2:
        assert (ref.line_ok[:, :, e0:] == 1).sum() >= 8 and (ref.line_ok[:, :, e0:] == 2).sum() >= 1
    else:   # the binding tether: the rows move the optimum (checked against the oracle below)
        plain = ReplanResult.empty(b)
        assert oracle.replan_batch(crafted.without_tether_rows(b), plain, 1) == 0
        assert np.abs(ref.coeff_out[0] - plain.coeff_out[0]).max() > 0.1
    assert np.array_equal(res.line_ok, ref.line_ok) and np.array_equal(res.status, ref.status)
    m = ref.line_ok == 1
    assert np.abs(res.lines[m] - ref.lines[m]).max() <= 1e-9 * max(1.0, np.abs(ref.lines[m]).max())
    assert np.abs(res.coeff_out - ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
    assert np.abs(res.obj - ref.obj).max() <= 1e-8 * max(1.0, np.abs(ref.obj).max())
    _check_solution_properties(par, b, res)
    s.close()


@pytest.mark.parametrize("kind", ["box", "vel", "n2box"])
def test_both_solves_infeasible_is_status_2(capi, oracle, kind):
    """optimize()'s failure path (solver_gurobi_poly.cpp:856-859): start state outside the box / above v_max (and the
    n == 2 single-point case) => status 2 on GPU == oracle, coeff_out == coeff_init bit for bit, objective 0.
    HiGHS calls the same models infeasible in tests/test_crafted_branches.py."""
    from tests import crafted
    par, b = crafted.infeasible_batch(kind)
    s = capi.Solver(par)
    res = s.replan(b)
    ref = ReplanResult.empty(b)
    assert oracle.replan_batch(b, ref, 1) == 0
    assert (ref.status == 2).all() and np.array_equal(res.status, ref.status)
    assert np.array_equal(res.line_ok, ref.line_ok)
    assert np.array_equal(res.coeff_out, b.coeff_init) and (res.obj == 0).all()
    s.close()


def test_golden_fixtures_through_the_abi(capi):
    """Every committed fixture (tests/golden/*.npz, crafted branches included) through nb_replan_batch against the
    outputs recorded from the oracle -- no oracle call at run time."""
    from tests.golden_util import golden_files, load
    files = golden_files()
    assert len(files) >= 10
    for path in files:
        par, b, z = load(path)
        s = capi.Solver(par)
        if par.num_of_static_obst:
            s.set_static(b.st_ptr, b.st_xy, z["strep"])
        res = s.replan(b)
        assert np.array_equal(res.line_ok, z["orc_line_ok"]), path
        m = z["orc_line_ok"] == 1
        assert np.abs(res.lines[m] - z["orc_lines"][m]).max(initial=0) <= 1e-9 * max(1.0, np.abs(z["orc_lines"][m]).max(initial=0))
        assert np.array_equal(res.status, z["orc_status"]), path
        assert np.abs(res.coeff_out - z["orc_coeff"]).max() <= 1e-6 * max(1.0, np.abs(z["orc_coeff"]).max())
        assert np.abs(res.obj - z["orc_obj"]).max() <= 1e-8 * max(1.0, np.abs(z["orc_obj"]).max())
        s.close()


def test_invalid_n_int_and_agent_id_are_rejected(capi):
    """nb_replan_batch validates n_int (1..num_pol) and agent_id (1..N) for host arguments."""
    par = config("mtlp5")
    sc = make_scene(par, 2002, sync=False)
    s = capi.Solver(par)
    for field, bad in (("n_int", 0), ("n_int", 9), ("agent_id", 0), ("agent_id", par.num_of_agents + 1)):
        arr = getattr(sc.batch, field).copy()
        arr[1] = bad
        with pytest.raises(capi.NbError, match=field):
            s.replan(dataclasses.replace(sc.batch, **{field: arr}))
    s.replan(sc.batch)   # the handle stays usable
    s.close()


def test_separate_degenerate_sets(capi, oracle):
    par = config("mtlp5")
    s = capi.Solver(par)
    B = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    cases = [np.array([[3.0, 0.5]]),                                   # a single point
             np.array([[3.0, 0.5], [3.0, 0.5]]),                       # duplicated point
             np.array([[3.0, -1.0], [3.0, 0.0], [3.0, 2.0]]),          # collinear
             np.array([[0.5, 0.5]]),                                   # inside B: inseparable
             np.array([[1.0, 0.5], [2.0, 0.5]]),                       # touching B: inseparable (margin 0)
             np.array([[1.0 + 1e-4, 0.5], [2.0, 0.5]]),                # 0.1 mm gap: separable, |n| = 2e4
             np.array([[1.0 + 1e-7, 0.5], [2.0, 0.5]])]                # 0.1 um gap: the +-1 margins cannot be verified
                                                                       # to 1e-9 with |n| = 2e7 -> reported unsolved
    a_ptr = np.concatenate([[0], np.cumsum([len(a) for a in cases])])
    b_ptr = np.arange(len(cases) + 1) * 4
    for poly in (False, True):
        ok, line = s.separate(a_ptr, np.concatenate(cases), b_ptr, np.tile(B, (len(cases), 1)), poly)
        for i, A in enumerate(cases):
            ok_o, l_o = oracle.separate(A, B)
            assert ok[i] == ok_o, (i, poly)
            if ok_o:
                assert np.abs(line[i] - l_o).max() <= 1e-9 * max(1.0, np.abs(l_o).max())
    assert list(ok) == [True, True, True, False, False, True, False]
    s.close()


def test_empty_batch_is_a_noop(capi):
    par = config("mtlp5")
    s = capi.Solver(par)
    a = capi.NbReplanArgs()
    a.B, a.space = 0, 0
    s.replan_args(a)
    s.close()


@pytest.mark.gpu
def test_compose_records_matches_oracle(capi, oracle):
    """nb_compose_records_batch (mu::composePieceWisePol, utils.cpp:318-402): bit-exact against the oracle on
    every branch of the reference function; more than 16 pieces is a loud capacity error."""
    from tests import compose_util as cu
    par = config("mtlp5")
    sv = capi.Solver(par)
    rng = np.random.default_rng(17)
    ts, prevs, nows, outs, nps = [], [], [], [], []
    for it in range(2100):
        t, p1, p2 = cu.random_case(rng, cu.KINDS[it % len(cu.KINDS)])
        n, out, _, _ = oracle.compose_records(t, par.dc, p1, p2)
        ts.append(t), prevs.append(p1), nows.append(p2), outs.append(out), nps.append(n)
    has_prev = np.ones(len(ts), np.uint8)
    has_prev[::13] = 0
    npc, go = sv.compose(ts, has_prev, np.stack(prevs), np.stack(nows))
    for b in range(len(ts)):
        if has_prev[b]:
            assert npc[b] == nps[b] and np.array_equal(go[b], outs[b]), b
        else:
            assert npc[b] == int(nows[b][0]) and np.array_equal(go[b], nows[b])
    p1 = cu.rec_from(0.1 * np.arange(17), np.ones((3, 16, 4)))
    p2 = cu.rec_from(1.55 + 0.5 * np.arange(9), np.ones((3, 8, 4)))
    with pytest.raises(RuntimeError, match="16 pieces"):
        sv.compose([0.05], [1], p1[None], p2[None])
    sv.compose(ts[:4], has_prev[:4], np.stack(prevs[:4]), np.stack(nows[:4]))   # the handle stays usable

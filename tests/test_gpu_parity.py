"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs.  Bars: LP solved flags and QP status paths bit-exact; separating lines
within 1e-9 relative; QP coefficients within 1e-6 * max(1, |x|_inf) at equal objective (1e-8 rel)."""
import numpy as np
import pytest

from neptune_b200 import config
from neptune_b200.batch import ReplanResult
from neptune_b200.scenes import make_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from neptune_b200 import capi
    capi.lib()
    return capi


def _solver(capi, par, sc=None):
    s = capi.Solver(par)
    if par.num_of_static_obst:
        s.set_static(sc.batch.st_ptr, sc.batch.st_xy, sc.strep)
    return s


def test_separate_batch_matches_oracle(capi, oracle):
    rng = np.random.default_rng(7)
    par = config("mtlp5")
    s = capi.Solver(par)
    A, B = [], []
    for _ in range(4000):
        A.append(oracle.convex_hull(rng.normal(size=(rng.integers(1, 14), 2)) * 1.5 + rng.normal(size=2) * 2))
        B.append(rng.normal(size=(4, 2)) + rng.normal(size=2) * 2)
    a_ptr = np.concatenate([[0], np.cumsum([len(a) for a in A])])
    b_ptr = np.concatenate([[0], np.cumsum([len(b) for b in B])])
    for poly in (True, False):
        ok, line = s.separate(a_ptr, np.concatenate(A), b_ptr, np.concatenate(B), poly)
        for i in range(len(A)):
            ok_o, l_o = oracle.separate(A[i], B[i])
            assert ok[i] == ok_o
            if ok_o:
                assert np.abs(line[i] - l_o).max() <= 1e-9 * max(1.0, np.abs(l_o).max())
                assert (A[i] @ line[i, :2] + line[i, 2] >= 1 - 1e-9).all()
                assert (B[i] @ line[i, :2] + line[i, 2] <= -1 + 1e-9).all()
    s.close()


@pytest.mark.parametrize("cfg,seeds,kw", [
    ("single", range(1001, 1004), dict(n_fixed=3)),
    ("mtlp5", range(2002, 2008), dict(sync=False)),
    ("obst8", range(3003, 3007), dict(sync=False)),
])
def test_replan_matches_oracle(capi, oracle, cfg, seeds, kw):
    par = config(cfg)
    for seed in seeds:
        sc = make_scene(par, seed, **kw)
        s = _solver(capi, par, sc)
        got = s.replan(sc.batch)
        ref = ReplanResult.empty(sc.batch)
        assert oracle.replan_batch(sc.batch, ref, 1) == 0
        assert (got.line_ok == ref.line_ok).all()
        m = ref.line_ok == 1
        assert np.abs(got.lines[m] - ref.lines[m]).max(initial=0) <= 1e-9 * max(1.0, np.abs(ref.lines[m]).max(initial=0))
        assert (got.status == ref.status).all(), (got.status, ref.status)
        tol = 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
        assert np.abs(got.coeff_out - ref.coeff_out).max() <= tol
        assert np.abs(got.obj - ref.obj).max() <= 1e-8 * max(1.0, np.abs(ref.obj).max())
        s.close()


def test_generate_traj_matches_oracle(capi, oracle):
    par = config("mtlp5")
    sc = make_scene(par, 2002, sync=False)
    s = capi.Solver(par)
    states, ns = s.generate_traj(sc.batch.n_int, sc.batch.coeff_init, par.dc)
    for b in range(sc.batch.B):
        ref = oracle.generate_traj(sc.batch.coeff_init[b], int(sc.batch.n_int[b]), par.T_span, par.dc)
        assert ns[b] == len(ref)
        # every operation is rounded on its own in the reference's order: bit-exact
        assert np.array_equal(states[b, :ns[b]], ref)
    s.close()


def test_missing_static_is_an_error(capi):
    par = config("obst8")
    sc = make_scene(par, 3003, sync=False)
    s = capi.Solver(par)
    with pytest.raises(capi.NbError):
        s.replan(sc.batch)
    s.close()


@pytest.mark.parametrize("cfg,seeds", [("mtlp5", (2002, 2005)), ("obst8", (3003, 3004, 3006))])
def test_entangle_chain_bit_exact(capi, oracle, cfg, seeds):
    """K3 vs oracle: alphas, active_cases, bendPointsIdx bit-exact (betas too: same IEEE operations)
    for the history walk, PredictAlphasBetas, the front-end chain and entangleCheckGivenPwp."""
    from tests.ent_backends import OracleEntBackend
    par = config(cfg)
    for seed in seeds:
        ref = make_scene(par, seed, sync=False, ent_backend=OracleEntBackend(oracle))
        s = capi.Solver(par)
        if par.num_of_static_obst:
            s.set_static(ref.batch.st_ptr, ref.batch.st_xy, ref.strep)
        dev = capi.DeviceEntBackend(s)
        got = make_scene(par, seed, sync=False, ent_backend=dev)
        for k in ("es0_cnt", "es0_alpha", "es0_beta", "es0_bend", "es0_active", "esA_cnt", "esA_alpha", "esA_beta",
                  "esA_bend", "esA_active"):
            assert np.array_equal(getattr(got, k), getattr(ref, k)), k
        for k in ("esv_cnt", "esv_alpha", "esv_active"):
            assert np.array_equal(getattr(got.batch, k), getattr(ref.batch, k)), k
        # post-check on the optimised trajectory (a17), both sides from the same state
        res = s.replan(ref.batch)
        b = ref.batch
        args = (par, b.agent_id, b.n_int, res.coeff_out, ref.samp, ref.known, ref.strep, b.bp_cnt, b.bp_xy,
                ref.esA_cnt, ref.esA_alpha, ref.esA_beta, ref.esA_bend, ref.esA_active)
        g = dev.check_batch(*args)
        r = OracleEntBackend(oracle).check_batch(*args)
        for x, y in zip(g, r):
            assert np.array_equal(x, y)
        s.close()


@pytest.mark.parametrize("cfg,seed", [("mtlp5", 2002), ("obst8", 3003), ("mtlp5", 2004)])
def test_hulls_samples_postcheck_bit_exact(capi, oracle, cfg, seed):
    """K1 / K5 vs oracle: hull vertices and order, vertex counts, interval indices, nih0 and samples
    bit-exact (same operands, same operation order, no FMA contraction); collision flags equal."""
    par = config(cfg)
    sc = make_scene(par, seed, sync=False)
    s = _solver(capi, par, sc)
    recs = capi.make_records(sc.committed)
    delta = 2 * par.drone_radius
    out = s.hulls(sc.t_start, recs, sc.known, delta)
    b = sc.batch
    for bi in range(b.B):
        for j in range(par.num_of_agents):
            if not sc.known[bi, j]:
                assert out["hull_cnt"][bi, j].sum() == 0 and np.isnan(out["nih0"][bi, j]).all()
                continue
            tm, cx, cy, _ = sc.committed[j]
            for i in range(par.num_pol):
                t0 = sc.t_start[bi] + i * par.T_span
                h, h2, idx = oracle.hull_of_interval(tm, cx, cy, t0, sc.t_start[bi] + (i + 1) * par.T_span, par.T_span, [delta] * 3)
                c = out["hull_cnt"][bi, j, i]
                assert c == len(h) and np.array_equal(out["hull_xy"][bi, j, i, :c], h)
                assert np.array_equal(out["nih0"][bi, j, i], h2[0])
                assert np.array_equal(out["idx"][bi, j, i], idx)
            smp, _ = oracle.sample_points(tm, cx, cy, sc.t_start[bi], sc.t_start[bi] + par.T_span * par.num_pol,
                                          par.num_pol, par.num_sample_per_interval)
            assert np.array_equal(out["samp"][bi, j], smp)
    # device-generated hulls straight into the back end == oracle on the packed scene hulls
    import ctypes as C
    res = ReplanResult.empty(b)
    a = capi.host_args(b, res)
    hx, hc, hp, n0 = (np.ascontiguousarray(out[k]) for k in ("hull_xy", "hull_cnt", "hull_ptr", "nih0"))
    a.hull_xy, a.hull_cnt, a.hull_ptr, a.nih0 = (x.ctypes.data_as(C.c_void_p) for x in (hx, hc, hp, n0))
    a.hull_nvert = hx.shape[0] * hx.shape[1] * hx.shape[2] * hx.shape[3]
    s.replan_args(a)
    ref = ReplanResult.empty(b)
    assert oracle.replan_batch(b, ref, 1) == 0
    assert (res.line_ok == ref.line_ok).all() and (res.status == ref.status).all()
    assert np.abs(res.coeff_out - ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
    # post-check of the optimised trajectories against everybody's committed trajectory
    col = s.postcheck(b.n_int, res.coeff_out, sc.t_start, recs, sc.known, delta)
    want = np.zeros(b.B, np.int32)
    for bi in range(b.B):
        for j in range(par.num_of_agents):
            if sc.known[bi, j]:
                tm, cx, cy, _ = sc.committed[j]
                if oracle.pwp_collides(res.coeff_out[bi], int(b.n_int[bi]), sc.t_start[bi], par.T_span, tm, cx, cy, [delta] * 3):
                    want[bi] = 1
    assert np.array_equal(col, want)
    s.close()


def _cycle_fetch_states(cyc, par):
    B, cap, NA = cyc.B, par.ent_cap, par.NA
    return dict(cnt=cyc.fetch("esA_cnt", (B, 2), np.int32), alpha=cyc.fetch("esA_alpha", (B, cap, 2), np.int32),
                beta=cyc.fetch("esA_beta", (B, cap), np.float64), bend=cyc.fetch("esA_bend", (B, cap), np.int32),
                active=cyc.fetch("esA_active", (B, NA), np.int32))


@pytest.mark.parametrize("cfg,seed,all_late", [("mtlp5", 2005, True), ("obst8", 3004, True), ("obst8", 3003, False)])
def test_replan_cycle_matches_oracle_composite(capi, oracle, cfg, seed, all_late):
    """The whole device-resident cycle of the C++ library (nb_cycle_*: K1 -> K3 predict -> K2/K4 -> K5 + gated K3
    post-check -> commit with DynTraj header into the record ring) against the same chain executed by the oracle.  The
    third case has late != known: a random subset of the trajectories arrives during the optimisation, changed."""
    from neptune_b200.cycle import ReplanCycle
    from tests import postcheck_util as pu
    from tests.ent_backends import OracleEntBackend
    par = config(cfg)
    ob = OracleEntBackend(oracle)
    sc = make_scene(par, seed, sync=False, ent_backend=ob)
    b = sc.batch
    cyc = ReplanCycle(par, b.agent_id - 1, "cuda:0", static=(b.st_ptr, b.st_xy, sc.strep), planned=np.ones(par.num_of_agents, np.uint8))
    known_recs = cyc.records_of(sc, seq=0)
    if all_late:
        late, late_committed, late_recs, bc_l, bx_l = sc.known, sc.committed, known_recs, b.bp_cnt, b.bp_xy
    else:
        late, late_committed, late_recs, bc_l, bx_l = next(pu.late_cases(par, sc, np.random.default_rng(seed), 1))
    cyc.seed_records(known_recs, late_recs)
    hin, hout = cyc.host_inputs(sc, late=late), cyc.host_outputs()
    cyc.step_from_host(hin, hout)
    cyc.check_errors()
    esA = _cycle_fetch_states(cyc, par)
    for k in ("cnt", "alpha", "beta", "bend", "active"):
        assert np.array_equal(esA[k], getattr(sc, "esA_" + k)), k
    # trajCB bookkeeping from the record headers
    assert np.array_equal(cyc.fetch("bp_cnt", (par.num_of_agents,), np.int32), b.bp_cnt)
    ref = ReplanResult.empty(b)
    assert oracle.replan_batch(b, ref, 2) == 0
    assert np.array_equal(hout["status"], ref.status)
    assert np.abs(hout["coeff_out"] - ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
    delta = 2 * par.drone_radius
    want_col = np.zeros(b.B, np.int32)
    for bi in range(b.B):
        for j in range(par.num_of_agents):
            if late[bi, j] and j != int(b.agent_id[bi]) - 1:
                tm, cx, cy, _ = late_committed[j]
                if oracle.pwp_collides(ref.coeff_out[bi], int(b.n_int[bi]), sc.t_start[bi], par.T_span, tm, cx, cy, [delta] * 3):
                    want_col[bi] = 1
    assert np.array_equal(hout["collide"], want_col)
    # safetyCheckAfterReplan, entanglement half: gated on late arrivals, fresh PredictAlphasBetas on the updated samples
    es0 = (sc.es0_cnt, sc.es0_alpha, sc.es0_beta, sc.es0_bend, sc.es0_active)
    ent = pu.oracle_postcheck_entangle(oracle, par, sc.strep, b.agent_id, sc.known, late, b.bp_cnt, b.bp_xy, bc_l, bx_l, es0,
                                       sc.prev_pos, sc.prev_pos_agent, sc.state_A[:, 0, :2], b.n_int, ref.coeff_out, sc.t_start,
                                       sc.samp, late_committed)
    assert np.array_equal(hout["entangled"], ent)
    # published records: pwp_now (times shifted by t_start, generatePwpOut :898) composed with the agent's previous
    # record at time_now (replanFull neptune.cpp:1689-1699) + the DynTraj header; a rejected replan keeps the previous one
    rec = cyc.records("new")
    PW = capi.NB_REC_PWP_DOUBLES
    for bi in range(b.B):
        n, me = int(b.n_int[bi]), int(b.agent_id[bi]) - 1
        now = np.zeros(capi.NB_REC_DOUBLES)
        now[0] = n
        now[1:2 + n] = sc.t_start[bi] + par.T_span * np.arange(n + 1)
        now[18:PW].reshape(3, 16, 4)[:, :n] = hout["coeff_out"][bi, :, :n]
        prev = late_recs[me]
        if hout["status"][bi] >= 2 or hout["entangled"][bi] or hout["collide"][bi]:
            assert np.array_equal(rec[me], prev)
            continue
        npc, want, _, _ = oracle.compose_records(hin["t_now"][bi], par.dc, prev, now)
        assert npc == hout["n_pieces"][bi] and npc >= n
        assert np.array_equal(rec[me, :PW], want[:PW])
        assert rec[me, 1] == hin["t_now"][bi] and rec[me, 1 + npc] == now[1 + n]
        # header (publishOwnTraj, neptune_ros.cpp:436-480)
        assert rec[me, capi.REC_ID] == me + 1 and rec[me, capi.REC_ISAGENT] == 1 and rec[me, capi.REC_SEQ] == 0
        assert (rec[me, capi.REC_BBOX:capi.REC_BBOX + 3] == 2 * par.drone_radius).all()
        assert np.array_equal(rec[me, capi.REC_POS:capi.REC_POS + 3], want[18:PW].reshape(3, 16, 4)[:, 0, 3])
        nb = int(rec[me, capi.REC_NBEND])
        exp_bend = [np.asarray(par.pb[me], float)]
        for q in range(int(sc.es0_cnt[bi, 1])):
            aid, cs = sc.es0_alpha[bi, sc.es0_bend[bi, q]]
            exp_bend.append(np.asarray(par.pb[aid - 1], float) if aid <= par.num_of_agents
                            else np.asarray(sc.strep, float).reshape(-1, 2, 2)[aid - par.num_of_agents - 1, cs])
        assert nb == len(exp_bend) and np.array_equal(rec[me, capi.REC_BEND:capi.REC_BEND + 2 * nb].reshape(nb, 2), np.stack(exp_bend))
    assert cyc.index == 1
    cyc.close()


def test_replan_cycle_feedback_and_graph(capi, oracle):
    """Records fed back through the ring: cycle k plans against the records of cycle k - 2 and post-checks against
    those of k - 1, identically with plain launches and with CUDA-graph replay."""
    from neptune_b200.cycle import ReplanCycle
    par = config("obst8")
    sc = make_scene(par, 3003, sync=False, ent_backend=capi.DeviceEntBackend(_solver(capi, par, make_scene(par, 3003, sync=False))))
    b = sc.batch
    runs = []
    for graph in (False, True):
        cyc = ReplanCycle(par, b.agent_id - 1, "cuda:0", static=(b.st_ptr, b.st_xy, sc.strep), planned=np.ones(par.num_of_agents, np.uint8))
        cyc.seed_records(cyc.records_of(sc))
        hin, hout = cyc.host_inputs(sc), cyc.host_outputs()
        cyc.upload(hin)
        outs = []
        for k in range(6):
            if graph and k == 1:
                cyc.capture()          # runs cycle 1 itself, then captures
            else:
                cyc.step()
            cyc.download(hout)
            cyc.stream.synchronize()
            outs.append((hout["coeff_out"].copy(), hout["status"].copy(), cyc.records("new").copy()))
        cyc.check_errors()
        assert cyc.index == 6
        runs.append(outs)
        cyc.close()
    for (c0, s0, r0), (c1, s1, r1) in zip(*runs):
        assert np.array_equal(c0, c1) and np.array_equal(s0, s1) and np.array_equal(r0, r1)
    # the feedback matters: later cycles plan against different records than the first
    assert not np.array_equal(runs[0][0][2], runs[0][3][2])
    seq = runs[0][5][2][:, capi.REC_SEQ]      # committed in cycle 5, or an older record of a rejected replan
    assert (seq <= 5).all() and (seq == 5).sum() >= par.num_of_agents // 2


def _shim_input(par, b, a, with_ent):
    """Text input of tests/cpp/test_shim.cpp for agent a of a batch."""
    n, N, NH = int(b.n_int[a]), par.num_of_agents, b.n_hull_slots
    lines = [f"{N} {int(b.agent_id[a])} {n} {NH} {par.T_span!r} {par.weight!r}",
             " ".join(repr(float(v)) for v in (par.x_min, par.x_max, par.y_min, par.y_max, par.z_min, par.z_max)),
             f"{par.v_max!r} {par.a_max!r}"]
    lines += [f"{float(p[0])!r} {float(p[1])!r}" for p in np.asarray(par.pb, float)]
    for ax in range(3):
        for i in range(n):
            lines.append(" ".join(repr(float(v)) for v in b.coeff_init[a, ax, i]))
    for s in range(NH):
        for i in range(n):
            k = (a * NH + s) * 8 + i
            pts = b.hull_xy[b.hull_ptr[k]:b.hull_ptr[k + 1]]
            lines.append(str(len(pts)) + " " + " ".join(f"{float(p[0])!r} {float(p[1])!r}" for p in pts))
    if with_ent:
        lines.append("ENT")
        for i in range(n + 1):
            na = int(b.esv_cnt[a, i, 0])
            lines.append(str(na) + " " + " ".join(f"{int(x)} {int(y)}" for x, y in b.esv_alpha[a, i, :na]))
            lines.append(" ".join(str(int(v)) for v in b.esv_active[a, i, :N]))
        for j in range(N):
            nb = int(b.bp_cnt[j])
            lines.append(str(nb) + " " + " ".join(f"{float(p[0])!r} {float(p[1])!r}" for p in b.bp_xy[j, :nb]))
        for j in range(N):
            for i in range(n):
                x, y = b.nih0[a, j, i]
                lines.append("0 0 0" if np.isnan(x) else f"1 {float(x)!r} {float(y)!r}")
    return "\n".join(lines) + "\n"


def _shim_binaries():
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bd = os.path.join(root, "tests", "cpp", "_build")
    if not os.path.exists(os.path.join(bd, "test_shim")):
        subprocess.run(["make", "-C", os.path.join(root, "tests", "cpp")], check=True, capture_output=True)
    # test_shim_ref: the same program compiled BESIDE the reference's own headers (tests/cpp/compile_against_reference.cpp);
    # built where /root/reference exists and shipped with the tree
    return [os.path.join(bd, x) for x in ("test_shim", "test_shim_ref") if os.path.exists(os.path.join(bd, x))]


def _run_shim(exe, text, tmp_path, tag):
    import subprocess
    fn = tmp_path / f"{tag}.txt"
    fn.write_text(text)
    return subprocess.run([exe, str(fn)], check=True, capture_output=True, text=True).stdout.split("\n")


def test_cpp_shim_polysolvergurobi_and_separator(capi, oracle, tmp_path):
    """The C++ drop-in classes (PolySolverGurobi / separator::Separator with the reference's method
    names and call order) through the C-ABI, against the oracle on the same agent -- with the repo's stand-in types and,
    where the binary was built, with the reference's own mader_types.hpp / entangle_utils.hpp in the translation unit."""
    par = config("mtlp5")
    sc = make_scene(par, 2002, sync=False)
    b = sc.batch
    ref = ReplanResult.empty(b)
    assert oracle.replan_batch(b, ref, 1) == 0
    exes = _shim_binaries()
    assert exes
    for exe in exes:
        for a in (0, 2):
            n, NH = int(b.n_int[a]), b.n_hull_slots
            out = _run_shim(exe, _shim_input(par, b, a, False), tmp_path, f"agent{a}")
            ok, status, obj, ntraj = out[0].split()
            assert int(status) == ref.status[a] and int(ok) == int(ref.status[a] != 2)
            co = np.array([[float(v) for v in ln.split()] for ln in out[1:1 + 3 * n]]).reshape(3, n, 4)
            assert np.abs(co - ref.coeff_out[a, :, :n]).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out[a]).max())
            assert abs(float(obj) - ref.obj[a]) <= 1e-8 * max(1.0, abs(ref.obj[a]))
            assert abs(int(ntraj) - n * par.T_span / par.dc) <= 1.5
            s_ok = int(out[1 + 3 * n].split()[0])      # separator::Separator::solveModel on hulls[0][0]
            k0 = a * NH * 8
            first = b.hull_xy[b.hull_ptr[k0]:b.hull_ptr[k0 + 1]]
            far = np.array([[100.0, 100.0], [101.0, 100.0], [101.0, 101.0], [100.0, 101.0]])
            assert s_ok == int(len(first) > 0 and oracle.separate(first, far)[0])


def test_cpp_shim_tether_constraints_and_failure_path(capi, oracle, tmp_path):
    """Through the C++ drop-in: the crafted tether-constraint scene (addEntangleConstraintForIJCase) equals the oracle,
    and an infeasible replan returns optimize() == false with pwp_out == pwp_init (solver_gurobi_poly.cpp:856-859)."""
    from tests import crafted
    for exe in _shim_binaries():
        par, b = crafted.ent_lp_batch(2)
        ref = ReplanResult.empty(b)
        assert oracle.replan_batch(b, ref, 1) == 0
        plain = ReplanResult.empty(b)
        assert oracle.replan_batch(crafted.without_tether_rows(b), plain, 1) == 0
        a, n = 0, int(b.n_int[0])
        out = _run_shim(exe, _shim_input(par, b, a, True), tmp_path, "ent")
        ok, status, obj, _ = out[0].split()
        co = np.array([[float(v) for v in ln.split()] for ln in out[1:1 + 3 * n]]).reshape(3, n, 4)
        assert int(ok) == 1 and int(status) == ref.status[a]
        assert np.abs(co - ref.coeff_out[a, :, :n]).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out[a]).max())
        # the tether rows matter in this scene: without them the optimum is a different one
        assert np.abs(ref.coeff_out[a] - plain.coeff_out[a]).max() > 0.1
        par, b = crafted.infeasible_batch("box")
        out = _run_shim(exe, _shim_input(par, b, 2, False), tmp_path, "box")
        ok, status, obj, _ = out[0].split()
        n = int(b.n_int[2])
        co = np.array([[float(v) for v in ln.split()] for ln in out[1:1 + 3 * n]]).reshape(3, n, 4)
        assert int(ok) == 0 and int(status) == 2 and np.array_equal(co, b.coeff_init[2, :, :n])


def test_entangle_random_walks_bend_points(capi, oracle):
    """K3 on walks that create and release bend points: bit-exact against the oracle."""
    from tests.ent_backends import OracleEntBackend
    from tests.ent_walks import compare_backends
    par = config("obst8")
    sc = make_scene(par, 3003, sync=False)
    s = _solver(capi, par, sc)
    mx_a, mx_b = compare_backends(par, sc, OracleEntBackend(oracle), capi.DeviceEntBackend(s), trials=20)
    assert mx_a >= 8 and mx_b >= 2
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,seed", [("obst8", 3004), ("mtlp5", 2005)])
def test_cycle_sparse_upload_equals_dense(capi, oracle, cfg, seed, monkeypatch):
    """Large worlds upload only the used prefix of every entanglement-list row and rebuild active_cases on the device
    (nb_cycle_upload_from).  Forced on a small world here: the rebuilt arrays equal the host's, every output of the cycle
    equals the dense upload's, and fewer bytes cross PCIe."""
    from neptune_b200.cycle import ReplanCycle
    from tests.ent_backends import OracleEntBackend
    par = config(cfg)
    sc = make_scene(par, seed, sync=False, ent_backend=OracleEntBackend(oracle))
    b = sc.batch
    outs = {}
    for mode in ("dense", "sparse"):
        monkeypatch.delenv("NB_CYCLE_DENSE_UPLOAD", raising=False)
        monkeypatch.delenv("NB_CYCLE_SPARSE_UPLOAD", raising=False)
        monkeypatch.setenv("NB_CYCLE_DENSE_UPLOAD" if mode == "dense" else "NB_CYCLE_SPARSE_UPLOAD", "1")
        cyc = ReplanCycle(par, b.agent_id - 1, "cuda:0", static=(b.st_ptr, b.st_xy, sc.strep), planned=np.ones(par.num_of_agents, np.uint8))
        cyc.seed_records(cyc.records_of(sc, seq=0))
        hin, hout = cyc.host_inputs(sc), cyc.host_outputs()
        moved, _ = cyc.step_from_host(hin, hout)
        cyc.check_errors()
        B, NA = cyc.B, par.NA
        outs[mode] = dict(moved=moved, esv_active=cyc.fetch("in_esv_active", (B, 9, NA), np.int32), es_active=cyc.fetch("in_es_active", (B, NA), np.int32),
                          rec=cyc.records("new").copy(), **{k: hout[k].copy() for k in ("coeff_out", "status", "entangled", "collide", "n_pieces")})
        cyc.close()
    d, sp = outs["dense"], outs["sparse"]
    assert np.array_equal(d["esv_active"], b.esv_active.reshape(d["esv_active"].shape))      # the dense upload carries the host's arrays
    assert np.array_equal(sp["esv_active"], d["esv_active"]) and np.array_equal(sp["es_active"], d["es_active"])
    for k in ("coeff_out", "status", "entangled", "collide", "n_pieces", "rec"):
        assert np.array_equal(sp[k], d[k]), k
    assert sp["moved"] < d["moved"]

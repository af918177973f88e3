"""Online entanglement tracker (NeptuneRos::updateEntStateStaticObs, reference neptune/src/neptune_ros.cpp:798-850,
with the 9-argument eu::entangleHSigToAddAgentInd, entangle_utils.cpp:820-1127).

Integer outputs (the signature word, active cases, bend-point indices, result codes) and the copied FP64 state
(previousCheckingPos_, betas) are compared bit for bit: device code on one host lane and the CUDA kernel through
the C-ABI against the oracle, over multi-tick walks in which tethers gain and lose contact points."""
import numpy as np
import pytest

from neptune_b200 import config
from neptune_b200.capi import EntArrays
from neptune_b200.scenes import make_scene


def _walk(par, seed, ticks=40, p_toggle=0.15):
    """A random multi-tick scenario: every agent random-walks, other tethers gain / lose a contact point now and
    then, some agents are silent (x = -1000: no message yet), some ticks fall under the 5 cm / 100 ms gate."""
    rng = np.random.default_rng(seed)
    N, bpm = par.num_of_agents, par.bp_max
    pb = np.asarray(par.pb, float)
    pos = pb + rng.normal(0, 3.0, size=(N, 2))
    bp_cnt = np.ones(N, np.int32)
    bp_xy = np.zeros((N, bpm, 2))
    bp_xy[:, 0] = pb
    silent = rng.random(N) < 0.15
    frames = []
    for t in range(ticks):
        prev_cnt, prev_xy = bp_cnt.copy(), bp_xy.copy()
        for j in range(N):
            if rng.random() < p_toggle:
                if bp_cnt[j] > 1 and rng.random() < 0.5:
                    bp_cnt[j] -= 1
                elif bp_cnt[j] < 3:
                    bp_xy[j, bp_cnt[j]] = 0.5 * (pb[j] + pos[j]) + rng.normal(0, 1.5, size=2)
                    bp_cnt[j] += 1
        tiny = rng.random(N) < 0.1
        step = rng.normal(0, 0.6, size=(N, 2))
        step[tiny] *= 0.01
        pos = pos + step
        elapsed = np.where(rng.random(N) < 0.5, 20.0, 150.0)
        latest = pos.copy()
        latest[silent] = -1000.0
        frames.append(dict(bp_cnt=bp_cnt.copy(), bp_xy=bp_xy.copy(), bp_cnt_prev=prev_cnt, bp_xy_prev=prev_xy,
                           cur=pos.copy(), latest=latest, elapsed=elapsed))
    start = pb + rng.normal(0, 3.0, size=(N, 2))
    return start, silent, frames


def _init(par, start, silent):
    N = par.num_of_agents
    st = EntArrays(par, N)
    prev_pos = np.repeat(start[:, None, :], N + 1, axis=1).copy()           # setUpCheckingPosAndStaticObs :854-857
    prev_pos_agent = np.repeat(start[None, :, :], N, axis=0).copy()
    prev_pos_agent[:, silent] = -1000.0
    return st, np.ascontiguousarray(prev_pos), np.ascontiguousarray(prev_pos_agent)


def _oracle_tick(oracle, par, strep, st, pp, ppa, fr):
    N = par.num_of_agents
    res = np.zeros(N, np.int32)
    for b in range(N):
        es = oracle.EntState(par.ent_cap, par.NA)
        es.n_alpha, es.n_bend = int(st.cnt[b, 0]), int(st.cnt[b, 1])
        es.alpha[:], es.beta[:], es.bend[:], es.active[:] = st.alpha[b], st.beta[b], st.bend[b], st.active[b]
        cx = oracle.EntCtx(par, b, strep, fr["bp_cnt"], fr["bp_xy"])
        latest_b = np.ascontiguousarray(fr["latest"])
        res[b] = oracle.track(es, cx, fr["bp_cnt_prev"], fr["bp_xy_prev"], pp[b], ppa[b], latest_b, fr["cur"][b], float(fr["elapsed"][b]))
        st.cnt[b] = [es.n_alpha, es.n_bend]
        st.alpha[b], st.beta[b], st.bend[b], st.active[b] = es.alpha, es.beta, es.bend, es.active
    return res


def _run(oracle, par, strep, seed, device_tick):
    start, silent, frames = _walk(par, seed)
    N = par.num_of_agents
    ids = np.arange(1, N + 1, dtype=np.int32)
    st_o, pp_o, ppa_o = _init(par, start, silent)
    st_d, pp_d, ppa_d = _init(par, start, silent)
    seen = {0: 0, 1: 0, "neg": 0, "changed": 0}
    for fr in frames:
        latest = np.repeat(fr["latest"][None], N, axis=0)
        res_o = _oracle_tick(oracle, par, strep, st_o, pp_o, ppa_o, fr)
        res_d, st_d, pp_d, ppa_d = device_tick(ids, fr["bp_cnt"], fr["bp_xy"], fr["bp_cnt_prev"], fr["bp_xy_prev"], st_d,
                                               pp_d, ppa_d, latest, fr["cur"], fr["elapsed"])
        assert np.array_equal(res_o, res_d)
        ok = res_o >= 0   # where the reference would have exited the state is no longer defined
        assert np.array_equal(st_o.cnt[ok], st_d.cnt[ok])
        for b in np.flatnonzero(ok):
            na, nb = st_o.cnt[b]
            assert np.array_equal(st_o.alpha[b, :na], st_d.alpha[b, :na]) and np.array_equal(st_o.beta[b, :na], st_d.beta[b, :na])
            assert np.array_equal(st_o.bend[b, :nb], st_d.bend[b, :nb]) and np.array_equal(st_o.active[b], st_d.active[b])
        assert np.array_equal(pp_o[ok], pp_d[ok]) and np.array_equal(ppa_o[ok], ppa_d[ok])
        # keep both sides on the oracle's state (stop cases included) so that later ticks stay comparable
        st_d, pp_d, ppa_d = st_o.copy(), pp_o.copy(), ppa_o.copy()
        seen[0] += int((res_o == 0).sum()); seen[1] += int((res_o == 1).sum()); seen["neg"] += int((res_o < 0).sum())
        seen["changed"] += int((fr["bp_cnt"] != fr["bp_cnt_prev"]).sum())
    assert seen[0] > 50 and seen[1] > 5 and seen["changed"] > 10
    assert int(st_o.cnt[:, 0].max()) >= 1    # the walks do cross tethers
    return seen


def _scene(oracle, cfg, seed):
    par = config(cfg)
    sc = make_scene(par, seed, sync=True)
    return par, sc.strep


def test_tracker_without_contact_changes_equals_predict(oracle):
    """With bendPtsForAgents_ unchanged the 9-argument test is the 8-argument one (entangle_utils.cpp:826-830), so a
    tracker tick equals PredictAlphasBetas (neptune.cpp:976-1008) on the same positions: two code paths of the oracle."""
    par, strep = _scene(oracle, "obst8", 3003)
    start, silent, frames = _walk(par, 5, ticks=25, p_toggle=0.0)
    N = par.num_of_agents
    st, pp, ppa = _init(par, start, silent)
    for fr in frames:
        fr["elapsed"][:] = 1000.0
        st_p, pp_before, ppa_before = st.copy(), pp.copy(), ppa.copy()
        _oracle_tick(oracle, par, strep, st, pp, ppa, fr)
        for b in range(N):
            es = oracle.EntState(par.ent_cap, par.NA)
            es.n_alpha, es.n_bend = int(st_p.cnt[b, 0]), int(st_p.cnt[b, 1])
            es.alpha[:], es.beta[:], es.bend[:], es.active[:] = st_p.alpha[b], st_p.beta[b], st_p.bend[b], st_p.active[b]
            cx = oracle.EntCtx(par, b, strep, fr["bp_cnt"], fr["bp_xy"])
            known = (ppa_before[b, :, 0] > -900).astype(np.uint8)
            assert oracle.predict(es, cx, pp_before[b], ppa_before[b], fr["cur"][b], fr["latest"], known) == 0
            assert (es.n_alpha, es.n_bend) == tuple(st.cnt[b])
            assert np.array_equal(es.alpha[:es.n_alpha], st.alpha[b, :es.n_alpha]) and np.array_equal(es.active, st.active[b])


@pytest.mark.parametrize("cfg,seed", [("obst8", 11), ("obst8", 12), ("mtlp5", 13)])
def test_emulated_tracker_matches_oracle(oracle, cfg, seed):
    from tests.emul import emul
    par, strep = _scene(oracle, cfg, 3003)
    _run(oracle, par, strep, seed, lambda *a: emul.track(par, strep, *a))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,seed", [("obst8", 21), ("obst8", 22), ("mtlp5", 23), ("grid64", 24)])
def test_gpu_tracker_matches_oracle(oracle, cfg, seed):
    from neptune_b200 import capi
    par, strep = _scene(oracle, cfg, 3003)
    s = capi.Solver(par)
    if par.num_of_static_obst:
        sc = make_scene(par, 3003, sync=True)
        s.set_static(sc.batch.st_ptr, sc.batch.st_xy, sc.strep)
    _run(oracle, par, strep, seed, lambda *a: s.entangle_track(*a))
    s.close()


def test_static_obstacle_representation(oracle):
    """NeptuneRos::setUpCheckingPosAndStaticObs (neptune_ros.cpp:852-1019): the product's host function against the
    oracle (bit-exact: same libm on the same host) and against what the reference's construction guarantees -- both
    points lie on the obstacle's boundary on a common-slope line through its centre, on opposite sides of it, the
    parallel lines are at least 1.6 voxels apart, none cuts the straight tether."""
    import ctypes as C
    from neptune_b200 import capi
    from neptune_b200.params import multi_obstacle_squares
    lib = capi.lib()
    f = lib.nb_static_obst_rep
    f.argtypes = [C.c_int32] + [C.c_void_p] * 4 + [C.c_double] + [C.c_void_p] * 2
    g = oracle.lib().orc_static_obst_rep
    g.restype = C.c_int
    rng = np.random.default_rng(3)
    sq = np.array([[0.25, 0.25], [0.25, -0.25], [-0.25, -0.25], [-0.25, 0.25]])
    n_ok = 0
    for trial in range(40):
        if trial == 0:
            polys = multi_obstacle_squares()
        else:
            polys = []
            while len(polys) < int(rng.integers(1, 12)):
                c = rng.uniform(-12, 12, size=2)
                if all(np.linalg.norm(c - q.mean(axis=0)) > 1.5 for q in polys):
                    polys.append(c + sq * rng.uniform(0.6, 2.0) if rng.random() < 0.7 else
                                 c + np.array([[0.4, 0.0], [0.0, 0.5], [-0.5, 0.1], [-0.1, -0.6]]))
        M = len(polys)
        ptr = np.concatenate([[0], np.cumsum([len(q) for q in polys])]).astype(np.int64)
        xy = np.ascontiguousarray(np.concatenate(polys), np.float64)
        base, pos = rng.uniform(-14, 14, size=2), rng.uniform(-14, 14, size=2)
        out = [np.zeros((M, 2, 2)), np.zeros((M, 2)), np.zeros((M, 2, 2)), np.zeros((M, 2))]
        rc_p = f(M, ptr.ctypes.data, xy.ctypes.data, base.ctypes.data, pos.ctypes.data, 0.2, out[0].ctypes.data, out[1].ctypes.data)
        rc_o = g(C.c_int(M), ptr.ctypes.data_as(C.c_void_p), xy.ctypes.data_as(C.c_void_p), base.ctypes.data_as(C.c_void_p),
                 pos.ctypes.data_as(C.c_void_p), C.c_double(0.2), out[2].ctypes.data_as(C.c_void_p), out[3].ctypes.data_as(C.c_void_p))
        assert (rc_p == 0) == (rc_o == 0)
        if rc_o != 0:
            continue
        n_ok += 1
        assert np.array_equal(out[0], out[2]) and np.array_equal(out[1], out[3])
        strep, longest = out[0], out[1]
        d = strep[:, 1] - strep[:, 0]
        ang = np.arctan2(d[:, 1], d[:, 0])
        assert np.abs(np.sin(ang - ang[0])).max() < 1e-9                      # one common slope
        nrm = np.array([-np.sin(ang[0]), np.cos(ang[0])])
        cen = np.array([q.mean(axis=0) for q in polys])
        off = cen @ nrm
        assert np.abs((strep[:, 0] - cen) @ nrm).max() < 1e-9                 # through the centre
        assert (((strep[:, 0] - cen) * (strep[:, 1] - cen)).sum(axis=1) < 0).all()   # on opposite sides of it
        if M > 1:
            gaps = np.abs(off[:, None] - off[None, :])[np.triu_indices(M, 1)]
            assert gaps.min() >= 0.2 * 1.6 - 1e-9
        assert np.abs(off - base @ nrm).min() >= 0.2 * 1.6 - 1e-9
        for m in range(M):                                                    # on the boundary; longest distance is a vertex distance
            for c in range(2):
                q = polys[m]
                def seg_dist(k):
                    e, w = q[(k + 1) % len(q)] - q[k], strep[m, c] - q[k]
                    return abs(e[0] * w[1] - e[1] * w[0]) / np.linalg.norm(e)
                dist = min(seg_dist(k) for k in range(len(q)))
                assert dist < 1e-9
                assert abs(longest[m, c] - np.linalg.norm(q - strep[m, c], axis=1).max()) < 1e-12
    assert n_ok >= 25


@pytest.mark.gpu
def test_gpu_tracker_device_pointers(oracle):
    """NB_DEVICE form of nb_entangle_track_batch (state and checking positions resident in HBM, asynchronous on the
    stream) gives the same bits as the NB_HOST form."""
    import ctypes as C
    import torch
    from neptune_b200 import capi
    par, strep = _scene(oracle, "obst8", 3003)
    sc = make_scene(par, 3003, sync=True)
    s = capi.Solver(par)
    s.set_static(sc.batch.st_ptr, sc.batch.st_xy, sc.strep)
    start, silent, frames = _walk(par, 31, ticks=6)
    N = par.num_of_agents
    ids = np.arange(1, N + 1, dtype=np.int32)
    st, pp, ppa = _init(par, start, silent)
    dev = torch.device("cuda", 0)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d_st = {k: to(getattr(st, k)) for k in ("cnt", "alpha", "beta", "bend", "active")}
    d_pp, d_ppa, d_ids = to(pp), to(ppa), to(ids)
    f = capi.lib().nb_entangle_track_batch
    P = C.c_void_p
    f.argtypes = [P, C.c_int32, C.c_int32, P, P, P, P, P, capi.NbEntState, P, P, P, P, P, P, P]
    for fr in frames:
        latest = np.repeat(fr["latest"][None], N, axis=0)
        res_h, st, pp, ppa = s.entangle_track(ids, fr["bp_cnt"], fr["bp_xy"], fr["bp_cnt_prev"], fr["bp_xy_prev"], st, pp, ppa,
                                              latest, fr["cur"], fr["elapsed"])
        keep = [to(fr["bp_cnt"]), to(fr["bp_xy"]), to(fr["bp_cnt_prev"]), to(fr["bp_xy_prev"]), to(latest), to(fr["cur"]),
                to(fr["elapsed"]), torch.zeros(N, dtype=torch.int32, device=dev)]
        es = capi.NbEntState()
        es.cnt, es.alpha, es.beta, es.bend, es.active = (P(d_st[k].data_ptr()) for k in ("cnt", "alpha", "beta", "bend", "active"))
        p = lambda t: P(t.data_ptr())  # noqa: E731
        rc = f(s.handle, N, capi.NB_DEVICE, p(d_ids), p(keep[0]), p(keep[1]), p(keep[2]), p(keep[3]), es, p(d_pp), p(d_ppa),
               p(keep[4]), p(keep[5]), p(keep[6]), p(keep[7]), P(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        torch.cuda.synchronize()
        assert np.array_equal(keep[7].cpu().numpy(), res_h)
        ok = res_h >= 0
        assert np.array_equal(d_st["cnt"].cpu().numpy()[ok], st.cnt[ok])
        assert np.array_equal(d_pp.cpu().numpy()[ok], pp[ok]) and np.array_equal(d_ppa.cpu().numpy()[ok], ppa[ok])
        for b in np.flatnonzero(ok):
            na = st.cnt[b, 0]
            assert np.array_equal(d_st["alpha"].cpu().numpy()[b, :na], st.alpha[b, :na])
            assert np.array_equal(d_st["active"].cpu().numpy()[b], st.active[b])
        # keep both sides on the same state where the reference would have exited
        for k in d_st:
            d_st[k].copy_(to(getattr(st, k)))
        d_pp.copy_(to(pp)), d_ppa.copy_(to(ppa))
    s.close()

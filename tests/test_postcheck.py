"""The post-check of a replan against trajectories that arrived during the optimisation: the entanglement half of
``Neptune::safetyCheckAfterReplan`` (reference neptune/src/neptune.cpp:735-752) with its gating (only agents with a late
trajectory; samples re-drawn over the optimised trajectory's time span; bend points of the late message; a fresh
PredictAlphasBetas), and the message side of the exchange (DynTraj header of the committed-trajectory records).
CPU: single-lane emulation of the kernel against the oracle's pieces composed in the reference's order.  GPU: the
library through the C-ABI against the same."""
import numpy as np
import pytest

from neptune_b200 import capi, config
from neptune_b200.batch import ReplanResult
from neptune_b200.capi import EntArrays
from neptune_b200.scenes import make_scene
from tests import postcheck_util as pu
from tests.ent_backends import OracleEntBackend


def _cases(oracle, cfg, seed, trials):
    par = config(cfg)
    sc = make_scene(par, seed, sync=False, ent_backend=OracleEntBackend(oracle))
    res = ReplanResult.empty(sc.batch)
    assert oracle.replan_batch(sc.batch, res, 2) == 0
    rng = np.random.default_rng(seed)
    es = pu.with_history(par, sc, rng)
    for late, committed, late_recs, bc, bx in pu.late_cases(par, sc, rng, trials):
        b = sc.batch
        want = pu.oracle_postcheck_entangle(oracle, par, sc.strep, b.agent_id, sc.known, late, b.bp_cnt, b.bp_xy, bc, bx, es,
                                            sc.prev_pos, sc.prev_pos_agent, sc.state_A[:, 0, :2], b.n_int, res.coeff_out,
                                            sc.t_start, sc.samp, committed)
        yield par, sc, res, es, late, late_recs, bc, bx, want


@pytest.mark.parametrize("cfg,seed", [("obst8", 3003), ("mtlp5", 2005)])
def test_emulated_postcheck_entangle_matches_oracle(oracle, cfg, seed):
    from tests.emul import emul
    n_ent = n_skip = 0
    for par, sc, res, es, late, late_recs, bc, bx, want in _cases(oracle, cfg, seed, 25):
        b = sc.batch
        got = emul.postcheck_entangle(par, sc.strep, b.agent_id, sc.known, late, b.bp_cnt, b.bp_xy, bc, bx, EntArrays.of(par, *es),
                                      sc.prev_pos, sc.prev_pos_agent, np.ascontiguousarray(sc.state_A[:, 0, :2]), b.n_int,
                                      res.coeff_out, sc.t_start, sc.samp, late_recs)
        assert np.array_equal(got, want)
        n_ent += int(want.sum())
        n_skip += int((late.sum(axis=1) == 0).sum())
    assert n_ent >= 5 and n_skip >= 1      # entangling outcomes and gated agents both occur


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,seed", [("obst8", 3003), ("mtlp5", 2005)])
def test_gpu_postcheck_entangle_matches_oracle(oracle, cfg, seed):
    n_ent = 0
    s = None
    for par, sc, res, es, late, late_recs, bc, bx, want in _cases(oracle, cfg, seed, 25):
        b = sc.batch
        if s is None:
            s = capi.Solver(par)
            if par.num_of_static_obst:
                s.set_static(b.st_ptr, b.st_xy, sc.strep)
        got = s.postcheck_entangle(b.agent_id, sc.known, late, b.bp_cnt, b.bp_xy, bc, bx, EntArrays.of(par, *es), sc.prev_pos,
                                   sc.prev_pos_agent, np.ascontiguousarray(sc.state_A[:, 0, :2]), b.n_int, res.coeff_out,
                                   sc.t_start, sc.samp, late_recs)
        assert np.array_equal(got, want)
        n_ent += int(want.sum())
    assert n_ent >= 5
    s.close()


@pytest.mark.gpu
def test_gpu_records_carry_the_dyntraj_header(oracle):
    """nb_unpack_records_batch (trajCB: bendpt[], pos) inverts what capi.make_records packs (publishOwnTraj)."""
    par = config("obst8")
    sc = make_scene(par, 3004, sync=False)
    bc, bx = sc.batch.bp_cnt.copy(), sc.batch.bp_xy.copy()
    bx[2, bc[2]] = [1.5, -2.5]
    bc[2] += 1
    recs = capi.make_records(sc.committed, par, bc, bx, seq=7)
    assert recs.shape[1] == 256 and (recs[:, capi.REC_ID] == np.arange(1, par.num_of_agents + 1)).all()
    s = capi.Solver(par)
    cnt, xy, pos = s.unpack_records(recs)
    assert np.array_equal(cnt, bc)
    for j in range(par.num_of_agents):
        assert np.array_equal(xy[j, :bc[j]], bx[j, :bc[j]]) and (xy[j, bc[j]:] == 0).all()
        assert np.array_equal(pos[j], [sc.committed[j][1][0][3], sc.committed[j][2][0][3]])
    s.close()


@pytest.mark.gpu
def test_gpu_static_representation_per_agent(oracle):
    """nb_set_static_rep_per_agent: every agent its own staticObsRep_ (NeptuneRos::setUpCheckingPosAndStaticObs,
    neptune_ros.cpp:852-1019, computed by nb_static_obst_rep from the agent's base and position).  PredictAlphasBetas, the
    chain behind entStateVec and the post-check of every agent then equal the oracle run with THAT agent's representation,
    and differ from the shared-representation run where the representations differ."""
    import ctypes as C
    from neptune_b200.capi import EntArrays
    par = config("obst8")
    sc = make_scene(par, 3003, sync=False, ent_backend=OracleEntBackend(oracle))
    b = sc.batch
    N, M = par.num_of_agents, par.num_of_static_obst
    lib = capi.lib()
    f = lib.nb_static_obst_rep
    f.argtypes = [C.c_int32] + [C.c_void_p] * 4 + [C.c_double] + [C.c_void_p] * 2
    polys = sc.static_raw
    ptr = np.concatenate([[0], np.cumsum([len(q) for q in polys])]).astype(np.int64)
    xy = np.ascontiguousarray(np.concatenate(polys), np.float64)
    strep_all, longest_all = np.zeros((N, M, 2, 2)), np.zeros((N, M, 2))
    rng = np.random.default_rng(5)
    for j in range(N):
        base = np.ascontiguousarray(par.pb[j], np.float64)
        for attempt in range(50):     # a position from which a representation exists (the reference exits otherwise)
            pos = np.ascontiguousarray(base + rng.normal(size=2) * 2.0)
            if f(M, ptr.ctypes.data, xy.ctypes.data, base.ctypes.data, pos.ctypes.data, par.a_star_fraction_voxel_size,
                 strep_all[j].ctypes.data, longest_all[j].ctypes.data) == 0:
                break
        else:
            raise AssertionError("no feasible static representation found")
    assert np.abs(strep_all - strep_all[0]).max() > 1e-3          # the agents' representations really differ
    s = capi.Solver(par)
    s.set_static(b.st_ptr, b.st_xy, sc.strep)
    s.set_static_rep_per_agent(strep_all, longest_all)
    es = EntArrays.of(par, sc.es0_cnt, sc.es0_alpha, sc.es0_beta, sc.es0_bend, sc.es0_active)
    cur = np.ascontiguousarray(sc.state_A[:, 0, :2])
    samp0 = np.ascontiguousarray(sc.samp[:, :, 0, 0, :])
    got = s.entangle_predict(b.agent_id, sc.known, b.bp_cnt, b.bp_xy, es, sc.prev_pos, sc.prev_pos_agent, cur, samp0)
    done, roll = s.entangle_rollout(b.agent_id, sc.known, b.bp_cnt, b.bp_xy, got, b.n_int, b.coeff_init, sc.samp)
    ob = OracleEntBackend(oracle)
    n_diff = 0
    for a in range(b.B):
        me = int(b.agent_id[a]) - 1
        sl = slice(a, a + 1)
        want = ob.predict_batch(par, b.agent_id[sl], sc.prev_pos[sl], sc.prev_pos_agent[sl], cur[sl], samp0[sl], sc.known[sl],
                                strep_all[me], b.bp_cnt, b.bp_xy, sc.es0_cnt[sl], sc.es0_alpha[sl], sc.es0_beta[sl], sc.es0_bend[sl],
                                sc.es0_active[sl])
        for x, y in zip((got.cnt[sl], got.alpha[sl], got.beta[sl], got.bend[sl], got.active[sl]), want):
            assert np.array_equal(x, y), a
        wr = ob.rollout_batch(par, b.agent_id[sl], b.n_int[sl], b.coeff_init[sl], sc.samp[sl], sc.known[sl], strep_all[me], b.bp_cnt,
                              b.bp_xy, *want)
        assert wr[0][0] == done[a]
        for x, y in zip((roll.cnt[sl], roll.alpha[sl], roll.beta[sl], roll.bend[sl], roll.active[sl]), wr[1:]):
            assert np.array_equal(x, y), a
        shared = ob.rollout_batch(par, b.agent_id[sl], b.n_int[sl], b.coeff_init[sl], sc.samp[sl], sc.known[sl], sc.strep, b.bp_cnt,
                                  b.bp_xy, *want)
        n_diff += int(not all(np.array_equal(x, y) for x, y in zip(wr[1:], shared[1:])))
    s.close()

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

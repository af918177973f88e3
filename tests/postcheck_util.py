"""The entanglement half of ``Neptune::safetyCheckAfterReplan`` (reference neptune/src/neptune.cpp:735-752) composed from
the oracle's own pieces -- SamplePointsOfIntervals, PredictAlphasBetas, entangleCheckGivenPwp, each pinned to the
reference's code on its own -- in the order the reference calls them, plus a generator of late-trajectory cases.
Test infrastructure."""
from __future__ import annotations

import numpy as np

from neptune_b200 import capi


def oracle_postcheck_entangle(orc, par, strep, agent_id, known, late, bp_cnt, bp_xy, bp_cnt_late, bp_xy_late, es_tuple,
                              prev_pos, prev_pos_agent, cur, n_int, coeff, t_start, samp, late_committed):
    """entangled [B].  late_committed[j] = (times, cx, cy, cz) of the late trajectory of agent j."""
    B, N, P, S = len(agent_id), par.num_of_agents, par.num_pol, par.num_sample_per_interval
    cnt, alpha, beta, bend, active = es_tuple
    out = np.zeros(B, np.int32)
    for b in range(B):
        me = int(agent_id[b]) - 1
        lt = np.array([bool(late[b, j]) and j != me for j in range(N)])
        if not lt.any():                                   # need_to_rerun_entanglecheck stays false (:722, :743)
            continue
        n = int(n_int[b])
        smp = np.array(samp[b], np.float64, copy=True)
        kn = np.array(known[b], np.uint8, copy=True)
        bc, bx = np.array(bp_cnt, np.int32, copy=True), np.array(bp_xy, np.float64, copy=True)
        for j in np.flatnonzero(lt):
            tm, cx, cy, _ = late_committed[j]
            s, _ = orc.sample_points(tm, cx, cy, t_start[b], t_start[b] + n * par.T_span, P, S)     # :737-738
            smp[j], kn[j] = s, 1
            bc[j], bx[j] = bp_cnt_late[j], bp_xy_late[j]                                                # :740-741
        es = orc.EntState(par.ent_cap, par.NA)
        es.n_alpha, es.n_bend = int(cnt[b, 0]), int(cnt[b, 1])
        es.alpha[:], es.beta[:], es.bend[:], es.active[:] = alpha[b], beta[b], bend[b], active[b]
        ecx = orc.EntCtx(par, me, strep, bc, bx)
        rc = orc.predict(es, ecx, prev_pos[b], prev_pos_agent[b], cur[b], np.ascontiguousarray(smp[:, 0, 0, :]), kn)   # :747-749
        assert rc == 0
        cxy = np.ascontiguousarray(coeff[b, :2, :n, :])
        out[b] = 1 if orc.entangle_check_pwp(es, ecx, n, cxy, smp, kn) > 0 else 0                       # :750
    return out


def late_cases(par, sc, rng, trials):
    """(late [B][N], late_committed, late_recs, bp_cnt_late, bp_xy_late) per trial: a random subset of the known agents
    re-published, with a trajectory that starts where the old one is at t_start but heads somewhere else, and with a
    tether that may have gained a bend point."""
    N = par.num_of_agents
    for _ in range(trials):
        late = (rng.random((sc.batch.B, N)) < 0.5) & (sc.known > 0)
        if rng.random() < 0.2:
            late[rng.integers(sc.batch.B)] = False                      # an agent without any late trajectory
        committed = []
        centre = sc.state_A[int(rng.integers(sc.batch.B)), 0, :2]
        for j in range(N):
            tm, cx, cy, cz = sc.committed[j]
            cx2, cy2 = np.array(cx, copy=True), np.array(cy, copy=True)
            if rng.random() < 0.5:      # another velocity from piece k on
                k = int(rng.integers(len(cx2)))
                cx2[k:, 2] += rng.normal() * 1.5
                cy2[k:, 2] += rng.normal() * 1.5
            else:                       # a fast straight pass close to one planning agent's start: tethers sweep across paths
                p0, v = centre + rng.normal(size=2) * 2.0, rng.normal(size=2) * 4.0
                for k in range(len(cx2)):
                    dt = float(tm[k] - tm[0])
                    cx2[k], cy2[k] = [0, 0, v[0], p0[0] + v[0] * dt], [0, 0, v[1], p0[1] + v[1] * dt]
            committed.append((np.array(tm, copy=True), cx2, cy2, np.array(cz, copy=True)))
        bc, bx = sc.batch.bp_cnt.copy(), sc.batch.bp_xy.copy()
        for j in range(N):
            if rng.random() < 0.4 and bc[j] < par.bp_max:
                bx[j, bc[j]] = bx[j, 0] + rng.normal(size=2) * 3.0
                bc[j] += 1
        yield late.astype(np.uint8), committed, capi.make_records(committed, par, bc, bx, seq=1), bc, bx


def with_history(par, sc, rng, frac=0.6):
    """entangle_state_ of the scene plus one earlier crossing (id, case) for a fraction of the other agents: a state
    from which a second, different crossing of the same tether makes active_cases reach 2 (entangled)."""
    cnt, alpha, beta, bend, active = (np.array(x, copy=True) for x in (sc.es0_cnt, sc.es0_alpha, sc.es0_beta, sc.es0_bend, sc.es0_active))
    N = par.num_of_agents
    for b in range(sc.batch.B):
        me = int(sc.batch.agent_id[b]) - 1
        for j in range(N):
            if j == me or active[b, j] != 0 or rng.random() > frac or cnt[b, 0] >= par.ent_cap - 1:
                continue
            q = cnt[b, 0]
            alpha[b, q] = [j + 1, int(rng.integers(0, 3))]
            beta[b, q] = 0.0
            active[b, j] = 1
            cnt[b, 0] += 1
    return cnt, alpha, beta, bend, active

"""Crafted replan inputs that reach the two branches no generated scene reaches:

* non-entangling separating lines -- ``addEntangleConstraintForIJCase`` (reference
  ``neptune/src/solver_gurobi_poly.cpp:620-642``, ``:715-784``): agents with ``active_cases[j] == 1`` whose tether (bend
  points + the ray beyond the agent) passes within one control-polygon length of the planning agent's path;
* the failure path of ``optimize`` (``:832-861``): both solves infeasible => ``pwp_out = pwp_init``, status 2.

Used by the CPU tests (oracle vs the single-lane emulation vs HiGHS), by the GPU tests (CUDA vs oracle) and by
``tests/golden/make_golden.py``.  Everything is derived from seeded generated scenes, so the inputs are reproducible.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from neptune_b200 import config
from neptune_b200.scenes import make_scene


def ent_lp_batch(variant: int = 0, ent_slots: int = 16):
    """mtlp5 scene 2002 with three other agents marked 'one active case' for the planning agents.

    variant 0: two extra bend points per tether near agent 0's start (segments AND the ray are close): case ids
               1, 2 and 3 so that every skip rule (k == case_id) is exercised;
    variant 1: the same for EVERY planning agent, tethers with 2..4 bend points, one tether far away (distance gate
               :743-745 rejects its segments) and one agent with case id 0 ("not our concern", :632).
    variant 2: ONE tether (agent 1, case id 5, bend points base -> P1 -> P2) whose segment P1-P2 lies where agent 0's
               unconstrained optimum wants to go (found by search, hard-coded): its rows are ACTIVE at the optimum, the
               optimiser's answer moves by ~0.5 m -- the test that the tether rows really enter the QP.
    Returns (par, batch)."""
    if variant == 2:
        return _binding_tether_batch(ent_slots)
    par = dataclasses.replace(config("mtlp5"), ent_slots=ent_slots)
    par.pb = config("mtlp5").pb
    sc = make_scene(par, 2002, sync=False)
    b = sc.batch
    N = par.num_of_agents
    act, cntv, alpha = b.esv_active.copy(), b.esv_cnt.copy(), b.esv_alpha.copy()
    bp_cnt, bp_xy, nih0 = b.bp_cnt.copy(), b.bp_xy.copy(), b.nih0.copy()
    rng = np.random.default_rng(77 + variant)
    planners = [0] if variant == 0 else list(range(b.B))
    # tethers of the "other" agents: base, then bend points around the start point of planning agent 0
    p0 = b.coeff_init[0, :2, 0, 3].copy()
    for j in range(N):
        nb = 3 if variant == 0 else 2 + (j % 3)
        bp_cnt[j] = nb
        for k in range(1, nb):
            off = np.array([0.6 * np.cos(1.3 * j + 2.1 * k), 0.6 * np.sin(1.3 * j + 2.1 * k)]) * (1.0 + 0.4 * k)
            bp_xy[j, k] = p0 + off
        if variant == 1 and j == N - 1:     # a far tether: every segment is rejected by the distance gate
            for k in range(1, nb):
                bp_xy[j, k] = p0 + np.array([40.0 + k, 35.0])
    for a in planners:
        me = int(b.agent_id[a]) - 1
        others = [j for j in range(N) if j != me][:4 if variant == 1 else 3]
        act[a] = 0
        cntv[a, :, 0] = len(others)
        for q, j in enumerate(others):
            act[a, :, j] = 1
            case_id = [3, 1, 2, 0][q % 4] if variant == 1 else [5, 1, 2][q]
            alpha[a, :, q] = [j + 1, case_id]
            # hullsNoInflation_[j][i].col(0): the position of agent j over interval i
            nih0[a, j] = b.coeff_init[a, :2, 0, 3] + np.array([1.0 + 0.3 * q, -0.8 + 0.5 * q]) + 0.05 * rng.normal(size=(8, 2))
    out = dataclasses.replace(b, par=par, esv_active=act, esv_cnt=cntv, esv_alpha=alpha, bp_cnt=bp_cnt, bp_xy=bp_xy,
                              nih0=np.ascontiguousarray(nih0))
    out.validate()
    return par, out


def _binding_tether_batch(ent_slots: int):
    par, b = ent_lp_batch(0, ent_slots)      # positions (nih0) of the other agents as in variant 0
    a, me = 0, int(b.agent_id[0]) - 1
    j = [jj for jj in range(par.num_of_agents) if jj != me][0]
    bp_cnt, bp_xy = b.bp_cnt.copy(), b.bp_xy.copy()
    c, dv = np.array([3.24255028, -7.07521061]), np.array([1.16518879, 0.90327916])
    bp_cnt[j] = 3
    bp_xy[j, 1], bp_xy[j, 2] = c + dv, c - dv
    act, cnt, alpha = b.esv_active.copy(), b.esv_cnt.copy(), b.esv_alpha.copy()
    act[a] = 0
    act[a, :, j] = 1
    cnt[a, :, 0] = 1
    alpha[a, :, 0] = [j + 1, 5]
    out = dataclasses.replace(b, par=par, esv_active=act, esv_cnt=cnt, esv_alpha=alpha, bp_cnt=bp_cnt, bp_xy=bp_xy)
    out.validate()
    return par, out


def without_tether_rows(b):
    """The same batch with every active case cleared (no addEntangleConstraintForIJCase call)."""
    return dataclasses.replace(b, esv_active=np.zeros_like(b.esv_active), esv_cnt=np.zeros_like(b.esv_cnt))


def infeasible_batch(kind: str):
    """Replans whose two solves (direct, fallback) are both infeasible => status 2 (solver_gurobi_poly.cpp:856-859).

    'box'   : the start position lies outside the position box (bound rows on the first control point of interval 0,
              which the initial-condition equalities pin, can never hold) -- every n;
    'vel'   : the start velocity exceeds v_max (first velocity control point pinned by the equalities);
    'n2box' : n == 2 agents only (the equalities leave a single point in the direct solve), start outside the box.
    Returns (par, batch)."""
    par = config("mtlp5")
    if kind == "n2box":
        for seed in range(2002, 2040):
            sc = make_scene(par, seed, sync=False, n_fixed=2)
            if (sc.batch.n_int == 2).any():
                break
    else:
        sc = make_scene(par, 2003, sync=False)
    b = sc.batch
    ci = b.coeff_init.copy()
    if kind in ("box", "n2box"):
        ci[:, 0, :, 3] += (par.x_max - par.x_min) + 5.0     # shift the whole x polynomial out of the box
    elif kind == "vel":
        ci[:, 0, 0, 2] = 3.0 * par.v_max                       # c0 = start velocity
    else:
        raise KeyError(kind)
    out = dataclasses.replace(b, coeff_init=np.ascontiguousarray(ci))
    out.validate()
    return par, out

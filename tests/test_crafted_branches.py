"""CPU: the two branches of the back end that no generated scene reaches (tests/crafted.py), on the oracle, on the
single-lane emulation of the kernels and against HiGHS:

* non-entangling separating lines, ``addEntangleConstraintForIJCase`` (solver_gurobi_poly.cpp:620-642, :715-784);
* both solves infeasible => status 2, ``pwp_out = pwp_init`` (:856-859).
The same inputs are committed as golden fixtures (tests/golden/mtlp5_crafted-*.npz) and run on the GPU by
tests/test_gpu_edges.py."""
import numpy as np
import pytest

from neptune_b200.batch import ReplanResult
from neptune_b200.minvo import solver_basis
from neptune_b200.params import long_length
from tests import crafted
from tests.highs_util import qp_highs


def expected_ent_lps(par, b, a, i):
    """The (pointA, pointB) pairs the reference hands to the 4-argument solveModel for agent a, interval i, in the
    order of its loops -- a plain-Python reading of solver_gurobi_poly.cpp:620-642 and :715-751, written
    independently of the oracle's C and of the kernel's list walk."""
    Ainv, _, _ = solver_basis(par.T_span)
    cp = np.stack([b.coeff_init[a, 0, i] @ Ainv, b.coeff_init[a, 1, i] @ Ainv], axis=1)   # ctrlPtsInit_[i] columns
    hulldist = sum(np.linalg.norm(cp[k + 1] - cp[k]) for k in range(3))
    LL = long_length(par)
    alphas = [tuple(b.esv_alpha[a, i, q]) for q in range(int(b.esv_cnt[a, i, 0]))]
    out = []
    for j in range(par.num_of_agents):
        if j == int(b.agent_id[a]) - 1 or b.esv_active[a, i, j] != 1:
            continue
        case_id = 0
        for (aid, c) in alphas:
            if aid == j + 1:
                case_id = c
        if case_id == 0:
            continue
        bend = [b.bp_xy[j, k] for k in range(int(b.bp_cnt[j]))]
        pos = b.nih0[a, j, i]
        for k in range(1, len(bend) + 1):
            if k == case_id:
                continue
            if k == 1:
                pA, pB = (1 - LL) * bend[-1] + LL * pos, pos
            else:
                pA, pB = bend[k - 2], bend[k - 1]
            if np.linalg.norm(pA - cp[0]) - hulldist > 0 and np.linalg.norm(pB - cp[0]) - hulldist > 0:
                continue
            out.append((pA, pB))
    return cp, out


@pytest.mark.parametrize("variant", [0, 1])
def test_non_entangling_lines_oracle_emulation_highs(oracle, variant):
    from tests.emul import emul
    par, b = crafted.ent_lp_batch(variant)
    ref = ReplanResult.empty(b)
    assert oracle.replan_batch(b, ref, 1) == 0
    got = emul.replan(b)
    e0 = b.n_hull_slots + par.num_of_agents + par.num_of_static_obst
    # which LPs are attempted, and on which point sets: an independent reading of the reference's loops
    n_lp = 0
    for a in range(b.B):
        for i in range(int(b.n_int[a])):
            cp, want = expected_ent_lps(par, b, a, i)
            flags = ref.line_ok[a, i, e0:]
            assert (flags[:len(want)] > 0).all() and (flags[len(want):] == 0).all(), (a, i)
            for q, (pA, pB) in enumerate(want):
                n_lp += 1
                sep_ok, line = oracle.separate(np.stack([pA, pB]), cp)
                assert flags[q] == (1 if sep_ok else 2)
                if sep_ok:
                    l = ref.lines[a, i, e0 + q]
                    assert np.abs(l - line).max() <= 1e-12 * max(1.0, np.abs(line).max())
                    assert min(pA @ l[:2] + l[2], pB @ l[:2] + l[2]) >= 1 - 1e-9      # tether segment on the A side
                    assert (cp @ l[:2] + l[2]).max() <= -1 + 1e-9                       # my control points on the B side
    assert n_lp >= 10 and (ref.line_ok[:, :, e0:] == 1).sum() >= 8 and (ref.line_ok[:, :, e0:] == 2).sum() >= 1
    # kernels (one host lane) == oracle
    assert np.array_equal(got.line_ok, ref.line_ok) and np.array_equal(got.status, ref.status)
    m = ref.line_ok == 1
    assert np.abs(got.lines[m] - ref.lines[m]).max() <= 1e-9 * max(1.0, np.abs(ref.lines[m]).max())
    assert np.abs(got.coeff_out - ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
    # the lines matter: without them agent 0's optimum is different (the constraint is active or at least present)
    # and HiGHS on the exported rows (tether rows included) returns the same minimiser
    n_checked = 0
    for a in range(b.B):
        if ref.status[a] == 2:
            continue
        mdl = oracle.export_qp(b, a, ref.status[a] == 1, ref.lines[a], ref.line_ok[a])
        if mdl["has_qc"]:
            continue
        assert mdl["G"].shape[0] == 48 * mdl["n"] + 4 * int((ref.line_ok[a, :mdl["n"]] == 1).sum())
        st, x, f = qp_highs(mdl["P"], mdl["q"], mdl["Aeq"], mdl["beq"], mdl["G"], mdl["h"])
        assert st == "Optimal"
        n = mdl["n"]
        xo = np.concatenate([ref.coeff_out[a, ax, i] for i in range(n) for ax in range(3)])
        assert np.abs(xo - x).max() <= 1e-6 * max(1.0, np.abs(x).max())
        assert ref.obj[a] <= (f + mdl["c0"]) * (1 + 1e-6) + 1e-9
        n_checked += 1
    assert n_checked >= 2


def test_tether_rows_bind_at_the_optimum(oracle):
    """Variant 2: the tether segment lies where the unconstrained optimum goes -- with the rows the optimiser's answer
    moves by decimetres, every row holds, and HiGHS on the exported rows (tether rows included) agrees."""
    from tests.emul import emul
    par, b = crafted.ent_lp_batch(2)
    ref, plain = ReplanResult.empty(b), ReplanResult.empty(b)
    assert oracle.replan_batch(b, ref, 1) == 0
    assert oracle.replan_batch(crafted.without_tether_rows(b), plain, 1) == 0
    e0 = b.n_hull_slots + par.num_of_agents + par.num_of_static_obst
    assert (ref.line_ok[0, :, e0:] == 1).sum() >= 3 and ref.status[0] == 0 and plain.status[0] == 0
    assert np.abs(ref.coeff_out[0] - plain.coeff_out[0]).max() > 0.1
    assert ref.obj[0] > plain.obj[0] * (1 + 1e-6)            # a tighter feasible set costs something
    got = emul.replan(b)
    assert np.array_equal(got.line_ok, ref.line_ok) and np.array_equal(got.status, ref.status)
    assert np.abs(got.coeff_out - ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
    mdl = oracle.export_qp(b, 0, False, ref.lines[0], ref.line_ok[0])
    st, x, f = qp_highs(mdl["P"], mdl["q"], mdl["Aeq"], mdl["beq"], mdl["G"], mdl["h"])
    n = mdl["n"]
    xo = np.concatenate([ref.coeff_out[0, ax, i] for i in range(n) for ax in range(3)])
    assert st == "Optimal" and np.abs(xo - x).max() <= 1e-6 * max(1.0, np.abs(x).max())
    # some row is active at the optimum that is inactive at the optimum without the tether rows
    xp = np.concatenate([plain.coeff_out[0, ax, i] for i in range(n) for ax in range(3)])
    assert (mdl["G"] @ xp - mdl["h"]).max() > 1e-3            # the unconstrained optimum violates a tether row
    assert (mdl["G"] @ xo - mdl["h"]).max() <= 1e-6


@pytest.mark.parametrize("kind", ["box", "vel", "n2box"])
def test_both_solves_infeasible_is_status_2(oracle, kind):
    from tests.emul import emul
    par, b = crafted.infeasible_batch(kind)
    ref = ReplanResult.empty(b)
    assert oracle.replan_batch(b, ref, 1) == 0
    got = emul.replan(b)
    assert (ref.status == 2).all() and np.array_equal(got.status, ref.status)
    assert np.array_equal(ref.coeff_out, b.coeff_init) and np.array_equal(got.coeff_out, b.coeff_init)   # :858
    assert (ref.obj == 0).all() and (got.obj == 0).all()
    if kind == "n2box":
        assert (b.n_int == 2).any()
    for a in range(b.B):      # HiGHS: infeasible with and without the terminal v/a rows
        for fb in (False, True):
            mdl = oracle.export_qp(b, a, fb, ref.lines[a], ref.line_ok[a])
            st, _, _ = qp_highs(mdl["P"], mdl["q"], mdl["Aeq"], mdl["beq"], mdl["G"], mdl["h"])
            assert st == "Infeasible", (a, fb, st)


# ---------------------------------------------------------------------------------------------- crowded, inconsistent worlds
def _crowded_worlds(oracle, n_scenes=6):
    """Paths and entanglement states of one 64-agent grid world against the hulls of ANOTHER one (same bases, other
    goals): what a closed loop drifts into when the committed trajectories move on and the front-end paths do not.  Hulls
    nearly touch control polygons along parallel edges, first solves are infeasible in ways that are not structural
    (they run and diverge), and crowded intervals carry dozens of redundant lines."""
    import dataclasses

    import bench
    from tests.ent_backends import OracleEntBackend
    par = bench.world_params(1)
    agents = bench.rank_agents(par, 1, 0, "grid64")
    _, scenes = bench.make_world(1, 0, n_scenes, OracleEntBackend(oracle), "grid64", agents=agents)
    for a in range(n_scenes):
        for b in range(n_scenes):
            if a != b:
                A, B = scenes[a].batch, scenes[b].batch
                yield (a, b), dataclasses.replace(A, hull_ptr=B.hull_ptr, hull_xy=B.hull_xy, nih0=B.nih0)


def test_crowded_worlds_emulation_equals_oracle(oracle):
    """30 mixed worlds x 64 agents: solved flags of all ~500 LPs per agent identical, status path identical, coefficients
    within tolerance -- including the first solves that RUN and fail (divergence test) and the feasible ones on which the
    oracle's unpruned full-space model needs up to 128 iterations against 12-18 for the product's pruned one.  One of the
    latter is also checked against HiGHS: the model is feasible and the optimum is the product's."""
    from tests.emul import emul
    from tests.highs_util import qp_highs
    n_hard = n_slow = 0
    checked_highs = False
    for (a, b), mix in _crowded_worlds(oracle):
        ref = ReplanResult.empty(mix)
        assert oracle.replan_batch(mix, ref, 8) == 0
        got = emul.replan(mix, with_lines=True)
        assert np.array_equal(got.line_ok, ref.line_ok), (a, b)
        assert np.array_equal(got.status, ref.status), (a, b, np.flatnonzero(got.status != ref.status))
        assert np.abs(got.coeff_out - ref.coeff_out).max() <= 1e-6 * max(1.0, np.abs(ref.coeff_out).max()), (a, b)
        n_hard += int(((ref.status >= 1) & (ref.iters[:, 0] > 0)).sum())
        slow = np.flatnonzero((ref.status == 0) & (ref.iters[:, 0] > 30))
        n_slow += len(slow)
        if len(slow) and not checked_highs:
            w = int(slow[0])
            m = oracle.export_qp(mix, w, False, ref.lines[w], ref.line_ok[w])
            st, _, f = qp_highs(m["P"], m["q"], m["Aeq"], m["beq"], m["G"], m["h"])
            assert st == "Optimal"
            assert abs(f + m["c0"] - got.obj[w]) <= 1e-6 * max(1.0, abs(got.obj[w]))
            assert got.iters[w, 0] <= 30
            checked_highs = True
    assert n_hard >= 3 and n_slow >= 1 and checked_highs

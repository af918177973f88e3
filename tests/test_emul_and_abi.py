"""CPU: host logic, the C-ABI surface, and the device kernels compiled for one host lane
(tests/emul: a debugging harness, not a fallback) against the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from neptune_b200 import capi, config
from neptune_b200.batch import ReplanResult
from neptune_b200.scenes import make_scene
from tests.golden_util import golden_files, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    """libneptune_b200.so loads without a GPU and exports every function include/neptune_b200.h declares."""
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(capi.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "neptune_b200.h")).read()
    names = set(re.findall(r"\b(nb_[a-z_0-9]+)\s*\(", hdr))
    assert {"nb_create", "nb_replan_batch", "nb_separate_batch", "nb_entangle_predict_batch", "nb_hulls_batch", "nb_compose_records_batch",
            "nb_postcheck_batch", "nb_destroy"} <= names
    for n in sorted(names):
        assert hasattr(lib, n), n


def test_no_cpu_fallback_without_device():
    """Without a CUDA device nb_create must fail loudly (NB_ERR_NO_DEVICE), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.NbError):
        capi.Solver(config("mtlp5"))


def test_product_does_not_touch_oracle():
    """The package must never import, link or execute anything under oracle/ (or tests/emul)."""
    for root, _, files in os.walk(os.path.join(ROOT, "neptune_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(root, f)).read()
                for needle in ("from oracle", "import oracle", "liboracle", "neptune_oracle", "libnb_emul", "tests.emul"):
                    assert needle not in src, (f, needle)


@pytest.mark.parametrize("path", golden_files())
def test_emulated_kernels_match_golden(path):
    from tests.emul import emul
    par, b, z = load(path)
    got = emul.replan(b)
    assert np.array_equal(got.line_ok, z["orc_line_ok"])
    m = z["orc_line_ok"] == 1
    assert np.abs(got.lines[m] - z["orc_lines"][m]).max(initial=0) <= 1e-9 * max(1.0, np.abs(z["orc_lines"][m]).max(initial=0))
    assert np.array_equal(got.status, z["orc_status"])
    assert np.abs(got.coeff_out - z["orc_coeff"]).max() <= 1e-6 * max(1.0, np.abs(z["orc_coeff"]).max())
    assert np.abs(got.obj - z["orc_obj"]).max() <= 1e-8 * max(1.0, np.abs(z["orc_obj"]).max())


def test_emulated_pruning_is_exact(oracle):
    """Dropping redundant lines must not change the optimum: pruned vs unpruned vs oracle (all rows)."""
    from tests.emul import emul
    par = config("obst8")
    sc = make_scene(par, 3005, sync=False)
    n0, n1 = np.zeros(sc.batch.B, np.int32), np.zeros(sc.batch.B, np.int32)
    full = emul.replan(sc.batch, prune=False, n_lines=n0)
    pruned = emul.replan(sc.batch, prune=True, n_lines=n1)
    ref = ReplanResult.empty(sc.batch)
    oracle.replan_batch(sc.batch, ref, 2)
    assert (n1 <= n0).all() and n1.sum() < n0.sum()
    assert np.array_equal(full.status, ref.status) and np.array_equal(pruned.status, ref.status)
    assert np.abs(pruned.coeff_out - ref.coeff_out).max() <= 1e-6
    assert np.abs(full.coeff_out - ref.coeff_out).max() <= 1e-6


def test_emulated_entangle_hulls_postcheck(oracle):
    from tests.emul import emul
    from tests.ent_backends import EmulEntBackend, OracleEntBackend
    par = config("obst8")
    ref = make_scene(par, 3004, sync=False, ent_backend=OracleEntBackend(oracle))
    got = make_scene(par, 3004, sync=False, ent_backend=EmulEntBackend())
    for k in ("es0_cnt", "es0_alpha", "es0_beta", "es0_bend", "es0_active", "esA_cnt", "esA_alpha", "esA_beta",
              "esA_bend", "esA_active"):
        assert np.array_equal(getattr(got, k), getattr(ref, k)), k
    for k in ("esv_cnt", "esv_alpha", "esv_active"):
        assert np.array_equal(getattr(got.batch, k), getattr(ref.batch, k)), k
    assert ref.es0_cnt[:, 0].sum() > 0  # the history walk did produce signature entries
    recs = capi.make_records(ref.committed)
    delta = 2 * par.drone_radius
    out = emul.hulls(par, ref.t_start, recs, ref.known, delta)
    for bi in range(ref.batch.B):
        for j in range(par.num_of_agents):
            if not ref.known[bi, j]:
                continue
            tm, cx, cy, _ = ref.committed[j]
            for i in range(par.num_pol):
                h, h2, idx = oracle.hull_of_interval(tm, cx, cy, ref.t_start[bi] + i * par.T_span,
                                                     ref.t_start[bi] + (i + 1) * par.T_span, par.T_span, [delta] * 3)
                c = out["hull_cnt"][bi, j, i]
                assert c == len(h) and np.array_equal(out["hull_xy"][bi, j, i, :c], h)
                assert np.array_equal(out["idx"][bi, j, i], idx) and np.array_equal(out["nih0"][bi, j, i], h2[0])
    col = emul.postcheck(par, ref.batch.n_int, ref.batch.coeff_init, ref.t_start, recs, ref.known, delta)
    want = np.zeros(ref.batch.B, np.int32)
    for bi in range(ref.batch.B):
        for j in range(par.num_of_agents):
            if ref.known[bi, j]:
                tm, cx, cy, _ = ref.committed[j]
                if oracle.pwp_collides(ref.batch.coeff_init[bi], int(ref.batch.n_int[bi]), ref.t_start[bi], par.T_span,
                                       tm, cx, cy, [delta] * 3):
                    want[bi] = 1
    assert np.array_equal(col, want)


def test_scene_is_deterministic_and_valid():
    par = config("mtlp5")
    a, b = make_scene(par, 2002, sync=False), make_scene(par, 2002, sync=False)
    assert np.array_equal(a.batch.coeff_init, b.batch.coeff_init) and np.array_equal(a.batch.hull_xy, b.batch.hull_xy)
    assert a.batch.algorithmic_bytes() > 0 and (a.batch.n_int >= 1).all()


def test_emulated_entangle_random_walks_bend_points(oracle):
    from tests.ent_backends import EmulEntBackend, OracleEntBackend
    from tests.ent_walks import compare_backends
    par = config("obst8")
    sc = make_scene(par, 3003, sync=False)
    mx_a, mx_b = compare_backends(par, sc, OracleEntBackend(oracle), EmulEntBackend())
    assert mx_a >= 8 and mx_b >= 2   # the walks do exercise long words and bend points


def test_compose_records_oracle_restatement_and_emulation(oracle):
    """mu::composePieceWisePol (utils.cpp:318-402): the C oracle agrees with a list-based restatement on
    every branch, and the device source (compiled for the host) agrees with the oracle bit for bit."""
    from tests import compose_util as cu
    from tests.emul import emul
    rng = np.random.default_rng(5)
    ts, prevs, nows, outs, nps = [], [], [], [], []
    seen = set()
    for it in range(700):
        kind = cu.KINDS[it % len(cu.KINDS)]
        t, p1, p2 = cu.random_case(rng, kind)
        n, out, _, _ = oracle.compose_records(t, 0.05, p1, p2)
        times, pieces = cu.compose_lists(t, cu.rec_to(p1), cu.rec_to(p2))
        assert n == max(len(times) - 1, 0), kind
        if n > 0:
            ot, oc = cu.rec_to(out)
            assert ot == [float(x) for x in times]
            assert np.array_equal(oc, np.stack(pieces, axis=1))
            assert all(ot[i] < ot[i + 1] for i in range(n)) or kind in ("same", "knot")
        else:
            assert not out.any()
        seen.add((kind, n > 0))
        ts.append(t), prevs.append(p1), nows.append(p2), outs.append(out), nps.append(n)
    assert ("mid", True) in seen and ("stale", False) in seen and ("gap", True) in seen
    has_prev = np.ones(len(ts), np.uint8)
    has_prev[::11] = 0
    npc, eo = emul.compose(ts, has_prev, np.stack(prevs), np.stack(nows))
    for b in range(len(ts)):
        if has_prev[b]:
            assert npc[b] == nps[b] and np.array_equal(eo[b], outs[b]), b
        else:
            assert np.array_equal(eo[b], nows[b])


def test_compose_records_overflow_is_reported(oracle):
    from tests import compose_util as cu
    from tests.emul import emul
    t1 = 0.1 * np.arange(17)
    p1 = cu.rec_from(t1, np.ones((3, 16, 4)))
    p2 = cu.rec_from(1.55 + 0.5 * np.arange(9), np.ones((3, 8, 4)))
    n, _, _, _ = oracle.compose_records(0.05, 0.05, p1, p2)
    assert n == -1
    npc, _ = emul.compose([0.05], [1], p1[None], p2[None])
    assert npc[0] == -1


def test_dropin_headers_compile_beside_the_reference():
    """neptune_b200/cpp/*.hpp in ONE translation unit with the reference's own mader_types.hpp, entangle_utils.hpp,
    kinodynamic_search.hpp, solver_gurobi_poly.hpp and utils.hpp (tests/cpp/compile_against_reference.cpp), built the way
    INTEGRATION.md describes: nothing the reference defines is defined twice, the reference's solver-facing headers drop
    out through their own include guards, and every member Neptune calls resolves against the B200 classes."""
    import subprocess
    if not os.path.isdir("/root/reference/neptune/include"):
        pytest.skip("no /root/reference in this environment")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_shim_ref")
    if os.path.exists(exe):
        os.remove(exe)
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "_build/test_shim_ref"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(exe)
    syms = subprocess.run(["nm", "-C", exe], capture_output=True, text=True).stdout
    for needle in ("PolySolverGurobi::optimize", "KinodynamicSearch::run", "nb_separate_batch", "nb_replan_batch", "nb_search_batch"):
        assert needle in syms, needle
    assert "GRBModel" not in syms and "glp_" not in syms     # neither Gurobi nor GLPK is referenced any more


def test_record_wire_format_matches_reference_messages():
    """neptune_b200/cpp/records_b200.hpp against the reference's own mu::pwp2PwpMsg / mu::pwpMsg2Pwp (utils.cpp compiled
    where it lies): the trajectory part of a record is the PieceWisePolTraj message field for field, the round trip is the
    reference's, the DynTraj header survives, updateTrajObstacles replaces by id (tests/cpp/records_check.cpp)."""
    import subprocess
    if not os.path.isdir("/root/reference/neptune/include"):
        pytest.skip("no /root/reference in this environment")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "_build/records_check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = subprocess.run([os.path.join(ROOT, "tests", "cpp", "_build", "records_check")], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok 500", (out.returncode, out.stdout, out.stderr)

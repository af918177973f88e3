"""ctypes binding of tests/emul/_build/libnb_emul.so: the device kernels compiled for ONE host lane.
CPU-side debugging harness for pytest -m "not gpu"; never used by the neptune_b200 package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from neptune_b200.batch import ReplanResult
from neptune_b200.capi import host_args, make_nb_params

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libnb_emul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_LIB)
    return _lib


def replan(batch, with_lines=True, prune=True, n_lines=None) -> ReplanResult:
    par = batch.par
    res = ReplanResult.empty(batch, with_lines)
    a = host_args(batch, res)
    nbp = make_nb_params(par)
    pb = np.ascontiguousarray(par.pb, np.float64)
    rc = lib().emul_replan_batch(C.byref(nbp), pb.ctypes.data_as(C.c_void_p), batch.st_ptr.ctypes.data_as(C.c_void_p),
                                 batch.st_xy.ctypes.data_as(C.c_void_p), C.byref(a), int(prune),
                                 None if n_lines is None else n_lines.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return res


def separate(A, B, a_polygon):
    A, B = np.ascontiguousarray(A, np.float64), np.ascontiguousarray(B, np.float64)
    out = np.zeros(3)
    ok = lib().emul_separate(A.ctypes.data_as(C.c_void_p), len(A), int(a_polygon), B.ctypes.data_as(C.c_void_p), len(B),
                             out.ctypes.data_as(C.c_void_p))
    return bool(ok), out


def hulls(par, t_start, recs, known, delta):
    from neptune_b200.capi import _hulls_impl
    nbp = make_nb_params(par)

    def call(B, ts, rc, kn, dl, hx, hc, hp, n0, sm, ix):
        f = lib().emul_hulls
        f.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_double] + [C.c_void_p] * 6
        assert f(C.addressof(nbp), B, ts, rc, kn, dl, hx, hc, hp, n0, sm, ix) == 0
    return _hulls_impl(call, par, t_start, recs, known, delta)


def postcheck(par, n_int, coeff, t_start, recs, late, delta):
    nbp = make_nb_params(par)
    arrs = [np.ascontiguousarray(n_int, np.int32), np.ascontiguousarray(coeff, np.float64),
            np.ascontiguousarray(t_start, np.float64), np.ascontiguousarray(recs, np.float64),
            np.ascontiguousarray(late, np.uint8)]
    col = np.zeros(len(arrs[0]), np.int32)
    f = lib().emul_postcheck
    f.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_double, C.c_void_p]
    f(C.addressof(nbp), len(arrs[0]), *[a.ctypes.data_as(C.c_void_p) for a in arrs], delta, col.ctypes.data_as(C.c_void_p))
    return col


def compose(t, has_prev, prev, now):
    arrs = [np.ascontiguousarray(t, np.float64), np.ascontiguousarray(has_prev, np.uint8),
            np.ascontiguousarray(prev, np.float64), np.ascontiguousarray(now, np.float64)]
    out, npc = np.zeros_like(arrs[3]), np.zeros(len(arrs[0]), np.int32)
    f = lib().emul_compose
    f.argtypes = [C.c_int] + [C.c_void_p] * 6
    f(len(arrs[0]), *[a.ctypes.data_as(C.c_void_p) for a in arrs], out.ctypes.data_as(C.c_void_p), npc.ctypes.data_as(C.c_void_p))
    return npc, out


def search(sb):
    """K0 front-end search, the device code on one host lane."""
    from neptune_b200.capi import host_search_args, make_search_params
    from neptune_b200.search import SearchResult
    par = sb.par
    res = SearchResult.empty(sb)
    a = host_search_args(sb, res)
    nbp, sp = make_nb_params(par), make_search_params(par)
    pb = np.ascontiguousarray(par.pb, np.float64)
    M = par.num_of_static_obst
    strep = np.ascontiguousarray(sb.strep, np.float64) if M else np.zeros((1, 2, 2))
    longest = np.ascontiguousarray(sb.st_longest, np.float64) if M else np.zeros((1, 2))
    st_xy = np.ascontiguousarray(sb.st_xy, np.float64) if M else np.zeros((1, 2))
    f = lib().emul_search_batch
    f.argtypes = [C.c_void_p] * 8
    rc = f(C.addressof(nbp), C.addressof(sp), pb.ctypes.data, sb.st_ptr.ctypes.data, st_xy.ctypes.data, strep.ctypes.data,
           longest.ctypes.data, C.addressof(a))
    assert rc == 0, rc
    return res


def track(par, strep, agent_id, bp_cnt, bp_xy, bp_cnt_prev, bp_xy_prev, state, prev_pos, prev_pos_agent, latest, cur, elapsed_ms):
    """k_entangle mode 3 (online tracker tick) on one host lane; same return convention as Solver.entangle_track."""
    nbp = make_nb_params(par)
    st = state.copy()
    pp, ppa = np.ascontiguousarray(prev_pos, np.float64).copy(), np.ascontiguousarray(prev_pos_agent, np.float64).copy()
    arrs = [np.ascontiguousarray(par.pb, np.float64),
            np.ascontiguousarray(strep, np.float64) if par.num_of_static_obst else np.zeros((1, 2, 2)),
            np.ascontiguousarray(agent_id, np.int32), np.ascontiguousarray(bp_cnt, np.int32), np.ascontiguousarray(bp_xy, np.float64),
            np.ascontiguousarray(bp_cnt_prev, np.int32), np.ascontiguousarray(bp_xy_prev, np.float64),
            np.ascontiguousarray(latest, np.float64), np.ascontiguousarray(cur, np.float64), np.ascontiguousarray(elapsed_ms, np.float64)]
    B = len(arrs[2])
    res = np.zeros(B, np.int32)
    from neptune_b200.capi import NbEntState
    f = lib().emul_track_batch
    P = C.c_void_p
    f.argtypes = [P, P, P, C.c_int, P, P, P, P, P, NbEntState, P, P, P, P, P, P]
    d = lambda x: x.ctypes.data_as(P)  # noqa: E731
    rc = f(C.addressof(nbp), d(arrs[0]), d(arrs[1]), B, d(arrs[2]), d(arrs[3]), d(arrs[4]), d(arrs[5]), d(arrs[6]), st.c(),
           d(pp), d(ppa), d(arrs[7]), d(arrs[8]), d(arrs[9]), d(res))
    assert rc == 0, rc
    return res, st, pp, ppa


def postcheck_entangle(par, strep, agent_id, known, late, bp_cnt, bp_xy, bp_cnt_late, bp_xy_late, state, prev_pos, prev_pos_agent,
                       cur, n_int, coeff, t_start, samp, late_recs):
    """k_entangle mode 4 (entanglement half of Neptune::safetyCheckAfterReplan) on one host lane: entangled [B]."""
    nbp = make_nb_params(par)
    arrs = [np.ascontiguousarray(x, dt) for x, dt in
            ((par.pb, np.float64), (strep if par.num_of_static_obst else np.zeros((1, 2, 2)), np.float64), (agent_id, np.int32),
             (known, np.uint8), (late, np.uint8), (bp_cnt, np.int32), (bp_xy, np.float64), (bp_cnt_late, np.int32),
             (bp_xy_late, np.float64), (prev_pos, np.float64), (prev_pos_agent, np.float64), (cur, np.float64),
             (n_int, np.int32), (coeff, np.float64), (t_start, np.float64), (samp, np.float64), (late_recs, np.float64))]
    B = len(arrs[2])
    ent = np.zeros(B, np.int32)
    from neptune_b200.capi import NbEntState
    f = lib().emul_postcheck_entangle
    P = C.c_void_p
    f.argtypes = [P, P, P, C.c_int] + [P] * 7 + [NbEntState] + [P] * 9
    d = [a.ctypes.data_as(P) for a in arrs]
    rc = f(C.addressof(nbp), d[0], d[1], B, d[2], d[3], d[4], d[5], d[6], d[7], d[8], state.c(), d[9], d[10], d[11], d[12], d[13],
           d[14], d[15], d[16], ent.ctypes.data_as(P))
    assert rc == 0, rc
    return ent

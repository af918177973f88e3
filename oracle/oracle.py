"""ctypes binding of the CPU ORACLE (oracle/neptune_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from neptune_b200/.  Parity status: the entanglement
chain and GJK are pinned against the reference's own sources (oracle/_ref, tests/test_reference_pin.py); the LP/QP are
"parity unpinned" against Gurobi/GLPK (not installable here) and pinned by HiGHS and known answers in tests/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def _cpu_tag() -> str:
    """Identifies the instruction set of this host: the library is built -march=native (the CPU arm is
    compiled for the machine it is timed on), so a copy built on another CPU model must be rebuilt, not loaded."""
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        model = next((ln.split(":", 1)[1].strip() for ln in txt.splitlines() if ln.startswith("model name")), "?")
        flags = next((ln.split(":", 1)[1] for ln in txt.splitlines() if ln.startswith("flags")), "")
        return model + " | " + " ".join(sorted(x for x in flags.split() if x.startswith(("avx", "fma", "bmi", "sse4"))))
    except OSError:
        return "unknown"


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("neptune_oracle.c", "neptune_search.c", "neptune_oracle.h", "Makefile")]
    tag_path = os.path.join(_HERE, "_build", "built_on.txt")
    tag = _cpu_tag()
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in srcs)
    other_host = (not os.path.exists(tag_path)) or open(tag_path).read() != tag
    if force or stale or other_host:
        if os.path.exists(_LIB_PATH):
            os.remove(_LIB_PATH)
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        with open(tag_path, "w") as f:
            f.write(tag)
    return _LIB_PATH


class OrcParams(C.Structure):
    _fields_ = [("num_pol", C.c_int), ("num_agents", C.c_int), ("num_static", C.c_int), ("samples", C.c_int),
                ("T_span", C.c_double), ("weight", C.c_double), ("lim_min", C.c_double * 3),
                ("lim_max", C.c_double * 3), ("v_max", C.c_double), ("a_max", C.c_double),
                ("drone_radius", C.c_double), ("ent_cap", C.c_int), ("bp_max", C.c_int), ("ent_slots", C.c_int),
                ("ipm_max_iter", C.c_int), ("ipm_tol", C.c_double)]


_P = C.c_void_p


class OrcReplanIn(C.Structure):
    _fields_ = [("agent_id", C.c_int), ("n", C.c_int), ("coeff_init", _P), ("n_hull_slots", C.c_int),
                ("hull_ptr", _P), ("hull_xy", _P), ("nih0", _P), ("st_ptr", _P), ("st_xy", _P),
                ("esv_cnt", _P), ("esv_alpha", _P), ("esv_active", _P), ("bp_cnt", _P), ("bp_xy", _P), ("pb", _P)]


class OrcBatch(C.Structure):
    _fields_ = [("B", C.c_int), ("agent_id", _P), ("n_int", _P), ("coeff_init", _P), ("n_hull_slots", C.c_int),
                ("hull_ptr", _P), ("hull_xy", _P), ("nih0", _P), ("st_ptr", _P), ("st_xy", _P), ("esv_cnt", _P),
                ("esv_alpha", _P), ("esv_active", _P), ("bp_shared", C.c_int), ("bp_cnt", _P), ("bp_xy", _P),
                ("pb", _P), ("coeff_out", _P), ("obj", _P), ("status", _P), ("iters", _P), ("lines", _P),
                ("line_ok", _P)]


class OrcEnt(C.Structure):
    _fields_ = [("n_alpha", C.c_int), ("n_bend", C.c_int), ("alpha", _P), ("beta", _P), ("bend", _P),
                ("active", _P)]


class OrcECtx(C.Structure):
    _fields_ = [("N", C.c_int), ("M", C.c_int), ("self", C.c_int), ("cap", C.c_int), ("pb", _P), ("strep", _P),
                ("bp_cnt", _P), ("bp_xy", _P), ("bp_max", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_separate.restype = C.c_int
        _lib.orc_lp_separable.restype = C.c_int
        _lib.orc_convex_hull_2d.restype = C.c_int
        _lib.orc_gjk_collision.restype = C.c_int
        _lib.orc_replan_batch.restype = C.c_int
        _lib.orc_export_qp.restype = C.c_int
        _lib.orc_generate_traj.restype = C.c_int
        _lib.orc_predict.restype = C.c_int
        _lib.orc_entangle_check_pwp.restype = C.c_int
        _lib.orc_entangle_rollout.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_P)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def make_params(par) -> OrcParams:
    op = OrcParams()
    op.num_pol, op.num_agents, op.num_static = par.num_pol, par.num_of_agents, par.num_of_static_obst
    op.samples, op.T_span, op.weight = par.num_sample_per_interval, par.T_span, par.weight
    op.lim_min[:] = [par.x_min, par.y_min, par.z_min]
    op.lim_max[:] = [par.x_max, par.y_max, par.z_max]
    op.v_max, op.a_max, op.drone_radius = par.v_max, par.a_max, par.drone_radius
    op.ent_cap, op.bp_max, op.ent_slots = par.ent_cap, par.bp_max, par.ent_slots
    op.ipm_max_iter, op.ipm_tol = par.ipm_max_iter, par.ipm_tol
    return op


# --------------------------------------------------------------------------- primitives
def basis(T: float):
    Ainv, V, A01 = np.zeros(16), np.zeros(9), np.zeros(16)
    lib().orc_basis(C.c_double(T), _p(Ainv), _p(V), _p(A01))
    return Ainv.reshape(4, 4), V.reshape(3, 3), A01.reshape(4, 4)


def separate(A, B):
    A, B = _c(A, np.float64), _c(B, np.float64)
    out = np.zeros(3)
    ok = lib().orc_separate(_p(A), C.c_int(A.shape[0]), _p(B), C.c_int(B.shape[0]), _p(out))
    return bool(ok), out


def lp_separable(A, B) -> bool:
    A, B = _c(A, np.float64), _c(B, np.float64)
    return bool(lib().orc_lp_separable(_p(A), C.c_int(A.shape[0]), _p(B), C.c_int(B.shape[0]), C.c_int(A.shape[1])))


def convex_hull(pts):
    pts = _c(pts, np.float64)
    out = np.zeros((max(pts.shape[0], 1), 2))
    n = lib().orc_convex_hull_2d(_p(pts), C.c_int(pts.shape[0]), _p(out))
    return out[:n].copy()


def hull_of_interval(times, cx, cy, t_start, t_end, T_span, delta):
    times, cx, cy = _c(times, np.float64), _c(cx, np.float64), _c(cy, np.float64)
    delta = _c(delta, np.float64)
    hull, hull2 = np.zeros((ORC_HMAX, 2)), np.zeros((ORC_HMAX, 2))
    hn, h2n, idx = C.c_int(0), C.c_int(0), np.zeros(2, np.int32)
    lib().orc_hull_of_interval(_p(times), C.c_int(times.shape[0]), _p(cx), _p(cy), C.c_double(t_start),
                               C.c_double(t_end), C.c_double(T_span), _p(delta), _p(hull), C.byref(hn), _p(hull2),
                               C.byref(h2n), _p(idx))
    return hull[:hn.value].copy(), hull2[:h2n.value].copy(), idx


ORC_HMAX = 64


def sample_points(times, cx, cy, t_start, t_end, num_pol, S):
    times, cx, cy = _c(times, np.float64), _c(cx, np.float64), _c(cy, np.float64)
    out = np.zeros((num_pol, S + 1, 2))
    idx = np.zeros((num_pol, S + 1), np.int32)
    lib().orc_sample_interval_points(_p(times), C.c_int(times.shape[0]), _p(cx), _p(cy), C.c_double(t_start),
                                     C.c_double(t_end), C.c_int(num_pol), C.c_int(S), _p(out), _p(idx))
    return out, idx


def gjk_collision(v1, v2) -> bool:
    v1, v2 = _c(v1, np.float64), _c(v2, np.float64)
    return bool(lib().orc_gjk_collision(_p(v1), C.c_int(v1.shape[0]), _p(v2), C.c_int(v2.shape[0])))


def pwp_collides(coeff, n, t_start, T_span, times, cx, cy, delta) -> bool:
    coeff, times, cx, cy = _c(coeff, np.float64), _c(times, np.float64), _c(cx, np.float64), _c(cy, np.float64)
    delta = _c(delta, np.float64)
    f = lib().orc_pwp_collides
    f.restype = C.c_int
    return bool(f(_p(coeff), C.c_int(n), C.c_double(t_start), C.c_double(T_span), _p(times), C.c_int(times.shape[0]),
                  _p(cx), _p(cy), _p(delta)))


def generate_traj(coeff, n, T, dc):
    coeff = _c(coeff, np.float64)
    mx = int(n * T / dc) + 8
    st = np.zeros((mx, 12))
    k = lib().orc_generate_traj(_p(coeff), C.c_int(n), C.c_double(T), C.c_double(dc), _p(st), C.c_int(mx))
    return st[:k].copy()


# --------------------------------------------------------------------------- entangle chain
class EntState:
    """eu::ent_state (entangle_utils.hpp:23-29) with fixed storage capacity."""

    def __init__(self, cap, NA):
        self.cap, self.NA = cap, NA
        self.alpha = np.zeros((cap, 2), np.int32)
        self.beta = np.zeros(cap)
        self.bend = np.zeros(cap, np.int32)
        self.active = np.zeros(NA, np.int32)
        self.n_alpha = 0
        self.n_bend = 0

    def copy(self):
        o = EntState(self.cap, self.NA)
        o.alpha[:], o.beta[:], o.bend[:], o.active[:] = self.alpha, self.beta, self.bend, self.active
        o.n_alpha, o.n_bend = self.n_alpha, self.n_bend
        return o

    def _c(self):
        e = OrcEnt()
        e.n_alpha, e.n_bend = self.n_alpha, self.n_bend
        e.alpha, e.beta, e.bend, e.active = _p(self.alpha), _p(self.beta), _p(self.bend), _p(self.active)
        return e

    def _back(self, e):
        self.n_alpha, self.n_bend = e.n_alpha, e.n_bend


class EntCtx:
    def __init__(self, par, self_idx, strep, bp_cnt, bp_xy):
        self.par = par
        self.pb = _c(par.pb, np.float64)
        self.strep = _c(strep, np.float64).reshape(-1, 2, 2) if par.num_of_static_obst else np.zeros((1, 2, 2))
        self.bp_cnt = _c(bp_cnt, np.int32)
        self.bp_xy = _c(bp_xy, np.float64)
        c = OrcECtx()
        c.N, c.M, c.self, c.cap = par.num_of_agents, par.num_of_static_obst, self_idx, par.ent_cap
        c.pb, c.strep, c.bp_cnt, c.bp_xy, c.bp_max = _p(self.pb), _p(self.strep), _p(self.bp_cnt), _p(self.bp_xy), par.bp_max
        self.c = c


def predict(es: EntState, cx: EntCtx, prev_pos, prev_pos_agent, cur, samp0, known) -> int:
    prev_pos, prev_pos_agent = _c(prev_pos, np.float64), _c(prev_pos_agent, np.float64)
    cur, samp0, known = _c(cur, np.float64), _c(samp0, np.float64), _c(known, np.uint8)
    e = es._c()
    rc = lib().orc_predict(C.byref(e), C.byref(cx.c), _p(prev_pos), _p(prev_pos_agent), _p(cur), _p(samp0), _p(known))
    es._back(e)
    return rc


def track(es: EntState, cx: EntCtx, bp_cnt_prev, bp_xy_prev, prev_pos, prev_pos_agent, latest, cur, elapsed_ms: float) -> int:
    """orc_track: one tick of the online tracker; es, prev_pos, prev_pos_agent are updated in place (float64 arrays)."""
    bp_cnt_prev, bp_xy_prev = _c(bp_cnt_prev, np.int32), _c(bp_xy_prev, np.float64)
    latest, cur = _c(latest, np.float64), _c(cur, np.float64)
    assert prev_pos.dtype == np.float64 and prev_pos.flags["C_CONTIGUOUS"] and prev_pos_agent.flags["C_CONTIGUOUS"]
    e = es._c()
    f = lib().orc_track
    f.restype = C.c_int
    rc = f(C.byref(e), C.byref(cx.c), _p(bp_cnt_prev), _p(bp_xy_prev), _p(prev_pos), _p(prev_pos_agent), _p(latest), _p(cur),
           C.c_double(elapsed_ms))
    es._back(e)
    return rc


def entangle_check_pwp(es: EntState, cx: EntCtx, n, cxy, samp, known) -> int:
    par = cx.par
    cxy, samp, known = _c(cxy, np.float64), _c(samp, np.float64), _c(known, np.uint8)
    e = es._c()
    r = lib().orc_entangle_check_pwp(C.byref(e), C.byref(cx.c), C.c_int(n), _p(cxy), _p(samp), _p(known),
                                     C.c_int(par.num_pol), C.c_int(par.num_sample_per_interval), C.c_double(par.T_span))
    es._back(e)
    return r


def entangle_rollout(es0: EntState, cx: EntCtx, n, cxy, samp, known):
    par = cx.par
    cap, NA = par.ent_cap, par.NA
    cxy, samp, known = _c(cxy, np.float64), _c(samp, np.float64), _c(known, np.uint8)
    cnt = np.zeros((n + 1, 2), np.int32)
    alpha = np.zeros((n + 1, cap, 2), np.int32)
    beta = np.zeros((n + 1, cap))
    bend = np.zeros((n + 1, cap), np.int32)
    active = np.zeros((n + 1, NA), np.int32)
    e = es0._c()
    done = lib().orc_entangle_rollout(C.byref(e), C.byref(cx.c), C.c_int(n), _p(cxy), _p(samp), _p(known),
                                      C.c_int(par.num_pol), C.c_int(par.num_sample_per_interval),
                                      C.c_double(par.T_span), _p(cnt), _p(alpha), _p(beta), _p(bend), _p(active))
    return done, cnt, alpha, beta, bend, active


# --------------------------------------------------------------------------- back end
def replan_batch(batch, result, nthreads: int = 1) -> int:
    """orc_replan_batch over a neptune_b200.batch.ReplanBatch; fills a ReplanResult."""
    par = batch.par
    op = make_params(par)
    pb = _c(par.pb, np.float64)
    b = OrcBatch()
    b.B = batch.B
    b.agent_id, b.n_int, b.coeff_init = _p(batch.agent_id), _p(batch.n_int), _p(batch.coeff_init)
    b.n_hull_slots = batch.n_hull_slots
    b.hull_ptr, b.hull_xy, b.nih0 = _p(batch.hull_ptr), _p(batch.hull_xy), _p(batch.nih0)
    b.st_ptr, b.st_xy = _p(batch.st_ptr), _p(batch.st_xy)
    b.esv_cnt, b.esv_alpha, b.esv_active = _p(batch.esv_cnt), _p(batch.esv_alpha), _p(batch.esv_active)
    b.bp_shared, b.bp_cnt, b.bp_xy, b.pb = 1, _p(batch.bp_cnt), _p(batch.bp_xy), _p(pb)
    b.coeff_out, b.obj, b.status, b.iters = _p(result.coeff_out), _p(result.obj), _p(result.status), _p(result.iters)
    b.lines, b.line_ok = _p(result.lines), _p(result.line_ok)
    return lib().orc_replan_batch(C.byref(op), C.byref(b), C.c_int(nthreads))


def export_qp(batch, a: int, fallback: bool, lines, line_ok):
    """Dense full-space QP (P, q, c0, Aeq, beq, G, h, has_qc) of agent `a` for HiGHS cross-checks."""
    par = batch.par
    op = make_params(par)
    pb = _c(par.pb, np.float64)
    N, NH, cap, NA = par.num_of_agents, batch.n_hull_slots, par.ent_cap, par.NA
    LS = batch.line_slots
    n = int(batch.n_int[a])
    nv = 12 * n
    hp = _c(batch.hull_ptr[a * NH * 8:(a + 1) * NH * 8 + 1], np.int64)
    i = OrcReplanIn()
    keep = [_c(batch.coeff_init[a], np.float64), _c(batch.nih0[a], np.float64), _c(batch.esv_cnt[a], np.int32),
            _c(batch.esv_alpha[a], np.int32), _c(batch.esv_active[a], np.int32)]
    i.agent_id, i.n, i.coeff_init, i.n_hull_slots = int(batch.agent_id[a]), n, _p(keep[0]), NH
    i.hull_ptr, i.hull_xy, i.nih0 = _p(hp), _p(batch.hull_xy), _p(keep[1])
    i.st_ptr, i.st_xy = _p(batch.st_ptr), _p(batch.st_xy)
    i.esv_cnt, i.esv_alpha, i.esv_active = _p(keep[2]), _p(keep[3]), _p(keep[4])
    i.bp_cnt, i.bp_xy, i.pb = _p(batch.bp_cnt), _p(batch.bp_xy), _p(pb)
    lines, line_ok = _c(lines, np.float64), _c(line_ok, np.uint8)
    max_rows = 48 * n + 4 * int((line_ok == 1).sum()) + 8
    P, q, c0 = np.zeros((nv, nv)), np.zeros(nv), C.c_double(0)
    Aeq, beq, n_eq = np.zeros((9 * n + 6, nv)), np.zeros(9 * n + 6), C.c_int(0)
    G, h, has_qc = np.zeros((max_rows, nv)), np.zeros(max_rows), C.c_int(0)
    m = lib().orc_export_qp(C.byref(op), C.byref(i), C.c_int(int(fallback)), _p(lines), _p(line_ok), C.c_int(LS),
                            _p(P), _p(q), C.byref(c0), _p(Aeq), _p(beq), C.byref(n_eq), _p(G), _p(h),
                            C.c_int(max_rows), C.byref(has_qc))
    assert m >= 0
    ne = n_eq.value
    return dict(P=P, q=q, c0=c0.value, Aeq=Aeq[:ne].copy(), beq=beq[:ne].copy(), G=G[:m].copy(), h=h[:m].copy(),
                has_qc=bool(has_qc.value), n=n)


def cycle_batch(scene, recs, nthreads: int = 1, do_entangle: bool = True, t_now=None, late=None, late_recs=None,
                bp_cnt_late=None, bp_xy_late=None):
    """orc_cycle_batch over a neptune_b200.scenes.Scene: the whole per-agent cycle on the CPU.  late [B][N] / late_recs /
    bp_*_late: what arrives during the optimisation (default: every known trajectory, unchanged)."""
    b, par = scene.batch, scene.par
    op = make_params(par)
    B = b.B
    strep = _c(scene.strep, np.float64) if par.num_of_static_obst else np.zeros((1, 2, 2))
    arrs = dict(agent_id=_c(b.agent_id, np.int32), n_int=_c(b.n_int, np.int32), coeff_init=_c(b.coeff_init, np.float64),
                t_start=_c(scene.t_start, np.float64), recs=_c(recs, np.float64), known=_c(scene.known, np.uint8),
                pb=_c(par.pb, np.float64), st_ptr=_c(b.st_ptr, np.int64), st_xy=_c(b.st_xy, np.float64), strep=strep,
                bp_cnt=_c(b.bp_cnt, np.int32), bp_xy=_c(b.bp_xy, np.float64), esv_cnt=_c(b.esv_cnt, np.int32),
                esv_alpha=_c(b.esv_alpha, np.int32), esv_active=_c(b.esv_active, np.int32),
                es_cnt=_c(scene.es0_cnt, np.int32), es_alpha=_c(scene.es0_alpha, np.int32),
                es_beta=_c(scene.es0_beta, np.float64), es_bend=_c(scene.es0_bend, np.int32),
                es_active=_c(scene.es0_active, np.int32), prev_pos=_c(scene.prev_pos, np.float64),
                prev_pos_agent=_c(scene.prev_pos_agent, np.float64), cur=_c(scene.state_A[:, 0, :2], np.float64))
    out = dict(coeff_out=np.zeros((B, 3, 8, 4)), obj=np.zeros(B), status=np.zeros(B, np.int32),
               iters=np.zeros((B, 2), np.int32), entangled=np.zeros(B, np.int32), collide=np.zeros(B, np.int32))
    f = lib().orc_cycle_batch
    f.restype = C.c_int
    order = ("agent_id", "n_int", "coeff_init", "t_start", "recs", "known", "pb", "st_ptr", "st_xy", "strep", "bp_cnt",
             "bp_xy", "esv_cnt", "esv_alpha", "esv_active", "es_cnt", "es_alpha", "es_beta", "es_bend", "es_active",
             "prev_pos", "prev_pos_agent", "cur")
    rc = f(C.byref(op), C.c_int(B), *[_p(arrs[k]) for k in order], C.c_double(2 * par.drone_radius),
           C.c_int(int(do_entangle)), _p(out["coeff_out"]), _p(out["obj"]), _p(out["status"]), _p(out["iters"]),
           _p(out["entangled"]), _p(out["collide"]), C.c_int(nthreads),
           None if late is None else _p(_c(late, np.uint8)), None if late_recs is None else _p(_c(late_recs, np.float64)),
           None if bp_cnt_late is None else _p(_c(bp_cnt_late, np.int32)),
           None if bp_xy_late is None else _p(_c(bp_xy_late, np.float64)))
    if rc == 0 and t_now is not None:
        out["new_recs"], out["new_pieces"] = np.zeros((B, 256)), np.zeros(B, np.int32)
        g = lib().orc_commit_compose_batch
        g.restype = C.c_int
        rc = g(C.byref(op), C.c_int(B), _p(arrs["agent_id"]), _p(arrs["n_int"]), _p(out["coeff_out"]), _p(arrs["t_start"]),
               _p(_c(t_now, np.float64)), _p(arrs["recs"]), _p(out["status"]), _p(out["entangled"]), _p(out["collide"]),
               _p(out["new_recs"]), _p(out["new_pieces"]))
    return rc, out


def compose_records(t, dc, p1, p2):
    """mu::composePieceWisePol on records; returns (n_pieces, out, p1_modified, p2_modified)."""
    p1, p2 = _c(p1, np.float64).copy(), _c(p2, np.float64).copy()
    out = np.zeros_like(p1)
    f = lib().orc_compose_records
    f.restype = C.c_int
    n = f(C.c_double(t), C.c_double(dc), _p(p1), _p(p2), _p(out))
    return n, out, p1, p2


# --------------------------------------------------------------------------- front end (neptune_search.c)
class OrcSearchPar(C.Structure):
    _fields_ = [("num_pol", C.c_int), ("N", C.c_int), ("M", C.c_int), ("S", C.c_int), ("T", C.c_double),
                ("x_min", C.c_double), ("x_max", C.c_double), ("y_min", C.c_double), ("y_max", C.c_double),
                ("v_max", C.c_double), ("a_max", C.c_double), ("j_max", C.c_double), ("num_samples", C.c_int),
                ("voxel_size", C.c_double), ("bias", C.c_double), ("goal_size", C.c_double), ("tether", C.c_double),
                ("enable_entangle", C.c_int), ("use_not_reaching", C.c_int), ("max_nodes", C.c_int),
                ("max_expansions", C.c_int), ("ecap", C.c_int), ("out_cap", C.c_int), ("bp_max", C.c_int)]


class OrcSearchBatch(C.Structure):
    _fields_ = [("B", C.c_int), ("agent_id", _P), ("init", _P), ("goal", _P), ("coeffs_z", _P), ("group", _P),
                ("hull_xy", _P), ("hull_cnt", _P), ("samp", _P), ("known", _P), ("st_ptr", _P), ("st_xy", _P),
                ("strep", _P), ("st_longest", _P), ("pb", _P), ("bp_cnt", _P), ("bp_xy", _P), ("es_cap", C.c_int),
                ("es_cnt", _P), ("es_alpha", _P), ("es_beta", _P), ("es_bend", _P), ("es_active", _P),
                ("comb_shared", C.c_int), ("comb", _P), ("status", _P), ("solved", _P), ("n_int", _P), ("coeff", _P),
                ("esv_cnt", _P), ("esv_alpha", _P), ("esv_beta", _P), ("esv_bend", _P), ("esv_active", _P),
                ("stats", _P), ("cost", _P)]


def make_search_par(par) -> OrcSearchPar:
    sp = OrcSearchPar()
    sp.num_pol, sp.N, sp.M, sp.S, sp.T = par.num_pol, par.num_of_agents, par.num_of_static_obst, par.num_sample_per_interval, par.T_span
    sp.x_min, sp.x_max, sp.y_min, sp.y_max = par.x_min, par.x_max, par.y_min, par.y_max
    sp.v_max, sp.a_max, sp.j_max, sp.num_samples = par.v_max, par.a_max, par.j_max, par.a_star_samp_x
    sp.voxel_size, sp.bias, sp.goal_size, sp.tether = par.a_star_fraction_voxel_size, par.a_star_bias, par.goal_radius, par.tetherLength
    sp.enable_entangle, sp.use_not_reaching = int(par.enable_entangle_check), int(par.use_not_reaching_soln)
    sp.max_nodes, sp.max_expansions, sp.ecap = par.search_max_nodes, par.search_max_expansions, par.search_ecap
    sp.out_cap, sp.bp_max = par.ent_cap, par.bp_max
    return sp


def search_batch_struct(sb, res):
    """(orc_search_par, orc_search_batch_t, keep-alive) for a neptune_b200.search.SearchBatch / SearchResult pair."""
    par = sb.par
    sp = make_search_par(par)
    M = par.num_of_static_obst
    keep = dict(pb=_c(par.pb, np.float64), hull_cnt=_c(sb.hull_cnt, np.int32),
                strep=_c(sb.strep, np.float64) if M else np.zeros((1, 2, 2)),
                longest=_c(sb.st_longest, np.float64) if M else np.zeros((1, 2)),
                st_ptr=_c(sb.st_ptr, np.int64), st_xy=_c(sb.st_xy, np.float64) if M else np.zeros((1, 2)))
    b = OrcSearchBatch()
    b.B = sb.B
    b.agent_id, b.init, b.goal, b.coeffs_z, b.group = _p(sb.agent_id), _p(sb.init), _p(sb.goal), _p(sb.coeffs_z), _p(sb.group)
    b.hull_xy, b.hull_cnt, b.samp, b.known = _p(sb.hull_xy), _p(keep["hull_cnt"]), _p(sb.samp), _p(sb.known)
    b.st_ptr, b.st_xy, b.strep, b.st_longest = _p(keep["st_ptr"]), _p(keep["st_xy"]), _p(keep["strep"]), _p(keep["longest"])
    b.pb, b.bp_cnt, b.bp_xy = _p(keep["pb"]), _p(sb.bp_cnt), _p(sb.bp_xy)
    b.es_cap = par.ent_cap
    b.es_cnt, b.es_alpha, b.es_beta, b.es_bend, b.es_active = _p(sb.es_cnt), _p(sb.es_alpha), _p(sb.es_beta), _p(sb.es_bend), _p(sb.es_active)
    b.comb_shared, b.comb = int(sb.comb.ndim == 1), _p(sb.comb)
    b.status, b.solved, b.n_int, b.coeff = _p(res.status), _p(res.solved), _p(res.n_int), _p(res.coeff)
    b.esv_cnt, b.esv_alpha, b.esv_beta, b.esv_bend, b.esv_active = _p(res.esv_cnt), _p(res.esv_alpha), _p(res.esv_beta), _p(res.esv_bend), _p(res.esv_active)
    b.stats, b.cost = _p(res.stats), _p(res.cost)
    return sp, b, keep


def search_batch(sb, res, nthreads: int = 1) -> int:
    """orc_search_batch over a neptune_b200.search.SearchBatch; fills a SearchResult."""
    sp, b, keep = search_batch_struct(sb, res)
    f = lib().orc_search_batch
    f.restype = C.c_int
    return f(C.byref(sp), C.byref(b), C.c_int(nthreads))

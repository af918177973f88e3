"""ctypes access to oracle/_ref/libneptune_ref.so: the REFERENCE's own entangle_utils.cpp and gjk.cpp compiled where
they lie against the Eigen stand-in (oracle/Makefile target _ref).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libneptune_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB) or os.path.isdir("/root/reference/neptune/src")


def lib():
    global _lib
    if _lib is None:
        if os.path.isdir("/root/reference/neptune/src"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref"], check=True, capture_output=True)
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def gjk(v1, v2) -> bool:
    v1, v2 = _c(v1, np.float64), _c(v2, np.float64)
    return bool(lib().ref_gjk_collision(_p(v1), len(v1), _p(v2), len(v2)))


def hsig_agent(pk, pk1, pik, pik1, pb, bend, agent_id, prev=None):
    arrs = [_c(x, np.float64) for x in (pk, pk1, pik, pik1, pb, bend)]
    out = np.zeros((64, 2), np.int32)
    if prev is None:
        n = lib().ref_hsig_agent(_p(out), 64, *[_p(a) for a in arrs], len(arrs[5]), agent_id)
    else:
        pv = _c(prev, np.float64)
        n = lib().ref_hsig_agent9(_p(out), 64, *[_p(a) for a in arrs], len(arrs[5]), _p(pv), len(pv), agent_id)
    return out[:n].copy()


def hsig_static(pk, pk1, strep, N):
    pk, pk1, strep = _c(pk, np.float64), _c(pk1, np.float64), _c(strep, np.float64)
    out = np.zeros((len(strep) + 4, 2), np.int32)
    n = lib().ref_hsig_static(_p(out), len(out), _p(pk), _p(pk1), _p(strep), len(strep), N)
    return out[:n].copy()


def chain(par, self_idx, strep, longest, bp_cnt, bp_xy, known, samp, n, cxy, cnt0, alpha0, beta0, bend0, active0):
    """The reference's chain along a path: returns (done, cnt [n+1][2], alpha, beta, bend, active, tether lengths [n])."""
    N, M, NA, cap = par.num_of_agents, par.num_of_static_obst, par.NA, par.ent_cap
    S = par.num_sample_per_interval
    arrs = dict(pb=_c(par.pb, np.float64), strep=_c(strep, np.float64) if M else np.zeros((1, 2, 2)),
                longest=_c(longest, np.float64) if M else np.zeros((1, 2)), bp_cnt=_c(bp_cnt, np.int32),
                bp_xy=_c(bp_xy, np.float64), known=_c(known, np.uint8), samp=_c(samp, np.float64), cxy=_c(cxy, np.float64),
                cnt0=_c(cnt0, np.int32), alpha0=_c(alpha0, np.int32), beta0=_c(beta0, np.float64), bend0=_c(bend0, np.int32),
                active0=_c(active0, np.int32))
    cnt, alpha = np.zeros((n + 1, 2), np.int32), np.zeros((n + 1, cap, 2), np.int32)
    beta, bend = np.zeros((n + 1, cap)), np.zeros((n + 1, cap), np.int32)
    active, length = np.zeros((n + 1, NA), np.int32), np.zeros(max(n, 1))
    f = lib().ref_chain
    f.restype = C.c_int
    f.argtypes = [C.c_int] * 3 + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int, C.c_int, C.c_double, C.c_int,
                  C.c_void_p, C.c_int] + [C.c_void_p] * 11
    done = f(N, M, self_idx, _p(arrs["pb"]), _p(arrs["strep"]), _p(arrs["longest"]), _p(arrs["bp_cnt"]), _p(arrs["bp_xy"]),
             par.bp_max, _p(arrs["known"]), _p(arrs["samp"]), par.num_pol, S, par.T_span, n, _p(arrs["cxy"]), cap,
             _p(arrs["cnt0"]), _p(arrs["alpha0"]), _p(arrs["beta0"]), _p(arrs["bend0"]), _p(arrs["active0"]), _p(cnt), _p(alpha),
             _p(beta), _p(bend), _p(active), _p(length))
    return done, cnt, alpha, beta, bend, active, length


def search(sb, res):
    """The reference's own KinodynamicSearch on every agent of a SearchBatch (oracle/ref_wrap_search.cpp).
    Returns (comb [B][ns^2] the jerk order each search drew, info [B][4]: node_num_max_, node_used_num_, goal_occupied_,
    states in entStateVec)."""
    from oracle import oracle
    sp, b, keep = oracle.search_batch_struct(sb, res)
    ns2 = sb.par.a_star_samp_x ** 2
    comb, info = np.zeros((sb.B, ns2), np.uint8), np.zeros((sb.B, 4), np.int32)
    f = lib().ref_search_batch
    f.restype = C.c_int
    f(C.byref(sp), C.byref(b), _p(comb), _p(info))
    return comb, info


def entangle_check_pwp(par, self_idx, strep, bp_cnt, bp_xy, known, samp, n, cxy, cnt0, alpha0, beta0, bend0, active0):
    """The reference's KinodynamicSearch::entangleCheckGivenPwp on a real object: (entangled, cnt, alpha, beta, bend, active)."""
    N, M, cap = par.num_of_agents, par.num_of_static_obst, par.ent_cap
    a = dict(pb=_c(par.pb, np.float64), strep=_c(strep, np.float64) if M else np.zeros((1, 2, 2)), bp_cnt=_c(bp_cnt, np.int32),
             bp_xy=_c(bp_xy, np.float64), known=_c(known, np.uint8), samp=_c(samp, np.float64), cxy=_c(cxy, np.float64))
    cnt, alpha, beta = np.array(cnt0, np.int32).copy(), np.array(alpha0, np.int32).copy(), np.array(beta0, np.float64).copy()
    bend, active = np.array(bend0, np.int32).copy(), np.array(active0, np.int32).copy()
    f = lib().ref_entangle_check_pwp
    f.restype = C.c_int
    f.argtypes = [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p,
                  C.c_int] + [C.c_void_p] * 5
    ent = f(N, M, self_idx, _p(a["pb"]), _p(a["strep"]), _p(a["bp_cnt"]), _p(a["bp_xy"]), par.bp_max, _p(a["known"]), _p(a["samp"]),
            par.num_pol, par.num_sample_per_interval, par.T_span, n, _p(a["cxy"]), cap, _p(cnt), _p(alpha), _p(beta), _p(bend), _p(active))
    return ent, cnt, alpha, beta, bend, active


def generate_traj(coeff, n, T, dc, t_start=0.0, max_states=4096):
    """The reference's generatePwpOut (KinodynamicSearch's copy of the code): (states [K][12], shifted knot times)."""
    coeff = _c(coeff, np.float64)
    st, times = np.zeros((max_states, 12)), np.zeros(n + 1)
    f = lib().ref_generate_traj
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
    k = f(_p(coeff), n, T, dc, t_start, _p(st), max_states, _p(times))
    assert k <= max_states
    return st[:k].copy(), times


def track(par, self_idx, strep, bp_cnt, bp_xy, bp_cnt_prev, bp_xy_prev, cnt, alpha, beta, bend, active, prev_pos, prev_pos_agent,
          latest, cur, elapsed_ms):
    """One tracker tick through the reference's functions; the state arrays and positions are updated in place."""
    N, M = par.num_of_agents, par.num_of_static_obst
    a = dict(pb=_c(par.pb, np.float64), strep=_c(strep, np.float64) if M else np.zeros((1, 2, 2)), bp_cnt=_c(bp_cnt, np.int32),
             bp_xy=_c(bp_xy, np.float64), bp_cnt_prev=_c(bp_cnt_prev, np.int32), bp_xy_prev=_c(bp_xy_prev, np.float64),
             latest=_c(latest, np.float64), cur=_c(cur, np.float64))
    for x in (cnt, alpha, beta, bend, active, prev_pos, prev_pos_agent):
        assert x.flags["C_CONTIGUOUS"]
    f = lib().ref_track
    f.restype = C.c_int
    f.argtypes = [C.c_int] * 3 + [C.c_void_p] * 6 + [C.c_int, C.c_int] + [C.c_void_p] * 9 + [C.c_double]
    return f(N, M, self_idx, _p(a["pb"]), _p(a["strep"]), _p(a["bp_cnt"]), _p(a["bp_xy"]), _p(a["bp_cnt_prev"]), _p(a["bp_xy_prev"]),
             par.bp_max, par.ent_cap, _p(cnt), _p(alpha), _p(beta), _p(bend), _p(active), _p(prev_pos), _p(prev_pos_agent),
             _p(a["latest"]), _p(a["cur"]), float(elapsed_ms))


def _count_exits(path):
    """Script mode: fork once per pickled ref.track argument tuple; print how many children ended with exit status 255."""
    import pickle
    import sys
    sys.path.insert(0, ROOT)
    with open(path, "rb") as f:
        cases = pickle.load(f)
    lib()
    n = 0
    for args in cases:
        pid = os.fork()
        if pid == 0:
            devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(devnull, 1)
            os.dup2(devnull, 2)
            try:
                track(*args)
            finally:
                os._exit(0)
        _, status = os.waitpid(pid, 0)
        n += os.WIFEXITED(status) and os.WEXITSTATUS(status) == 255
    print(n)


if __name__ == "__main__":
    import sys
    _count_exits(sys.argv[1])


def predict(par, self_idx, strep, bp_cnt, bp_xy, cnt, alpha, beta, bend, active, prev_pos, prev_pos_agent, cur, samp0, known):
    """PredictAlphasBetas through the reference's functions; cnt / alpha / beta / bend / active are updated in place."""
    N, M = par.num_of_agents, par.num_of_static_obst
    a = dict(pb=_c(par.pb, np.float64), strep=_c(strep, np.float64) if M else np.zeros((1, 2, 2)), bp_cnt=_c(bp_cnt, np.int32),
             bp_xy=_c(bp_xy, np.float64), pp=_c(prev_pos, np.float64), ppa=_c(prev_pos_agent, np.float64), cur=_c(cur, np.float64),
             samp0=_c(samp0, np.float64), known=_c(known, np.uint8))
    f = lib().ref_predict
    f.restype = C.c_int
    f.argtypes = [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_void_p] * 10
    return f(N, M, self_idx, _p(a["pb"]), _p(a["strep"]), _p(a["bp_cnt"]), _p(a["bp_xy"]), par.bp_max, par.ent_cap, _p(cnt), _p(alpha),
             _p(beta), _p(bend), _p(active), _p(a["pp"]), _p(a["ppa"]), _p(a["cur"]), _p(a["samp0"]), _p(a["known"]))


def compose_records(t, dc, p1, p2):
    """The reference's mu::composePieceWisePol on two records: (n_pieces, times [n+1], coeff [3][n][4], p1 times, p2 times)."""
    p1, p2 = _c(p1, np.float64), _c(p2, np.float64)
    times, coeff, t1, t2 = np.zeros(64), np.zeros((3, 64, 4)), np.zeros(17), np.zeros(17)
    f = lib().ref_compose_records
    f.restype = C.c_int
    f.argtypes = [C.c_double, C.c_double] + [C.c_void_p] * 6
    n = f(t, dc, _p(p1), _p(p2), _p(times), _p(coeff), _p(t1), _p(t2))
    return n, times[:n + 1].copy() if n else times[:0], coeff[:, :n].copy(), t1, t2


_LP_CB = None
LP_MODELS = []   # (G rows as the reference built them, lower, upper) of every LP solved since the last clear


def _install_highs():
    """Give the GLPK stand-in an LP engine: HiGHS (scipy) on exactly the model the reference's separator built."""
    global _LP_CB
    if _LP_CB is not None:
        return
    from scipy.optimize import linprog
    I, D = C.POINTER(C.c_int), C.POINTER(C.c_double)
    proto = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, I, D, D, I, D, D, D, C.c_int, C.c_int, I, I, D, D)

    def solve(rows, cols, rt, rlb, rub, ct, clb, cub, obj, direction, ne, ia, ja, ar, x):
        G = np.zeros((rows, cols))
        for k in range(ne):
            G[ia[k] - 1, ja[k] - 1] += ar[k]
        lo = np.array([rlb[i] if rt[i] in (2, 4, 5) else -np.inf for i in range(rows)])
        up = np.array([rub[i] if rt[i] in (3, 4) else (rlb[i] if rt[i] == 5 else np.inf) for i in range(rows)])
        LP_MODELS.append((G, lo, up))
        A_ub = np.vstack([G[np.isfinite(up)], -G[np.isfinite(lo)]])
        b_ub = np.concatenate([up[np.isfinite(up)], -lo[np.isfinite(lo)]])
        bounds = [(clb[j] if ct[j] in (2, 4, 5) else None, cub[j] if ct[j] in (3, 4) else (clb[j] if ct[j] == 5 else None)) for j in range(cols)]
        c = np.array([obj[j] for j in range(cols)]) * (-1.0 if direction == 2 else 1.0)
        r = linprog(c, A_ub=A_ub if len(A_ub) else None, b_ub=b_ub if len(b_ub) else None, bounds=bounds, method="highs")
        if r.status == 0:
            for j in range(cols):
                x[j] = r.x[j]
            return 5    # GLP_OPT
        return 6 if r.status == 3 else 4   # GLP_UNBND / GLP_NOFEAS

    _LP_CB = proto(solve)
    lib().ref_set_lp_solver(_LP_CB)


def separator_solve(variant, A, B, Aplus=None):
    """The reference's separator::Separator::solveModel (2-D variants) with HiGHS as its LP engine: (solved, n[3])."""
    _install_highs()
    A, B = _c(A, np.float64), _c(B, np.float64)
    Ap = _c(Aplus, np.float64) if Aplus is not None else np.zeros((0, 2))
    n = np.zeros(3)
    f = lib().ref_separator_solve
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    ok = f(variant, _p(A), len(A), _p(Ap) if len(Ap) else None, len(Ap), _p(B), len(B), _p(n))
    return bool(ok), n


def separator_solve3d(A, B):
    """The 3-D solveModel that the reference's test_separator.cpp calls: (solved, n[3], d)."""
    _install_highs()
    A, B = _c(A, np.float64), _c(B, np.float64)
    out = np.zeros(4)
    f = lib().ref_separator_solve3d
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    ok = f(_p(A), len(A), _p(B), len(B), _p(out))
    return bool(ok), out[:3].copy(), float(out[3])


# --------------------------------------------------------------------------------------------------------------------
# The reference's own PolySolverGurobi (neptune/src/solver_gurobi_poly.cpp) with recording stand-ins underneath
class RefQpModel(C.Structure):
    """ref_qp_model of oracle/ref_stubs/gurobi_c++.h"""
    _D, _S = C.POINTER(C.c_double), C.POINTER(C.c_char)
    _fields_ = [("nvar", C.c_int), ("lb", _D), ("ub", _D), ("Q", _D), ("c", _D), ("c0", C.c_double),
                ("nlin", C.c_int), ("A", _D), ("sense", _S), ("rhs", _D),
                ("nquad", C.c_int), ("Qc", _D), ("qc", _D), ("qsense", _S), ("qrhs", _D),
                ("time_limit", C.c_double), ("non_convex", C.c_int)]


QP_MODELS = []    # one dict per GRBModel::optimize() call of the reference since the last clear
_QP_CB = None
_LP_CANON_CB = None


def _arr(ptr, shape):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape)
    return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape).copy()


def _solve_recorded(m):
    """Solve one recorded model: HiGHS (convex QP); with the quadratic terminal row, SLSQP on the reduced space.
    Returns (GRB status, x)."""
    from tests.highs_util import qp_highs
    nv = m["nvar"]
    used = (np.abs(m["Q"]).sum(0) + np.abs(m["Q"]).sum(1) + np.abs(m["c"]) + np.abs(m["A"]).sum(0)) > 0
    if m["nquad"]:
        used |= (np.abs(m["Qc"]).sum((0, 1)) + np.abs(m["Qc"]).sum((0, 2)) + np.abs(m["qc"]).sum(0)) > 0
    idx = np.flatnonzero(used)
    P = (m["Q"] + m["Q"].T)[np.ix_(idx, idx)]
    q = m["c"][idx]
    A = m["A"][:, idx]
    eq = m["sense"] == b"="
    le, ge = m["sense"] == b"<", m["sense"] == b">"
    G = np.vstack([A[le], -A[ge]])
    h = np.concatenate([m["rhs"][le], -m["rhs"][ge]])
    x = np.zeros(nv)
    if m["nquad"] == 0:
        st, xs, _ = qp_highs(P, q, A[eq], m["rhs"][eq], G, h)
        if st != "Optimal":
            return 3, x
        x[idx] = xs
        return 2, x
    # quadratic row(s): eliminate the equalities, then SLSQP from the solution without the quadratic row
    from scipy.optimize import minimize
    Aeq, beq = A[eq], m["rhs"][eq]
    U, sv, Vt = np.linalg.svd(Aeq, full_matrices=True)
    rank = int((sv > 1e-10 * sv[0]).sum())
    xp = Vt[:rank].T @ ((U[:, :rank].T @ beq) / sv[:rank])
    if np.abs(Aeq @ xp - beq).max() > 1e-8 * (1 + np.abs(beq).max()):
        return 3, x
    Z = Vt[rank:].T
    Qc = [(m["Qc"][k] + m["Qc"][k].T)[np.ix_(idx, idx)] * 0.5 for k in range(m["nquad"])]
    qc = [m["qc"][k][idx] for k in range(m["nquad"])]

    def full(w):
        return xp + Z @ w
    if Z.shape[1] == 0:
        xs = xp
        ok = (G @ xs - h).max(initial=-1) <= 1e-7 and all(xs @ Qc[k] @ xs + qc[k] @ xs <= m["qrhs"][k] + 1e-9 for k in range(m["nquad"]))
        x[idx] = xs
        return (2 if ok else 3), x
    cons = [{"type": "ineq", "fun": lambda w: h - G @ full(w), "jac": lambda w: -G @ Z}]
    for k in range(m["nquad"]):
        cons.append({"type": "ineq", "fun": (lambda w, k=k: m["qrhs"][k] - full(w) @ Qc[k] @ full(w) - qc[k] @ full(w)),
                     "jac": (lambda w, k=k: -(2 * Qc[k] @ full(w) + qc[k]) @ Z)})
    st0, x0, _ = qp_highs(P, q, Aeq, beq, G, h)
    w0 = Z.T @ (x0 - xp) if st0 == "Optimal" else np.zeros(Z.shape[1])
    best = None
    for scale in (1.0, 0.0):
        r = minimize(lambda w: 0.5 * full(w) @ P @ full(w) + q @ full(w), w0 * scale, jac=lambda w: Z.T @ (P @ full(w) + q),
                     constraints=cons, method="SLSQP", options={"ftol": 1e-15, "maxiter": 500})
        xs = full(r.x)
        feas = (G @ xs - h).max(initial=-1) <= 1e-7 and all(xs @ Qc[k] @ xs + qc[k] @ xs <= m["qrhs"][k] + 1e-8 for k in range(m["nquad"]))
        if feas and (best is None or r.fun < best[0]):
            best = (r.fun, xs)
    if best is None:
        return 3, x
    x[idx] = best[1]
    return 2, x


def install_qp_hooks(oracle):
    """Engines for the reference's two third-party solvers: the separator's LP returns the canonical (minimum-norm) line
    of the oracle -- with a zero objective any feasible vertex is a legitimate GLPK answer, and this one makes the QP
    rows comparable with the oracle's -- and GRBModel::optimize records the model and solves it with HiGHS."""
    global _QP_CB, _LP_CANON_CB, _LP_CB
    I, D = C.POINTER(C.c_int), C.POINTER(C.c_double)
    lp_proto = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, I, D, D, I, D, D, D, C.c_int, C.c_int, I, I, D, D)

    def lp_solve(rows, cols, rt, rlb, rub, ct, clb, cub, obj, direction, ne, ia, ja, ar, x):
        G = np.zeros((rows, cols))
        for k in range(ne):
            G[ia[k] - 1, ja[k] - 1] += ar[k]
        isA = np.array([rt[i] == 2 for i in range(rows)])      # GLP_LO: n.a + d >= 1
        A, B = G[isA][:, :2], G[~isA][:, :2]
        assert cols == 3 and np.all(G[:, 2] == 1.0)
        ok, line = oracle.separate(A, B)
        if not ok:
            return 4
        for j in range(3):
            x[j] = line[j]
        return 5
    _LP_CANON_CB = lp_proto(lp_solve)
    _LP_CB = None        # a later separator_solve() re-installs HiGHS
    lib().ref_set_lp_solver(_LP_CANON_CB)
    qp_proto = C.CFUNCTYPE(C.c_int, C.POINTER(RefQpModel), D, I, C.c_void_p)

    def qp_solve(mp, x, sol_count, user):
        m = mp.contents
        nv, nl, nq = m.nvar, m.nlin, m.nquad
        rec = dict(nvar=nv, nlin=nl, nquad=nq, Q=_arr(m.Q, (nv, nv)), c=_arr(m.c, (nv,)), c0=m.c0, A=_arr(m.A, (nl, nv)),
                   sense=np.array([m.sense[i] for i in range(nl)]), rhs=_arr(m.rhs, (nl,)),
                   Qc=_arr(m.Qc, (nq, nv, nv)), qc=_arr(m.qc, (nq, nv)), qsense=np.array([m.qsense[i] for i in range(nq)]),
                   qrhs=_arr(m.qrhs, (nq,)), lb=_arr(m.lb, (nv,)), ub=_arr(m.ub, (nv,)), time_limit=m.time_limit)
        status, xs = _solve_recorded(rec)
        rec["status"], rec["x"] = status, xs
        QP_MODELS.append(rec)
        for j in range(nv):
            x[j] = xs[j]
        sol_count[0] = 1 if status == 2 else 0
        return status
    _QP_CB = qp_proto(qp_solve)
    lib().ref_set_qp_solver.argtypes = [qp_proto, C.c_void_p]
    lib().ref_set_qp_solver(_QP_CB, None)


def qp_replan(batch, a: int, replans: int = 1, t_start: float = 1.25):
    """The reference's PolySolverGurobi on agent `a` of a ReplanBatch, driven as Neptune drives it.
    Returns (ok, coeff_out [3][8][4], times [n+1], objective, n_states); the models of every optimize() are in QP_MODELS."""
    par = batch.par
    N, M, NH, cap = par.num_of_agents, par.num_of_static_obst, batch.n_hull_slots, par.ent_cap
    n = int(batch.n_int[a])
    hp = _c(batch.hull_ptr[a * NH * 8:(a + 1) * NH * 8 + 1], np.int64)
    arrs = dict(pb=_c(par.pb, np.float64), lim=_c([par.x_min, par.x_max, par.y_min, par.y_max, par.z_min, par.z_max], np.float64),
                st_ptr=_c(batch.st_ptr, np.int64), st_xy=_c(batch.st_xy if len(batch.st_xy) else np.zeros((1, 2)), np.float64),
                ci=_c(batch.coeff_init[a], np.float64), hxy=_c(batch.hull_xy if len(batch.hull_xy) else np.zeros((1, 2)), np.float64),
                nih0=_c(batch.nih0[a], np.float64), ecnt=_c(batch.esv_cnt[a], np.int32), ealpha=_c(batch.esv_alpha[a], np.int32),
                eact=_c(batch.esv_active[a], np.int32), bpc=_c(batch.bp_cnt, np.int32), bpx=_c(batch.bp_xy, np.float64))
    co, times, obj, ns = np.zeros((3, 8, 4)), np.zeros(9), C.c_double(0), C.c_int(0)
    f = lib().ref_qp_replan
    f.restype = C.c_int
    P, D, Ic = C.c_void_p, C.c_double, C.c_int
    f.argtypes = [Ic, Ic, Ic, D, D, P, P, D, D, D, D, D, Ic, P, P, Ic, P, Ic, P, P, P, P, P, P, Ic, P, P, Ic, Ic, D, D, P, P,
                  C.POINTER(C.c_double), C.POINTER(C.c_int)]
    ok = f(N, int(batch.agent_id[a]), par.num_pol, par.T_span, par.weight, _p(arrs["pb"]), _p(arrs["lim"]), par.v_max, par.a_max,
           par.j_max, par.runtime_opt, par.tetherLength, M, _p(arrs["st_ptr"]), _p(arrs["st_xy"]), n, _p(arrs["ci"]), NH, _p(hp),
           _p(arrs["hxy"]), _p(arrs["nih0"]), _p(arrs["ecnt"]), _p(arrs["ealpha"]), _p(arrs["eact"]), cap, _p(arrs["bpc"]),
           _p(arrs["bpx"]), par.bp_max, replans, t_start, par.dc, _p(co), _p(times), C.byref(obj), C.byref(ns))
    return bool(ok), co, times[:n + 1].copy(), obj.value, ns.value

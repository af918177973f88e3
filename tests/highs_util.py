"""HiGHS (bundled with scipy) as the independent LP/QP cross-check named in SURVEY.md section 8c."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def qp_highs(P, q, Aeq, beq, G, h):
    """min 0.5 x'Px + q'x  s.t. Aeq x = beq, G x <= h.  Returns (status_str, x, objective).

    The equalities are eliminated with an SVD null-space basis first: HiGHS' active-set QP solver is
    unreliable on the full-space model (its Hessian is only positive definite on the null space)."""
    Aeq = np.asarray(Aeq, float)
    if Aeq.shape[0]:
        U, sv, Vt = np.linalg.svd(Aeq, full_matrices=True)
        rank = int((sv > 1e-10 * sv[0]).sum())
        xp = Vt[:rank].T @ ((U[:, :rank].T @ beq) / sv[:rank])
        if np.abs(Aeq @ xp - beq).max() > 1e-8 * (1 + np.abs(beq).max()):
            return "Infeasible", xp, np.inf
        Z = Vt[rank:].T
        if Z.shape[1] == 0:
            ok = (G @ xp - h).max() <= 1e-7
            return ("Optimal" if ok else "Infeasible"), xp, 0.5 * xp @ P @ xp + q @ xp
        st, w, _ = _qp_highs_ineq(Z.T @ P @ Z, Z.T @ (P @ xp + q), G @ Z, h - G @ xp)
        x = xp + Z @ w
        return st, x, 0.5 * x @ P @ x + q @ x
    return _qp_highs_ineq(P, q, G, h)


def _qp_highs_ineq(P, q, G, h):
    Aeq, beq = np.zeros((0, len(q))), np.zeros(0)
    import scipy.optimize._highspy._core as hc

    n = len(q)
    A = sp.csc_matrix(np.vstack([Aeq, G]))
    lo = np.concatenate([beq, np.full(len(h), -hc.kHighsInf)])
    hi = np.concatenate([beq, h])
    hs = hc._Highs()
    hs.setOptionValue("output_flag", False)
    hs.setOptionValue("time_limit", 5.0)
    hs.setOptionValue("primal_feasibility_tolerance", 1e-9)
    hs.setOptionValue("dual_feasibility_tolerance", 1e-9)
    lp = hc.HighsLp()
    lp.num_col_, lp.num_row_ = n, A.shape[0]
    lp.col_cost_ = np.asarray(q, float)
    lp.col_lower_ = np.full(n, -hc.kHighsInf)
    lp.col_upper_ = np.full(n, hc.kHighsInf)
    lp.row_lower_, lp.row_upper_ = lo, hi
    lp.a_matrix_.format_ = hc.MatrixFormat.kColwise
    lp.a_matrix_.start_ = A.indptr.astype(np.int32)
    lp.a_matrix_.index_ = A.indices.astype(np.int32)
    lp.a_matrix_.value_ = A.data.astype(float)
    hs.passModel(lp)
    Pl = sp.csc_matrix(np.tril(P))
    hess = hc.HighsHessian()
    hess.dim_ = n
    hess.format_ = hc.HessianFormat.kTriangular
    hess.start_ = Pl.indptr.astype(np.int32)
    hess.index_ = Pl.indices.astype(np.int32)
    hess.value_ = Pl.data.astype(float)
    hs.passHessian(hess)
    hs.run()
    st = hs.modelStatusToString(hs.getModelStatus())
    sol = hs.getSolution()
    x = np.array(sol.col_value)
    return st, x, 0.5 * x @ P @ x + q @ x

"""Generates tests/golden/*.npz: seeded scenes (inputs), the oracle's outputs and HiGHS' solution of
the same QPs.  The reference ships no golden vectors and cannot be run here (no Gurobi / GLPK / CGAL),
so these fixtures pin the ORACLE against an independent solver (HiGHS via scipy) instead.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from highs_util import qp_highs  # noqa: E402
from neptune_b200 import config  # noqa: E402
from neptune_b200.batch import ReplanResult  # noqa: E402
from neptune_b200.scenes import make_scene  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.ent_backends import OracleEntBackend  # noqa: E402

CASES = [("single", 1001, dict(n_fixed=3)), ("mtlp5", 2002, dict(sync=False)), ("mtlp5", 2005, dict(sync=False)),
         ("obst8", 3003, dict(sync=False)), ("obst8", 3004, dict(sync=False)),
         # crafted inputs (tests/crafted.py) for the branches no generated scene reaches: non-entangling lines
         # (solver_gurobi_poly.cpp:715-784) and the failure path pwp_out = pwp_init (:856-859)
         ("mtlp5", "crafted-ent0", None), ("mtlp5", "crafted-ent1", None), ("mtlp5", "crafted-ent2", None),
         ("mtlp5", "crafted-box", None), ("mtlp5", "crafted-vel", None), ("mtlp5", "crafted-n2box", None)]
BATCH_KEYS = ("agent_id", "n_int", "coeff_init", "hull_ptr", "hull_xy", "nih0", "st_ptr", "st_xy", "esv_cnt",
              "esv_alpha", "esv_active", "bp_cnt", "bp_xy")


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for cfg, seed, kw in CASES:
        sc = None
        if kw is None:
            from tests import crafted
            kind = seed.split("-")[1]
            par, b = crafted.ent_lp_batch(int(kind[3:])) if kind.startswith("ent") else crafted.infeasible_batch(kind)
        else:
            par = config(cfg)
            sc = make_scene(par, seed, ent_backend=OracleEntBackend(orc), **kw)
            b = sc.batch
        res = ReplanResult.empty(b)
        assert orc.replan_batch(b, res, 1) == 0
        hx = np.full((b.B, 2, 96), np.nan)
        hstat = np.zeros((b.B, 2), np.int32)  # 1 optimal, 0 infeasible, -1 other
        hobj = np.full((b.B, 2), np.nan)
        has_qc = np.zeros(b.B, np.uint8)
        for a in range(b.B):
            for fb in (0, 1):
                m = orc.export_qp(b, a, bool(fb), res.lines[a], res.line_ok[a])
                has_qc[a] = m["has_qc"]
                st, x, f = qp_highs(m["P"], m["q"], m["Aeq"], m["beq"], m["G"], m["h"])
                hstat[a, fb] = 1 if st == "Optimal" else (0 if st == "Infeasible" else -1)
                if st == "Optimal":
                    n = m["n"]
                    for i in range(n):
                        for ax in range(3):
                            hx[a, fb, ax * 32 + 4 * i:ax * 32 + 4 * i + 4] = x[i * 12 + ax * 4:i * 12 + ax * 4 + 4]
                    hobj[a, fb] = f + m["c0"]
        d = {k: getattr(b, k) for k in BATCH_KEYS}
        if sc is not None:
            d.update(strep=sc.strep, t_start=sc.t_start, samp=sc.samp, known=sc.known,
                     esA_cnt=sc.esA_cnt, esA_alpha=sc.esA_alpha, esA_beta=sc.esA_beta, esA_bend=sc.esA_bend,
                     esA_active=sc.esA_active)
        d.update(n_hull_slots=b.n_hull_slots,
                 orc_coeff=res.coeff_out, orc_obj=res.obj, orc_status=res.status, orc_lines=res.lines,
                 orc_line_ok=res.line_ok, highs_x=hx, highs_status=hstat, highs_obj=hobj, has_qc=has_qc)
        path = os.path.join(out_dir, f"{cfg}_{seed}.npz")
        np.savez_compressed(path, **d)
        print(path, os.path.getsize(path), "status", res.status, "highs", hstat.tolist())


if __name__ == "__main__":
    main()

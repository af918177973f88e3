"""Records golden vectors of the entanglement chain and of whole front-end searches from the REFERENCE library
(oracle/_ref/libneptune_ref.so: the reference's entangle_utils.cpp, gjk.cpp and kinodynamic_search.cpp compiled against the
Eigen stand-in) for tests/test_reference_pin.py.
Run in a container that has /root/reference:  python tests/golden/make_ref_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from neptune_b200.search import static_longest_dist  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests import ref_pin_util as ref  # noqa: E402
from tests.test_reference_pin import SEARCH_CASES, _chain_cases, _reference_search, _search_case  # noqa: E402


def main():
    out, k = {}, 0
    for par, sc, b, n, cxy, st0 in _chain_cases(orc):
        if par.num_of_agents > 8:
            continue
        M = par.num_of_static_obst
        longest = static_longest_dist(sc.static_raw, np.asarray(sc.strep).reshape(M, 2, 2)) if M else np.zeros((0, 2))
        done, cnt, alpha, beta, bend, active, length = ref.chain(par, int(sc.batch.agent_id[b]) - 1, sc.strep, longest, sc.batch.bp_cnt,
                                                                 sc.batch.bp_xy, sc.known[b], sc.samp[b], n, cxy, *st0)
        mx = max(1, int(cnt[:, 0].max()))
        out.update({f"cxy_{k}": cxy, f"done_{k}": np.int32(done), f"cnt_{k}": cnt, f"alpha_{k}": alpha[:, :mx], f"beta_{k}": beta[:, :mx],
                    f"bend_{k}": bend[:, :mx], f"active_{k}": active, f"len_{k}": length})
        k += 1
    out["n_cases"] = np.int32(k)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference", "ref_chain.npz"), **out)
    print("wrote", k, "cases")

    out = {"n_cases": np.int32(len(SEARCH_CASES))}
    for k, (cfg, seed, mods, mb) in enumerate(SEARCH_CASES):
        par, sb = _search_case(orc, cfg, seed, mods, mb)
        g = _reference_search(sb)
        mx = max(1, int(g["esv_cnt"][:, :, 0].max()), int(g["esv_cnt"][:, :, 1].max()))
        for f in ("esv_alpha", "esv_beta", "esv_bend"):
            g[f] = g[f][:, :, :mx]
        out.update({f"{f}_{k}": v for f, v in g.items()})
        print(cfg, seed, "status", g["status"].tolist(), "nodes", g["nodes"].tolist())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference", "ref_search.npz"), **out)
    make_qp_golden(orc)


def make_qp_golden(orc):
    """Outputs of the reference's own PolySolverGurobi (solver_gurobi_poly.cpp, HiGHS under a recording Gurobi stand-in,
    canonical lines from its separator) on every agent of every fixture in tests/golden/: status path, pwp_out, objective.
    The recorded models are checked against the oracle's restatement while recording (tests/test_reference_pin.py)."""
    from neptune_b200.batch import ReplanResult
    from tests.golden_util import golden_files, load
    from tests.test_reference_pin import run_reference_qp
    ref.install_qp_hooks(orc)
    out = {}
    for path in golden_files():
        name = os.path.basename(path)[:-4]
        par, b, z = load(path)
        res = ReplanResult.empty(b)
        assert orc.replan_batch(b, res, 2) == 0
        st, co, ob = np.zeros(b.B, np.int32), np.zeros((b.B, 3, 8, 4)), np.zeros(b.B)
        for a in range(b.B):
            st[a], co[a], ob[a] = run_reference_qp(orc, b, a, res)
        out[name + "/status"], out[name + "/coeff"], out[name + "/obj"] = st, co, ob
        print(name, "reference status", st.tolist(), "oracle", res.status.tolist(), "max|dcoeff|", np.abs(co - res.coeff_out).max())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference", "ref_qp.npz"), **out)


if __name__ == "__main__":
    main()

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


if __name__ == "__main__":
    main()

"""Random straight-line walks through the obstacle field of config 3: they cross many tethers and
static-obstacle lines, so the signature word grows, cancels, and bend points are created and released
(eu::updateBendPts) -- the part of the entanglement chain the planner scenes rarely reach."""
import numpy as np

from neptune_b200.batch import NPOL


def random_walks(par, B, rng, scale):
    T = par.T_span
    co = np.zeros((B, 3, NPOL, 4))
    for b in range(B):
        p = np.asarray(par.pb[b]) + rng.normal(0, 1.0, 2)
        for i in range(8):
            q = np.clip(p + rng.normal(0, scale, 2), -14, 14)
            v = (q - p) / T
            co[b, 0, i] = [0, 0, v[0], p[0]]
            co[b, 1, i] = [0, 0, v[1], p[1]]
            p = q
    return co


def compare_backends(par, sc, ref_backend, got_backend, trials=12, seed=0):
    """Runs the front-end chain and the post-check on random walks with both back ends; returns
    (max alphas length, max bend points) seen.  Asserts bit-exact equality of every output."""
    from neptune_b200.capi import EntArrays
    rng = np.random.default_rng(seed)
    B = sc.batch.B
    st = EntArrays(par, B)
    mx_a = mx_b = 0
    for trial in range(trials):
        co = random_walks(par, B, rng, 3.0 if trial % 2 else 6.0)
        n = np.full(B, 8, np.int32)
        args = (par, sc.batch.agent_id, n, co, sc.samp, sc.known, sc.strep, sc.batch.bp_cnt, sc.batch.bp_xy,
                st.cnt, st.alpha, st.beta, st.bend, st.active)
        r1, r2 = ref_backend.rollout_batch(*args), got_backend.rollout_batch(*args)
        for x, y in zip(r1, r2):
            assert np.array_equal(x, y)
        mx_a, mx_b = max(mx_a, int(r1[1][..., 0].max())), max(mx_b, int(r1[1][..., 1].max()))
        # post-check (interval 0 only) starting from the state reached after 4 intervals
        mid = tuple(np.ascontiguousarray(a[:, 4]) for a in r1[1:])
        args2 = (par, sc.batch.agent_id, n, co, sc.samp, sc.known, sc.strep, sc.batch.bp_cnt, sc.batch.bp_xy) + mid
        c1, c2 = ref_backend.check_batch(*args2), got_backend.check_batch(*args2)
        for x, y in zip(c1, c2):
            assert np.array_equal(x, y)
    return mx_a, mx_b

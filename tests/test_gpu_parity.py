"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs.  Bars: LP solved flags and QP status paths bit-exact; separating lines
within 1e-9 relative; QP coefficients within 1e-6 * max(1, |x|_inf) at equal objective (1e-8 rel)."""
import numpy as np
import pytest

from neptune_b200 import config
from neptune_b200.batch import ReplanResult
from neptune_b200.scenes import make_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from neptune_b200 import capi
    capi.lib()
    return capi


def _solver(capi, par, sc=None):
    s = capi.Solver(par)
    if par.num_of_static_obst:
        s.set_static(sc.batch.st_ptr, sc.batch.st_xy, sc.strep)
    return s


def test_separate_batch_matches_oracle(capi, oracle):
    rng = np.random.default_rng(7)
    par = config("mtlp5")
    s = capi.Solver(par)
    A, B = [], []
    for _ in range(4000):
        A.append(oracle.convex_hull(rng.normal(size=(rng.integers(1, 14), 2)) * 1.5 + rng.normal(size=2) * 2))
        B.append(rng.normal(size=(4, 2)) + rng.normal(size=2) * 2)
    a_ptr = np.concatenate([[0], np.cumsum([len(a) for a in A])])
    b_ptr = np.concatenate([[0], np.cumsum([len(b) for b in B])])
    for poly in (True, False):
        ok, line = s.separate(a_ptr, np.concatenate(A), b_ptr, np.concatenate(B), poly)
        for i in range(len(A)):
            ok_o, l_o = oracle.separate(A[i], B[i])
            assert ok[i] == ok_o
            if ok_o:
                assert np.abs(line[i] - l_o).max() <= 1e-9 * max(1.0, np.abs(l_o).max())
                assert (A[i] @ line[i, :2] + line[i, 2] >= 1 - 1e-9).all()
                assert (B[i] @ line[i, :2] + line[i, 2] <= -1 + 1e-9).all()
    s.close()


@pytest.mark.parametrize("cfg,seeds,kw", [
    ("single", range(1001, 1004), dict(n_fixed=3)),
    ("mtlp5", range(2002, 2008), dict(sync=False)),
    ("obst8", range(3003, 3007), dict(sync=False)),
])
def test_replan_matches_oracle(capi, oracle, cfg, seeds, kw):
    par = config(cfg)
    for seed in seeds:
        sc = make_scene(par, seed, **kw)
        s = _solver(capi, par, sc)
        got = s.replan(sc.batch)
        ref = ReplanResult.empty(sc.batch)
        assert oracle.replan_batch(sc.batch, ref, 1) == 0
        assert (got.line_ok == ref.line_ok).all()
        m = ref.line_ok == 1
        assert np.abs(got.lines[m] - ref.lines[m]).max(initial=0) <= 1e-9 * max(1.0, np.abs(ref.lines[m]).max(initial=0))
        assert (got.status == ref.status).all(), (got.status, ref.status)
        tol = 1e-6 * max(1.0, np.abs(ref.coeff_out).max())
        assert np.abs(got.coeff_out - ref.coeff_out).max() <= tol
        assert np.abs(got.obj - ref.obj).max() <= 1e-8 * max(1.0, np.abs(ref.obj).max())
        s.close()


def test_generate_traj_matches_oracle(capi, oracle):
    par = config("mtlp5")
    sc = make_scene(par, 2002, sync=False)
    s = capi.Solver(par)
    states, ns = s.generate_traj(sc.batch.n_int, sc.batch.coeff_init, par.dc)
    for b in range(sc.batch.B):
        ref = oracle.generate_traj(sc.batch.coeff_init[b], int(sc.batch.n_int[b]), par.T_span, par.dc)
        assert ns[b] == len(ref)
        # floating point (FMA contraction differs between nvcc and gcc): 1e-12 absolute
        assert np.abs(states[b, :ns[b]] - ref).max() <= 1e-12
    s.close()


def test_missing_static_is_an_error(capi):
    par = config("obst8")
    sc = make_scene(par, 3003, sync=False)
    s = capi.Solver(par)
    with pytest.raises(capi.NbError):
        s.replan(sc.batch)
    s.close()


@pytest.mark.parametrize("cfg,seeds", [("mtlp5", (2002, 2005)), ("obst8", (3003, 3004, 3006))])
def test_entangle_chain_bit_exact(capi, oracle, cfg, seeds):
    """K3 vs oracle: alphas, active_cases, bendPointsIdx bit-exact (betas too: same IEEE operations)
    for the history walk, PredictAlphasBetas, the front-end chain and entangleCheckGivenPwp."""
    from tests.ent_backends import OracleEntBackend
    par = config(cfg)
    for seed in seeds:
        ref = make_scene(par, seed, sync=False, ent_backend=OracleEntBackend(oracle))
        s = capi.Solver(par)
        if par.num_of_static_obst:
            s.set_static(ref.batch.st_ptr, ref.batch.st_xy, ref.strep)
        dev = capi.DeviceEntBackend(s)
        got = make_scene(par, seed, sync=False, ent_backend=dev)
        for k in ("es0_cnt", "es0_alpha", "es0_beta", "es0_bend", "es0_active", "esA_cnt", "esA_alpha", "esA_beta",
                  "esA_bend", "esA_active"):
            assert np.array_equal(getattr(got, k), getattr(ref, k)), k
        for k in ("esv_cnt", "esv_alpha", "esv_active"):
            assert np.array_equal(getattr(got.batch, k), getattr(ref.batch, k)), k
        # post-check on the optimised trajectory (a17), both sides from the same state
        res = s.replan(ref.batch)
        b = ref.batch
        args = (par, b.agent_id, b.n_int, res.coeff_out, ref.samp, ref.known, ref.strep, b.bp_cnt, b.bp_xy,
                ref.esA_cnt, ref.esA_alpha, ref.esA_beta, ref.esA_bend, ref.esA_active)
        g = dev.check_batch(*args)
        r = OracleEntBackend(oracle).check_batch(*args)
        for x, y in zip(g, r):
            assert np.array_equal(x, y)
        s.close()

"""ctypes binding of the C-ABI in include/neptune_b200.h (libneptune_b200.so).

This is the only compute path of the package.  There is no CPU fallback: if the shared library is
missing, or no CUDA device is present, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .batch import NPOL, ReplanBatch, ReplanResult
from .params import Params

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libneptune_b200.so")
NB_HOST, NB_DEVICE = 0, 1
_P = C.c_void_p


class NbParams(C.Structure):
    _fields_ = [("num_pol", C.c_int32), ("deg_pol", C.c_int32), ("num_agents", C.c_int32),
                ("num_static", C.c_int32), ("samples", C.c_int32), ("use_linear_constraints", C.c_int32),
                ("T_span", C.c_double), ("weight", C.c_double), ("lim_min", C.c_double * 3),
                ("lim_max", C.c_double * 3), ("v_max", C.c_double), ("a_max", C.c_double),
                ("drone_radius", C.c_double), ("tether_length", C.c_double), ("ent_cap", C.c_int32),
                ("bp_max", C.c_int32), ("ent_slots", C.c_int32), ("ipm_max_iter", C.c_int32),
                ("ipm_tol", C.c_double)]


class NbReplanArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("space", C.c_int32), ("agent_id", _P), ("n_int", _P), ("coeff_init", _P),
                ("n_hull_slots", C.c_int32), ("hull_ptr", _P), ("hull_xy", _P), ("hull_nvert", C.c_int64),
                ("hull_cnt", _P), ("nih0", _P), ("nih0_group", _P), ("hull_known", _P), ("esv_cnt", _P), ("esv_alpha", _P), ("esv_active", _P), ("bp_cnt", _P),
                ("bp_xy", _P), ("coeff_out", _P), ("obj", _P), ("status", _P), ("iters", _P), ("lines", _P),
                ("line_ok", _P)]


def make_nb_params(par: Params) -> NbParams:
    p = NbParams()
    p.num_pol, p.deg_pol, p.num_agents, p.num_static = par.num_pol, par.deg_pol, par.num_of_agents, par.num_of_static_obst
    p.samples, p.use_linear_constraints = par.num_sample_per_interval, int(par.use_linear_constraints)
    p.T_span, p.weight = par.T_span, par.weight
    p.lim_min[:] = [par.x_min, par.y_min, par.z_min]
    p.lim_max[:] = [par.x_max, par.y_max, par.z_max]
    p.v_max, p.a_max, p.drone_radius, p.tether_length = par.v_max, par.a_max, par.drone_radius, par.tetherLength
    p.ent_cap, p.bp_max, p.ent_slots = par.ent_cap, par.bp_max, par.ent_slots
    p.ipm_max_iter, p.ipm_tol = par.ipm_max_iter, par.ipm_tol
    return p


def _np(a):
    return None if a is None else a.ctypes.data_as(_P)


def host_args(batch: ReplanBatch, res: ReplanResult) -> NbReplanArgs:
    """nb_replan_args over host (numpy) buffers."""
    a = NbReplanArgs()
    a.B, a.space = batch.B, NB_HOST
    a.agent_id, a.n_int, a.coeff_init = _np(batch.agent_id), _np(batch.n_int), _np(batch.coeff_init)
    a.n_hull_slots, a.hull_ptr, a.hull_xy = batch.n_hull_slots, _np(batch.hull_ptr), _np(batch.hull_xy)
    a.hull_nvert = int(batch.hull_xy.shape[0])
    a.hull_cnt = None
    a.nih0_group = None
    a.hull_known = None
    a.nih0, a.esv_cnt, a.esv_alpha, a.esv_active = _np(batch.nih0), _np(batch.esv_cnt), _np(batch.esv_alpha), _np(batch.esv_active)
    a.bp_cnt, a.bp_xy = _np(batch.bp_cnt), _np(batch.bp_xy)
    a.coeff_out, a.obj, a.status, a.iters = _np(res.coeff_out), _np(res.obj), _np(res.status), _np(res.iters)
    a.lines, a.line_ok = _np(res.lines), _np(res.line_ok)
    return a


class NbEntState(C.Structure):
    _fields_ = [("cnt", _P), ("alpha", _P), ("beta", _P), ("bend", _P), ("active", _P)]


class EntArrays:
    """eu::ent_state for a batch: fixed-capacity numpy storage ([S] states)."""

    def __init__(self, par: Params, shape):
        shape = tuple(np.atleast_1d(shape))
        self.cnt = np.zeros(shape + (2,), np.int32)
        self.alpha = np.zeros(shape + (par.ent_cap, 2), np.int32)
        self.beta = np.zeros(shape + (par.ent_cap,), np.float64)
        self.bend = np.zeros(shape + (par.ent_cap,), np.int32)
        self.active = np.zeros(shape + (par.NA,), np.int32)

    @staticmethod
    def of(par, cnt, alpha, beta, bend, active):
        e = EntArrays.__new__(EntArrays)
        e.cnt, e.alpha, e.beta = np.ascontiguousarray(cnt, np.int32), np.ascontiguousarray(alpha, np.int32), np.ascontiguousarray(beta, np.float64)
        e.bend, e.active = np.ascontiguousarray(bend, np.int32), np.ascontiguousarray(active, np.int32)
        return e

    def copy(self):
        return EntArrays.of(None, self.cnt.copy(), self.alpha.copy(), self.beta.copy(), self.bend.copy(), self.active.copy())

    def c(self) -> NbEntState:
        s = NbEntState()
        s.cnt, s.alpha, s.beta, s.bend, s.active = _np(self.cnt), _np(self.alpha), _np(self.beta), _np(self.bend), _np(self.active)
        return s

    def tuple(self):
        return self.cnt, self.alpha, self.beta, self.bend, self.active


class NbSearchParams(C.Structure):
    _fields_ = [("num_samples", C.c_int32), ("j_max", C.c_double), ("voxel_size", C.c_double), ("bias", C.c_double),
                ("goal_size", C.c_double), ("enable_entangle_check", C.c_int32), ("use_not_reaching_soln", C.c_int32),
                ("max_nodes", C.c_int32), ("max_expansions", C.c_int32), ("ecap", C.c_int32)]


class NbSearchArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("space", C.c_int32), ("agent_id", _P), ("init", _P), ("goal", _P), ("coeffs_z", _P),
                ("n_groups", C.c_int32), ("group", _P), ("hull_xy", _P), ("hull_cnt", _P), ("samp", _P), ("known", _P),
                ("es", NbEntState), ("bp_cnt", _P), ("bp_xy", _P), ("comb", _P), ("comb_shared", C.c_int32),
                ("status", _P), ("solved", _P), ("n_int", _P), ("coeff", _P), ("esv", NbEntState), ("stats", _P),
                ("cost", _P)]


def make_search_params(par: Params) -> NbSearchParams:
    sp = NbSearchParams()
    sp.num_samples, sp.j_max, sp.voxel_size = par.a_star_samp_x, par.j_max, par.a_star_fraction_voxel_size
    sp.bias, sp.goal_size = par.a_star_bias, par.goal_radius
    sp.enable_entangle_check, sp.use_not_reaching_soln = int(par.enable_entangle_check), int(par.use_not_reaching_soln)
    sp.max_nodes, sp.max_expansions, sp.ecap = par.search_max_nodes, par.search_max_expansions, par.search_ecap
    return sp


def host_search_args(sb, res) -> NbSearchArgs:
    """nb_search_args over the host (numpy) buffers of a SearchBatch / SearchResult."""
    a = NbSearchArgs()
    a.B, a.space = sb.B, NB_HOST
    a.agent_id, a.init, a.goal, a.coeffs_z = _np(sb.agent_id), _np(sb.init), _np(sb.goal), _np(sb.coeffs_z)
    a.n_groups, a.group = sb.G, _np(sb.group)
    a.hull_xy, a.hull_cnt, a.samp, a.known = _np(sb.hull_xy), _np(sb.hull_cnt), _np(sb.samp), _np(sb.known)
    a.es.cnt, a.es.alpha, a.es.beta, a.es.bend, a.es.active = _np(sb.es_cnt), _np(sb.es_alpha), _np(sb.es_beta), _np(sb.es_bend), _np(sb.es_active)
    a.bp_cnt, a.bp_xy, a.comb, a.comb_shared = _np(sb.bp_cnt), _np(sb.bp_xy), _np(sb.comb), int(sb.comb.ndim == 1)
    a.status, a.solved, a.n_int, a.coeff = _np(res.status), _np(res.solved), _np(res.n_int), _np(res.coeff)
    a.esv.cnt, a.esv.alpha, a.esv.beta, a.esv.bend, a.esv.active = _np(res.esv_cnt), _np(res.esv_alpha), _np(res.esv_beta), _np(res.esv_bend), _np(res.esv_active)
    a.stats, a.cost = _np(res.stats), _np(res.cost)
    return a


_lib = None


def lib():
    """Load libneptune_b200.so; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). neptune_b200 has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.nb_last_error.restype = C.c_char_p
        _lib.nb_launch_count.restype = C.c_longlong
        _lib.nb_launch_count.argtypes = [_P]
        _lib.nb_create.argtypes = [C.POINTER(NbParams), _P, C.c_int, C.POINTER(_P)]
        _lib.nb_destroy.argtypes = [_P]
        _lib.nb_set_static.argtypes = [_P, _P, _P, _P]
        _lib.nb_replan_batch.argtypes = [_P, C.POINTER(NbReplanArgs), _P]
        _lib.nb_line_slots.argtypes = [_P, C.c_int]
        _lib.nb_search_configure.argtypes = [_P, C.POINTER(NbSearchParams)]
        _lib.nb_set_static_longest.argtypes = [_P, _P]
        _lib.nb_search_batch.argtypes = [_P, C.POINTER(NbSearchArgs), _P]
    return _lib


class NbError(RuntimeError):
    pass


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise NbError(f"{what} failed with nb_error {rc}: {lib().nb_last_error().decode()}")


class Solver:
    """Long-lived back-end object: the batched counterpart of one ``PolySolverGurobi`` per agent
    (reference ``neptune/include/neptune.hpp:183``; constructed at ``neptune.cpp:102-107``)."""

    def __init__(self, par: Params, device: int = -1):
        self.par = par
        self._h = _P()
        nbp = make_nb_params(par)
        pb = np.ascontiguousarray(par.pb, dtype=np.float64)
        assert pb.shape == (par.num_of_agents, 2)
        _check(lib().nb_create(C.byref(nbp), _np(pb), device, C.byref(self._h)), "nb_create")

    def close(self) -> None:
        if self._h:
            lib().nb_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def launch_count(self) -> int:
        return int(lib().nb_launch_count(self._h))

    def set_static(self, st_ptr: np.ndarray, st_xy: np.ndarray, strep: np.ndarray | None = None) -> None:
        """``setStaticObstVert`` (+ ``setStaticObstRep``) with the inflated static hulls."""
        st_ptr = np.ascontiguousarray(st_ptr, np.int64)
        st_xy = np.ascontiguousarray(st_xy, np.float64)
        strep = None if strep is None or len(strep) == 0 else np.ascontiguousarray(strep, np.float64)
        _check(lib().nb_set_static(self._h, _np(st_ptr), _np(st_xy), _np(strep)), "nb_set_static")

    def replan(self, batch: ReplanBatch, with_lines: bool = True, stream=None) -> ReplanResult:
        """setInitTrajectory .. optimize for every agent of the batch, host buffers in and out."""
        res = ReplanResult.empty(batch, with_lines)
        a = host_args(batch, res)
        _check(lib().nb_replan_batch(self._h, C.byref(a), _P(stream or 0)), "nb_replan_batch")
        return res

    # ---- front end (K0)
    def search_configure(self, par: Params | None = None) -> None:
        """One-time setters of ``KinodynamicSearch`` (``neptune.cpp:88-97``)."""
        sp = make_search_params(par or self.par)
        _check(lib().nb_search_configure(self._h, C.byref(sp)), "nb_search_configure")

    def set_static_longest(self, longest: np.ndarray) -> None:
        """``staticObsLongestDist`` of ``setStaticObstRep``."""
        if self.par.num_of_static_obst:
            longest = np.ascontiguousarray(longest, np.float64)
            _check(lib().nb_set_static_longest(self._h, _np(longest)), "nb_set_static_longest")

    def search(self, sb, stream=None):
        """``KinodynamicSearch::setUp`` + ``run`` for every agent of a SearchBatch, host buffers in and out."""
        from .search import SearchResult
        res = SearchResult.empty(sb)
        a = host_search_args(sb, res)
        _check(lib().nb_search_batch(self._h, C.byref(a), _P(stream or 0)), "nb_search_batch")
        return res

    def search_args(self, a: "NbSearchArgs", stream=None) -> None:
        _check(lib().nb_search_batch(self._h, C.byref(a), _P(stream or 0)), "nb_search_batch")

    def replan_args(self, a: NbReplanArgs, stream=None) -> None:
        """Raw call with caller-built arguments (device pointers for the HBM-resident path)."""
        _check(lib().nb_replan_batch(self._h, C.byref(a), _P(stream or 0)), "nb_replan_batch")

    def separate(self, a_ptr, a_xy, b_ptr, b_xy, a_polygon: bool):
        """L separating-line LPs: ``separator::Separator::solveModel`` 2-D, batched."""
        a_ptr, b_ptr = np.ascontiguousarray(a_ptr, np.int64), np.ascontiguousarray(b_ptr, np.int64)
        a_xy, b_xy = np.ascontiguousarray(a_xy, np.float64), np.ascontiguousarray(b_xy, np.float64)
        L = len(a_ptr) - 1
        line, ok = np.zeros((L, 3)), np.zeros(L, np.uint8)
        f = lib().nb_separate_batch
        f.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_int32, _P, _P, _P]
        _check(f(self._h, L, NB_HOST, _np(a_ptr), _np(a_xy), _np(b_ptr), _np(b_xy), int(a_polygon), _np(line), _np(ok),
                 None), "nb_separate_batch")
        return ok.astype(bool), line

    def generate_traj(self, n_int, coeff, dc: float):
        """``generatePwpOut`` sampling loop, batched -> (states [B][K][12], n_states [B])."""
        n_int, coeff = np.ascontiguousarray(n_int, np.int32), np.ascontiguousarray(coeff, np.float64)
        B = len(n_int)
        mx = int(self.par.num_pol * self.par.T_span / dc) + 8
        states, ns = np.zeros((B, mx, 12)), np.zeros(B, np.int32)
        f = lib().nb_generate_traj_batch
        f.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, C.c_double, C.c_int32, _P, _P, _P]
        _check(f(self._h, B, NB_HOST, _np(n_int), _np(coeff), dc, mx, _np(states), _np(ns), None),
               "nb_generate_traj_batch")
        return states, ns

    # ---- entanglement-signature chain (K3)
    def entangle_predict(self, agent_id, known, bp_cnt, bp_xy, state: EntArrays, prev_pos, prev_pos_agent, cur, samp0):
        """``Neptune::PredictAlphasBetas``: returns entangle_state_A for every agent (input untouched)."""
        st = state.copy()
        arrs = [np.ascontiguousarray(agent_id, np.int32), np.ascontiguousarray(known, np.uint8),
                np.ascontiguousarray(bp_cnt, np.int32), np.ascontiguousarray(bp_xy, np.float64),
                np.ascontiguousarray(prev_pos, np.float64), np.ascontiguousarray(prev_pos_agent, np.float64),
                np.ascontiguousarray(cur, np.float64), np.ascontiguousarray(samp0, np.float64)]
        f = lib().nb_entangle_predict_batch
        f.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, NbEntState, _P, _P, _P, _P, _P]
        _check(f(self._h, len(arrs[0]), NB_HOST, _np(arrs[0]), _np(arrs[1]), _np(arrs[2]), _np(arrs[3]), st.c(),
                 _np(arrs[4]), _np(arrs[5]), _np(arrs[6]), _np(arrs[7]), None), "nb_entangle_predict_batch")
        return st

    def entangle_track(self, agent_id, bp_cnt, bp_xy, bp_cnt_prev, bp_xy_prev, state: EntArrays, prev_pos, prev_pos_agent,
                       latest, cur, elapsed_ms):
        """``NeptuneRos::updateEntStateStaticObs`` (one tick of the online tracker): returns
        (result [B], entangle_state_, previousCheckingPos_, previousCheckingPosAgent_), inputs untouched."""
        st = state.copy()
        pp, ppa = np.ascontiguousarray(prev_pos, np.float64).copy(), np.ascontiguousarray(prev_pos_agent, np.float64).copy()
        arrs = [np.ascontiguousarray(agent_id, np.int32), np.ascontiguousarray(bp_cnt, np.int32),
                np.ascontiguousarray(bp_xy, np.float64), np.ascontiguousarray(bp_cnt_prev, np.int32),
                np.ascontiguousarray(bp_xy_prev, np.float64), np.ascontiguousarray(latest, np.float64),
                np.ascontiguousarray(cur, np.float64), np.ascontiguousarray(elapsed_ms, np.float64)]
        res = np.zeros(len(arrs[0]), np.int32)
        f = lib().nb_entangle_track_batch
        f.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, NbEntState, _P, _P, _P, _P, _P, _P, _P]
        _check(f(self._h, len(arrs[0]), NB_HOST, _np(arrs[0]), _np(arrs[1]), _np(arrs[2]), _np(arrs[3]), _np(arrs[4]), st.c(),
                 _np(pp), _np(ppa), _np(arrs[5]), _np(arrs[6]), _np(arrs[7]), _np(res), None), "nb_entangle_track_batch")
        return res, st, pp, ppa

    def entangle_rollout(self, agent_id, known, bp_cnt, bp_xy, state: EntArrays, n_int, coeff, samp, samp_shared=False):
        """Front-end chain along a path: (done [B], states after 0..n intervals as EntArrays [B][9])."""
        B = len(agent_id)
        out = EntArrays(self.par, (B, NPOL + 1))
        done = np.zeros(B, np.int32)
        arrs = [np.ascontiguousarray(agent_id, np.int32), np.ascontiguousarray(known, np.uint8),
                np.ascontiguousarray(bp_cnt, np.int32), np.ascontiguousarray(bp_xy, np.float64),
                np.ascontiguousarray(n_int, np.int32), np.ascontiguousarray(coeff, np.float64),
                np.ascontiguousarray(samp, np.float64)]
        f = lib().nb_entangle_rollout_batch
        f.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, NbEntState, _P, _P, _P, C.c_int32, NbEntState, _P, _P]
        _check(f(self._h, B, NB_HOST, _np(arrs[0]), _np(arrs[1]), _np(arrs[2]), _np(arrs[3]), state.c(), _np(arrs[4]),
                 _np(arrs[5]), _np(arrs[6]), int(samp_shared), out.c(), _np(done), None), "nb_entangle_rollout_batch")
        return done, out

    def entangle_check(self, agent_id, known, bp_cnt, bp_xy, state: EntArrays, n_int, coeff, samp, samp_shared=False):
        """``entangleCheckGivenPwp``: (entangled [B], updated state)."""
        B = len(agent_id)
        st = state.copy()
        ent = np.zeros(B, np.int32)
        arrs = [np.ascontiguousarray(agent_id, np.int32), np.ascontiguousarray(known, np.uint8),
                np.ascontiguousarray(bp_cnt, np.int32), np.ascontiguousarray(bp_xy, np.float64),
                np.ascontiguousarray(n_int, np.int32), np.ascontiguousarray(coeff, np.float64),
                np.ascontiguousarray(samp, np.float64)]
        f = lib().nb_entangle_check_batch
        f.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, NbEntState, _P, _P, _P, C.c_int32, _P, _P]
        _check(f(self._h, B, NB_HOST, _np(arrs[0]), _np(arrs[1]), _np(arrs[2]), _np(arrs[3]), st.c(), _np(arrs[4]),
                 _np(arrs[5]), _np(arrs[6]), int(samp_shared), _np(ent), None), "nb_entangle_check_batch")
        return ent, st


def solver_postcheck_entangle(self, agent_id, known, late, bp_cnt, bp_xy, bp_cnt_late, bp_xy_late, state: EntArrays, prev_pos,
                             prev_pos_agent, cur, n_int, coeff, t_start, samp, late_recs):
    """Entanglement half of ``Neptune::safetyCheckAfterReplan`` (nb_postcheck_entangle_batch): entangled [B]."""
    arrs = [np.ascontiguousarray(x, dt) for x, dt in
            ((agent_id, np.int32), (known, np.uint8), (late, np.uint8), (bp_cnt, np.int32), (bp_xy, np.float64),
             (bp_cnt_late, np.int32), (bp_xy_late, np.float64), (prev_pos, np.float64), (prev_pos_agent, np.float64),
             (cur, np.float64), (n_int, np.int32), (coeff, np.float64), (t_start, np.float64), (samp, np.float64),
             (late_recs, np.float64))]
    B = len(arrs[0])
    ent = np.zeros(B, np.int32)
    f = lib().nb_postcheck_entangle_batch
    f.argtypes = [_P, C.c_int32, C.c_int32] + [_P] * 7 + [NbEntState] + [_P] * 7 + [C.c_int32, _P, _P, _P, _P]
    p = [_np(a) for a in arrs]
    _check(f(self._h, B, NB_HOST, p[0], p[1], p[2], p[3], p[4], p[5], p[6], state.c(), p[7], p[8], p[9], p[10], p[11], p[12],
             p[13], 0, None, p[14], _np(ent), None), "nb_postcheck_entangle_batch")
    return ent


def solver_unpack_records(self, recs):
    """``NeptuneRos::trajCB`` bookkeeping for all records: (bp_cnt [N], bp_xy [N][bp_max][2], latest_pos [N][2])."""
    recs = np.ascontiguousarray(recs, np.float64)
    N, bm = self.par.num_of_agents, self.par.bp_max
    cnt, xy, pos = np.zeros(N, np.int32), np.zeros((N, bm, 2)), np.zeros((N, 2))
    f = lib().nb_unpack_records_batch
    f.argtypes = [_P, C.c_int32, _P, _P, _P, _P, _P]
    _check(f(self._h, NB_HOST, _np(recs), _np(cnt), _np(xy), _np(pos), None), "nb_unpack_records_batch")
    return cnt, xy, pos


def solver_set_static_rep_per_agent(self, strep_all, longest_all=None):
    """One staticObsRep_ (and staticObsLongestDist_) per agent: [N][M][2][2], [N][M][2]."""
    sa = np.ascontiguousarray(strep_all, np.float64)
    la = None if longest_all is None else np.ascontiguousarray(longest_all, np.float64)
    f = lib().nb_set_static_rep_per_agent
    f.argtypes = [_P, _P, _P]
    _check(f(self._h, _np(sa), _np(la)), "nb_set_static_rep_per_agent")


NB_REC_PWP_DOUBLES = 1 + 17 + 3 * 16 * 4   # trajectory part of a record
NB_REC_DOUBLES = 256                         # record stride: trajectory + DynTraj header (include/neptune_b200.h)
REC_ID, REC_ISAGENT, REC_BBOX, REC_POS, REC_NBEND, REC_BEND, REC_SEQ = 210, 211, 212, 215, 218, 219, 235
NB_HULL_STRIDE = 24


def make_records(committed, par: Params | None = None, bp_cnt=None, bp_xy=None, seq: int = 0) -> np.ndarray:
    """Committed-trajectory records (the payload of the per-cycle exchange) from (times, cx, cy, cz) tuples; with
    `par` the DynTraj header is filled as NeptuneRos::publishOwnTraj fills the message (neptune_ros.cpp:436-480):
    id, is_agent, bbox = 2 drone_radius, pos = first point of the trajectory, bendpt[] = the agent's tether."""
    recs = np.zeros((len(committed), NB_REC_DOUBLES))
    for j, (tm, cx, cy, cz) in enumerate(committed):
        n = len(cx)
        assert n <= 16
        recs[j, 0] = n
        recs[j, 1:2 + n] = tm
        co = recs[j, 18:NB_REC_PWP_DOUBLES].reshape(3, 16, 4)
        co[0, :n], co[1, :n], co[2, :n] = cx, cy, cz
        if par is not None:
            recs[j, REC_ID], recs[j, REC_ISAGENT], recs[j, REC_SEQ] = j + 1, 1.0, seq
            recs[j, REC_BBOX:REC_BBOX + 3] = 2.0 * par.drone_radius
            recs[j, REC_POS:REC_POS + 3] = [cx[0][3], cy[0][3], cz[0][3]]
            if bp_cnt is not None:
                nb = int(bp_cnt[j])
                assert nb <= 8
                recs[j, REC_NBEND] = nb
                recs[j, REC_BEND:REC_BEND + 2 * nb] = np.asarray(bp_xy[j, :nb], np.float64).reshape(-1)
    return recs


def _hulls_impl(fn, par, t_start, recs, known, delta, want_idx=True):
    B, N, S = len(t_start), par.num_of_agents, par.num_sample_per_interval
    t_start, recs = np.ascontiguousarray(t_start, np.float64), np.ascontiguousarray(recs, np.float64)
    known = np.ascontiguousarray(known, np.uint8)
    out = dict(hull_xy=np.zeros((B, N, NPOL, NB_HULL_STRIDE, 2)), hull_cnt=np.zeros((B, N, NPOL), np.int32),
               hull_ptr=np.zeros(B * N * NPOL, np.int64), nih0=np.zeros((B, N, NPOL, 2)),
               samp=np.zeros((B, N, par.num_pol, S + 1, 2)), idx=np.zeros((B, N, NPOL, 2), np.int32))
    fn(B, _np(t_start), _np(recs), _np(known), C.c_double(delta), _np(out["hull_xy"]), _np(out["hull_cnt"]),
       _np(out["hull_ptr"]), _np(out["nih0"]), _np(out["samp"]), _np(out["idx"]) if want_idx else None)
    return out


def solver_hulls(self, t_start, recs, known, delta):
    """K1: hulls, nih0 and samples of every committed trajectory for every planning agent."""
    f = lib().nb_hulls_batch
    f.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, _P, C.c_double, _P, _P, _P, _P, _P, _P, _P]

    def call(B, ts, rc, kn, dl, hx, hc, hp, n0, sm, ix):
        _check(f(self._h, B, NB_HOST, ts, rc, kn, dl, hx, hc, hp, n0, sm, ix, None), "nb_hulls_batch")
    return _hulls_impl(call, self.par, t_start, recs, known, delta)


def solver_postcheck(self, n_int, coeff, t_start, recs, late, delta):
    """K5: ``trajsAndPwpAreInCollision2d`` of every agent's optimised pwp against the late trajectories."""
    n_int, coeff = np.ascontiguousarray(n_int, np.int32), np.ascontiguousarray(coeff, np.float64)
    t_start, recs = np.ascontiguousarray(t_start, np.float64), np.ascontiguousarray(recs, np.float64)
    late = np.ascontiguousarray(late, np.uint8)
    col = np.zeros(len(n_int), np.int32)
    f = lib().nb_postcheck_batch
    f.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_double, _P, _P]
    _check(f(self._h, len(n_int), NB_HOST, _np(n_int), _np(coeff), _np(t_start), _np(recs), _np(late), delta, _np(col),
             None), "nb_postcheck_batch")
    return col


def solver_compose(self, t, has_prev, prev, now):
    """``mu::composePieceWisePol`` on committed-trajectory records -> (n_pieces [B], out [B][256])."""
    arrs = [np.ascontiguousarray(t, np.float64), np.ascontiguousarray(has_prev, np.uint8),
            np.ascontiguousarray(prev, np.float64), np.ascontiguousarray(now, np.float64)]
    out, npc = np.zeros_like(arrs[3]), np.zeros(len(arrs[0]), np.int32)
    f = lib().nb_compose_records_batch
    f.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P]
    _check(f(self._h, len(arrs[0]), NB_HOST, _np(arrs[0]), _np(arrs[1]), _np(arrs[2]), _np(arrs[3]), _np(out), _np(npc),
             None), "nb_compose_records_batch")
    return npc, out


Solver.postcheck_entangle = solver_postcheck_entangle
Solver.unpack_records = solver_unpack_records
Solver.set_static_rep_per_agent = solver_set_static_rep_per_agent
Solver.compose = solver_compose
Solver.hulls = solver_hulls
Solver.postcheck = solver_postcheck


class DeviceEntBackend:
    """Entanglement back end for ``neptune_b200.scenes.fill_entangle`` running on the product's K3
    kernels (used by bench.py so that the workload generator never touches the oracle)."""

    def __init__(self, solver: Solver):
        self.s = solver

    def predict_batch(self, par, agent_id, prev_pos, prev_pos_agent, cur, samp0, known, strep, bp_cnt, bp_xy,
                      cnt, alpha, beta, bend, active):
        st = self.s.entangle_predict(agent_id, known, bp_cnt, bp_xy, EntArrays.of(par, cnt, alpha, beta, bend, active),
                                     prev_pos, prev_pos_agent, cur, samp0)
        return st.tuple()

    def rollout_batch(self, par, agent_id, n_int, coeff, samp, known, strep, bp_cnt, bp_xy, cnt, alpha, beta, bend,
                      active):
        done, out = self.s.entangle_rollout(agent_id, known, bp_cnt, bp_xy,
                                            EntArrays.of(par, cnt, alpha, beta, bend, active), n_int, coeff, samp)
        return (done,) + out.tuple()

    def check_batch(self, par, agent_id, n_int, coeff, samp, known, strep, bp_cnt, bp_xy, cnt, alpha, beta, bend, active):
        ent, st = self.s.entangle_check(agent_id, known, bp_cnt, bp_xy,
                                        EntArrays.of(par, cnt, alpha, beta, bend, active), n_int, coeff, samp)
        return (ent,) + st.tuple()

"""Entanglement back ends for neptune_b200.scenes.fill_entangle used by the tests: the CPU oracle and
the single-lane emulation of the device code.  (bench.py uses neptune_b200.capi.DeviceEntBackend.)"""
from __future__ import annotations

import ctypes as C

import numpy as np

from neptune_b200.batch import NPOL
from neptune_b200.capi import EntArrays, make_nb_params


class OracleEntBackend:
    def __init__(self, orc):
        self.orc = orc

    def _ctx(self, par, self_idx, strep, bp_cnt, bp_xy):
        return self.orc.EntCtx(par, self_idx, strep, bp_cnt, bp_xy)

    def _state(self, par, cnt, alpha, beta, bend, active, b):
        es = self.orc.EntState(par.ent_cap, par.NA)
        es.n_alpha, es.n_bend = int(cnt[b, 0]), int(cnt[b, 1])
        es.alpha[:], es.beta[:], es.bend[:], es.active[:] = alpha[b], beta[b], bend[b], active[b]
        return es

    def predict_batch(self, par, agent_id, prev_pos, prev_pos_agent, cur, samp0, known, strep, bp_cnt, bp_xy,
                      cnt, alpha, beta, bend, active):
        out = EntArrays(par, len(agent_id))
        for b in range(len(agent_id)):
            es = self._state(par, cnt, alpha, beta, bend, active, b)
            cx = self._ctx(par, int(agent_id[b]) - 1, strep, bp_cnt, bp_xy)
            rc = self.orc.predict(es, cx, prev_pos[b], prev_pos_agent[b], cur[b], samp0[b], known[b])
            assert rc == 0
            out.cnt[b] = [es.n_alpha, es.n_bend]
            out.alpha[b], out.beta[b], out.bend[b], out.active[b] = es.alpha, es.beta, es.bend, es.active
        return out.tuple()

    def rollout_batch(self, par, agent_id, n_int, coeff, samp, known, strep, bp_cnt, bp_xy, cnt, alpha, beta, bend,
                      active):
        B = len(agent_id)
        out = EntArrays(par, (B, NPOL + 1))
        done = np.zeros(B, np.int32)
        for b in range(B):
            n = int(n_int[b])
            es = self._state(par, cnt, alpha, beta, bend, active, b)
            cx = self._ctx(par, int(agent_id[b]) - 1, strep, bp_cnt, bp_xy)
            cxy = np.ascontiguousarray(coeff[b, :2, :n, :])
            d, c, a, be, bd, ac = self.orc.entangle_rollout(es, cx, n, cxy, samp[b], known[b])
            done[b] = d
            out.cnt[b, :n + 1], out.alpha[b, :n + 1], out.beta[b, :n + 1] = c, a, be
            out.bend[b, :n + 1], out.active[b, :n + 1] = bd, ac
            out.cnt[b, n + 1:], out.alpha[b, n + 1:], out.beta[b, n + 1:] = c[n], a[n], be[n]
            out.bend[b, n + 1:], out.active[b, n + 1:] = bd[n], ac[n]
        return (done,) + out.tuple()

    def check_batch(self, par, agent_id, n_int, coeff, samp, known, strep, bp_cnt, bp_xy, cnt, alpha, beta, bend, active):
        B = len(agent_id)
        out = EntArrays(par, B)
        ent = np.zeros(B, np.int32)
        for b in range(B):
            n = int(n_int[b])
            es = self._state(par, cnt, alpha, beta, bend, active, b)
            cx = self._ctx(par, int(agent_id[b]) - 1, strep, bp_cnt, bp_xy)
            cxy = np.ascontiguousarray(coeff[b, :2, :n, :])
            ent[b] = self.orc.entangle_check_pwp(es, cx, n, cxy, samp[b], known[b])
            out.cnt[b] = [es.n_alpha, es.n_bend]
            out.alpha[b], out.beta[b], out.bend[b], out.active[b] = es.alpha, es.beta, es.bend, es.active
        return (ent,) + out.tuple()


class EmulEntBackend:
    """Device entanglement code compiled for one host lane (tests/emul)."""

    def __init__(self):
        from tests.emul import emul
        self.lib = emul.lib()

    def _call(self, par, mode, agent_id, known, strep, bp_cnt, bp_xy, st, out, n_int=None, coeff=None, samp=None,
              prev_pos=None, prev_pos_agent=None, cur=None, samp0=None):
        P = C.c_void_p
        keep = [np.ascontiguousarray(x, dt) if x is not None else None for x, dt in
                ((agent_id, np.int32), (known, np.uint8), (bp_cnt, np.int32), (bp_xy, np.float64), (n_int, np.int32),
                 (coeff, np.float64), (samp, np.float64), (prev_pos, np.float64), (prev_pos_agent, np.float64),
                 (cur, np.float64), (samp0, np.float64), (par.pb, np.float64),
                 (strep if len(strep) else np.zeros((1, 2, 2)), np.float64))]
        p = [None if k is None else k.ctypes.data_as(P) for k in keep]
        B = len(agent_id)
        result = np.zeros(B, np.int32)
        nbp = make_nb_params(par)
        f = self.lib.emul_entangle
        from neptune_b200.capi import NbEntState
        f.argtypes = [P, P, P, C.c_int, C.c_int, P, P, P, P, NbEntState, NbEntState, P, P, P, C.c_int, P, P, P, P, P]
        rc = f(C.addressof(nbp), p[11], p[12], mode, B, p[0], p[1], p[2], p[3], st.c(), out.c(), p[4], p[5], p[6], 0,
               p[7], p[8], p[9], p[10], result.ctypes.data_as(P))
        assert rc == 0, rc
        return result

    def predict_batch(self, par, agent_id, prev_pos, prev_pos_agent, cur, samp0, known, strep, bp_cnt, bp_xy,
                      cnt, alpha, beta, bend, active):
        st = EntArrays.of(par, cnt.copy(), alpha.copy(), beta.copy(), bend.copy(), active.copy())
        self._call(par, 0, agent_id, known, strep, bp_cnt, bp_xy, st, st, prev_pos=prev_pos,
                   prev_pos_agent=prev_pos_agent, cur=cur, samp0=samp0)
        return st.tuple()

    def rollout_batch(self, par, agent_id, n_int, coeff, samp, known, strep, bp_cnt, bp_xy, cnt, alpha, beta, bend,
                      active):
        st = EntArrays.of(par, cnt.copy(), alpha.copy(), beta.copy(), bend.copy(), active.copy())
        out = EntArrays(par, (len(agent_id), NPOL + 1))
        done = self._call(par, 1, agent_id, known, strep, bp_cnt, bp_xy, st, out, n_int=n_int, coeff=coeff, samp=samp)
        return (done,) + out.tuple()

    def check_batch(self, par, agent_id, n_int, coeff, samp, known, strep, bp_cnt, bp_xy, cnt, alpha, beta, bend, active):
        st = EntArrays.of(par, cnt.copy(), alpha.copy(), beta.copy(), bend.copy(), active.copy())
        ent = self._call(par, 2, agent_id, known, strep, bp_cnt, bp_xy, st, st, n_int=n_int, coeff=coeff, samp=samp)
        return (ent,) + st.tuple()

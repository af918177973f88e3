"""One replan cycle for the agents of one rank: host-side mirror of ``Neptune::replanFull``'s hot part
(reference ``neptune/src/neptune.cpp:1430-1448`` hulls/samples + PredictAlphasBetas, ``:1512-1529``
back end, ``:1641-1647`` post-check) and of the inter-agent exchange (ROS topic ``/trajs``,
``neptune_ros.cpp:172-179, :434-480``) as ONE all-gather of committed-trajectory records per cycle.

PyTorch is used for device memory, streams and ``torch.distributed`` only; all math runs in
libneptune_b200.so through the C-ABI with device pointers.  With ``front_end=True`` the kinodynamic search
(``KinodynamicSearch::run``, ``neptune.cpp:1453``; K0) runs on the device too and its ``pwp_init`` /
``entStateVec`` feed the back end in place; otherwise they are inputs of the cycle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .batch import NPOL
from .params import Params

REC = capi.NB_REC_DOUBLES
HS = capi.NB_HULL_STRIDE


def shard_agents(n_agents: int, world: int, rank: int) -> np.ndarray:
    """Block partition of agent indices (0-based) over ranks; the last ranks may own one agent less."""
    base, extra = divmod(n_agents, world)
    start = rank * base + min(rank, extra)
    return np.arange(start, start + base + (1 if rank < extra else 0))


def gather_records(local, world: int, group=None, sizes=None):
    """All-gather of fixed-stride committed-trajectory records: [B_r][REC] per rank -> [sum B_r][REC]
    in rank order.  Works on CUDA tensors (NCCL over NVLink) and CPU tensors (gloo, used by the tests).
    sizes: per-rank record counts when the shards are uneven."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    local = local.contiguous()
    if sizes is None or len(set(sizes)) == 1:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    # uneven shards: pad to the largest shard (collectives need equal sizes), gather, drop the padding
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)], dim=0)


class ReplanCycle:
    """Device-resident state and launch sequence of one rank."""

    def __init__(self, par: Params, agents: np.ndarray, device, static=None, world: int = 1, group=None,
                 front_end: bool = False):
        import torch
        self.torch = torch
        self.front_end = front_end
        self.par, self.agents, self.world, self.group = par, np.asarray(agents), world, group
        self.dev = torch.device(device)
        self.B, self.N = len(agents), par.num_of_agents
        self.solver = capi.Solver(par, device=self.dev.index or 0)
        if par.num_of_static_obst:
            st_ptr, st_xy, strep = static[:3]
            self.solver.set_static(st_ptr, st_xy, strep)
            if front_end:
                self.solver.set_static_longest(static[3])
        if front_end:
            self.solver.search_configure()
        B, N, S, cap, NA = self.B, self.N, par.num_sample_per_interval, par.ent_cap, par.NA
        f64, i32, u8, i64 = torch.float64, torch.int32, torch.uint8, torch.int64

        def z(shape, dt):
            return torch.zeros(shape, dtype=dt, device=self.dev)
        # inputs of a cycle: views into ONE packed device buffer, mirrored by one pinned host buffer, so that
        # the end-to-end path pays one H2D copy per cycle instead of twenty small ones
        np_of = {f64: np.float64, i32: np.int32, u8: np.uint8}
        spec = dict(
            agent_id=(B, i32), n_int=(B, i32), coeff_init=((B, 3, NPOL, 4), f64), t_start=(B, f64),
            recs=((N, REC), f64), late_recs=((N, REC), f64), known=((B, N), u8), late=((B, N), u8),
            esv_cnt=((B, NPOL + 1, 2), i32), esv_alpha=((B, NPOL + 1, cap, 2), i32), esv_active=((B, NPOL + 1, NA), i32),
            bp_cnt=(N, i32), bp_xy=((N, par.bp_max, 2), f64),
            es_cnt=((B, 2), i32), es_alpha=((B, cap, 2), i32), es_beta=((B, cap), f64), es_bend=((B, cap), i32),
            es_active=((B, NA), i32), prev_pos=((B, N + 1, 2), f64), prev_pos_agent=((B, N, 2), f64), cur=((B, 2), f64),
            t_group=(B, f64), group=(B, i32), t_now=(B, f64))
        if front_end:   # start state A, goal, initial z polynomial, order of the jerk samples
            spec.update(fe_init=((B, 6), f64), fe_goal=((B, 2), f64), fe_coeffs_z=((B, NPOL, 4), f64),
                        fe_comb=((B, par.a_star_samp_x ** 2), u8))
        self.layout, off = {}, 0
        for k, (shape, dt) in spec.items():
            shape = (shape,) if isinstance(shape, int) else tuple(shape)
            nbytes = int(np.prod(shape)) * np.dtype(np_of[dt]).itemsize
            self.layout[k] = (off, nbytes, shape, dt)
            off += (nbytes + 255) // 256 * 256
        self.in_bytes = off
        self.in_buf = torch.zeros(off, dtype=u8, device=self.dev)
        self.d = {k: self.in_buf[o:o + nb].view(dt).view(shape) for k, (o, nb, shape, dt) in self.layout.items()}
        self._np_of = np_of
        # intermediates and outputs (group-shaped buffers are sized by _ensure_groups)
        self.o = dict(
            samp0=z((B, N, 2), f64),
            esA_cnt=z((B, 2), i32), esA_alpha=z((B, cap, 2), i32), esA_beta=z((B, cap), f64), esA_bend=z((B, cap), i32),
            esA_active=z((B, NA), i32),
            esC_cnt=z((B, 2), i32), esC_alpha=z((B, cap, 2), i32), esC_beta=z((B, cap), f64), esC_bend=z((B, cap), i32),
            esC_active=z((B, NA), i32),
            new_recs=z((B, REC), f64), new_pieces=z(B, i32))
        if front_end:   # outputs of the search, laid out like the back end's inputs
            self.o.update(fe_status=z(B, i32), fe_solved=z(B, i32), fe_n_int=z(B, i32), fe_coeff=z((B, 3, NPOL, 4), f64),
                          fe_esv_cnt=z((B, NPOL + 1, 2), i32), fe_esv_alpha=z((B, NPOL + 1, cap, 2), i32),
                          fe_esv_beta=z((B, NPOL + 1, cap), f64), fe_esv_bend=z((B, NPOL + 1, cap), i32),
                          fe_esv_active=z((B, NPOL + 1, NA), i32), fe_stats=z((B, 4), i32), fe_cost=z(B, f64))
        self.G = 0
        # packed outputs (one D2H copy)
        ospec = dict(coeff_out=((B, 3, NPOL, 4), f64), obj=(B, f64), status=(B, i32), iters=((B, 2), i32),
                     entangled=(B, i32), collide=(B, i32))
        self.olayout, off = {}, 0
        for k, (shape, dt) in ospec.items():
            shape = (shape,) if isinstance(shape, int) else tuple(shape)
            nbytes = int(np.prod(shape)) * np.dtype(np_of[dt]).itemsize
            self.olayout[k] = (off, nbytes, shape, dt)
            off += (nbytes + 255) // 256 * 256
        self.out_bytes = off
        self.out_buf = torch.zeros(off, dtype=u8, device=self.dev)
        for k, (o_, nb, shape, dt) in self.olayout.items():
            self.o[k] = self.out_buf[o_:o_ + nb].view(dt).view(shape)
        self.gathered = None
        self.delta = 2.0 * par.drone_radius  # bbox/2 + drone_radius with bbox = 2 drone_radius (neptune_ros.cpp:447-449)
        self._lib = capi.lib()
        self._sig()
        self.profile = False   # True: one stream, synchronise and time every stage of step() with CUDA events
        self.stage_ms = {}
        self.overlap = True    # independent stages on side streams (predict; late-trajectory hulls)
        self.side = [torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)] \
            if self.dev.type == "cuda" else []

    def _ensure_groups(self, G: int):
        """Buffers shaped by the number of distinct t_start values (window groups)."""
        if G == self.G:
            return
        torch, N, S = self.torch, self.N, self.par.num_sample_per_interval
        f64, i32, u8, i64 = torch.float64, torch.int32, torch.uint8, torch.int64

        def z(shape, dt):
            return torch.zeros(shape, dtype=dt, device=self.dev)
        self.G = G
        self.d["ones_g"] = torch.ones((G, N), dtype=u8, device=self.dev)
        for tag in ("", "_l"):   # planning-time trajectories, late trajectories
            self.o["hull_xy_g" + tag] = z((G, N, NPOL, HS, 2), f64)
            self.o["hull_cnt_g" + tag] = z((G, N, NPOL), i32)
            self.o["hull_ptr_g" + tag] = z(G * N * NPOL, i64)
            self.o["nih0_g" + tag] = z((G, N, NPOL, 2), f64)
            self.o["samp_g" + tag] = z((G, N, self.par.num_pol, S + 1, 2), f64)
        if G > 1:
            self.o["samp_b"] = z((self.B, N, self.par.num_pol, S + 1, 2), f64)

    def _sig(self):
        P, L = C.c_void_p, self._lib
        L.nb_hull_index_batch.argtypes = [P, C.c_int32, C.c_int32, P, P, P, P, P, P, P]
        L.nb_postcheck_hulls_batch.argtypes = [P, C.c_int32, C.c_int32, P, P, P, P, P, P, P, P]
        L.nb_hulls_batch.argtypes = [P, C.c_int32, C.c_int32, P, P, P, C.c_double, P, P, P, P, P, P, P]
        L.nb_entangle_predict_batch.argtypes = [P, C.c_int32, C.c_int32, P, P, P, P, capi.NbEntState, P, P, P, P, P]
        L.nb_entangle_check_batch.argtypes = [P, C.c_int32, C.c_int32, P, P, P, P, capi.NbEntState, P, P, P, C.c_int32, P, P]
        L.nb_postcheck_batch.argtypes = [P, C.c_int32, C.c_int32, P, P, P, P, P, C.c_double, P, P]
        L.nb_commit_compose_batch.argtypes = [P, C.c_int32, C.c_int32] + [P] * 13

    # ------------------------------------------------------------------ host <-> device
    DELTA_T_STEPS = 2   # t_start - time_now in units of dc (deltaT_ of neptune.cpp:1236-1262 for the synthetic world)
    OUT_KEYS = ("coeff_out", "obj", "status", "iters", "entangled", "collide")

    def host_inputs(self, scene, fe: dict | None = None) -> dict:
        """Pinned host copy of everything a cycle needs from the planner core / front end, laid out exactly
        like the packed device buffer.  Returns {"buf": pinned uint8 tensor, "G": groups, <name>: numpy views}."""
        torch = self.torch
        b = scene.batch
        recs = capi.make_records(scene.committed)
        uniq, inv = np.unique(np.asarray(scene.t_start, np.float64), return_inverse=True)
        tg = np.zeros(self.B)
        tg[:len(uniq)] = uniq
        src = dict(agent_id=b.agent_id, n_int=b.n_int, coeff_init=b.coeff_init, t_start=scene.t_start,
                   recs=recs, late_recs=recs, known=scene.known, late=scene.known,
                   esv_cnt=b.esv_cnt, esv_alpha=b.esv_alpha, esv_active=b.esv_active, bp_cnt=b.bp_cnt, bp_xy=b.bp_xy,
                   es_cnt=scene.es0_cnt, es_alpha=scene.es0_alpha, es_beta=scene.es0_beta, es_bend=scene.es0_bend,
                   es_active=scene.es0_active, prev_pos=scene.prev_pos, prev_pos_agent=scene.prev_pos_agent,
                   cur=np.ascontiguousarray(scene.state_A[:, 0, :2]), t_group=tg, group=inv.astype(np.int32),
                   t_now=np.asarray(scene.t_start, np.float64) - self.DELTA_T_STEPS * self.par.dc)
        if self.front_end:   # neptune_b200.scenes.search_host_inputs(scene, seed)
            src.update(fe_init=fe["init"], fe_goal=fe["goal"], fe_coeffs_z=fe["coeffs_z"], fe_comb=fe["comb"])
        buf = torch.zeros(self.in_bytes, dtype=torch.uint8)
        if torch.cuda.is_available():
            buf = buf.pin_memory()
        out = {"buf": buf, "G": len(uniq)}
        nb = buf.numpy()
        for k, (o, nbytes, shape, dt) in self.layout.items():
            view = nb[o:o + nbytes].view(self._np_of[dt]).reshape(shape)
            view[...] = np.asarray(src[k]).astype(self._np_of[dt]).reshape(shape)
            out[k] = view
        return out

    def upload(self, host: dict) -> int:
        """One H2D copy of the packed per-cycle inputs."""
        self._ensure_groups(int(host["G"]))
        self.in_buf.copy_(host["buf"], non_blocking=True)
        return self.in_bytes

    def download(self, host_out: dict) -> int:
        """One D2H copy of the packed results."""
        host_out["buf"].copy_(self.out_buf, non_blocking=True)
        return self.out_bytes

    def host_outputs(self) -> dict:
        buf = self.torch.zeros(self.out_bytes, dtype=self.torch.uint8).pin_memory()
        out = {"buf": buf}
        nb = buf.numpy()
        for k, (o, nbytes, shape, dt) in self.olayout.items():
            out[k] = nb[o:o + nbytes].view(self._np_of[dt]).reshape(shape)
        return out

    # ------------------------------------------------------------------ the cycle
    def capture(self):
        """Capture the launch sequence of one cycle (for the current window grouping) in a CUDA graph:
        the ~25 launches of a cycle are short, so replaying one graph removes the host launch cost."""
        torch = self.torch
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._step_impl()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        l0 = self.solver.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread may touch CUDA while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._step_impl()
        self.launches_per_cycle = self.solver.launch_count() - l0   # library kernels one replay launches
        self.graph_G = self.G

    def step(self, exchange: bool = True):
        """One cycle on the current CUDA stream (graph replay when captured), then the all-gather."""
        if getattr(self, "graph", None) is not None and self.graph_G == self.G and not self.profile:
            self.graph.replay()
        else:
            self._step_impl()
        if exchange and self.world > 1:
            sizes = [len(shard_agents(self.N, self.world, r)) for r in range(self.world)]
            self.gathered = gather_records(self.o["new_recs"], self.world, self.group, sizes)
        return self.o

    def _step_impl(self):
        """hulls/samples (K1) -> PredictAlphasBetas (K3) -> LPs + QP (K2, K4) -> post-check (K5, K3) ->
        commit records.  Everything is enqueued on the current CUDA stream (and two side streams)."""
        torch, L, h, d, o, B = self.torch, self._lib, self.solver.handle, self.d, self.o, self.B
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        DEV = capi.NB_DEVICE
        chk = capi._check
        marks = []

        def mark(name):
            if self.profile:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))
        mark("start")
        G = self.G
        main = torch.cuda.current_stream()
        par_streams = self.overlap and not self.profile and len(self.side) == 2
        sB, sC = (self.side if par_streams else (main, main))
        stB, stC = C.c_void_p(sB.cuda_stream), C.c_void_p(sC.cuda_stream)
        gl = d["group"].long()
        if par_streams:
            sC.wait_stream(main)
        # (stream C) hulls / samples of the late trajectories over the same windows (neptune.cpp:737-741, :792):
        # independent of the optimisation, so they overlap it
        with torch.cuda.stream(sC):
            chk(L.nb_hulls_batch(h, G, DEV, p(d["t_group"]), p(d["late_recs"]), p(d["ones_g"]), self.delta,
                                 p(o["hull_xy_g_l"]), p(o["hull_cnt_g_l"]), p(o["hull_ptr_g_l"]), p(o["nih0_g_l"]),
                                 p(o["samp_g_l"]), None, stC), "nb_hulls_batch")
            if G > 1:
                o["samp_b"].copy_(o["samp_g_l"].index_select(0, gl))
        mark("late_hulls")
        # (main) hulls / samples of the committed trajectories the agents plan against
        chk(L.nb_hulls_batch(h, G, DEV, p(d["t_group"]), p(d["recs"]), p(d["ones_g"]), self.delta, p(o["hull_xy_g"]),
                             p(o["hull_cnt_g"]), p(o["hull_ptr_g"]), p(o["nih0_g"]), p(o["samp_g"]), None, st), "nb_hulls_batch")
        mark("hulls_samples")
        # (stream B) entangle_state_A = PredictAlphasBetas(entangle_state_): needs the samples only
        if par_streams:
            sB.wait_stream(main)
        esA = capi.NbEntState()
        esA.cnt, esA.alpha, esA.beta, esA.bend, esA.active = (p(o["esA_" + k]) for k in ("cnt", "alpha", "beta", "bend", "active"))
        with torch.cuda.stream(sB):
            for k in ("cnt", "alpha", "beta", "bend", "active"):
                o["esA_" + k].copy_(d["es_" + k])
            o["samp0"].copy_(o["samp_g"][:, :, 0, 0, :].index_select(0, gl))
            chk(L.nb_entangle_predict_batch(h, B, DEV, p(d["agent_id"]), p(d["known"]), p(d["bp_cnt"]), p(d["bp_xy"]), esA,
                                            p(d["prev_pos"]), p(d["prev_pos_agent"]), p(d["cur"]), p(o["samp0"]), stB),
                "nb_entangle_predict_batch")
            for k in ("cnt", "alpha", "beta", "bend", "active"):   # copy for entangleCheckGivenPwp (works on a local)
                o["esC_" + k].copy_(o["esA_" + k])
        mark("predict")
        n_int_t, coeff_t = d["n_int"], d["coeff_init"]
        esv_t = (d["esv_cnt"], d["esv_alpha"], d["esv_active"])
        if self.front_end:
            # KinodynamicSearch::setUp + run (neptune.cpp:1450-1453) from entangle_state_A and the shared hulls /
            # samples; pwp_init and entStateVec stay on the device for the back end (:1509-1517)
            if par_streams:
                main.wait_stream(sB)
            sa = capi.NbSearchArgs()
            sa.B, sa.space = B, DEV
            sa.agent_id, sa.init, sa.goal, sa.coeffs_z = (d[k].data_ptr() for k in ("agent_id", "fe_init", "fe_goal", "fe_coeffs_z"))
            sa.n_groups, sa.group = G, d["group"].data_ptr()
            sa.hull_xy, sa.hull_cnt, sa.samp, sa.known = o["hull_xy_g"].data_ptr(), o["hull_cnt_g"].data_ptr(), o["samp_g"].data_ptr(), d["known"].data_ptr()
            sa.es = esA
            sa.bp_cnt, sa.bp_xy, sa.comb, sa.comb_shared = d["bp_cnt"].data_ptr(), d["bp_xy"].data_ptr(), d["fe_comb"].data_ptr(), 0
            sa.status, sa.solved, sa.n_int, sa.coeff = (o[k].data_ptr() for k in ("fe_status", "fe_solved", "fe_n_int", "fe_coeff"))
            esv = capi.NbEntState()
            esv.cnt, esv.alpha, esv.beta, esv.bend, esv.active = (p(o["fe_esv_" + k]) for k in ("cnt", "alpha", "beta", "bend", "active"))
            sa.esv = esv
            sa.stats, sa.cost = o["fe_stats"].data_ptr(), o["fe_cost"].data_ptr()
            chk(L.nb_search_batch(h, C.byref(sa), st), "nb_search_batch")
            # "returning with no solution" (neptune.cpp:1473-1478): the replan is rejected below; the back end
            # still gets a well-formed (host-provided) path for those agents so that the batch stays dense
            ok = o["fe_solved"] > 0
            n_int_t = torch.where(ok, o["fe_n_int"], d["n_int"])
            coeff_t = torch.where(ok[:, None, None, None], o["fe_coeff"], d["coeff_init"])
            esv_t = (torch.where(ok[:, None, None], o["fe_esv_cnt"], d["esv_cnt"]),
                     torch.where(ok[:, None, None, None], o["fe_esv_alpha"], d["esv_alpha"]),
                     torch.where(ok[:, None, None], o["fe_esv_active"], d["esv_active"]))
            self._fe_keep = (n_int_t, coeff_t, esv_t)   # referenced by the captured graph
            mark("front_end")
        a = capi.NbReplanArgs()
        a.B, a.space, a.n_hull_slots = B, DEV, self.N
        a.agent_id, a.n_int, a.coeff_init = d["agent_id"].data_ptr(), n_int_t.data_ptr(), coeff_t.data_ptr()
        # shared-window mode: the lines kernel reads hull (group[b], j, i) directly (own slot / unknown empty)
        a.hull_ptr, a.hull_xy, a.hull_cnt = None, o["hull_xy_g"].data_ptr(), o["hull_cnt_g"].data_ptr()
        a.hull_nvert = G * self.N * NPOL * HS
        a.nih0, a.nih0_group, a.hull_known = o["nih0_g"].data_ptr(), d["group"].data_ptr(), d["known"].data_ptr()
        a.esv_cnt, a.esv_alpha, a.esv_active = (t.data_ptr() for t in esv_t)
        a.bp_cnt, a.bp_xy = d["bp_cnt"].data_ptr(), d["bp_xy"].data_ptr()
        a.coeff_out, a.obj, a.status, a.iters = (o[k].data_ptr() for k in ("coeff_out", "obj", "status", "iters"))
        a.lines, a.line_ok = None, None
        chk(L.nb_replan_batch(h, C.byref(a), st), "nb_replan_batch")
        mark("lines_qp")
        # safetyCheckAfterReplan: geometric check against the late trajectories, then the entangle re-check
        if par_streams:
            main.wait_stream(sC)
            main.wait_stream(sB)
        chk(L.nb_postcheck_hulls_batch(h, B, DEV, p(n_int_t), p(o["coeff_out"]), p(d["group"]), p(o["hull_xy_g_l"]),
                                       p(o["hull_cnt_g_l"]), p(d["late"]), p(o["collide"]), st), "nb_postcheck_hulls_batch")
        if self.par.enable_entangle_check:
            esC = capi.NbEntState()
            esC.cnt, esC.alpha, esC.beta, esC.bend, esC.active = (p(o["esC_" + k]) for k in ("cnt", "alpha", "beta", "bend", "active"))
            samp_ptr, shared = (p(o["samp_g_l"]), 1) if G == 1 else (p(o["samp_b"]), 0)
            chk(L.nb_entangle_check_batch(h, B, DEV, p(d["agent_id"]), p(d["known"]), p(d["bp_cnt"]), p(d["bp_xy"]), esC,
                                          p(n_int_t), p(o["coeff_out"]), samp_ptr, shared, p(o["entangled"]), st),
                "nb_entangle_check_batch")
        if self.front_end:   # no front-end solution: replanFull returned before the back end (neptune.cpp:1473-1478)
            o["collide"].add_(1 - o["fe_solved"])
        mark("postcheck")
        # tail of replanFull (neptune.cpp:1685-1699): pwp_out = composePieceWisePol(time_now, dc, pwp_prev, pwp_now);
        # an agent whose replan was rejected keeps publishing its previous trajectory
        chk(L.nb_commit_compose_batch(h, B, DEV, p(n_int_t), p(o["coeff_out"]), p(d["t_start"]), p(d["t_now"]),
                                      p(d["recs"]), p(d["agent_id"]), None, p(o["status"]), p(o["entangled"]),
                                      p(o["collide"]), p(o["new_recs"]), p(o["new_pieces"]), st),
            "nb_commit_compose_batch")
        mark("commit")
        if self.profile:
            torch.cuda.synchronize()
            self.stage_ms = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
        return o

    def step_from_host(self, host_in: dict, host_out: dict, exchange: bool = True):
        """End-to-end cycle: pinned host inputs -> device, all kernels, results back to the host."""
        h2d = self.upload(host_in)
        self.step(exchange)
        d2h = self.download(host_out)
        self.torch.cuda.current_stream().synchronize()
        return h2d, d2h

    def check_errors(self):
        st = C.c_void_p(self.torch.cuda.current_stream().cuda_stream)
        capi._check(self._lib.nb_check_async_errors(self.solver.handle, st), "nb_check_async_errors")

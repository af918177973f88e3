"""One replan cycle for the agents of one rank: Python face of the library's device-resident cycle (``nb_cycle_*`` in
``include/neptune_b200.h``, ``csrc/nb_cycle.cu``) -- the hot part of ``Neptune::replanFull`` (reference
``neptune/src/neptune.cpp:1430-1448`` hulls / samples + PredictAlphasBetas, ``:1450-1510`` front end, ``:1512-1529``
back end, ``:1641-1647`` post-check, ``:1685-1699`` compose) and the message exchange around it
(``neptune_ros.cpp:379-480``).

The launch sequence, its streams, the CUDA graphs, the record ring and the peer-to-peer exchange all live in the
C++ library; this module only packs host inputs into the pinned buffer the library describes (``nb_cycle_layout``),
and bootstraps the exchange over ``torch.distributed`` (an all-gather of CUDA IPC handles, once).  PyTorch is used for
the stream the cycle is enqueued on and for that bootstrap -- no torch kernel runs inside a cycle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .batch import NPOL
from .params import Params

REC = capi.NB_REC_DOUBLES
HS = capi.NB_HULL_STRIDE
_P = C.c_void_p


def shard_agents(n_agents: int, world: int, rank: int) -> np.ndarray:
    """Block partition of agent indices (0-based) over ranks; the last ranks may own one agent less."""
    base, extra = divmod(n_agents, world)
    start = rank * base + min(rank, extra)
    return np.arange(start, start + base + (1 if rank < extra else 0))


def ring_phases(k: int) -> tuple[int, int, int]:
    """(new, late, known) slot of the three-slot record ring in cycle k: the records committed in cycle k go to slot
    k % 3; cycle k post-checks against those of cycle k - 1 and plans against those of cycle k - 2 (csrc/nb_cycle.cu)."""
    return k % 3, (k + 2) % 3, (k + 1) % 3


def gather_records(local, world: int, group=None, sizes=None):
    """All-gather of fixed-stride committed-trajectory records: [B_r][REC] per rank -> [sum B_r][REC] in rank order, on
    CUDA tensors (NCCL) or CPU tensors (gloo).  The collective form of the exchange: the device-resident cycle does the
    same by peer-to-peer stores from its commit kernel; this function remains for hosts that seed or inspect records
    across ranks and for the CPU tests of the sharding logic.  sizes: per-rank record counts when the shards are uneven."""
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


class NbCycleDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("agent_id", _P), ("front_end", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("planned", _P), ("bbox", C.c_double), ("delta", C.c_double)]


_IN_FIELDS = ("n_int", "coeff_init", "t_start", "t_now", "t_group", "group", "known", "late", "esv_cnt", "esv_alpha", "esv_active",
              "es_cnt", "es_alpha", "es_beta", "es_bend", "es_active", "prev_pos", "prev_pos_agent", "cur",
              "fe_init", "fe_goal", "fe_coeffs_z", "fe_comb")
_OUT_FIELDS = ("coeff_out", "obj", "status", "iters", "entangled", "collide", "n_pieces", "fe_status", "fe_solved", "fe_n_int",
               "fe_stats")


class NbCycleLayout(C.Structure):
    _fields_ = [(k, C.c_int64) for k in _IN_FIELDS] + [("in_bytes", C.c_int64)] + \
               [(k, C.c_int64) for k in _OUT_FIELDS] + [("out_bytes", C.c_int64)]


STAGES = ("late_hulls", "hulls_samples", "predict", "front_end", "lines_qp", "postcheck", "commit", "exchange_wait")


class PinnedBuffer:
    """Page-locked host memory from the library with named numpy views into it."""

    def __init__(self, nbytes: int, views: dict):
        L = capi.lib()
        L.nb_pinned_alloc.restype = _P
        L.nb_pinned_alloc.argtypes = [C.c_int64]
        L.nb_pinned_free.argtypes = [_P]
        self.ptr = L.nb_pinned_alloc(nbytes)
        if not self.ptr:
            raise capi.NbError("nb_pinned_alloc failed")
        self.nbytes = nbytes
        raw = (C.c_uint8 * nbytes).from_address(self.ptr)
        self.raw = np.frombuffer(raw, dtype=np.uint8)
        self.v = {}
        for k, (off, shape, dt) in views.items():
            if off >= 0:
                n = int(np.prod(shape)) * np.dtype(dt).itemsize
                self.v[k] = self.raw[off:off + n].view(dt).reshape(shape)

    def __getitem__(self, k):
        return self.v[k]

    def __contains__(self, k):
        return k in self.v

    def close(self):
        if self.ptr:
            self.v, self.raw = {}, None
            capi.lib().nb_pinned_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ReplanCycle:
    """Device-resident replan cycle of one rank (library object ``nb_cycle``)."""

    DELTA_T_STEPS = 2   # t_start - time_now in units of dc (deltaT_ of neptune.cpp:1236-1262 for the synthetic world)
    OUT_KEYS = ("coeff_out", "obj", "status", "iters", "entangled", "collide", "n_pieces")

    def __init__(self, par: Params, agents: np.ndarray, device, static=None, world: int = 1, group=None,
                 front_end: bool = False, rank: int = 0, planned=None, static_per_agent=None):
        import torch
        self.torch = torch
        self.front_end, self.par, self.world, self.rank, self.group = front_end, par, world, rank, group
        self.agents = np.asarray(agents)
        self.dev = torch.device(device)
        self.B, self.N = len(self.agents), par.num_of_agents
        self.solver = capi.Solver(par, device=self.dev.index or 0)
        if par.num_of_static_obst:
            st_ptr, st_xy, strep = static[:3]
            self.solver.set_static(st_ptr, st_xy, strep)
            if front_end:
                self.solver.set_static_longest(static[3])
            if static_per_agent is not None:      # (strep_all [N][M][2][2], longest_all [N][M][2] or None)
                self.solver.set_static_rep_per_agent(*static_per_agent)
        if front_end:
            self.solver.search_configure()
        self._lib = L = capi.lib()
        self._sig()
        self._ids = np.ascontiguousarray(self.agents + 1, np.int32)
        d = NbCycleDesc()
        d.B, d.agent_id, d.front_end, d.rank, d.world = self.B, self._ids.ctypes.data_as(_P), int(front_end), rank, world
        self._planned = None if planned is None else np.ascontiguousarray(planned, np.uint8)
        if self._planned is None and world > 1:
            self._planned = np.ones(self.N, np.uint8)     # a world whose agents are all planned by some rank
        d.planned = None if self._planned is None else self._planned.ctypes.data_as(_P)
        d.bbox = 2.0 * par.drone_radius                     # neptune_ros.cpp:447-449
        d.delta = self.delta = 2.0 * par.drone_radius       # bbox / 2 + drone_radius (neptune.cpp:340)
        self._h = _P()
        capi._check(L.nb_cycle_create(self.solver.handle, C.byref(d), C.byref(self._h)), "nb_cycle_create")
        lay = NbCycleLayout()
        capi._check(L.nb_cycle_layout_get(self._h, C.byref(lay)), "nb_cycle_layout_get")
        self.lay = lay
        B, N, cap, NA = self.B, self.N, par.ent_cap, par.NA
        f64, i32, u8 = np.float64, np.int32, np.uint8
        ns2 = par.a_star_samp_x ** 2
        shapes = dict(n_int=((B,), i32), coeff_init=((B, 3, NPOL, 4), f64), t_start=((B,), f64), t_now=((B,), f64),
                      t_group=((B,), f64), group=((B,), i32), known=((B, N), u8), late=((B, N), u8),
                      esv_cnt=((B, NPOL + 1, 2), i32), esv_alpha=((B, NPOL + 1, cap, 2), i32), esv_active=((B, NPOL + 1, NA), i32),
                      es_cnt=((B, 2), i32), es_alpha=((B, cap, 2), i32), es_beta=((B, cap), f64), es_bend=((B, cap), i32),
                      es_active=((B, NA), i32), prev_pos=((B, N + 1, 2), f64), prev_pos_agent=((B, N, 2), f64), cur=((B, 2), f64),
                      fe_init=((B, 6), f64), fe_goal=((B, 2), f64), fe_coeffs_z=((B, NPOL, 4), f64), fe_comb=((B, ns2), u8))
        oshapes = dict(coeff_out=((B, 3, NPOL, 4), f64), obj=((B,), f64), status=((B,), i32), iters=((B, 2), i32),
                       entangled=((B,), i32), collide=((B,), i32), n_pieces=((B,), i32), fe_status=((B,), i32),
                       fe_solved=((B,), i32), fe_n_int=((B,), i32), fe_stats=((B, 4), i32))
        self._in_views = {k: (int(getattr(lay, k)), shapes[k][0], shapes[k][1]) for k in _IN_FIELDS}
        self._out_views = {k: (int(getattr(lay, k)), oshapes[k][0], oshapes[k][1]) for k in _OUT_FIELDS}
        self.in_bytes, self.out_bytes = int(lay.in_bytes), int(lay.out_bytes)
        self.stream = torch.cuda.Stream(device=self.dev) if self.dev.type == "cuda" else None
        self.stage_ms = {}
        self.captured = False

    def _sig(self):
        L = self._lib
        L.nb_cycle_create.argtypes = [_P, C.POINTER(NbCycleDesc), C.POINTER(_P)]
        L.nb_cycle_destroy.argtypes = [_P]
        L.nb_cycle_layout_get.argtypes = [_P, C.POINTER(NbCycleLayout)]
        L.nb_cycle_seed_records.argtypes = [_P, _P, _P, _P]
        L.nb_cycle_upload_from.argtypes = [_P, _P, C.c_int32, _P]
        L.nb_cycle_download_to.argtypes = [_P, _P, _P]
        L.nb_cycle_step.argtypes = [_P, _P]
        L.nb_cycle_align.argtypes = [_P, _P]
        L.nb_cycle_last_upload_bytes.argtypes = [_P]
        L.nb_cycle_last_upload_bytes.restype = C.c_longlong
        L.nb_cycle_capture.argtypes = [_P, _P]
        L.nb_cycle_step_profiled.argtypes = [_P, _P, _P]
        L.nb_cycle_launches_per_step.argtypes = [_P]
        L.nb_cycle_launches_per_step.restype = C.c_longlong
        L.nb_cycle_index.argtypes = [_P]
        L.nb_cycle_index.restype = C.c_longlong
        L.nb_cycle_fetch.argtypes = [_P, C.c_char_p, _P, C.c_int64, _P]
        L.nb_cycle_ipc_handle.argtypes = [_P, _P]
        L.nb_cycle_open_peers.argtypes = [_P, _P]

    def _st(self):
        return _P(self.stream.cuda_stream)

    # ------------------------------------------------------------------ host <-> device
    def host_inputs(self, scene, fe: dict | None = None, late=None) -> PinnedBuffer:
        """Pinned host copy of everything a cycle needs from the planner core / front end besides the records, laid out
        like the packed device buffer.  late [B][N]: trajectories that arrive during the optimisation (default: every
        known one -- in a lock-step world each agent's newest plan reaches the others while they optimise)."""
        b = scene.batch
        uniq, inv = np.unique(np.asarray(scene.t_start, np.float64), return_inverse=True)
        tg = np.zeros(self.B)
        tg[:len(uniq)] = uniq
        src = dict(n_int=b.n_int, coeff_init=b.coeff_init, t_start=scene.t_start, known=scene.known,
                   late=scene.known if late is None else late,
                   esv_cnt=b.esv_cnt, esv_alpha=b.esv_alpha, esv_active=b.esv_active,
                   es_cnt=scene.es0_cnt, es_alpha=scene.es0_alpha, es_beta=scene.es0_beta, es_bend=scene.es0_bend,
                   es_active=scene.es0_active, prev_pos=scene.prev_pos, prev_pos_agent=scene.prev_pos_agent,
                   cur=np.ascontiguousarray(scene.state_A[:, 0, :2]), t_group=tg, group=inv.astype(np.int32),
                   t_now=np.asarray(scene.t_start, np.float64) - self.DELTA_T_STEPS * self.par.dc)
        if self.front_end:   # neptune_b200.scenes.search_host_inputs(scene, seed)
            src.update(fe_init=fe["init"], fe_goal=fe["goal"], fe_coeffs_z=fe["coeffs_z"], fe_comb=fe["comb"])
        buf = PinnedBuffer(self.in_bytes, self._in_views)
        for k, view in buf.v.items():
            view[...] = np.asarray(src[k]).astype(view.dtype).reshape(view.shape)
        buf.G = len(uniq)
        return buf

    def host_outputs(self) -> PinnedBuffer:
        return PinnedBuffer(self.out_bytes, self._out_views)

    def records_of(self, scene, seq: int = 0) -> np.ndarray:
        """Committed-trajectory records of every agent of the scene, DynTraj header included."""
        return capi.make_records(scene.committed, self.par, scene.batch.bp_cnt, scene.batch.bp_xy, seq=seq)

    def seed_records(self, known_recs: np.ndarray, late_recs: np.ndarray | None = None):
        """Records the agents know when the next cycle starts and those that arrive during it (default: the same)."""
        kr = np.ascontiguousarray(known_recs, np.float64)
        lr = kr if late_recs is None else np.ascontiguousarray(late_recs, np.float64)
        assert kr.shape == lr.shape == (self.N, REC)
        capi._check(self._lib.nb_cycle_seed_records(self._h, kr.ctypes.data_as(_P), lr.ctypes.data_as(_P), self._st()), "nb_cycle_seed_records")
        self.stream.synchronize()      # the numpy arrays may go away

    def upload(self, host: PinnedBuffer) -> int:
        """One H2D copy of the packed per-cycle inputs."""
        capi._check(self._lib.nb_cycle_upload_from(self._h, _P(host.ptr), int(host.G), self._st()), "nb_cycle_upload_from")
        return int(self._lib.nb_cycle_last_upload_bytes(self._h))

    def download(self, host_out: PinnedBuffer) -> int:
        """One D2H copy of the packed results."""
        capi._check(self._lib.nb_cycle_download_to(self._h, _P(host_out.ptr), self._st()), "nb_cycle_download_to")
        return self.out_bytes

    # ------------------------------------------------------------------ the cycle
    def capture(self):
        """Run one cycle, then capture the launch sequence (for the current window grouping) in CUDA graphs."""
        capi._check(self._lib.nb_cycle_capture(self._h, self._st()), "nb_cycle_capture")
        self.captured = True
        self.launches_per_cycle = int(self._lib.nb_cycle_launches_per_step(self._h))

    def align(self):
        """Rank barrier on the device, enqueued on the cycle's stream (measurement helper, nb_cycle_align)."""
        capi._check(self._lib.nb_cycle_align(self._h, self._st()), "nb_cycle_align")

    def step(self):
        """One cycle on the cycle's stream (graph replay when captured), exchange included."""
        capi._check(self._lib.nb_cycle_step(self._h, self._st()), "nb_cycle_step")
        self.launches_per_cycle = int(self._lib.nb_cycle_launches_per_step(self._h))

    def step_profiled(self) -> dict:
        """One cycle on a single stream with an event after every stage: {stage: ms}."""
        ms = (C.c_double * 8)()
        capi._check(self._lib.nb_cycle_step_profiled(self._h, self._st(), ms), "nb_cycle_step_profiled")
        self.stage_ms = {k: float(ms[i]) for i, k in enumerate(STAGES)}
        return self.stage_ms

    def step_from_host(self, host_in: PinnedBuffer, host_out: PinnedBuffer):
        """End-to-end cycle: pinned host inputs -> device, all kernels, results back to the host."""
        h2d = self.upload(host_in)
        self.step()
        d2h = self.download(host_out)
        self.stream.synchronize()
        return h2d, d2h

    @property
    def index(self) -> int:
        return int(self._lib.nb_cycle_index(self._h))

    def fetch(self, name: str, shape, dtype) -> np.ndarray:
        """A device array of the cycle on the host (tests, diagnostics); see nb_cycle_fetch for the names."""
        out = np.zeros(shape, dtype)
        capi._check(self._lib.nb_cycle_fetch(self._h, name.encode(), out.ctypes.data_as(_P), out.nbytes, self._st()), "nb_cycle_fetch")
        return out

    def records(self, which: str = "new") -> np.ndarray:
        """Ring slot of the last completed cycle: 'new' (committed in it), 'late', 'known'."""
        return self.fetch("ring_" + which, (self.N, REC), np.float64)

    def check_errors(self):
        capi._check(self._lib.nb_check_async_errors(self.solver.handle, self._st()), "nb_check_async_errors")

    # ------------------------------------------------------------------ exchange bootstrap
    def connect(self):
        """world > 1: all-gather the CUDA IPC handle of every rank's record ring (torch.distributed, once) and open the
        peers' rings; from then on the commit kernel of every cycle stores this rank's records into all rings."""
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist
        mine = (C.c_uint8 * 64)()
        capi._check(self._lib.nb_cycle_ipc_handle(self._h, mine), "nb_cycle_ipc_handle")
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.dev)
        allh = torch.empty(self.world * 64, dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, t, group=self.group)
        hb = allh.cpu().numpy().tobytes()
        capi._check(self._lib.nb_cycle_open_peers(self._h, hb), "nb_cycle_open_peers")
        dist.barrier(group=self.group)
        self.captured = False

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nb_cycle_destroy(self._h)
            self._h = None
        self.solver.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Host-side SoA containers for a batch of front-end searches.

One ``SearchBatch`` holds, for B agents, what ``Neptune::replanFull`` hands to its
``KinodynamicSearch`` per replan (reference ``neptune/src/neptune.cpp:1406-1453``): the start state A
(``setUp``), the goal, the initial z polynomial (``setInitZCoeffs``), the inflated hulls and the sampled
positions of the other agents' committed trajectories, ``entangle_state_A`` and the bend points; plus the
two inputs that replace non-deterministic state of the reference -- the order of the jerk samples
(``all_combinations_`` is shuffled with a wall-clock seed, ``kinodynamic_search.cpp:321-322``) and the
number of open-list pops that stands in for ``max_runtime_`` (``:1646``).  Arrays are plain numpy so the
same bytes go to the CUDA library (C-ABI ``nb_search_batch``) and to the CPU oracle.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .batch import NPOL
from .params import Params

HULL_STRIDE = 24


@dataclasses.dataclass
class SearchBatch:
    par: Params
    agent_id: np.ndarray     # [B] int32, 1-based
    init: np.ndarray         # [B][6] f64: px py vx vy ax ay (initial_, kinodynamic_search.cpp:205-207)
    goal: np.ndarray         # [B][2] f64 goal2d_
    coeffs_z: np.ndarray     # [B][8][4] f64 getInitialZPwp
    group: np.ndarray        # [B] int32 window group (agents with equal t_start share hulls and samples)
    hull_xy: np.ndarray      # [G][N][8][24][2] f64 inflated hulls per window, CCW
    hull_cnt: np.ndarray     # [G][N][8] int32 vertex counts
    samp: np.ndarray         # [G][N][num_pol][S+1][2] f64 SampledPointsForAll
    known: np.ndarray        # [B][N] uint8 (0 for the agent itself)
    es_cnt: np.ndarray       # [B][2] int32 entangle_state_A
    es_alpha: np.ndarray     # [B][ent_cap][2] int32
    es_beta: np.ndarray      # [B][ent_cap] f64
    es_bend: np.ndarray      # [B][ent_cap] int32
    es_active: np.ndarray    # [B][N+M] int32
    bp_cnt: np.ndarray       # [N] int32
    bp_xy: np.ndarray        # [N][bp_max][2] f64
    comb: np.ndarray         # [ns*ns] or [B][ns*ns] uint8: order of the jerk samples (value = jx*ns+jy)
    st_ptr: np.ndarray       # [M+1] int64 inflated static obstacles (setStaticObstVert)
    st_xy: np.ndarray        # [nsv][2]
    strep: np.ndarray        # [M][2][2] staticObsRep
    st_longest: np.ndarray   # [M][2] staticObsLongestDist

    @property
    def B(self) -> int:
        return int(self.agent_id.shape[0])

    @property
    def G(self) -> int:
        return int(self.hull_cnt.shape[0])

    def validate(self) -> None:
        p, B, N, G = self.par, self.B, self.par.num_of_agents, self.G
        ns2 = p.a_star_samp_x ** 2
        assert self.init.shape == (B, 6) and self.goal.shape == (B, 2) and self.coeffs_z.shape == (B, NPOL, 4)
        assert self.group.shape == (B,) and self.group.dtype == np.int32 and int(self.group.max()) < G
        assert self.hull_xy.shape == (G, N, NPOL, HULL_STRIDE, 2) and self.hull_cnt.shape == (G, N, NPOL)
        assert self.samp.shape == (G, N, p.num_pol, p.num_sample_per_interval + 1, 2)
        assert self.known.shape == (B, N) and self.known.dtype == np.uint8
        assert self.es_cnt.shape == (B, 2) and self.es_alpha.shape == (B, p.ent_cap, 2)
        assert self.es_beta.shape == (B, p.ent_cap) and self.es_bend.shape == (B, p.ent_cap)
        assert self.es_active.shape == (B, p.NA)
        assert self.comb.dtype == np.uint8 and self.comb.shape in ((ns2,), (B, ns2))
        assert sorted(np.atleast_2d(self.comb)[0].tolist()) == list(range(ns2))
        assert self.st_longest.shape == (p.num_of_static_obst, 2)
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            if isinstance(v, np.ndarray):
                assert v.flags["C_CONTIGUOUS"], f.name


@dataclasses.dataclass
class SearchResult:
    status: np.ndarray      # [B] int32: 0 runtime reached, 1 goal reached, 2 open list empty (:1637-1639)
    solved: np.ndarray      # [B] int32: return value of run()
    n_int: np.ndarray       # [B] int32 pieces of pwp_out_ (0 when not solved)
    coeff: np.ndarray       # [B][3][8][4] pwp_out_ (getPwpOut_0tstart)
    esv_cnt: np.ndarray     # [B][9][2] entStateVec (getEntStateVector)
    esv_alpha: np.ndarray   # [B][9][ent_cap][2]
    esv_beta: np.ndarray    # [B][9][ent_cap]
    esv_bend: np.ndarray    # [B][9][ent_cap]
    esv_active: np.ndarray  # [B][9][N+M]
    stats: np.ndarray       # [B][4] int32: nodes used, pops, index of the best node, goal_occupied
    cost: np.ndarray        # [B] g of the best node

    @staticmethod
    def empty(sb: SearchBatch) -> "SearchResult":
        B, p = sb.B, sb.par
        return SearchResult(
            status=np.full(B, -1, np.int32), solved=np.zeros(B, np.int32), n_int=np.zeros(B, np.int32),
            coeff=np.zeros((B, 3, NPOL, 4)), esv_cnt=np.zeros((B, NPOL + 1, 2), np.int32),
            esv_alpha=np.zeros((B, NPOL + 1, p.ent_cap, 2), np.int32), esv_beta=np.zeros((B, NPOL + 1, p.ent_cap)),
            esv_bend=np.zeros((B, NPOL + 1, p.ent_cap), np.int32), esv_active=np.zeros((B, NPOL + 1, p.NA), np.int32),
            stats=np.zeros((B, 4), np.int32), cost=np.zeros(B))


def static_longest_dist(static_raw, strep) -> np.ndarray:
    """``staticObsLongestDist_`` (reference ``neptune_ros.cpp:986-1002``): for each representative point of a
    static obstacle, the largest distance to a vertex of that obstacle."""
    out = np.zeros((len(static_raw), 2))
    for m, poly in enumerate(static_raw):
        for c in range(2):
            best = 0.0
            for v in poly:
                d = float(np.linalg.norm(v - strep[m, c]))
                if d > best:
                    best = d
            out[m, c] = best
    return out


def jerk_order(par: Params, seed: int, B: int | None = None) -> np.ndarray:
    """A seeded stand-in for the reference's wall-clock shuffle of ``all_combinations_``."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ns2 = par.a_star_samp_x ** 2
    if B is None:
        return rng.permutation(ns2).astype(np.uint8)
    return np.stack([rng.permutation(ns2) for _ in range(B)]).astype(np.uint8)

"""Host-side SoA containers for a batch of replans.

One ``ReplanBatch`` holds, for B agents, exactly what the reference hands to
``PolySolverGurobi`` per replan (``neptune/src/neptune.cpp:1514-1517``):
``pwp_init`` (setInitTrajectory), the inflated hulls of the other agents per interval (setHulls),
column 0 of the un-inflated hulls (setHullsNoInflation; the only column the solver reads,
``solver_gurobi_poly.cpp:722-723, :734``), the per-interval entanglement states and the bend points
(setEntStateVector), plus the static-obstacle hulls (setStaticObstVert).  Arrays are plain numpy so
the same bytes go to the CUDA library (through the C-ABI) and to the CPU oracle.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .params import Params

NPOL = 8  # storage stride for intervals (num_pol <= 8 in every shipped YAML)


@dataclasses.dataclass
class ReplanBatch:
    par: Params
    agent_id: np.ndarray      # [B] int32, 1-based id of the planning agent (par_.id)
    n_int: np.ndarray         # [B] int32, intervals of pwp_init (1..num_pol)
    coeff_init: np.ndarray    # [B][3][8][4] f64, [a b c d] per axis/interval, t in seconds
    n_hull_slots: int         # hull slots per agent (<= N); empty slots have zero vertices
    hull_ptr: np.ndarray      # [B*slots*8+1] int64 vertex offsets (slot-major, then interval)
    hull_xy: np.ndarray       # [nvert][2] f64 packed CCW polygons
    nih0: np.ndarray          # [B][N][8][2] f64, NaN where agent unknown / self
    st_ptr: np.ndarray        # [M+1] int64
    st_xy: np.ndarray         # [nsv][2] f64 inflated static hulls (Neptune::setStaticObst)
    esv_cnt: np.ndarray       # [B][9][2] int32 (n_alpha, n_bend) of entStateVec[i]
    esv_alpha: np.ndarray     # [B][9][cap][2] int32
    esv_active: np.ndarray    # [B][9][N+M] int32
    bp_cnt: np.ndarray        # [N] int32 (shared) bend points per agent, base included
    bp_xy: np.ndarray         # [N][bp_max][2] f64

    @property
    def B(self) -> int:
        return int(self.agent_id.shape[0])

    @property
    def line_slots(self) -> int:
        return self.par.line_slots(self.n_hull_slots)

    def validate(self) -> None:
        p, B, N = self.par, self.B, self.par.num_of_agents
        assert self.agent_id.dtype == np.int32 and self.n_int.dtype == np.int32
        assert self.coeff_init.shape == (B, 3, NPOL, 4) and self.coeff_init.dtype == np.float64
        assert self.hull_ptr.shape == (B * self.n_hull_slots * NPOL + 1,) and self.hull_ptr.dtype == np.int64
        assert self.hull_xy.dtype == np.float64 and self.hull_xy.ndim == 2
        assert self.nih0.shape == (B, N, NPOL, 2)
        assert self.st_ptr.shape == (p.num_of_static_obst + 1,) and self.st_ptr.dtype == np.int64
        assert self.esv_cnt.shape == (B, NPOL + 1, 2) and self.esv_cnt.dtype == np.int32
        assert self.esv_alpha.shape == (B, NPOL + 1, p.ent_cap, 2) and self.esv_alpha.dtype == np.int32
        assert self.esv_active.shape == (B, NPOL + 1, p.NA) and self.esv_active.dtype == np.int32
        assert self.bp_cnt.shape == (N,) and self.bp_cnt.dtype == np.int32
        assert self.bp_xy.shape == (N, p.bp_max, 2)
        for a in (self.agent_id, self.n_int, self.coeff_init, self.hull_ptr, self.hull_xy, self.nih0,
                  self.st_ptr, self.st_xy, self.esv_cnt, self.esv_alpha, self.esv_active, self.bp_cnt, self.bp_xy):
            assert a.flags["C_CONTIGUOUS"]

    def input_bytes(self) -> int:
        """Bytes of the per-replan inputs as laid out for the device (h2d bytes per step)."""
        arrs = (self.agent_id, self.n_int, self.coeff_init, self.hull_ptr, self.hull_xy, self.nih0,
                self.esv_cnt, self.esv_alpha, self.esv_active)
        return int(sum(a.nbytes for a in arrs))

    def algorithmic_bytes(self) -> int:
        """Compulsory bytes at the PolySolverGurobi boundary (SURVEY.md section 8d formula):
        pwp_init + real hull vertices + nih0 column + ent-state entries actually present +
        bend points + static hulls once per batch + outputs."""
        p = self.par
        n = self.n_int.astype(np.int64)
        b = int((8 * (12 * n + n + 1)).sum())
        b += 16 * int(self.hull_xy.shape[0])
        known = ~np.isnan(self.nih0[..., 0])
        b += 16 * int(known.sum())
        b += 16 * int(self.st_xy.shape[0])
        L = self.esv_cnt[..., 0].astype(np.int64)
        for i in range(self.B):
            ni = int(n[i])
            b += int((4 * (2 * L[i, :ni + 1] + p.NA)).sum())
        b += 16 * int(self.bp_cnt.sum())
        b += int((8 * 12 * n + 8 + 4).sum())
        return b


@dataclasses.dataclass
class ReplanResult:
    coeff_out: np.ndarray   # [B][3][8][4]
    obj: np.ndarray         # [B]
    status: np.ndarray      # [B] int32: 0 direct, 1 fallback, 2 failed (pwp_out = pwp_init)
    iters: np.ndarray       # [B][2] int32 IPM iterations (direct, fallback)
    lines: np.ndarray | None = None     # [B][8][LS][3]
    line_ok: np.ndarray | None = None   # [B][8][LS] uint8: 0 not attempted, 1 solved, 2 unsolved

    @staticmethod
    def empty(batch: ReplanBatch, with_lines: bool = True) -> "ReplanResult":
        B, LS = batch.B, batch.line_slots
        return ReplanResult(
            coeff_out=np.zeros((B, 3, NPOL, 4)), obj=np.zeros(B), status=np.full(B, -1, np.int32),
            iters=np.zeros((B, 2), np.int32),
            lines=np.zeros((B, NPOL, LS, 3)) if with_lines else None,
            line_ok=np.zeros((B, NPOL, LS), np.uint8) if with_lines else None)

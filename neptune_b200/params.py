"""Planner parameters on the replan hot path.

Mirror of the subset of ``mt::parameters`` (reference ``neptune/include/mader_types.hpp:582-670``)
that the back end reads through ``PolySolverGurobi``'s constructor and setters
(``neptune/src/neptune.cpp:102-107``), with the values of the shipped YAML files
(``neptune/param/*.yaml``).  Names follow the reference.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np


@dataclasses.dataclass
class Params:
    num_pol: int = 8                 # neptune_multi_obstacle.yaml:74
    deg_pol: int = 3
    num_of_agents: int = 5
    num_of_static_obst: int = 0
    num_sample_per_interval: int = 3  # :69
    T_span: float = 0.5               # :70
    weight: float = 1000.0            # :76
    x_min: float = -15.0
    x_max: float = 15.0
    y_min: float = -15.0
    y_max: float = 15.0
    z_min: float = -0.2
    z_max: float = 2.5
    v_max: float = 2.0               # par_.v_max(0) is applied to every axis (neptune.cpp:104-105)
    a_max: float = 3.0
    j_max: float = 5.0
    drone_radius: float = 0.6
    tetherLength: float = 25.0
    dc: float = 0.05
    runtime_opt: float = 0.08
    use_linear_constraints: bool = True
    enable_entangle_check: bool = True
    # front end (KinodynamicSearch), neptune_multi_obstacle.yaml:9, :43-44, :127; setBias(1.1) neptune.cpp:97
    goal_radius: float = 0.5
    a_star_samp_x: int = 5
    a_star_fraction_voxel_size: float = 0.2
    a_star_bias: float = 1.1
    use_not_reaching_soln: bool = True
    search_max_nodes: int = 4096      # node_num_max_ (the reference: 15 * area / voxel^2, kinodynamic_search.cpp:370)
    search_max_expansions: int = 600  # open-list pops allowed: stands in for the wall-clock max_runtime_ (:1646)
    search_ecap: int = 24             # entries an alphas list of a search node can hold (semantic bound N+M)
    # storage capacities of this implementation (the reference uses std::vector)
    ent_cap: int = 48                # entries an alphas list can hold (semantic bound stays 3(N+M))
    bp_max: int = 8                  # bend points per agent list, base included
    ent_slots: int = 16              # LP slots per interval reserved for non-entangling constraints
    ipm_max_iter: int = 30        # converged solves need <= 15 iterations on every scene; an infeasible first solve that does not blow up stalls the whole batch until the cap
    ipm_tol: float = 1e-9
    pb: np.ndarray | None = None     # [N][2] base positions (par_.pb)

    @property
    def NA(self) -> int:
        return self.num_of_agents + self.num_of_static_obst

    def line_slots(self, n_hull_slots: int) -> int:
        """LP slots per interval: other agents, bases, static obstacles, non-entangling."""
        return n_hull_slots + self.num_of_agents + self.num_of_static_obst + self.ent_slots


def circle_bases(n: int) -> np.ndarray:
    """``circle_init_bases`` formula, reference ``neptune/src/neptune_ros.cpp:138-150``
    (``cosf``/``sinf`` are evaluated in single precision there)."""
    pb = np.zeros((n, 2))
    one_slice = 3.1415927 * 2 / n
    for i in range(n):
        th = np.float32(one_slice * i)
        th2 = np.float32(1.571 - one_slice * i)
        pb[i, 0] = -10.0 * float(np.cos(th, dtype=np.float32)) - 2.5 * float(np.cos(th2, dtype=np.float32))
        pb[i, 1] = -10.0 * float(np.sin(th, dtype=np.float32)) + 2.5 * float(np.sin(th2, dtype=np.float32))
    return pb


def grid_bases(side: int, pitch: float = 8.0) -> np.ndarray:
    """Builder-defined base grid for configs 4 and 5."""
    xs = (np.arange(side) - (side - 1) / 2.0) * pitch
    gx, gy = np.meshgrid(xs, xs, indexing="ij")
    return np.stack([gx.ravel(), gy.ravel()], axis=1)


# the nine fixed 0.5 m squares of neptune_multi_obstacle.yaml:104-124
_MO_X = [-0.25, -0.25, 0.25, 0.25, 6.25, 5.75, 5.75, 6.25, -6.25, -5.75, -5.75, -6.25, 6.25, 5.75, 5.75, 6.25,
         -6.25, -5.75, -5.75, -6.25, 5.75, 5.75, 6.25, 6.25, -5.75, -5.75, -6.25, -6.25, -0.25, 0.25, 0.25, -0.25,
         -0.25, 0.25, 0.25, -0.25]
_MO_Y = [-0.25, 0.25, 0.25, -0.25, 6.25, 6.25, 5.75, 5.75, -6.25, -6.25, -5.75, -5.75, -6.25, -6.25, -5.75, -5.75,
         6.25, 6.25, 5.75, 5.75, -0.25, 0.25, 0.25, -0.25, -0.25, 0.25, 0.25, -0.25, 5.75, 5.75, 6.25, 6.25,
         -5.75, -5.75, -6.25, -6.25]


def multi_obstacle_squares() -> list[np.ndarray]:
    return [np.array([_MO_X[4 * i:4 * i + 4], _MO_Y[4 * i:4 * i + 4]]).T.copy() for i in range(9)]


def config(name: str) -> Params:
    """The five BASELINE.json configs."""
    if name == "single":        # config 1: neptune_single_benchmark.yaml, derived (M=0, n=3)
        p = Params(num_of_agents=1, x_min=-11.0, x_max=15.0, y_min=-15.0, y_max=15.0, z_min=-1.0, z_max=2.5,
                   drone_radius=0.5, tetherLength=27.0, runtime_opt=0.08)
        p.pb = np.array([[-10.0, 0.5]])
    elif name == "mtlp5":       # config 2: neptune_mtlp_benchmark.yaml
        p = Params(num_of_agents=5, x_min=-12.0, x_max=12.0, y_min=-12.0, y_max=12.0, z_min=-0.2, z_max=5.1,
                   tetherLength=40.0, runtime_opt=0.05)
        p.pb = circle_bases(5)
    elif name == "obst8":       # config 3: neptune_multi_obstacle.yaml with N raised to 8
        p = Params(num_of_agents=8, num_of_static_obst=9, tetherLength=25.0, runtime_opt=0.08)
        p.pb = circle_bases(8)
    elif name == "grid64":      # config 4: builder-defined 8x8 grid
        p = Params(num_of_agents=64, tetherLength=25.0)
        p.pb = grid_bases(8)
        half = 3.5 * 8.0 + 12.0
        p.x_min = p.y_min = -half
        p.x_max = p.y_max = half
    elif name == "grid1024":    # config 5: builder-defined 32x32 grid, 200 static squares
        p = Params(num_of_agents=1024, num_of_static_obst=200, tetherLength=25.0, ent_cap=1232,  # >= N + M
                   search_ecap=96)   # crowded world: longer signature words per search node
        p.pb = grid_bases(32)
        half = 15.5 * 8.0 + 12.0
        p.x_min = p.y_min = -half
        p.x_max = p.y_max = half
    else:
        raise KeyError(name)
    return p


def long_length(p: Params) -> float:
    """``long_length_`` = workspace diagonal (solver_gurobi_poly.cpp:173)."""
    return math.sqrt((p.x_max - p.x_min) ** 2 + (p.y_max - p.y_min) ** 2)

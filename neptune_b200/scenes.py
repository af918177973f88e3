"""Synthetic random-goal scenes for the five BASELINE.json configs.

This is the WORKLOAD GENERATOR: host-side numpy that produces the inputs the reference's planner
core hands to its back end (``neptune/src/neptune.cpp:1426-1517``) -- committed trajectories of the
other agents, start state A, a front-end path ``pwp_init``, hull/sample arrays, entanglement-state
vectors.  It is neither the product's hot path nor the oracle.  The ``pwp_init`` it produces comes from a
greedy walk over the reference's 5x5 constant-jerk lattice (``kinodynamic_search.cpp:1045-1140`` primitives
and admissibility tests): a cheap stand-in used for the previous plans of the other agents and as the back
end's input when the search is not run.  The reference's search itself is the product's K0 kernel
(``nb_search_batch``); ``make_search_batch`` / ``search_host_inputs`` below build its inputs.

Entanglement states are filled through an ``ent_backend`` (the product's device kernels in bench.py,
the oracle in tests/), so this module never touches oracle/.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

from .batch import NPOL, ReplanBatch
from .minvo import solver_basis
from .params import Params, multi_obstacle_squares


# --------------------------------------------------------------------------- geometry helpers
def hull2d(pts: np.ndarray) -> np.ndarray:
    """Strict convex hull, CCW from the lexicographically smallest point (convention adopted for
    ``cu::convexHullOfPoints2d``, reference ``cgal_utils.cpp:157-174``)."""
    p = sorted(set((float(x), float(y)) for x, y in pts))
    if len(p) <= 2:
        return np.array(p).reshape(-1, 2)

    def cr(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lo = []
    for q in p:
        while len(lo) >= 2 and cr(lo[-2], lo[-1], q) <= 0:
            lo.pop()
        lo.append(q)
    up = []
    for q in reversed(p):
        while len(up) >= 2 and cr(up[-2], up[-1], q) <= 0:
            up.pop()
        up.append(q)
    return np.array(lo[:-1] + up[:-1])


def interval_ctrl_points(times, cx, cy, t0, t1, T, A01inv):
    """Un-inflated MINVO control points of every piece overlapping [t0,t1] (``neptune.cpp:379-448``)."""
    npc = len(times) - 1
    first = int(np.searchsorted(times, t0, side="left")) - 1
    last = int(np.searchsorted(times, t1, side="right")) - 1
    first = min(max(first, 0), npc - 1)
    last = min(max(last, 0), npc - 1)
    out = []
    for i in range(first, last + 1):
        if i != last or t1 > times[i + 1]:
            t = times[i + 1] - times[i]
        else:
            t = t1 - times[i]
        t = min(max(t, 0.0), T)
        sc = np.array([t ** 3, t ** 2, t, 1.0])
        P = np.stack([cx[i] * sc, cy[i] * sc])
        out.append((P @ A01inv).T)
    return np.concatenate(out, axis=0), (first, last)


def hulls_of_interval(times, cx, cy, t0, t1, T, delta, A01inv):
    cps, _ = interval_ctrl_points(times, cx, cy, t0, t1, T, A01inv)
    corners = np.array([[1, 1], [1, -1], [-1, -1], [-1, 1]], float) * delta
    infl = (cps[:, None, :] + corners[None]).reshape(-1, 2)
    return hull2d(infl), hull2d(cps)


def sample_points(times, cx, cy, t_start, t_end, num_pol, S):
    """``Neptune::SamplePointsOfIntervals`` (``neptune.cpp:500-566``) -> [num_pol][S+1][2]."""
    dT = (t_end - t_start) / num_pol
    npc = len(times) - 1
    out = np.zeros((num_pol, S + 1, 2))
    for i in range(num_pol):
        for j in range(S + 1):
            ts = t_start + dT * i + dT / S * j
            low = int(np.searchsorted(times, ts, side="right"))
            if low != len(times):
                ii = min(max(low - 1, 0), npc - 1)
                te = min(max(ts - times[ii], 0.0), dT)
            else:
                ii = low - 2
                te = times[low - 1] - times[low - 2]
            tv = np.array([te ** 3, te ** 2, te, 1.0])
            out[i, j] = [cx[ii] @ tv, cy[ii] @ tv]
    return out


def eval_pwp(times, c, t):
    """position, velocity, acceleration of one axis of a piecewise cubic at absolute time t (clamped)."""
    npc = len(times) - 1
    i = min(max(int(np.searchsorted(times, t, side="right")) - 1, 0), npc - 1)
    u = min(max(t - times[i], 0.0), times[i + 1] - times[i])
    a, b, cc, d = c[i]
    return a * u ** 3 + b * u ** 2 + cc * u + d, 3 * a * u ** 2 + 2 * b * u + cc, 6 * a * u + 2 * b


def initial_z_pwp(par: Params, p0, v0, a0, z_final, v_max_z=2.0, a_max_z=9.6):
    """``Neptune::getInitialZPwp`` (``neptune.cpp:1727-1810``): clamped uniform-B-spline z guess."""
    n, T = par.num_pol, par.T_span
    v0 = min(max(v0, -v_max_z), v_max_z)
    a0 = min(max(a0, -a_max_z), a_max_z)
    q = np.zeros(n + 3)
    v = np.zeros(n + 2)
    q[0] = p0
    q[1] = p0 + T * v0 / 3
    q[2] = (9 * q[1] - 2 * T * (-a0 * T + v0) - 3 * (q[1] - 2 * T * v0)) / 6
    q[n] = z_final
    inc = (z_final - q[2]) / (n - 2)
    for i in range(3, n):
        q[i] = q[i - 1] + inc
    for i in range(3, n + 1):
        v[i - 1] = (q[i] - q[i - 1]) / T
        if v[i - 1] > v_max_z:
            q[i] = q[i - 1] + v_max_z * T
            v[i - 1] = v_max_z
        elif v[i - 1] < -v_max_z:
            q[i] = q[i - 1] - v_max_z * T
            v[i - 1] = -v_max_z
    for i in range(2, n):
        ai = (v[i] - v[i - 1]) / T
        if ai > a_max_z:
            v[i] = v[i - 1] + a_max_z * T
        elif ai < -a_max_z:
            v[i] = v[i - 1] - a_max_z * T
        q[i + 1] = q[i] + v[i] * T
    q[n + 1] = q[n]
    q[n + 2] = q[n]
    M = np.array([[1, 4, 1, 0], [-3, 0, 3, 0], [3, -6, 3, 0], [-1, 3, -3, 1]], float) / 6.0
    Bm = np.diag([1.0, 1.0 / T, T ** -2, T ** -3]) @ M
    return np.stack([(Bm @ q[i:i + 4])[::-1] for i in range(n)])


# --------------------------------------------------------------------------- front-end stub
_JERKS = None


def lattice_path(par: Params, Ainv, V, p, v, a, goal, base, boxes, n_max, goal_radius=0.5):
    """Greedy walk over the 5x5 constant-jerk lattice with the reference's admissibility tests
    (``kinodynamic_search.cpp:1053-1160``).  boxes[i] = [K][4] AABBs to stay clear of in interval i."""
    global _JERKS
    if _JERKS is None:
        jv = np.linspace(-par.j_max, par.j_max, 5)
        _JERKS = np.array([[jx, jy] for jx in jv for jy in jv])
    T = par.T_span
    J = _JERKS
    cxs, cys = [], []
    p, v, a = np.array(p, float), np.array(v, float), np.array(a, float)
    for i in range(n_max):
        pe = p + v * T + a * T * T / 2 + J * T ** 3 / 6
        ve = v + a * T + J * T * T / 2
        ae = a + J * T
        ok = (np.abs(ae) <= par.a_max + 1e-12).all(axis=1)
        ok &= np.linalg.norm(np.concatenate([pe - p, ve - v, ae - a], axis=1), axis=1) >= 1e-5
        Px = np.stack([J[:, 0] / 6, np.full(25, a[0] / 2), np.full(25, v[0]), np.full(25, p[0])], axis=1)
        Py = np.stack([J[:, 1] / 6, np.full(25, a[1] / 2), np.full(25, v[1]), np.full(25, p[1])], axis=1)
        Qx, Qy = Px @ Ainv, Py @ Ainv
        ok &= ((Qx >= par.x_min) & (Qx <= par.x_max) & (Qy >= par.y_min) & (Qy <= par.y_max)).all(axis=1)
        ok &= (np.hypot(Qx - base[0], Qy - base[1]) <= par.tetherLength).all(axis=1)
        Vx, Vy = Px[:, :3] @ V, Py[:, :3] @ V
        ok &= ((np.abs(Vx) <= par.v_max) & (np.abs(Vy) <= par.v_max)).all(axis=1)
        for d in range(2):  # future velocity admissibility (:1141-1160)
            up = (ae[:, d] > 0) & (ve[:, d] + 0.5 * ae[:, d] ** 2 / par.j_max > par.v_max)
            dn = (ae[:, d] < 0) & (ve[:, d] - 0.5 * ae[:, d] ** 2 / par.j_max < -par.v_max)
            ok &= ~(up | dn)
        if boxes is not None and len(boxes[i]):
            bx = boxes[i]
            lo_x, hi_x, lo_y, hi_y = Qx.min(1), Qx.max(1), Qy.min(1), Qy.max(1)
            ov = ((lo_x[:, None] <= bx[None, :, 2] + 0.05) & (hi_x[:, None] >= bx[None, :, 0] - 0.05) &
                  (lo_y[:, None] <= bx[None, :, 3] + 0.05) & (hi_y[:, None] >= bx[None, :, 1] - 0.05)).any(axis=1)
            ok &= ~ov
        if not ok.any():
            break
        stop = pe + ve * 0.7 + 0.5 * ae * 0.35
        cost = np.linalg.norm(stop - goal, axis=1) + 0.02 * np.linalg.norm(J, axis=1)
        cost[~ok] = np.inf
        k = int(np.argmin(cost))
        cxs.append(Px[k])
        cys.append(Py[k])
        p, v, a = pe[k], ve[k], ae[k]
        if np.linalg.norm(p - goal) < goal_radius and np.linalg.norm(v) < 0.3:
            break
    if not cxs:  # no admissible primitive: hold the current motion with zero jerk
        cxs.append(np.array([0.0, a[0] / 2, v[0], p[0]]))
        cys.append(np.array([0.0, a[1] / 2, v[1], p[1]]))
    return np.array(cxs), np.array(cys)


# --------------------------------------------------------------------------- scene
@dataclasses.dataclass
class Scene:
    par: Params
    batch: ReplanBatch
    t_start: np.ndarray            # [B]
    committed: list                # per agent (times, cx, cy, cz)
    goals: np.ndarray              # [N][3]
    static_raw: list               # un-inflated static polygons
    strep: np.ndarray              # [M][2][2] staticObsRep (col0, col1)
    samp: np.ndarray               # [B][N][8][S+1][2]
    known: np.ndarray              # [B][N] uint8
    prev_pos: np.ndarray           # [B][N+1][2] previousCheckingPos_
    prev_pos_agent: np.ndarray     # [B][N][2]   previousCheckingPosAgent_
    state_A: np.ndarray            # [B][3][3] pos/vel/acc of the start state
    es0_cnt: np.ndarray            # [B][2]  entangle_state_ before PredictAlphasBetas
    es0_alpha: np.ndarray          # [B][cap][2]
    es0_beta: np.ndarray           # [B][cap]
    es0_bend: np.ndarray           # [B][cap]
    es0_active: np.ndarray         # [B][N+M]
    esA_cnt: np.ndarray | None = None    # entangle_state_A
    esA_alpha: np.ndarray | None = None
    esA_beta: np.ndarray | None = None
    esA_bend: np.ndarray | None = None
    esA_active: np.ndarray | None = None
    group: np.ndarray | None = None        # [B] window group (index into the distinct t_start values)
    hull_g_xy: np.ndarray | None = None    # [G][N][8][24][2] inflated hulls per window group (fixed stride)
    hull_g_cnt: np.ndarray | None = None   # [G][N][8]
    samp_g: np.ndarray | None = None       # [G][N][num_pol][S+1][2]


def slice_scene(sc: Scene, idx) -> Scene:
    """The same world seen by a subset of its planning agents (rows idx of every per-agent array): what one rank of a
    sharded run plans for.  The packed host hulls are dropped (the device builds hulls from the records)."""
    idx = np.asarray(idx)
    b = sc.batch
    nb = dataclasses.replace(
        b, agent_id=np.ascontiguousarray(b.agent_id[idx]), n_int=np.ascontiguousarray(b.n_int[idx]),
        coeff_init=np.ascontiguousarray(b.coeff_init[idx]), hull_ptr=np.zeros(len(idx) * b.n_hull_slots * NPOL + 1, np.int64),
        hull_xy=np.zeros((0, 2)), nih0=np.ascontiguousarray(b.nih0[idx]), esv_cnt=np.ascontiguousarray(b.esv_cnt[idx]),
        esv_alpha=np.ascontiguousarray(b.esv_alpha[idx]), esv_active=np.ascontiguousarray(b.esv_active[idx]))
    per_agent = ("t_start", "samp", "known", "prev_pos", "prev_pos_agent", "state_A", "es0_cnt", "es0_alpha", "es0_beta", "es0_bend",
                 "es0_active", "esA_cnt", "esA_alpha", "esA_beta", "esA_bend", "esA_active")
    kw = {k: (None if getattr(sc, k) is None else np.ascontiguousarray(np.asarray(getattr(sc, k))[idx])) for k in per_agent}
    return dataclasses.replace(sc, batch=nb, group=None, hull_g_xy=None, hull_g_cnt=None, samp_g=None, **kw)


def _static_obstacles(par: Params, rng, pb):
    M = par.num_of_static_obst
    if M == 0:
        return []
    if M == 9:
        return multi_obstacle_squares()
    pts = []  # random 0.5 m squares with the spacing rule of neptune_ros.cpp:229-243
    while len(pts) < M:
        c = np.array([rng.uniform(par.x_min + 1, par.x_max - 1), rng.uniform(par.y_min + 1, par.y_max - 1)])
        okc = all(not (d < 0.6 or (2 * par.drone_radius < d < 0.566 + 8 * par.drone_radius))
                  for d in (np.linalg.norm(c - q) for q in pts))
        if okc and np.min(np.linalg.norm(pb - c, axis=1)) >= 2.5:
            pts.append(c)
    sq = np.array([[0.25, 0.25], [0.25, -0.25], [-0.25, -0.25], [-0.25, 0.25]])
    return [c + sq for c in pts]


def _static_rep(poly, theta):
    """Two representative points of a static obstacle: where the line through its centre at angle
    theta leaves the polygon (simplified ``setUpCheckingPosAndStaticObs``, neptune_ros.cpp:940-990)."""
    c = poly.mean(axis=0)
    d = np.array([math.cos(theta), math.sin(theta)])
    ts = []
    for k in range(len(poly)):
        u, w = poly[k], poly[(k + 1) % len(poly)]
        e = w - u
        den = d[0] * e[1] - d[1] * e[0]
        if abs(den) < 1e-12:
            continue
        t = ((u[0] - c[0]) * e[1] - (u[1] - c[1]) * e[0]) / den
        s = ((u[0] - c[0]) * d[1] - (u[1] - c[1]) * d[0]) / den
        if -1e-12 <= s <= 1 + 1e-12:
            ts.append(t)
    return np.stack([c + min(ts) * d, c + max(ts) * d])  # col(0)=min_vert, col(1)=max_vert


def make_scene(par: Params, seed: int, *, n_fixed: int | None = None, sync: bool = True,
               ent_backend=None, agents: np.ndarray | None = None, spread: float | None = None,
               pack_hulls: bool = True, group_hulls: bool = False) -> Scene:
    """Build one replan cycle's inputs for the planning agents `agents` (0-based, default all).
    pack_hulls=False skips the host-side packed hull arrays (large worlds: the device builds the hulls
    from the committed-trajectory records, K1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N, M, T, S = par.num_of_agents, par.num_of_static_obst, par.T_span, par.num_sample_per_interval
    pb = np.asarray(par.pb, float)
    Ainv, V, A01 = solver_basis(T)
    agents = np.arange(N) if agents is None else np.asarray(agents)
    B = len(agents)
    static_raw = _static_obstacles(par, rng, pb)
    sd = 2 * par.drone_radius + 0.2  # Neptune::setStaticObst (neptune.cpp:642)
    corners = np.array([[1, 1], [1, -1], [-1, -1], [-1, 1]], float) * sd
    static_infl = [hull2d((sq[:, None, :] + corners[None]).reshape(-1, 2)) for sq in static_raw]
    strep = np.stack([_static_rep(sq, math.radians(33.0)) for sq in static_raw]) if M else np.zeros((0, 2, 2))
    st_boxes = np.array([[h[:, 0].min(), h[:, 1].min(), h[:, 0].max(), h[:, 1].max()] for h in static_infl]) \
        if M else np.zeros((0, 4))

    # current positions and goals
    R = spread if spread is not None else min(0.45 * par.tetherLength, 9.0)
    pos = np.zeros((N, 2))
    for j in range(N):
        for _ in range(2000):
            ang, rad = rng.uniform(0, 2 * math.pi), R * math.sqrt(rng.uniform(0.02, 1))
            c = pb[j] + rad * np.array([math.cos(ang), math.sin(ang)])
            if not (par.x_min + 1 < c[0] < par.x_max - 1 and par.y_min + 1 < c[1] < par.y_max - 1):
                continue
            if j and np.min(np.linalg.norm(pos[:j] - c, axis=1)) < 3.2:
                continue
            if np.min(np.linalg.norm(pb - c, axis=1)) < 1.2:
                continue
            if M and ((st_boxes[:, 0] - 0.3 < c[0]) & (c[0] < st_boxes[:, 2] + 0.3) &
                      (st_boxes[:, 1] - 0.3 < c[1]) & (c[1] < st_boxes[:, 3] + 0.3)).any():
                continue
            break
        pos[j] = c
    goals = np.zeros((N, 3))
    for j in range(N):
        for _ in range(2000):
            ang, rad = rng.uniform(0, 2 * math.pi), 0.85 * par.tetherLength * math.sqrt(rng.uniform(0, 1))
            g = pb[j] + rad * np.array([math.cos(ang), math.sin(ang)])
            if not (par.x_min + 1 < g[0] < par.x_max - 1 and par.y_min + 1 < g[1] < par.y_max - 1):
                continue
            if np.linalg.norm(g - pos[j]) >= (5.0 if n_fixed is None else 2.0):
                break
        goals[j] = [g[0], g[1], 1.0 + rng.uniform(0, 1.0)]
    z0 = 1.0 + rng.uniform(0, 0.5, size=N)

    # committed trajectories: an earlier plan of every agent, started t_age seconds ago
    committed = []
    t_age = rng.uniform(0.15, 0.9, size=N) if not sync else np.full(N, 0.5)
    for j in range(N):
        cx, cy = lattice_path(par, Ainv, V, pos[j], [0, 0], [0, 0], goals[j, :2], pb[j], None, par.num_pol)
        cz = initial_z_pwp(par, z0[j], 0.0, 0.0, goals[j, 2])[:len(cx)]
        times = -t_age[j] + T * np.arange(len(cx) + 1)
        committed.append((times, cx, cy, cz))

    # per planning agent: start time, start state A
    k_index = np.full(B, 4) if sync else rng.integers(2, 9, size=B)
    t_start = k_index * par.dc
    state_A = np.zeros((B, 3, 3))
    for bi, b in enumerate(agents):
        tm, cx, cy, cz = committed[b]
        for ax, c in enumerate((cx, cy, cz)):
            state_A[bi, :, ax] = eval_pwp(tm, c, t_start[bi])

    # hulls of every agent's committed trajectory per window (shared when all t_start are equal)
    delta = np.array([2 * par.drone_radius] * 2)  # bbox/2 + drone_radius (neptune.cpp:340, neptune_ros.cpp:448)
    uniq = sorted(set(float(t) for t in t_start))
    hull_cache = {}
    for ts in uniq:
        for j in range(N):
            tm, cx, cy, _ = committed[j]
            hull_cache[(ts, j)] = [hulls_of_interval(tm, cx, cy, ts + i * T, ts + (i + 1) * T, T, delta, A01)
                                   for i in range(par.num_pol)]
    samp_cache = {(ts, j): sample_points(committed[j][0], committed[j][1], committed[j][2], ts,
                                         ts + T * par.num_pol, par.num_pol, S) for ts in uniq for j in range(N)}

    # pack hulls: slot j holds agent j (empty for self)
    NH = N
    cnt = np.zeros((B, NH, NPOL), np.int64)
    chunks = []
    nih0 = np.full((B, N, NPOL, 2), np.nan)
    samp = np.zeros((B, N, NPOL, S + 1, 2))
    known = np.ones((B, N), np.uint8)
    for bi, b in enumerate(agents):
        ts = float(t_start[bi])
        known[bi, b] = 0
        for j in range(N):
            if j == b:
                continue
            hs = hull_cache[(ts, j)]
            if pack_hulls:
                for i in range(par.num_pol):
                    cnt[bi, j, i] = len(hs[i][0])
                    chunks.append(hs[i][0])
                    nih0[bi, j, i] = hs[i][1][0]
            samp[bi, j] = samp_cache[(ts, j)]
    hull_ptr = np.zeros(B * NH * NPOL + 1, np.int64)
    np.cumsum(cnt.reshape(-1), out=hull_ptr[1:])
    hull_xy = np.ascontiguousarray(np.concatenate(chunks, axis=0)) if chunks else np.zeros((0, 2))

    # front-end stub: pwp_init per planning agent against the other agents' window AABBs
    coeff_init = np.zeros((B, 3, NPOL, 4))
    n_int = np.zeros(B, np.int32)
    box_cache = {ts: [np.array([[h[:, 0].min(), h[:, 1].min(), h[:, 0].max(), h[:, 1].max()]
                                 for j in range(N) for h in (hull_cache[(ts, j)][i][0],)]) for i in range(par.num_pol)]
                 for ts in uniq}
    bx = np.array([[q[0] - 0.7, q[1] - 0.7, q[0] + 0.7, q[1] + 0.7] for q in pb])
    for bi, b in enumerate(agents):
        ts = float(t_start[bi])
        boxes = [np.concatenate([np.delete(box_cache[ts][i], b, axis=0), bx, st_boxes], axis=0)
                 for i in range(par.num_pol)]
        nmax = par.num_pol if n_fixed is None else n_fixed
        cx, cy = lattice_path(par, Ainv, V, state_A[bi, 0, :2], state_A[bi, 1, :2], state_A[bi, 2, :2],
                              goals[b, :2], pb[b], boxes, nmax)
        n = len(cx)
        cz = initial_z_pwp(par, state_A[bi, 0, 2], state_A[bi, 1, 2], state_A[bi, 2, 2], goals[b, 2])[:n]
        n_int[bi] = n
        coeff_init[bi, 0, :n], coeff_init[bi, 1, :n], coeff_init[bi, 2, :n] = cx, cy, cz

    st_ptr = np.zeros(M + 1, np.int64)
    if M:
        np.cumsum([len(h) for h in static_infl], out=st_ptr[1:])
    st_xy = np.concatenate(static_infl, axis=0) if M else np.zeros((0, 2))

    cap, NA = par.ent_cap, par.NA
    bp_cnt = np.ones(N, np.int32)
    bp_xy = np.zeros((N, par.bp_max, 2))
    bp_xy[:, 0] = pb
    prev_pos = np.zeros((B, N + 1, 2))
    prev_pos_agent = np.zeros((B, N, 2))
    for bi, b in enumerate(agents):
        tm, cx, cy, _ = committed[b]
        prev_pos[bi, :] = [eval_pwp(tm, cx, 0.0)[0], eval_pwp(tm, cy, 0.0)[0]]
        for j in range(N):
            tj, jx, jy, _ = committed[j]
            prev_pos_agent[bi, j] = [eval_pwp(tj, jx, 0.0)[0], eval_pwp(tj, jy, 0.0)[0]]

    batch = ReplanBatch(
        par=par, agent_id=(agents + 1).astype(np.int32), n_int=n_int, coeff_init=coeff_init, n_hull_slots=NH,
        hull_ptr=hull_ptr, hull_xy=hull_xy, nih0=nih0, st_ptr=st_ptr, st_xy=np.ascontiguousarray(st_xy),
        esv_cnt=np.zeros((B, NPOL + 1, 2), np.int32), esv_alpha=np.zeros((B, NPOL + 1, cap, 2), np.int32),
        esv_active=np.zeros((B, NPOL + 1, NA), np.int32), bp_cnt=bp_cnt, bp_xy=bp_xy)
    sc = Scene(par=par, batch=batch, t_start=t_start, committed=committed, goals=goals, static_raw=static_raw,
               strep=strep, samp=samp, known=known, prev_pos=prev_pos, prev_pos_agent=prev_pos_agent,
               state_A=state_A, es0_cnt=np.zeros((B, 2), np.int32), es0_alpha=np.zeros((B, cap, 2), np.int32),
               es0_beta=np.zeros((B, cap)), es0_bend=np.zeros((B, cap), np.int32),
               es0_active=np.zeros((B, NA), np.int32))
    if group_hulls:  # fixed-stride hulls / samples per window group: the front-end search's view
        G = len(uniq)
        sc.group = np.array([uniq.index(float(t)) for t in t_start], np.int32)
        sc.hull_g_xy = np.zeros((G, N, NPOL, 24, 2))
        sc.hull_g_cnt = np.zeros((G, N, NPOL), np.int32)
        sc.samp_g = np.zeros((G, N, par.num_pol, S + 1, 2))
        for gi, ts in enumerate(uniq):
            for j in range(N):
                for i in range(par.num_pol):
                    hj = hull_cache[(ts, j)][i][0]
                    sc.hull_g_cnt[gi, j, i] = len(hj)
                    sc.hull_g_xy[gi, j, i, :len(hj)] = hj
                sc.samp_g[gi, j] = samp_cache[(ts, j)]
    if ent_backend is not None:
        fill_entangle(sc, ent_backend, rng)
    batch.validate()
    return sc


def search_host_inputs(sc: Scene, seed: int) -> dict:
    """Per-agent inputs of the front-end search that do not come from other kernels: start state A, goal,
    initial z polynomial (getInitialZPwp) and a seeded order of the jerk samples."""
    from .search import jerk_order

    par, b = sc.par, sc.batch
    B = b.B
    init = np.zeros((B, 6))
    init[:, 0:2], init[:, 2:4], init[:, 4:6] = sc.state_A[:, 0, :2], sc.state_A[:, 1, :2], sc.state_A[:, 2, :2]
    agents = b.agent_id - 1
    coeffs_z = np.zeros((B, NPOL, 4))
    for bi in range(B):
        coeffs_z[bi, :par.num_pol] = initial_z_pwp(par, sc.state_A[bi, 0, 2], sc.state_A[bi, 1, 2], sc.state_A[bi, 2, 2],
                                                   sc.goals[agents[bi], 2])
    return dict(init=init, goal=np.ascontiguousarray(sc.goals[agents, :2]), coeffs_z=coeffs_z,
                comb=jerk_order(par, seed, B))


def make_search_batch(sc: Scene, seed: int, per_agent_order: bool = False):
    """Inputs of KinodynamicSearch::setUp / run for every planning agent of a scene built with
    ``group_hulls=True`` and an ``ent_backend`` (``neptune.cpp:1437-1453``)."""
    from .search import SearchBatch, jerk_order, static_longest_dist

    par, b = sc.par, sc.batch
    assert sc.hull_g_xy is not None and sc.esA_cnt is not None
    B = b.B
    init = np.zeros((B, 6))
    init[:, 0:2], init[:, 2:4], init[:, 4:6] = sc.state_A[:, 0, :2], sc.state_A[:, 1, :2], sc.state_A[:, 2, :2]
    agents = b.agent_id - 1
    goal = np.ascontiguousarray(sc.goals[agents, :2])
    cz = np.stack([initial_z_pwp(par, sc.state_A[bi, 0, 2], sc.state_A[bi, 1, 2], sc.state_A[bi, 2, 2],
                                 sc.goals[agents[bi], 2]) for bi in range(B)])
    coeffs_z = np.zeros((B, NPOL, 4))
    coeffs_z[:, :par.num_pol] = cz
    M = par.num_of_static_obst
    strep = np.ascontiguousarray(sc.strep, np.float64).reshape(M, 2, 2)
    sb = SearchBatch(
        par=par, agent_id=b.agent_id.copy(), init=init, goal=goal, coeffs_z=coeffs_z, group=sc.group.copy(),
        hull_xy=sc.hull_g_xy, hull_cnt=sc.hull_g_cnt, samp=sc.samp_g, known=sc.known.copy(),
        es_cnt=np.ascontiguousarray(sc.esA_cnt, np.int32), es_alpha=np.ascontiguousarray(sc.esA_alpha, np.int32),
        es_beta=np.ascontiguousarray(sc.esA_beta, np.float64), es_bend=np.ascontiguousarray(sc.esA_bend, np.int32),
        es_active=np.ascontiguousarray(sc.esA_active, np.int32), bp_cnt=b.bp_cnt, bp_xy=b.bp_xy,
        comb=jerk_order(par, seed, B if per_agent_order else None), st_ptr=b.st_ptr, st_xy=b.st_xy, strep=strep,
        st_longest=static_longest_dist(sc.static_raw, strep) if M else np.zeros((0, 2)))
    sb.validate()
    return sb


def fill_entangle(sc: Scene, be, rng) -> None:
    """History walk -> entangle_state_, PredictAlphasBetas -> entangle_state_A, front-end chain ->
    entStateVec, all through the backend `be` (methods: rollout_batch, predict_batch)."""
    par, batch = sc.par, sc.batch
    B, N, T, S = batch.B, par.num_of_agents, par.T_span, par.num_sample_per_interval
    pb = np.asarray(par.pb, float)
    # history: a straight-line walk base-side waypoint -> random waypoint -> current position with
    # the other agents parked at their current positions
    hist = np.zeros((B, 3, NPOL, 4))
    hn = np.zeros(B, np.int32)
    samp_h = np.zeros_like(sc.samp)
    for bi in range(B):
        b = int(batch.agent_id[bi]) - 1
        cur = sc.prev_pos[bi, 0]
        start = pb[b] + 0.8 * (cur - pb[b]) / max(np.linalg.norm(cur - pb[b]), 1e-9)
        mid = 0.5 * (start + cur) + rng.normal(0, 3.0, size=2)
        wps = [start] + [start + (mid - start) * k / 4 for k in (1, 2, 3, 4)] + [mid + (cur - mid) * k / 4 for k in (1, 2, 3, 4)]
        for i in range(8):
            vel = (wps[i + 1] - wps[i]) / T
            hist[bi, 0, i] = [0, 0, vel[0], wps[i][0]]
            hist[bi, 1, i] = [0, 0, vel[1], wps[i][1]]
        hn[bi] = 8
        samp_h[bi] = sc.prev_pos_agent[bi][:, None, None, :]
    done, cnt, alpha, beta, bend, active = be.rollout_batch(
        par, batch.agent_id, hn, hist, samp_h, sc.known, sc.strep, batch.bp_cnt, batch.bp_xy,
        sc.es0_cnt, sc.es0_alpha, sc.es0_beta, sc.es0_bend, sc.es0_active)
    for bi in range(B):
        if done[bi] == 8:  # a non-entangling history: adopt its final state
            sc.es0_cnt[bi], sc.es0_alpha[bi], sc.es0_beta[bi] = cnt[bi, 8], alpha[bi, 8], beta[bi, 8]
            sc.es0_bend[bi], sc.es0_active[bi] = bend[bi, 8], active[bi, 8]
    # PredictAlphasBetas (neptune.cpp:976-1008)
    cur = np.ascontiguousarray(sc.state_A[:, 0, :2])
    samp0 = np.ascontiguousarray(sc.samp[:, :, 0, 0, :])
    sc.esA_cnt, sc.esA_alpha, sc.esA_beta, sc.esA_bend, sc.esA_active = be.predict_batch(
        par, batch.agent_id, sc.prev_pos, sc.prev_pos_agent, cur, samp0, sc.known, sc.strep, batch.bp_cnt,
        batch.bp_xy, sc.es0_cnt, sc.es0_alpha, sc.es0_beta, sc.es0_bend, sc.es0_active)
    # entStateVec along pwp_init (recoverEntStateVector, kinodynamic_search.cpp:582-603)
    done, cnt, alpha, beta, bend, active = be.rollout_batch(
        par, batch.agent_id, batch.n_int, batch.coeff_init, sc.samp, sc.known, sc.strep, batch.bp_cnt,
        batch.bp_xy, sc.esA_cnt, sc.esA_alpha, sc.esA_beta, sc.esA_bend, sc.esA_active)
    batch.esv_cnt[:], batch.esv_alpha[:], batch.esv_active[:] = cnt, alpha, active

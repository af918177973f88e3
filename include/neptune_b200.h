/*
 * neptune_b200.h -- C-ABI of the B200-native NEPTUNE replan hot path (libneptune_b200.so).
 *
 * Drop-in boundary: these entry points are what a cgo/ctypes/C++ binding of
 * the reference's back end would bind.  Plain pointers and sizes only; no torch, Eigen or CUDA types
 * in the signatures (streams travel as void*).  All functions return 0 on success or a negative
 * nb_error; nothing throws across the ABI.  Paths cited are relative to the reference tree.
 *
 * Batch convention: every call processes B independent replans ("one replan = one agent, one
 * cycle").  The reference runs B = 1 per process (one PolySolverGurobi per agent,
 * neptune/include/neptune.hpp:183); the C++ shim in neptune_b200/csrc/poly_solver_b200.hpp does the
 * same through this ABI, the benchmark harness passes B = number of agents on the rank.
 *
 * Memory spaces: pointers inside nb_replan_args / nb_* batches are HOST pointers when
 * `space == NB_HOST` (the library stages them through pinned memory and copies results back:
 * the end-to-end path) or DEVICE pointers when `space == NB_DEVICE` (inputs resident in HBM).
 */
#ifndef NEPTUNE_B200_H
#define NEPTUNE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_NPOL 8 /* storage stride for intervals (num_pol <= 8 in every shipped YAML) */

typedef enum nb_error
{
  NB_OK = 0,
  NB_ERR_CUDA = -1,       /* a CUDA call failed; nb_last_error() has the text */
  NB_ERR_ARG = -2,        /* invalid argument */
  NB_ERR_CAPACITY = -3,   /* a fixed-capacity list would overflow (ent_slots / ent_cap / bp_max) */
  NB_ERR_NO_DEVICE = -4   /* no CUDA device: there is no CPU fallback */
} nb_error;

typedef enum nb_space { NB_HOST = 0, NB_DEVICE = 1 } nb_space;

/* status of one replan == PolySolverGurobi::optimize status path (solver_gurobi_poly.cpp:832-861) */
#define NB_STATUS_OK 0       /* first solve accepted */
#define NB_STATUS_FALLBACK 1 /* terminal v/a rows dropped + soft cost, re-solve accepted (:838-847) */
#define NB_STATUS_FAILED 2   /* both failed: pwp_out = pwp_init, optimize() == false (:856-859) */

/*
 * Replaces the constructor + one-time setters of PolySolverGurobi
 * (solver_gurobi_poly.cpp:25-134 ctor; :140-175 setMaxValues; :177-185 setTetherLength/setMaxRuntime;
 * call site neptune.cpp:102-107).
 */
typedef struct nb_params
{
  int32_t num_pol;      /* ctor num_pol (<= NB_NPOL) */
  int32_t deg_pol;      /* ctor deg_pol; only 3 is supported, as in the reference */
  int32_t num_agents;   /* pb.size() */
  int32_t num_static;   /* setStaticObstVert: number of static obstacles */
  int32_t samples;      /* num_sample_per_interval (entangle check) */
  int32_t use_linear_constraints; /* ctor flag; only 1 (the shipped mode) is supported */
  double T_span;        /* ctor T_span */
  double weight;        /* ctor weight_term */
  double lim_min[3];    /* setMaxValues x_min,y_min,z_min */
  double lim_max[3];    /* setMaxValues x_max,y_max,z_max */
  double v_max, a_max;  /* setMaxValues v_max, a_max (j_max is not a QP constraint) */
  double drone_radius;  /* hull inflation (neptune.cpp:340) */
  double tether_length; /* setTetherLength (stored; not a QP constraint in the reference either) */
  int32_t ent_cap;      /* storage capacity of an alphas list */
  int32_t bp_max;       /* storage capacity of a bend-point list (base included) */
  int32_t ent_slots;    /* LP slots per interval reserved for non-entangling constraints */
  int32_t ipm_max_iter; /* stands in for setMaxRuntime: iteration cap of the interior-point solve */
  double ipm_tol;
} nb_params;

typedef struct nb_handle nb_handle;

/* pb: [num_agents][2] base positions (ctor argument pb). device < 0 selects the current device. */
int nb_create(const nb_params* par, const double* pb, int device, nb_handle** out);
void nb_destroy(nb_handle* h);
const char* nb_last_error(void);
/* number of kernels this library has launched since nb_create (bench.py's gpu_launches) */
long long nb_launch_count(const nb_handle* h);
/* Measurement hooks (no reference counterpart).  With profiling on, every kernel of
 * nb_replan_batch is bracketed by CUDA events on the launching stream; nb_kernel_times waits for
 * them and writes the durations of the last call in ms: [0] line generation, [1] QP. */
int nb_set_profiling(nb_handle* h, int on);
int nb_kernel_times(nb_handle* h, double* ms, int n);
/* NB_DEVICE calls are asynchronous: capacity overflows are latched on the device; this call
 * synchronises `stream` and returns NB_ERR_CAPACITY if any call since the last check overflowed. */
int nb_check_async_errors(nb_handle* h, void* stream);

/*
 * Replaces PolySolverGurobi::setStaticObstVert (solver_gurobi_poly.cpp:316-320; caller
 * Neptune::setStaticObst neptune.cpp:639-664 passes the INFLATED hulls).  Host pointers.
 * st_ptr: [M+1] vertex offsets, st_xy: [st_ptr[M]][2].  strep: [M][2][2] staticObsRep_ (col0,col1)
 * used by the entanglement chain (Neptune::setStaticObstRep neptune.cpp:666-671), may be NULL if M==0.
 */
int nb_set_static(nb_handle* h, const int64_t* st_ptr, const double* st_xy, const double* strep);

/*
 * One batch of back-end replans.  Replaces, per agent, the per-replan call sequence
 *   setInitTrajectory (:187-244) -> setHulls (:246-281) -> setHullsNoInflation (:283-288) ->
 *   setEntStateVector (:307-314) -> optimize (:804-887)            [neptune.cpp:1514-1519]
 * including every separator::Separator::solveModel call optimize() makes
 * (separator_glpk.cpp:248-373 and :375-498; call sites solver_gurobi_poly.cpp:480, :543, :584, :751).
 */
typedef struct nb_replan_args
{
  int32_t B;
  int32_t space;              /* nb_space of every pointer below */
  const int32_t* agent_id;    /* [B] 1-based id (ctor argument id) */
  const int32_t* n_int;       /* [B] intervals of pwp_init */
  const double* coeff_init;   /* [B][3][8][4] pwp_init coeff_x/y/z, [a b c d], t in seconds */
  int32_t n_hull_slots;       /* hull slots per agent; slot s, interval i -> polygon hull_ptr[(b*slots+s)*8+i] */
  const int64_t* hull_ptr;    /* [B*slots*8+1] vertex offsets; empty polygon = slot unused */
  const double* hull_xy;      /* [hull_ptr[last]][2] CCW convex polygons (setHulls) */
  int64_t hull_nvert;         /* total vertices in hull_xy */
  const int32_t* hull_cnt;    /* optional [B*slots*8] vertex counts; NULL: count = hull_ptr[k+1]-hull_ptr[k] (CSR).
                                 With counts, hull_ptr needs only B*slots*8 entries (nb_hulls_batch output). */
  const double* nih0;         /* [B][N][8][2] col(0) of hullsNoInflation_[j][i]; NaN = unknown */
  const int32_t* nih0_group;  /* optional [B]: nih0 is then [G][N][8][2] and agent b reads block nih0_group[b]
                                 (agents planning over the same windows share it, see nb_hull_index_batch) */
  const uint8_t* hull_known;  /* optional [B][N], with nih0_group: shared-window mode for the hulls too --
                                 hull_xy / hull_cnt are the group-shaped outputs of nb_hulls_batch
                                 ([G][N][8][NB_HULL_STRIDE][2], [G][N][8]), hull_ptr is ignored, slot j of agent
                                 b is hull (nih0_group[b], j, i), empty when j is b itself or hull_known[b][j]
                                 == 0 (what nb_hull_index_batch would materialise) */
  const int32_t* esv_cnt;     /* [B][9][2] (alphas.size(), bendPointsIdx.size()) of entStateVec[i] */
  const int32_t* esv_alpha;   /* [B][9][ent_cap][2] */
  const int32_t* esv_active;  /* [B][9][N+M] active_cases */
  const int32_t* bp_cnt;      /* [N] bendPtsForAgents_[j].size() (shared by the batch) */
  const double* bp_xy;        /* [N][bp_max][2] */
  /* outputs */
  double* coeff_out;          /* [B][3][8][4] pwp_out_ coefficients */
  double* obj;                /* [B] objective_value */
  int32_t* status;            /* [B] NB_STATUS_* */
  int32_t* iters;             /* [B][2] interior-point iterations (direct, fallback) */
  double* lines;              /* optional [B][8][LS][3] every separating line (n0,n1,d) by slot */
  uint8_t* line_ok;           /* optional [B][8][LS]: 0 not attempted, 1 solved, 2 unsolved */
} nb_replan_args;

/* Measurement hook: lines the pruning kept per (agent, interval) in the last nb_replan_batch, out [B][8] (host). */
int nb_kept_lines(nb_handle* h, int32_t* out, int32_t B);

/* LS = n_hull_slots + num_agents + num_static + ent_slots: agents | bases | static | non-entangling */
int nb_line_slots(const nb_handle* h, int n_hull_slots);

/* stream: cudaStream_t as void* (NULL = default stream).  Synchronous w.r.t. the host for
 * NB_HOST arguments; asynchronous on `stream` for NB_DEVICE arguments. */
int nb_replan_batch(nb_handle* h, const nb_replan_args* args, void* stream);

/*
 * Replaces separator::Separator::solveModel, 2-D overloads (separator_glpk.cpp:248, :375, :500), L LPs
 * at once.  Point set A of LP l is a_xy[a_ptr[l] .. a_ptr[l+1]) (for the 4-arg overload the caller
 * appends pointsAPlus to pointsA), B likewise.  a_polygon != 0 promises every A is a convex polygon
 * in cyclic order (what CGAL returns); 0 makes no assumption.  out_line: [L][3] (n0,n1,d),
 * out_ok: [L] 1 = GLP_OPT/GLP_FEAS, 0 = infeasible.
 */
int nb_separate_batch(nb_handle* h, int32_t L, int32_t space, const int64_t* a_ptr, const double* a_xy,
                      const int64_t* b_ptr, const double* b_xy, int32_t a_polygon, double* out_line,
                      uint8_t* out_ok, void* stream);

/* PolySolverGurobi::generatePwpOut sampling loop (:911-934): states [B][max_states][12]
 * (pos, vel, accel, jerk), n_states[B]. */
int nb_generate_traj_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* n_int, const double* coeff,
                           double dc, int32_t max_states, double* states, int32_t* n_states, void* stream);

/*
 * Committed-trajectory record: the fixed-stride payload of the per-cycle exchange, the batched form of
 * mader_msgs/msg/DynTraj.msg:1-9 + PieceWisePolTraj.msg (what NeptuneRos::publishOwnTraj sends, neptune_ros.cpp:436-480,
 * and NeptuneRos::trajCB reads, :379-432): NB_REC_DOUBLES = 256 doubles (2 KB) per agent,
 *   [0] n_pieces (<= 16), [1..17] times, [18..209] coeff[3][16][4] ([a b c d] per axis and piece)   -- pwp
 *   [210] id, [211] is_agent, [212..214] bbox, [215..217] pos, [218] bendpt.size(), [219..234] bendpt[8] (x, y),
 *   [235] sequence number of the commit, [236..255] reserved (zero).
 */
#define NB_REC_PIECES 16
#define NB_REC_PWP_DOUBLES (1 + (NB_REC_PIECES + 1) + 3 * NB_REC_PIECES * 4)
#define NB_REC_DOUBLES 256
#define NB_REC_OFF_ID 210
#define NB_REC_OFF_ISAGENT 211
#define NB_REC_OFF_BBOX 212
#define NB_REC_OFF_POS 215
#define NB_REC_OFF_NBEND 218
#define NB_REC_OFF_BEND 219
#define NB_REC_OFF_SEQ 235
#define NB_HULL_STRIDE 24 /* vertices reserved per hull in nb_hulls_batch output */

/*
 * Replaces Neptune::convexHullsOfCurves2d (neptune.cpp:224-267 and callees :269-452, with
 * cu::convexHullOfPoints2d cgal_utils.cpp:157-174) and Neptune::SamplePointsOfCurves
 * (neptune.cpp:463-566) for B planning agents against the N committed trajectories `recs`.
 * Windows are [t_start[b] + i T, t_start[b] + (i+1) T], i < num_pol.  delta = bbox/2 + drone_radius
 * (neptune.cpp:340).  known [B][N]: 0 = no trajectory of j at agent b (or j is b itself).
 * Outputs (fixed stride, directly usable as nb_replan_args.hull_* with n_hull_slots = N):
 *   hull_xy [B][N][8][NB_HULL_STRIDE][2], hull_cnt [B][N][8] (0 when unknown),
 *   hull_ptr [B*N*8] (= k * NB_HULL_STRIDE), nih0 [B][N][8][2] (NaN when unknown),
 *   samp [B][N][8][S+1][2], idx [B][N][8][2] = (index_first_interval, index_last_interval), optional.
 */
int nb_hulls_batch(nb_handle* h, int32_t B, int32_t space, const double* t_start, const double* recs,
                   const uint8_t* known, double delta, double* hull_xy, int32_t* hull_cnt, int64_t* hull_ptr,
                   double* nih0, double* samp, int32_t* idx, void* stream);

/*
 * Window sharing.  Agents whose replans start at the same t_start look at the same windows of every
 * other agent, so hulls / samples need to be built once per distinct t_start ("group"), not once per
 * planning agent: call nb_hulls_batch with B = G groups, then this function to give every planning agent
 * its own (hull_ptr, hull_cnt) view of the shared hulls: slot j of agent b points at hull (group[b], j, i)
 * and is empty when j is b itself or known[b][j] == 0.  Results are identical to the per-agent call.
 * hull_cnt_g [G][N][8] -> hull_ptr [B][N][8], hull_cnt [B][N][8].
 */
int nb_hull_index_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id, const int32_t* group,
                        const uint8_t* known, const int32_t* hull_cnt_g, int64_t* hull_ptr, int32_t* hull_cnt,
                        void* stream);

/* nb_postcheck_batch on hulls already built by nb_hulls_batch from the late records, per group:
 * hull_xy_g [G][N][8][NB_HULL_STRIDE][2], hull_cnt_g [G][N][8]. */
int nb_postcheck_hulls_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* n_int, const double* coeff,
                             const int32_t* group, const double* hull_xy_g, const int32_t* hull_cnt_g,
                             const uint8_t* late, int32_t* collide, void* stream);

/*
 * Replaces the geometric half of Neptune::safetyCheckAfterReplan (neptune.cpp:719-765):
 * trajsAndPwpAreInCollision2d (:767-806) with gjk::collision (gjk.cpp:76-148), for every (b, j) with
 * late[b][j] != 0 (trajectory of j received after time_init_opt_).  collide [B] = 1 if any collides.
 */
int nb_postcheck_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* n_int, const double* coeff,
                       const double* t_start, const double* recs, const uint8_t* late, double delta,
                       int32_t* collide, void* stream);

/*
 * Committed-trajectory records of this rank's agents for the all-gather: the pwp_now of
 * PolySolverGurobi::generatePwpOut (times shifted by t_start, solver_gurobi_poly.cpp:892-907).
 * recs_out [B][NB_REC_DOUBLES].  mu::composePieceWisePol with the previous plan (utils.cpp:318-402) is
 * applied by nb_commit_compose_batch / nb_compose_records_batch below.
 */
int nb_commit_records_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* n_int, const double* coeff,
                            const double* t_start, double* recs_out, void* stream);

/*
 * Fused commit of the resident cycle, the tail of Neptune::replanFull (neptune.cpp:1685-1699): pwp_now from
 * coeff/t_start as above, then recs_out[b] = composePieceWisePol(t_now[b], dc, prev[prev_agent[b] - 1], pwp_now)
 * (prev_agent: 1-based agent ids as in nb_replan_args, NULL: prev[b]; has_prev NULL: all have a previous plan).  Where status[b] >= 2, entangled[b] != 0
 * or collide[b] != 0 (each nullable) the replan is rejected and recs_out[b] = the previous record, as
 * replanFull returns before publishing.  Device pointers only.  Overflow of the 16-piece record is reported
 * by nb_check_async_errors.  n_pieces [B] nullable.
 */
int nb_commit_compose_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* n_int, const double* coeff,
                            const double* t_start, const double* t_now, const double* prev, const int32_t* prev_agent,
                            const uint8_t* has_prev, const int32_t* status, const int32_t* entangled,
                            const int32_t* collide, double* recs_out, int32_t* n_pieces, void* stream);

/*
 * Replaces mu::composePieceWisePol (neptune/src/utils.cpp:318-402) as used by Neptune::replanFull
 * (neptune.cpp:1689-1699): out[b] = pieces of prev[b] between t[b] and the start of now[b], then now[b];
 * where has_prev[b] == 0, out[b] = now[b] (exists_previous_pwp_ == false).  All [B][NB_REC_DOUBLES].
 * n_pieces [B]: pieces of out (0 = the reference's empty "dummy" result).  SURVEY section 8(f) "next #2".
 */
int nb_compose_records_batch(nb_handle* h, int32_t B, int32_t space, const double* t, const uint8_t* has_prev,
                             const double* prev, const double* now, double* out, int32_t* n_pieces, void* stream);

/*
 * Entanglement-signature chain, batched (one agent per warp).  State = eu::ent_state
 * (neptune/include/entangle_utils.hpp:23-29) with fixed storage: cnt [B][2] = (alphas.size(),
 * bendPointsIdx.size()), alpha [B][ent_cap][2], beta [B][ent_cap], bend [B][ent_cap], active [B][N+M].
 * known: [B][N] 1 where SampledPointsForAll[j] is non-empty.  samp: sampled positions of the other
 * agents, [B][N][8][S+1][2] or, when samp_shared != 0, [N][8][S+1][2] shared by the batch.
 * Static representation and base points come from nb_set_static / nb_create.
 */
typedef struct nb_ent_state
{
  int32_t* cnt;
  int32_t* alpha;
  double* beta;
  int32_t* bend;
  int32_t* active;
} nb_ent_state;

/* Replaces Neptune::PredictAlphasBetas (neptune.cpp:976-1008) and its eu:: callees
 * (entangle_utils.cpp:1129-1228, :1231-1277, :1402-1534, :1536-1604): state is updated in place.
 * prev_pos [B][N+1][2] previousCheckingPos_, prev_pos_agent [B][N][2], cur [B][2] (start point A),
 * samp0 [B][N][2] = SampledPointsForAll[j][0].col(0). */
int nb_entangle_predict_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id, const uint8_t* known,
                              const int32_t* bp_cnt, const double* bp_xy, nb_ent_state st, const double* prev_pos,
                              const double* prev_pos_agent, const double* cur, const double* samp0, void* stream);

/* Replaces NeptuneRos::updateEntStateStaticObs (neptune_ros.cpp:798-850), one odometry tick of the online tracker
 * that maintains entangle_state_, with the 9-argument eu::entangleHSigToAddAgentInd (entangle_utils.cpp:820-1127:
 * a bend point of the other tether added or released since the last check).  SURVEY section 8(f) "next #3".
 * state, prev_pos [B][N+1][2] (previousCheckingPos_) and prev_pos_agent [B][N][2] (previousCheckingPosAgent_; x < -900 =
 * no message from that agent yet) are updated in place; latest_pos_agent [B][N][2] = latestCheckingPosAgent_;
 * bp_*_prev = bendPtsForAgents_prev_ (the caller sets prev = current after the call, :818); elapsed_ms [B] replaces
 * the wall-clock timer of :803-804.  result [B]: 0 updated, 1 skipped by the gate, -k where the reference prints
 * "stop k" and calls exit(-1). */
int nb_entangle_track_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id, const int32_t* bp_cnt,
                            const double* bp_xy, const int32_t* bp_cnt_prev, const double* bp_xy_prev, nb_ent_state st,
                            double* prev_pos, double* prev_pos_agent, const double* latest_pos_agent, const double* cur,
                            const double* elapsed_ms, int32_t* result, void* stream);

/* Replaces the per-interval chain of KinodynamicSearch::entanglesWithOtherAgents
 * (kinodynamic_search.cpp:707-895; list bound N+M, tether-length test excluded) that produces
 * entStateVec (recoverEntStateVector :582-603).  in: state at A.  out: states after 0..n intervals,
 * [B][9][...] (entries past n_int are copies of the last), done [B] = intervals before the first
 * entangling step (n if none). */
int nb_entangle_rollout_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id, const uint8_t* known,
                              const int32_t* bp_cnt, const double* bp_xy, nb_ent_state in, const int32_t* n_int,
                              const double* coeff, const double* samp, int32_t samp_shared, nb_ent_state out,
                              int32_t* done, void* stream);

/* Replaces KinodynamicSearch::entangleCheckGivenPwp (kinodynamic_search.cpp:897-985), including its
 * quirk of evaluating interval 0 only.  state updated in place; entangled [B] = its return value. */
int nb_entangle_check_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id, const uint8_t* known,
                            const int32_t* bp_cnt, const double* bp_xy, nb_ent_state st, const int32_t* n_int,
                            const double* coeff, const double* samp, int32_t samp_shared, int32_t* entangled,
                            void* stream);

/*
 * Entanglement half of Neptune::safetyCheckAfterReplan (neptune.cpp:735-752), faithful to its gating: for an agent with
 * at least one late trajectory (late[b][j] != 0: received after time_init_opt_), SampledPointsForAll[j] of every late j
 * is re-sampled from late_recs over [t_start, t_start + n T] (SamplePointsOfIntervals, :737-738), the front end's copy
 * of j's bend points becomes the late message's (updateSPocAndbendPtsForAgent :740-741; bp_*_late), PredictAlphasBetas
 * is re-run from entangle_state_ `st` (read only; :747-749) and entangleCheckGivenPwp decides (:750).  Agents without
 * a late trajectory: entangled = 0 (need_to_rerun_entanglecheck stays false).  samp: the planning-time samples,
 * [B][N][num_pol][S+1][2], or shared / group-indexed like nb_entangle_check_batch (samp_group [B], device only).
 */
int nb_postcheck_entangle_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id, const uint8_t* known,
                                const uint8_t* late, const int32_t* bp_cnt, const double* bp_xy, const int32_t* bp_cnt_late,
                                const double* bp_xy_late, nb_ent_state st, const double* prev_pos, const double* prev_pos_agent,
                                const double* cur, const int32_t* n_int, const double* coeff, const double* t_start,
                                const double* samp, int32_t samp_shared, const int32_t* samp_group, const double* late_recs,
                                int32_t* entangled, void* stream);

/*
 * Replaces the bookkeeping of NeptuneRos::trajCB besides updateTrajObstacles (neptune_ros.cpp:413-431) for all N
 * records at once: bendPtsForAgents_[j] = msg.bendpt -> bp_cnt [N], bp_xy [N][bp_max][2]; latestCheckingPosAgent_[j] =
 * msg.pos -> latest_pos [N][2] (nullable).  SURVEY section 8(f) #2.
 */
int nb_unpack_records_batch(nb_handle* h, int32_t space, const double* recs, int32_t* bp_cnt, double* bp_xy,
                            double* latest_pos, void* stream);

/*
 * Replaces the tail of Neptune::replanFull (neptune.cpp:1685-1699) followed by NeptuneRos::publishOwnTraj
 * (neptune_ros.cpp:436-480): nb_commit_compose_batch plus the DynTraj header -- id, is_agent, bbox (3 x `bbox`), pos =
 * start of the composed trajectory, bendpt[] = the tether base and the contact point of every entry of
 * entangle_state_.bendPointsIdx (`es`, [B] states; es.cnt == NULL: base only), `seq` in the sequence slot.
 * fe_solved (nullable): 0 = the front end returned no path, the replan is rejected (neptune.cpp:1473-1478).  t_now NULL:
 * no composition (recs_out = pwp_now + header).  Device pointers only.
 */
int nb_publish_records_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id, const int32_t* n_int,
                             const double* coeff, const double* t_start, const double* t_now, const double* prev,
                             const uint8_t* has_prev, const int32_t* status, const int32_t* entangled, const int32_t* collide,
                             const int32_t* fe_solved, nb_ent_state es, double bbox, double seq, double* recs_out,
                             int32_t* n_pieces, void* stream);

/*
 * Every reference process derives its own staticObsRep_ / staticObsLongestDist_ from its base and start position
 * (NeptuneRos::setUpCheckingPosAndStaticObs, neptune_ros.cpp:852-1019; nb_static_obst_rep computes one agent's).  This
 * call gives the batched kernels one representation PER AGENT: strep_all [N][M][2][2], longest_all [N][M][2] (nullable
 * when the front end is not used), indexed by agent id - 1.  It replaces the `strep` of nb_set_static, which stays the
 * shared-representation form.  Host pointers.
 */
int nb_set_static_rep_per_agent(nb_handle* h, const double* strep_all, const double* longest_all);

/*
 * ---- Front end: KinodynamicSearch (neptune/src/kinodynamic_search.cpp), SURVEY.md section 8(f) #1 ----
 *
 * Replaces the one-time setters of KinodynamicSearch (setMaxValuesAndSamples :272-327, setXYZMinMaxAndRa
 * :329-344, setBias :346, setGoalSize :360, setRunTime :356, constructor flags :29-138; call site
 * neptune.cpp:88-97).  Bounds, v_max, a_max, T_span, num_pol, samples and the tether length come from
 * nb_params.  Two fields replace non-deterministic state of the reference: max_expansions (pops of the open
 * list allowed, for the wall-clock max_runtime_ of :1646) and the jerk-sample order passed per call.
 */
typedef struct nb_search_params
{
  int32_t num_samples;        /* a_star_samp_x (jerk samples per axis, <= 5) */
  double j_max;
  double voxel_size;          /* a_star_fraction_voxel_size */
  double bias;                /* setBias (1.1 in neptune.cpp:97) */
  double goal_size;           /* goal_radius */
  int32_t enable_entangle_check;
  int32_t use_not_reaching_soln;
  int32_t max_nodes;          /* node_num_max_ (:370-371) */
  int32_t max_expansions;     /* stands in for max_runtime_ */
  int32_t ecap;               /* storage capacity of a search node's alphas list (semantic bound N+M) */
} nb_search_params;

int nb_search_configure(nb_handle* h, const nb_search_params* sp);

/*
 * Replaces NeptuneRos::setUpCheckingPosAndStaticObs (neptune_ros.cpp:852-1019), the one-time choice of the two
 * representative points of every static obstacle (staticObsRep_) and of staticObsLongestDist_ for ONE agent.  Host
 * function (no device work, like the reference's).  poly_*: the raw (un-inflated) obstacles, CSR; base / pos: the
 * agent's base and current position; voxel_size: a_star_fraction_voxel_size.  strep [M][2][2], longest [M][2].
 * Returns NB_ERR_ARG where the reference prints "cannot find a feasible vertex representation" and exits.
 * SURVEY section 8(f) "next #3".
 */
int nb_static_obst_rep(int32_t M, const int64_t* poly_ptr, const double* poly_xy, const double* base, const double* pos,
                       double voxel_size, double* strep, double* longest);

/* Second argument of KinodynamicSearch::setStaticObstRep (:385-390): staticObsLongestDist [M][2]. Host pointer. */
int nb_set_static_longest(nb_handle* h, const double* longest);

/*
 * One batch of front-end searches.  Replaces, per agent, KinodynamicSearch::setUp (:190-249) + run
 * (:1629-1827) + getPwpOut_0tstart / getEntStateVector (:610-618)          [neptune.cpp:1453-1510].
 * Hulls and samples are the group-shaped outputs of nb_hulls_batch (agents planning over the same
 * windows share them): agent b reads group[b] (NULL: b), the slot of b itself and of agents with
 * known[b][j] == 0 is empty.  The outputs coeff / n_int / esv are laid out exactly like
 * nb_replan_args.coeff_init / n_int / esv_* so the back end can consume them in place.
 */
typedef struct nb_search_args
{
  int32_t B;
  int32_t space;
  const int32_t* agent_id;   /* [B] 1-based */
  const double* init;        /* [B][6] px py vx vy ax ay of the start state A */
  const double* goal;        /* [B][2] */
  const double* coeffs_z;    /* [B][8][4] setInitZCoeffs (getInitialZPwp) */
  int32_t n_groups;          /* G */
  const int32_t* group;      /* [B] or NULL */
  const double* hull_xy;     /* [G][N][8][NB_HULL_STRIDE][2] */
  const int32_t* hull_cnt;   /* [G][N][8] */
  const double* samp;        /* [G][N][num_pol][S+1][2] */
  const uint8_t* known;      /* [B][N] */
  nb_ent_state es;           /* entangle_state_A, [B] states, stride ent_cap */
  const int32_t* bp_cnt;     /* [N] */
  const double* bp_xy;       /* [N][bp_max][2] */
  const uint8_t* comb;       /* [ns*ns] (comb_shared != 0) or [B][ns*ns]: order of the jerk samples, jx*ns+jy */
  int32_t comb_shared;
  /* outputs */
  int32_t* status;           /* [B] 0 runtime reached, 1 goal reached, 2 open list empty (:1637-1639) */
  int32_t* solved;           /* [B] return value of run() */
  int32_t* n_int;            /* [B] pieces of pwp_out_ */
  double* coeff;             /* [B][3][8][4] pwp_out_ */
  nb_ent_state esv;          /* entStateVec: [B][9] states, stride ent_cap */
  int32_t* stats;            /* [B][4] nodes used, pops, index of the best node, goal_occupied */
  double* cost;              /* [B] g of the best node */
} nb_search_args;

int nb_search_batch(nb_handle* h, const nb_search_args* args, void* stream);

/* Measurement hook (no reference counterpart): with nb_set_profiling on, SM cycles thread 0 of every search CTA
 * spent per phase in the last nb_search_batch, out [B][16]: [0] children, [1] sequential resolve, [2] pool copy,
 * [3] open-list pop, [4] collision tests, [5] endpoint tests, [6] set-up; [8..14] child 0 only: primitive, state copy,
 * entanglement chain, key + lookup, step geometry, crossing tests, list automaton. */
int nb_search_phase_cycles(nb_handle* h, long long* out, int B);

/*
 * ---- The replan cycle of one rank, resident on the device --------------------------------------------------------------
 *
 * Host side of the hot part of Neptune::replanFull for the B agents a rank plans (neptune.cpp:1430-1448 hulls, samples,
 * PredictAlphasBetas; :1450-1510 front end; :1512-1529 back end; :1641-1647 safetyCheckAfterReplan; :1685-1699 compose) and
 * of the message exchange around it (publishOwnTraj / trajCB, neptune_ros.cpp:379-480), as one stream-ordered launch
 * sequence owned by the library: side streams for the independent stages, the whole sequence captured in CUDA graphs,
 * committed-trajectory records kept in a three-slot ring on the device (slot k % 3 receives the records committed in
 * cycle k; cycle k plans against the records of cycle k - 2 and post-checks against those of cycle k - 1, i.e. every
 * other agent's newest trajectory arrives while this agent optimises).  With world > 1 the commit kernel stores every
 * record into the ring of EVERY rank over NVLink (peer memory opened from CUDA IPC handles) and raises a flag there; the
 * cycle ends when the flags of all peers for this cycle are up -- the all-gather of the reference's /trajs topic fused
 * into the commit, no collective call and no host in the loop.
 *
 * Per-cycle inputs travel in ONE packed pinned host buffer -> ONE device buffer (nb_cycle_upload), results likewise
 * (nb_cycle_download); nb_cycle_layout tells where each array lives.  Records never travel through the host once seeded.
 */
typedef struct nb_cycle nb_cycle;

typedef struct nb_cycle_desc
{
  int32_t B;                 /* agents planned by this rank */
  const int32_t* agent_id;   /* [B] 1-based ids (ctor argument id of each agent's solver) */
  int32_t front_end;         /* 1: KinodynamicSearch::run inside the cycle (nb_search_configure must have been called);
                                2: Neptune::replanKinodynamic (neptune.cpp:1010-1300) -- the search, then post-check and
                                commit of the FRONT-END path itself, no LPs / QP (status 0, coeff_out = the search's path) */
  int32_t rank, world;       /* position of this rank in the exchange (world == 1: no peers) */
  const uint8_t* planned;    /* [num_agents] 1 = some rank plans this agent (its ring slot is written by a commit every
                                cycle); 0 = nobody does (its record is carried forward).  NULL: all planned by this rank
                                when world == 1 */
  double bbox;               /* DynTraj bbox edge of every agent (2 drone_radius, neptune_ros.cpp:447-449) */
  double delta;              /* hull inflation bbox / 2 + drone_radius (neptune.cpp:340) */
} nb_cycle_desc;

/* byte offsets of the arrays inside the packed buffers (-1: absent) */
typedef struct nb_cycle_layout
{
  /* inputs */
  int64_t n_int, coeff_init, t_start, t_now, t_group, group, known, late, esv_cnt, esv_alpha, esv_active;
  int64_t es_cnt, es_alpha, es_beta, es_bend, es_active, prev_pos, prev_pos_agent, cur;
  int64_t fe_init, fe_goal, fe_coeffs_z, fe_comb;
  int64_t in_bytes;
  /* outputs */
  int64_t coeff_out, obj, status, iters, entangled, collide, n_pieces, fe_status, fe_solved, fe_n_int, fe_stats;
  int64_t out_bytes;
} nb_cycle_layout;

int nb_cycle_create(nb_handle* h, const nb_cycle_desc* d, nb_cycle** out);
void nb_cycle_destroy(nb_cycle* c);
int nb_cycle_layout_get(const nb_cycle* c, nb_cycle_layout* out);
void* nb_cycle_host_in(nb_cycle* c);    /* pinned, in_bytes */
void* nb_cycle_host_out(nb_cycle* c);   /* pinned, out_bytes */
/* records the agents know when cycle k starts and those that arrive during it: host [num_agents][NB_REC_DOUBLES] each */
int nb_cycle_seed_records(nb_cycle* c, const double* known_recs, const double* late_recs, void* stream);
int nb_cycle_upload(nb_cycle* c, int32_t n_groups, void* stream);    /* packed inputs host -> device */
int nb_cycle_step(nb_cycle* c, void* stream);                         /* one cycle (graph replay once captured) */
int nb_cycle_download(nb_cycle* c, void* stream);                     /* packed results device -> host */
/* the same with caller-owned PINNED buffers of in_bytes / out_bytes (several prepared input sets, nb_pinned_alloc) */
int nb_cycle_upload_from(nb_cycle* c, const void* host_in, int32_t n_groups, void* stream);
int nb_cycle_download_to(nb_cycle* c, void* host_out, void* stream);
/* bytes the last upload moved over PCIe: in_bytes for small worlds; for large ones (packed buffer >= 4 MB) only the used
 * prefix of every entanglement-list row travels and the active-case arrays are rebuilt on the device from the lists */
long long nb_cycle_last_upload_bytes(const nb_cycle* c);
void* nb_pinned_alloc(int64_t bytes);   /* page-locked host memory, zeroed; NULL on failure */
void nb_pinned_free(void* p);
/* runs ONE real cycle (k advances), then captures the launch sequence for the current grouping in three CUDA graphs (one
 * per ring phase); needs a non-default stream.  With world > 1 every rank must call it at the same cycle. */
int nb_cycle_capture(nb_cycle* c, void* stream);
/* one cycle on a single stream with an event after every stage; ms [8]: late hulls, hulls + samples, predict, front end,
 * lines + QP, post-check, commit, exchange wait */
int nb_cycle_step_profiled(nb_cycle* c, void* stream, double* ms);
long long nb_cycle_launches_per_step(const nb_cycle* c);             /* kernels one step launches */
long long nb_cycle_index(const nb_cycle* c);                          /* k: cycles completed */
/* copy a device array of the cycle to the host (tests, diagnostics): "ring_new", "ring_late", "ring_known" (records,
 * [num_agents][NB_REC_DOUBLES], as of the LAST completed cycle for ring_new), "esA_cnt", "esA_alpha", "esA_beta",
 * "esA_bend", "esA_active", "hull_cnt", "samp", "bp_cnt", "bp_xy", "latest_pos" */
int nb_cycle_fetch(nb_cycle* c, const char* name, void* dst, int64_t bytes, void* stream);
/* exchange set-up for world > 1: every rank exports one IPC handle (64 bytes) of its ring; after an all-gather of the
 * handles (the caller's job: torch.distributed, MPI ...) every rank opens its peers' rings */
int nb_cycle_ipc_handle(nb_cycle* c, void* handle64);
/* Measurement helper (no reference counterpart): a rank barrier ON THE DEVICE, enqueued on `stream` -- every rank's stream
 * continues only when all ranks have reached this point (flags in the same peer memory as the commit).  bench.py places
 * one before each timed cycle so that the untimed input upload of one rank is not counted as exchange wait by another.
 * Every rank must call it the same number of times.  No-op when world == 1. */
int nb_cycle_align(nb_cycle* c, void* stream);
int nb_cycle_open_peers(nb_cycle* c, const void* handles /* [world][64] */);

/* Measurement hook (no reference counterpart): with nb_set_profiling on, SM cycles lane 0 of every QP warp spent per
 * phase in the last nb_replan_batch, out [B][16]: [0] set-up, [1] residual sweep, [2] dual residual + stopping test,
 * [3] normal-matrix assembly, [4] factorisation, [5] predictor solve, [6] predictor sweep, [7] corrector solve,
 * [8] final sweep. */
int nb_qp_phase_cycles(nb_handle* h, long long* out, int B);

#ifdef __cplusplus
}
#endif
#endif /* NEPTUNE_B200_H */

/*
 * neptune_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the algorithm on NEPTUNE's per-agent replan hot path
 * (SURVEY.md section 8a).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (neptune_b200/) never links, imports or executes anything in oracle/.
 *
 * PARITY STATUS, by part:
 *  - front-end search (neptune_search.c): PINNED against the reference's own
 *    kinodynamic_search.cpp, whole runs compared field by field; trajectory
 *    sampling and composePieceWisePol: PINNED against kinodynamic_search.cpp /
 *    utils.cpp.
 *  - entanglement chain (crossing tests with 8 and 9 arguments and for static
 *    obstacles, addAlphaBetaToList, updateBendPts, getLengthToContactPoints,
 *    the per-interval loop around them) and gjk::collision: PINNED against the
 *    reference's own entangle_utils.cpp and gjk.cpp, compiled where they lie
 *    (oracle/Makefile target _ref, Eigen replaced by oracle/eigen_shim) --
 *    tests/test_reference_pin.py, recorded vectors tests/golden/reference/ref_chain.npz.
 *  - separating-line LP: model and solved flag PINNED against the reference's
 *    own separator_glpk.cpp run with HiGHS as its LP engine (GLPK is absent;
 *    oracle/ref_stubs/glpk.h records the model); the vertex GLPK would return
 *    for the zero objective is "parity unpinned" (oracle and product return the
 *    minimum-norm line).
 *  - trajectory QP, hull order: "parity unpinned" against a recorded
 *    Gurobi / CGAL run -- the reference ships no golden vectors and none of
 *    Gurobi 9.1.2, CGAL 4.14.2, GLPK 4.65, Eigen3 or ROS exist in this image
 *    (see DESIGN.md).  These are checked instead by HiGHS (scipy) as an
 *    independent QP solver on every golden scene and by analytic cases
 *    (tests/test_oracle.py).
 *
 * Every function cites the reference file:line it restates (paths relative
 * to /root/reference).
 */
#ifndef NEPTUNE_ORACLE_H
#define NEPTUNE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NPOL_MAX 8
#define ORC_HMAX 64 /* max hull vertices the oracle handles per polygon */

/* status path of PolySolverGurobi::optimize (solver_gurobi_poly.cpp:832-861) */
#define ORC_STATUS_OK 0       /* first solve accepted */
#define ORC_STATUS_FALLBACK 1 /* terminal v/a rows dropped, soft cost, re-solve accepted */
#define ORC_STATUS_FAILED 2   /* both solves failed: pwp_out = pwp_init */

typedef struct orc_params
{
  int num_pol;      /* par_.num_pol */
  int num_agents;   /* N = pb.size() */
  int num_static;   /* M */
  int samples;      /* num_sample_per_interval */
  double T_span;
  double weight;    /* weight_term */
  double lim_min[3], lim_max[3]; /* x,y,z box (setMaxValues) */
  double v_max, a_max;
  double drone_radius;
  int ent_cap;      /* storage capacity of alphas lists (entries) */
  int bp_max;       /* storage capacity of a bend-point list */
  int ent_slots;    /* LP slots reserved per interval for entangle constraints */
  int ipm_max_iter;
  double ipm_tol;
} orc_params;

/* ---- constants: solver_gurobi_poly.cpp:25-134, mader_types.hpp:152-162 ---- */
void orc_basis(double T, double Ainv[16], double V[9], double Ainv01[16]);

/* ---- separator: separator_glpk.cpp:248-373 / :375-498 (canonical vertex) ---- */
int orc_separate(const double* A, int nA, const double* B, int nB, double out[3]);
/* generic LP feasibility of the same rows (phase-1 simplex); dim = 2 or 3 */
int orc_lp_separable(const double* A, int nA, const double* B, int nB, int dim);

/* ---- hulls / samples: neptune.cpp:224-452, :463-566; cgal_utils.cpp:157-174 ---- */
int orc_convex_hull_2d(const double* pts, int n, double* out);
void orc_hull_of_interval(const double* times, int nt, const double* cx, const double* cy, double t_start,
                          double t_end, double T_span, const double delta[3], double* hull, int* hull_n,
                          double* hull2, int* hull2_n, int idx[2]);
void orc_sample_interval_points(const double* times, int nt, const double* cx, const double* cy,
                                double t_start, double t_end, int num_pol, int S, double* out, int* idx_out);

/* ---- GJK: gjk.cpp:76-148 ---- */
int orc_gjk_collision(const double* v1, int n1, const double* v2, int n2);

/* Neptune::trajsAndPwpAreInCollision2d neptune.cpp:767-806 */
int orc_pwp_collides(const double* coeff /*[3][8][4]*/, int n, double t_start, double T_span, const double* times,
                     int nt, const double* cx, const double* cy, const double delta[3]);

/* ---- entanglement chain: entangle_utils.cpp:1129-1722 ---- */
typedef struct orc_ent
{
  int n_alpha, n_bend;
  int* alpha;   /* [cap][2] */
  double* beta; /* [cap]    */
  int* bend;    /* [cap]    */
  int* active;  /* [N+M]    */
} orc_ent;

typedef struct orc_ectx
{
  int N, M, self;      /* self = 0-based index of the planning agent */
  int cap;             /* storage capacity of lists */
  const double* pb;    /* [N][2] */
  const double* strep; /* [M][2 cols][2] staticObsRep */
  const int* bp_cnt;   /* [N] */
  const double* bp_xy; /* [N][bp_max][2] */
  int bp_max;
} orc_ectx;

int orc_hsig_agent(int* toadd, int nadd, const double pk[2], const double pk1[2], const double pik[2],
                   const double pik1[2], const double pb_self[2], const double* bend, int nbend, int agent_id);
int orc_hsig_static(int* toadd, int nadd, const double pk[2], const double pk1[2], const double* strep,
                    int M, int N);
int orc_add_alpha_beta(int* toadd, int nadd, orc_ent* es, const double pk[2], const orc_ectx* cx);
void orc_update_bend_pts(orc_ent* es, const double pk1[2], const orc_ectx* cx);

/* eu::entangleHSigToAddAgentInd, 9-arg (entangle_utils.cpp:820-1127) */
int orc_hsig_agent9(int* toadd, int nadd, const double pk[2], const double pk1[2], const double pik[2],
                    const double pik1[2], const double pb_self[2], const double* bend, int nbend, const double* prev,
                    int nprev, int agent_id, int* stop);
/* NeptuneRos::updateEntStateStaticObs neptune_ros.cpp:798-850 (one tick of the online tracker) */
int orc_track(orc_ent* es, const orc_ectx* cx, const int* bp_cnt_prev, const double* bp_xy_prev, double* prev_pos,
              double* prev_pos_agent, const double* latest, const double cur[2], double elapsed_ms);
/* Neptune::PredictAlphasBetas neptune.cpp:976-1008 */
int orc_predict(orc_ent* es, const orc_ectx* cx, const double* prev_pos /*[N+1][2]*/,
                const double* prev_pos_agent /*[N][2]*/, const double cur[2],
                const double* samp0 /*[N][2] first sample of each agent*/, const unsigned char* known);
/* KinodynamicSearch::entangleCheckGivenPwp kinodynamic_search.cpp:897-985 */
int orc_entangle_check_pwp(orc_ent* es, const orc_ectx* cx, int n, const double* cxy /*[2][n][4]*/,
                           const double* samp /*[N][num_pol][S+1][2]*/, const unsigned char* known,
                           int num_pol, int S, double T);
/* per-interval chain of KinodynamicSearch::entanglesWithOtherAgents :707-895
 * (tether-length test omitted); writes the state after every interval. */
int orc_entangle_rollout(const orc_ent* es0, const orc_ectx* cx, int n, const double* cxy,
                         const double* samp, const unsigned char* known, int num_pol, int S, double T,
                         int* out_cnt /*[n+1][2]*/, int* out_alpha /*[n+1][cap][2]*/,
                         double* out_beta, int* out_bend, int* out_active /*[n+1][N+M]*/);

/* ---- back end: solver_gurobi_poly.cpp:187-936 ---- */
typedef struct orc_replan_in
{
  int agent_id;                 /* 1-based */
  int n;                        /* intervals of pwp_init */
  const double* coeff_init;     /* [3][8][4] */
  int n_hull_slots;             /* slots per interval for other agents' hulls */
  const long long* hull_ptr;    /* [slots*8+1] vertex offsets, slot-major then interval */
  const double* hull_xy;
  const double* nih0;           /* [N][8][2] col(0) of hullsNoInflation_ (NaN = empty) */
  const long long* st_ptr;      /* [M+1] */
  const double* st_xy;
  const int* esv_cnt;           /* [9][2] */
  const int* esv_alpha;         /* [9][cap][2] */
  const int* esv_active;        /* [9][N+M] */
  const int* bp_cnt;            /* [N] */
  const double* bp_xy;          /* [N][bp_max][2] */
  const double* pb;             /* [N][2] */
} orc_replan_in;

typedef struct orc_replan_out
{
  double* coeff_out; /* [3][8][4] */
  double* obj;       /* [1] */
  int* status;       /* [1] */
  int* iters;        /* [2] */
  double* lines;     /* [8][LS][3]  LS = n_hull_slots + N + M + ent_slots */
  unsigned char* line_ok; /* [8][LS] 0 = not attempted, 1 = solved, 2 = attempted, unsolved */
} orc_replan_out;

int orc_replan(const orc_params* par, const orc_replan_in* in, orc_replan_out* out);

/* dense export of the full-space QP the reference hands to Gurobi (for HiGHS cross-checks).
 * Returns number of inequality rows written; P is 12n x 12n, rows are dense 12n. */
int orc_export_qp(const orc_params* par, const orc_replan_in* in, int fallback, const double* lines,
                  const unsigned char* line_ok, int LS, double* P, double* q, double* c0, double* Aeq,
                  double* beq, int* n_eq, double* G, double* h, int max_rows, int* has_qc);

/* PolySolverGurobi::generatePwpOut :889-936 ; returns number of states written */
int orc_generate_traj(const double* coeff /*[3][8][4]*/, int n, double T, double dc, double* states, int max_states);

/* batch driver used as CPU baseline: arrays are the same SoA buffers the C-ABI takes. */
typedef struct orc_batch
{
  int B;
  const int* agent_id;        /* [B] */
  const int* n_int;           /* [B] */
  const double* coeff_init;   /* [B][3][8][4] */
  int n_hull_slots;
  const long long* hull_ptr;  /* [B*slots*8+1] */
  const double* hull_xy;
  const double* nih0;         /* [B][N][8][2] */
  const long long* st_ptr;
  const double* st_xy;
  const int* esv_cnt;         /* [B][9][2] */
  const int* esv_alpha;       /* [B][9][cap][2] */
  const int* esv_active;      /* [B][9][N+M] */
  int bp_shared;              /* 1: bp arrays are [N]..., 0: [B][N]... */
  const int* bp_cnt;
  const double* bp_xy;
  const double* pb;
  double* coeff_out;          /* [B][3][8][4] */
  double* obj;                /* [B] */
  int* status;                /* [B] */
  int* iters;                 /* [B][2] */
  double* lines;              /* [B][8][LS][3] or NULL */
  unsigned char* line_ok;     /* [B][8][LS] or NULL */
} orc_batch;

int orc_replan_batch(const orc_params* par, const orc_batch* b, int nthreads);

/* mu::composePieceWisePol utils.cpp:318-402 on 210-double records; p1, p2 modified like the reference's arguments */
int orc_compose_records(double t, double dc, double* p1, double* p2, double* out);

/* tail of replanFull (neptune.cpp:1685-1699): pwp_now + composePieceWisePol with the previous record */
int orc_commit_compose_batch(const orc_params* par, int B, const int* agent_id, const int* n_int, const double* coeff_out,
                             const double* t_start, const double* t_now, const double* recs, const int* status,
                             const int* entangled, const int* collide, double* new_recs, int* n_pieces);

/* whole cycle per agent (hulls/samples, predict, back end, post-check): the CPU baseline of bench.py */
int orc_cycle_batch(const orc_params* par, int B, const int* agent_id, const int* n_int, const double* coeff_init,
                    const double* t_start, const double* recs, const unsigned char* known, const double* pb,
                    const long long* st_ptr, const double* st_xy, const double* strep, const int* bp_cnt,
                    const double* bp_xy, const int* esv_cnt, const int* esv_alpha, const int* esv_active,
                    const int* es_cnt, const int* es_alpha, const double* es_beta, const int* es_bend,
                    const int* es_active, const double* prev_pos, const double* prev_pos_agent, const double* cur,
                    double delta, int do_entangle, double* coeff_out, double* obj, int* status, int* iters,
                    int* entangled, int* collide, int nthreads, const unsigned char* late, const double* late_recs,
                    const int* bp_cnt_late, const double* bp_xy_late);

/* ---- front end: KinodynamicSearch (kinodynamic_search.cpp), neptune_search.c ---- */
#define ORC_SEARCH_HSTRIDE 24 /* vertices reserved per hull in the fixed-stride hull arrays */

typedef struct orc_search_par
{
  int num_pol, N, M, S;
  double T;
  double x_min, x_max, y_min, y_max;
  double v_max, a_max, j_max;
  int num_samples;     /* a_star_samp_x: jerk samples per axis */
  double voxel_size;   /* a_star_fraction_voxel_size */
  double bias;         /* setBias(1.1), neptune.cpp:97 */
  double goal_size;    /* goal_radius */
  double tether;       /* tetherLength */
  int enable_entangle; /* enable_entangle_check */
  int use_not_reaching;/* use_not_reaching_soln */
  int max_nodes;       /* node_num_max_ */
  int max_expansions;  /* stands in for max_runtime_ */
  int ecap;            /* storage capacity of a node's alphas list */
  int out_cap;         /* stride of the esv_* outputs (= the back end's ent_cap) */
  int bp_max;
} orc_search_par;

typedef struct orc_search_in
{
  int agent_id;            /* 1-based */
  double init[6];          /* px py vx vy ax ay */
  double goal[2];
  const double* coeffs_z;  /* [8][4] getInitialZPwp */
  const double* hull_xy;   /* [N][8][ORC_SEARCH_HSTRIDE][2] inflated hulls per window */
  const int* hull_cnt;     /* [N][8], 0 = no hull (unknown agent or self) */
  const double* samp;      /* [N][num_pol][S+1][2] */
  const unsigned char* known; /* [N] */
  const long long* st_ptr; /* [M+1] inflated static obstacles */
  const double* st_xy;
  const double* strep;     /* [M][2][2] */
  const double* st_longest;/* [M][2] staticObsLongestDist */
  const double* pb;        /* [N][2] */
  const int* bp_cnt;       /* [N] */
  const double* bp_xy;     /* [N][bp_max][2] */
  const int* es_cnt;       /* [2] entangle_state_A */
  const int* es_alpha;
  const double* es_beta;
  const int* es_bend;
  const int* es_active;    /* [N+M] */
  const unsigned char* comb; /* [num_samples^2] order of the jerk samples, value = jx*num_samples+jy */
} orc_search_in;

typedef struct orc_search_out
{
  int* status;     /* 0 runtime reached, 1 goal reached, 2 open list empty (kinodynamic_search.cpp:1637-1639) */
  int* solved;     /* return value of run() */
  int* n_int;      /* pieces of pwp_out_ */
  double* coeff;   /* [3][8][4] */
  int* esv_cnt;    /* [9][2] entStateVec */
  int* esv_alpha;  /* [9][out_cap][2] */
  double* esv_beta;/* [9][out_cap] */
  int* esv_bend;   /* [9][out_cap] */
  int* esv_active; /* [9][N+M] */
  int* stats;      /* [4] nodes used, pops, index of the best node, goal_occupied */
  double* cost;    /* g of the best node */
} orc_search_out;

int orc_search(const orc_search_par* par, const orc_search_in* in, orc_search_out* out);

typedef struct orc_search_batch_t
{
  int B;
  const int* agent_id;       /* [B] */
  const double* init;        /* [B][6] */
  const double* goal;        /* [B][2] */
  const double* coeffs_z;    /* [B][8][4] */
  const int* group;          /* [B] window group of each agent, NULL: group = b */
  const double* hull_xy;     /* [G][N][8][24][2] */
  const int* hull_cnt;       /* [G][N][8] */
  const double* samp;        /* [G][N][num_pol][S+1][2] */
  const unsigned char* known;/* [B][N] */
  const long long* st_ptr;
  const double* st_xy;
  const double* strep;
  const double* st_longest;
  const double* pb;
  const int* bp_cnt;
  const double* bp_xy;
  int es_cap;                /* stride of es_alpha / es_beta / es_bend */
  const int* es_cnt;         /* [B][2] */
  const int* es_alpha;       /* [B][es_cap][2] */
  const double* es_beta;
  const int* es_bend;
  const int* es_active;      /* [B][N+M] */
  int comb_shared;
  const unsigned char* comb; /* [num_samples^2] or [B][num_samples^2] */
  int* status;
  int* solved;
  int* n_int;
  double* coeff;             /* [B][3][8][4] */
  int* esv_cnt;              /* [B][9][2] */
  int* esv_alpha;            /* [B][9][out_cap][2] */
  double* esv_beta;
  int* esv_bend;
  int* esv_active;           /* [B][9][N+M] */
  int* stats;                /* [B][4] */
  double* cost;              /* [B] */
} orc_search_batch_t;

int orc_search_batch(const orc_search_par* par, const orc_search_batch_t* b, int nthreads);
/* NeptuneRos::setUpCheckingPosAndStaticObs neptune_ros.cpp:852-1019 (static-obstacle representation of one agent) */
int orc_static_obst_rep(int M, const long long* ptr, const double* xy, const double base[2], const double pos[2],
                        double voxel, double* strep, double* longest);
/* eu::getTetherLength (entangle_utils.cpp:1724-1743) on an explicit state (test hook) */
double orc_tether_length_state(const orc_ent* es, const orc_ectx* cx, const double* st_longest, const double pk1[2]);
/* test hook: open-list script through the heap restatement (compared with the real std::priority_queue) */
int orc_heap_replay(int n_ops, const int* ops, const double* vals, int n_ids, double bias, int* out);

#ifdef __cplusplus
}
#endif
#endif

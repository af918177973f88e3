// ref_wrap_search.cpp -- a C entry point around the REFERENCE's own KinodynamicSearch (TEST INFRASTRUCTURE).
//
// oracle/Makefile (target _ref) compiles neptune/src/kinodynamic_search.cpp where it lies under /root/reference,
// unmodified, against the Eigen stand-in (oracle/eigen_shim) and the header stand-ins in oracle/ref_stubs (ROS clock and
// message structs, bspline_utils.hpp, exprtk.hpp, glpk.h: nothing of theirs is used by the search), together with this file.
// ref_search() drives the class exactly as Neptune does (neptune.cpp:89-97 construction, :662 / :670 static obstacles,
// :1310 clearProcess, :1419-1452 setRunTime / setInitZCoeffs / setUp / run, :1509-1510 getters) on the oracle's own
// plain-array inputs and writes the oracle's output layout, so tests/test_reference_pin.py can compare field by field.
//
// The reference shuffles the jerk samples with a wall-clock seed inside run() (kinodynamic_search.cpp:1462-1463); the
// order it drew is read back from the object afterwards (comb_out) and handed to the oracle as its `comb` input.  The
// wall-clock budget is set far above what the test cases need, so the search ends by reaching the goal, by emptying the
// open list or by exhausting the node pool (node_num_max_, returned in info[0], = the oracle's max_nodes).
#include <queue>
#include <sstream>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>
#include <Eigen/Dense>
#include "mader_types.hpp"

// access to all_combinations_, node_used_num_, goal_occupied_ and best_node_ptr_ (read only); the one member written from
// here are num_of_static_obst_ in ref_entangle_check_pwp (setStaticObstVert would set it together with a node pool) and
// pwp_out_ in ref_generate_traj (recoverPwpOut would, after a search)
#define private public
#include "kinodynamic_search.hpp"
#undef private

#include "neptune_oracle.h"

typedef Eigen::Vector2d V2;

static mt::Polygon_Std polygon(const double* xy, int n)
{
  mt::Polygon_Std p(2, n);
  for (int i = 0; i < n; i++) p(0, i) = xy[2 * i], p(1, i) = xy[2 * i + 1];
  return p;
}

// info[0] node_num_max_, info[1] node_used_num_, info[2] goal_occupied_
extern "C" int ref_search(const orc_search_par* par, const orc_search_in* in, orc_search_out* out, unsigned char* comb_out, int* info)
{
  const int N = par->N, M = par->M, NA = N + M, S = par->S, np = par->num_pol, cap = par->out_cap;
  std::vector<V2> pb;
  for (int i = 0; i < N; i++) pb.push_back(V2(in->pb[2 * i], in->pb[2 * i + 1]));

  std::streambuf* keep = std::cout.rdbuf();  // the reference prints progress lines; keep the test output readable
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());

  KinodynamicSearch ks(np, 3, in->agent_id, 1.0, par->T, S, pb, par->use_not_reaching != 0, par->enable_entangle != 0);
  ks.setTetherLength(par->tether);
  ks.setMaxValuesAndSamples(par->v_max, par->a_max, par->j_max, par->num_samples);
  ks.setXYZMinMaxAndRa(par->x_min, par->x_max, par->y_min, par->y_max, -10.0, 10.0, 5.0, par->voxel_size);
  ks.setBias(par->bias);
  ks.setGoalSize(par->goal_size);

  std::vector<mt::Polygon_Std> statics;
  std::vector<Eigen::Matrix<double, 2, 2>> rep;
  std::vector<V2> longest;
  for (int m = 0; m < M; m++)
  {
    statics.push_back(polygon(in->st_xy + 2 * in->st_ptr[m], (int)(in->st_ptr[m + 1] - in->st_ptr[m])));
    Eigen::Matrix<double, 2, 2> r;  // column c = representative point c (neptune_ros.cpp:961-983)
    r(0, 0) = in->strep[4 * m + 0], r(1, 0) = in->strep[4 * m + 1], r(0, 1) = in->strep[4 * m + 2], r(1, 1) = in->strep[4 * m + 3];
    rep.push_back(r);
    longest.push_back(V2(in->st_longest[2 * m], in->st_longest[2 * m + 1]));
  }
  ks.setStaticObstVert(statics);
  ks.setStaticObstRep(rep, longest);
  info[0] = ks.node_num_max_;

  ks.clearProcess();
  ks.setRunTime(1.0e6);
  std::vector<Eigen::Matrix<double, 4, 1>> cz;
  for (int i = 0; i < ORC_NPOL_MAX; i++)
    cz.push_back(Eigen::Matrix<double, 4, 1>(in->coeffs_z[4 * i], in->coeffs_z[4 * i + 1], in->coeffs_z[4 * i + 2], in->coeffs_z[4 * i + 3]));
  ks.setInitZCoeffs(cz);

  mt::state A;
  A.setPos(in->init[0], in->init[1], 0.0);
  A.setVel(in->init[2], in->init[3], 0.0);
  A.setAccel(in->init[4], in->init[5], 0.0);
  Eigen::Vector3d goal(in->goal[0], in->goal[1], 0.0);

  mt::ConvexHullsOfCurves_Std2d hulls;  // one entry per KNOWN trajectory (neptune.cpp:1436-1441), intervals 0..num_pol-1
  for (int o = 0; o < N; o++)
  {
    if (in->hull_cnt[o * ORC_NPOL_MAX] <= 0) continue;
    mt::ConvexHullsOfCurve_Std2d one;
    for (int i = 0; i < np; i++)
      one.push_back(polygon(in->hull_xy + ((size_t)(o * ORC_NPOL_MAX + i) * ORC_SEARCH_HSTRIDE) * 2, in->hull_cnt[o * ORC_NPOL_MAX + i]));
    hulls.push_back(one);
  }
  mt::SampledPointsofCurves spoc(N);  // indexed by agent id - 1, empty = unknown
  for (int a = 0; a < N; a++)
  {
    if (!in->known[a]) continue;
    for (int i = 0; i < np; i++) spoc[a].push_back(polygon(in->samp + ((size_t)(a * np + i) * (S + 1)) * 2, S + 1));
  }
  eu::ent_state es;
  for (int i = 0; i < in->es_cnt[0]; i++)
    es.alphas.push_back(Eigen::Vector2i(in->es_alpha[2 * i], in->es_alpha[2 * i + 1])), es.betas.push_back(in->es_beta[i]);
  for (int i = 0; i < in->es_cnt[1]; i++) es.bendPointsIdx.push_back(in->es_bend[i]);
  for (int i = 0; i < NA; i++) es.active_cases.push_back(in->es_active[i]);
  std::vector<std::vector<V2>> bends(N);
  for (int a = 0; a < N; a++)
    for (int i = 0; i < in->bp_cnt[a]; i++)
      bends[a].push_back(V2(in->bp_xy[((size_t)a * par->bp_max + i) * 2], in->bp_xy[((size_t)a * par->bp_max + i) * 2 + 1]));

  ks.setUp(A, goal, hulls, spoc, es, bends);
  std::vector<Eigen::Vector3d> path;
  int status = -1;
  const bool solved = ks.run(path, status);
  std::cout.rdbuf(keep);

  for (size_t i = 0; i < ks.all_combinations_.size(); i++)
    comb_out[i] = (unsigned char)(std::get<0>(ks.all_combinations_[i]) * par->num_samples + std::get<1>(ks.all_combinations_[i]));
  info[1] = ks.node_used_num_;
  info[2] = ks.goal_occupied_ ? 1 : 0;

  out->status[0] = status;
  out->solved[0] = solved ? 1 : 0;
  out->n_int[0] = 0;
  if (solved)
  {
    mt::PieceWisePol pwp;
    ks.getPwpOut_0tstart(pwp);
    std::vector<eu::ent_state> esv;
    ks.getEntStateVector(esv);
    const int n = (int)pwp.coeff_x.size();
    out->n_int[0] = n;
    for (int i = 0; i < n && i < ORC_NPOL_MAX; i++)
      for (int k = 0; k < 4; k++)
      {
        out->coeff[(0 * ORC_NPOL_MAX + i) * 4 + k] = pwp.coeff_x[i](k);
        out->coeff[(1 * ORC_NPOL_MAX + i) * 4 + k] = pwp.coeff_y[i](k);
        out->coeff[(2 * ORC_NPOL_MAX + i) * 4 + k] = pwp.coeff_z[i](k);
      }
    info[3] = (int)esv.size();
    for (size_t s = 0; s < esv.size() && s <= ORC_NPOL_MAX; s++)
    {
      const eu::ent_state& e = esv[s];
      out->esv_cnt[2 * s] = (int)e.alphas.size(), out->esv_cnt[2 * s + 1] = (int)e.bendPointsIdx.size();
      for (size_t i = 0; i < e.alphas.size() && (int)i < cap; i++)
      {
        out->esv_alpha[(s * cap + i) * 2] = e.alphas[i](0), out->esv_alpha[(s * cap + i) * 2 + 1] = e.alphas[i](1);
        out->esv_beta[s * cap + i] = e.betas[i];
      }
      for (size_t i = 0; i < e.bendPointsIdx.size() && (int)i < cap; i++) out->esv_bend[s * cap + i] = e.bendPointsIdx[i];
      for (int i = 0; i < NA && i < (int)e.active_cases.size(); i++) out->esv_active[s * NA + i] = e.active_cases[i];
    }
    out->cost[0] = ks.best_node_ptr_ ? ks.best_node_ptr_->g : 0.0;  // getCost() is declared but never defined
  }
  return 0;
}

// The same slicing of a batch into per-agent inputs as orc_search_batch (oracle/neptune_search.c), one reference search
// per agent.  comb_out [B][num_samples^2], info [B][4].
extern "C" int ref_search_batch(const orc_search_par* par, const orc_search_batch_t* b, unsigned char* comb_out, int* info)
{
  const int N = par->N, M = par->M, NA = N + M, S = par->S, np = par->num_pol, ocap = par->out_cap;
  const int nchild = par->num_samples * par->num_samples;
  for (int i = 0; i < b->B; i++)
  {
    orc_search_in in;
    orc_search_out out;
    const int g = b->group ? b->group[i] : i;
    in.agent_id = b->agent_id[i];
    for (int k = 0; k < 6; k++) in.init[k] = b->init[6 * (size_t)i + k];
    in.goal[0] = b->goal[2 * i], in.goal[1] = b->goal[2 * i + 1];
    in.coeffs_z = b->coeffs_z + (size_t)i * ORC_NPOL_MAX * 4;
    std::vector<int> hc(b->hull_cnt + (size_t)g * N * ORC_NPOL_MAX, b->hull_cnt + (size_t)(g + 1) * N * ORC_NPOL_MAX);
    const unsigned char* known = b->known + (size_t)i * N;
    for (int o = 0; o < N; o++)
      if (o == in.agent_id - 1 || !known[o])
        for (int k = 0; k < ORC_NPOL_MAX; k++) hc[o * ORC_NPOL_MAX + k] = 0;
    in.hull_cnt = hc.data();
    in.hull_xy = b->hull_xy + (size_t)g * N * ORC_NPOL_MAX * ORC_SEARCH_HSTRIDE * 2;
    in.samp = b->samp + (size_t)g * N * np * (S + 1) * 2;
    in.known = known;
    in.st_ptr = b->st_ptr, in.st_xy = b->st_xy, in.strep = b->strep, in.st_longest = b->st_longest;
    in.pb = b->pb, in.bp_cnt = b->bp_cnt, in.bp_xy = b->bp_xy;
    in.es_cnt = b->es_cnt + 2 * (size_t)i;
    in.es_alpha = b->es_alpha + (size_t)i * 2 * b->es_cap;
    in.es_beta = b->es_beta + (size_t)i * b->es_cap;
    in.es_bend = b->es_bend + (size_t)i * b->es_cap;
    in.es_active = b->es_active + (size_t)i * NA;
    in.comb = 0;
    out.status = b->status + i, out.solved = b->solved + i, out.n_int = b->n_int + i;
    out.coeff = b->coeff + (size_t)i * 3 * ORC_NPOL_MAX * 4;
    out.esv_cnt = b->esv_cnt + (size_t)i * 9 * 2;
    out.esv_alpha = b->esv_alpha + (size_t)i * 9 * 2 * ocap;
    out.esv_beta = b->esv_beta + (size_t)i * 9 * ocap;
    out.esv_bend = b->esv_bend + (size_t)i * 9 * ocap;
    out.esv_active = b->esv_active + (size_t)i * 9 * NA;
    out.stats = b->stats + 4 * (size_t)i;
    out.cost = b->cost + i;
    ref_search(par, &in, &out, comb_out + (size_t)i * nchild, info + 4 * (size_t)i);
  }
  return 0;
}

// KinodynamicSearch::entangleCheckGivenPwp (kinodynamic_search.cpp:897-985) called on a real object: the post-check of
// an optimised trajectory.  State in / out as in ref_chain (oracle/ref_wrap.cpp); returns the reference's answer.
extern "C" int ref_entangle_check_pwp(int N, int M, int self, const double* pb, const double* strep, const int* bp_cnt,
                                      const double* bp_xy, int bp_max, const unsigned char* known, const double* samp, int num_pol,
                                      int S, double T, int n, const double* cxy /*[2][n][4]*/, int cap, int* cnt, int* alpha,
                                      double* beta, int* bend, int* active)
{
  std::vector<V2> vpb;
  for (int i = 0; i < N; i++) vpb.push_back(V2(pb[2 * i], pb[2 * i + 1]));
  std::streambuf* keep = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  KinodynamicSearch ks(num_pol, 3, self + 1, 1.0, T, S, vpb, true, true);
  std::vector<Eigen::Matrix<double, 2, 2>> rep;
  std::vector<V2> longest(M, V2(0.0, 0.0));
  for (int m = 0; m < M; m++)
  {
    Eigen::Matrix<double, 2, 2> r;
    r(0, 0) = strep[4 * m + 0], r(1, 0) = strep[4 * m + 1], r(0, 1) = strep[4 * m + 2], r(1, 1) = strep[4 * m + 3];
    rep.push_back(r);
  }
  ks.setStaticObstRep(rep, longest);
  ks.num_of_static_obst_ = M;  // normally set by setStaticObstVert (:367), which also allocates the node pool
  mt::SampledPointsofCurves spoc(N);
  for (int a = 0; a < N; a++)
  {
    if (!known[a]) continue;
    for (int i = 0; i < num_pol; i++) spoc[a].push_back(polygon(samp + ((size_t)(a * num_pol + i) * (S + 1)) * 2, S + 1));
  }
  std::vector<std::vector<V2>> bends(N);
  for (int a = 0; a < N; a++)
    for (int i = 0; i < bp_cnt[a]; i++) bends[a].push_back(V2(bp_xy[((size_t)a * bp_max + i) * 2], bp_xy[((size_t)a * bp_max + i) * 2 + 1]));
  eu::ent_state es;
  for (int i = 0; i < cnt[0]; i++) es.alphas.push_back(Eigen::Vector2i(alpha[2 * i], alpha[2 * i + 1])), es.betas.push_back(beta[i]);
  for (int i = 0; i < cnt[1]; i++) es.bendPointsIdx.push_back(bend[i]);
  for (int i = 0; i < N + M; i++) es.active_cases.push_back(active[i]);
  mt::state A;
  Eigen::Vector3d goal(0.0, 0.0, 0.0);
  mt::ConvexHullsOfCurves_Std2d no_hulls;
  eu::ent_state es_setup = es;
  ks.setUp(A, goal, no_hulls, spoc, es_setup, bends);

  mt::PieceWisePol pwp;
  for (int i = 0; i < n; i++)
  {
    pwp.coeff_x.push_back(Eigen::Matrix<double, 4, 1>(cxy[4 * i], cxy[4 * i + 1], cxy[4 * i + 2], cxy[4 * i + 3]));
    pwp.coeff_y.push_back(Eigen::Matrix<double, 4, 1>(cxy[4 * (n + i)], cxy[4 * (n + i) + 1], cxy[4 * (n + i) + 2], cxy[4 * (n + i) + 3]));
  }
  const bool ent = n > 0 ? ks.entangleCheckGivenPwp(pwp, es) : false;  // with no piece the reference falls off the end (no return value)
  std::cout.rdbuf(keep);
  cnt[0] = (int)es.alphas.size(), cnt[1] = (int)es.bendPointsIdx.size();
  for (size_t i = 0; i < es.alphas.size() && (int)i < cap; i++) alpha[2 * i] = es.alphas[i](0), alpha[2 * i + 1] = es.alphas[i](1), beta[i] = es.betas[i];
  for (size_t i = 0; i < es.bendPointsIdx.size() && (int)i < cap; i++) bend[i] = es.bendPointsIdx[i];
  for (int i = 0; i < N + M; i++) active[i] = es.active_cases[i];
  return ent ? 1 : 0;
}

// KinodynamicSearch::generatePwpOut (kinodynamic_search.cpp:621-668), the same code as PolySolverGurobi::generatePwpOut
// (solver_gurobi_poly.cpp:889-936, which needs Gurobi to compile): pwp_out_ is set here the way recoverPwpOut leaves it
// (n pieces, n + 1 knots at multiples of T), then the reference samples it every dc.  states [max_states][12].
extern "C" int ref_generate_traj(const double* coeff /*[3][8][4]*/, int n, double T, double dc, double t_start, double* states, int max_states,
                                 double* times_out /*[n+1]*/)
{
  std::vector<V2> vpb(1, V2(0.0, 0.0));
  std::streambuf* keep = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  KinodynamicSearch ks(ORC_NPOL_MAX, 3, 1, 1.0, T, 10, vpb, true, true);
  ks.pwp_out_.clear();
  for (int i = 0; i <= n; i++) ks.pwp_out_.times.push_back(i * T);
  for (int i = 0; i < n; i++)
  {
    const double *x = coeff + 4 * i, *y = coeff + 32 + 4 * i, *z = coeff + 64 + 4 * i;
    ks.pwp_out_.coeff_x.push_back(Eigen::Matrix<double, 4, 1>(x[0], x[1], x[2], x[3]));
    ks.pwp_out_.coeff_y.push_back(Eigen::Matrix<double, 4, 1>(y[0], y[1], y[2], y[3]));
    ks.pwp_out_.coeff_z.push_back(Eigen::Matrix<double, 4, 1>(z[0], z[1], z[2], z[3]));
  }
  mt::PieceWisePol pwp;
  std::vector<mt::state> traj;
  ks.generatePwpOut(pwp, traj, t_start, dc);
  std::cout.rdbuf(keep);
  for (size_t k = 0; k < pwp.times.size() && (int)k <= n; k++) times_out[k] = pwp.times[k];
  int cnt = 0;
  for (size_t k = 0; k < traj.size() && cnt < max_states; k++, cnt++)
    for (int a = 0; a < 3; a++)
    {
      states[12 * k + a] = traj[k].pos(a), states[12 * k + 3 + a] = traj[k].vel(a);
      states[12 * k + 6 + a] = traj[k].accel(a), states[12 * k + 9 + a] = traj[k].jerk(a);
    }
  return (int)traj.size();
}

// poly_solver_b200.hpp -- source-compatible replacement of the reference's back-end classes on top of
// the C-ABI of include/neptune_b200.h (B = 1 per call, host buffers).
//
//   class PolySolverGurobi       <- neptune/include/solver_gurobi_poly.hpp:27-52 (public section)
//   class separator::Separator   <- submodules/separator/include/separator.hpp:18-48 (2-D overloads)
//
// Same method names, argument meaning, call order (neptune.cpp:102-107 once; :1514-1527 per replan)
// and error behaviour: optimize() returns false and leaves pwp_out = pwp_init when both solves fail;
// solveModel() returns false for inseparable sets.  The header leaks no Gurobi types, so it is source-
// not ABI-compatible: relink neptune_test_node against libneptune_b200.so.
//
// Coexistence with the reference's tree: this header defines the include guard of neptune/include/solver_gurobi_poly.hpp
// and is meant to be seen first (`g++ -include poly_solver_b200.hpp`, or one edited #include in neptune.hpp), so the
// reference's own header -- and with it gurobi_c++.h -- drops out without touching a reference file; separator.hpp
// (#pragma once, found through the include path) is shadowed by neptune_b200/cpp/dropin/separator.hpp.
#pragma once
#ifndef SOLVER_GUROBI_POLY_HPP
#define SOLVER_GUROBI_POLY_HPP  // include guard of the reference's header: keeps it out of the translation unit
#endif
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/neptune_b200.h"
#include "nb_types.hpp"

namespace nb_detail
{
inline void check(int rc, const char* what)
{
  if (rc != NB_OK) throw std::runtime_error(std::string(what) + ": " + nb_last_error());
}
}  // namespace nb_detail

class PolySolverGurobi
{
public:
  PolySolverGurobi(int num_pol, int deg_pol, int id, double T_span, std::vector<Eigen::Vector2d> pb, double weight_term,
                   double rad_term, bool use_linear_constraints)
    : id_(id), pb_(pb)
  {
    par_ = nb_params();
    par_.num_pol = num_pol, par_.deg_pol = deg_pol, par_.num_agents = (int)pb.size(), par_.num_static = 0;
    par_.samples = 3, par_.use_linear_constraints = use_linear_constraints ? 1 : 0;
    par_.T_span = T_span, par_.weight = weight_term;
    // storage capacities follow the semantic bounds of the reference's vectors and grow on demand (optimize):
    // an alphas list holds at most 3 (N + M) entries (kinodynamic_search.cpp:938-942)
    par_.ent_cap = 3 * (int)pb.size() + 16, par_.bp_max = 8, par_.ent_slots = 16, par_.ipm_max_iter = 30, par_.ipm_tol = 1e-9;
    par_.drone_radius = 0.0, par_.tether_length = 0.0;
    (void)rad_term;  // rad_term_ is stored but unused by the reference in linear mode (:701-706)
  }
  ~PolySolverGurobi()
  {
    if (h_) nb_destroy(h_);
  }
  PolySolverGurobi(const PolySolverGurobi&) = delete;
  PolySolverGurobi& operator=(const PolySolverGurobi&) = delete;

  // Gurobi "TimeLimit" (:811-812).  The interior-point kernel is capped by iterations: an iteration costs about 20 us
  // on a B200, so the 60-iteration cap (1.2 ms) is inside every budget the shipped YAMLs set (15 - 80 ms); a smaller
  // budget lowers the cap proportionally.  Like a barrier solve that hits TimeLimit before its first incumbent
  // (SolCount == 0, :832-836), a solve that hits the cap counts as failed.
  void setMaxRuntime(double runtime)
  {
    max_runtime_ = runtime;
    const int cap = (int)(runtime / 20e-6);
    const int it = cap < 5 ? 5 : (cap > 30 ? 30 : cap);
    if (it != par_.ipm_max_iter) par_.ipm_max_iter = it, reset();
  }
  // stored and never read by the reference's linear mode either (solver_gurobi_poly.cpp:290-305, :786-801)
  void setBetasVector(std::vector<std::vector<Eigen::Vector3d>>& vecOfAgents) { vecOfAgentsBetas_ = vecOfAgents; }
  void setMaxValues(double x_min, double x_max, double y_min, double y_max, double z_min, double z_max, double v_max,
                    double a_max, double j_max)
  {
    par_.lim_min[0] = x_min, par_.lim_min[1] = y_min, par_.lim_min[2] = z_min;
    par_.lim_max[0] = x_max, par_.lim_max[1] = y_max, par_.lim_max[2] = z_max;
    par_.v_max = v_max, par_.a_max = a_max;
    (void)j_max;  // not a QP constraint in the reference either
    reset();
  }
  void setTetherLength(double tetherLength)
  {
    par_.tether_length = tetherLength;
  }
  void setStaticObstVert(std::vector<mt::Polygon_Std>& convexHullOfStaticObs)
  {
    st_ptr_.assign(1, 0);
    st_xy_.clear();
    for (auto& p : convexHullOfStaticObs)
    {
      for (int c = 0; c < p.cols(); c++) st_xy_.push_back(p(0, c)), st_xy_.push_back(p(1, c));
      st_ptr_.push_back((int64_t)st_xy_.size() / 2);
    }
    par_.num_static = (int)convexHullOfStaticObs.size();
    par_.ent_cap = 3 * (par_.num_agents + par_.num_static) + 16;
    reset();
  }
  void setInitTrajectory(mt::PieceWisePol pwp_init)
  {
    pwp_init_ = pwp_init;
    pwp_out_ = pwp_init;
  }
  void setHulls(mt::ConvexHullsOfCurves_Std2d& hulls) { hulls_ = hulls; }
  void setHullsNoInflation(mt::ConvexHullsOfCurves_Std2d& hulls) { hullsNoInflation_ = hulls; }
  void setEntStateVector(std::vector<eu::ent_state>& entStateVec, std::vector<std::vector<Eigen::Vector2d>>& bendPtsForAgents)
  {
    entStateVec_ = entStateVec;
    bendPtsForAgents_ = bendPtsForAgents;
  }

  // Failure behaviour of the reference: optimize() == false with pwp_out_ = pwp_init_, nothing thrown (:856-859; the
  // caller flies the front-end path, neptune.cpp:1519-1527).  Library errors (no device, out of memory) take the same
  // exit, with the text kept in lastError(); invalid input (n outside 1..8) is reported the same way.
  bool optimize(double& objective_value)
  {
    last_error_.clear();
    try
    {
      return optimizeImpl(objective_value);
    }
    catch (const std::exception& e)
    {
      last_error_ = e.what();
      std::fprintf(stderr, "PolySolverGurobi (neptune_b200): %s -- returning the initial trajectory\n", e.what());
      pwp_out_ = pwp_init_;
      last_status_ = NB_STATUS_FAILED;
      total_replannings_++;
      return false;
    }
  }
  const std::string& lastError() const { return last_error_; }

private:
  bool optimizeImpl(double& objective_value)
  {
    const int n = (int)pwp_init_.coeff_x.size();
    if (n < 1 || n > NB_NPOL) throw std::runtime_error("pwp_init with " + std::to_string(n) + " intervals (1..8 supported)");
    for (auto& bl : bendPtsForAgents_)   // lists longer than the current storage: rebuild the handle once, larger
      if ((int)bl.size() > par_.bp_max) par_.bp_max = 2 * (int)bl.size(), reset();
    for (auto& e : entStateVec_)
      if ((int)e.alphas.size() > par_.ent_cap) par_.ent_cap = 2 * (int)e.alphas.size(), reset();
    for (int attempt = 0;; attempt++)
    {
      const int rc = replanOnce(objective_value);
      if (rc == NB_OK) break;
      if (rc == NB_ERR_CAPACITY && attempt < 4)
      {  // more tether constraints in one interval than LP slots: the reference's vectors just grow
        par_.ent_slots *= 2, reset();
        continue;
      }
      throw std::runtime_error(std::string("nb_replan_batch: ") + nb_last_error());
    }
    if (last_status_ == NB_STATUS_FAILED) return false;  // pwp_out_ == pwp_init_ (:856-859)
    solutions_found_++;
    return true;
  }
  int replanOnce(double& objective_value)
  {
    ensure();
    const int N = par_.num_agents, M = par_.num_static, NA = N + M, cap = par_.ent_cap;
    const int n = (int)pwp_init_.coeff_x.size();
    const int NH = (int)hulls_.size();
    std::vector<double> ci(96, 0.0), nih0((size_t)N * 16, std::nan("")), hxy, bp_xy((size_t)N * par_.bp_max * 2, 0.0);
    std::vector<int64_t> hptr((size_t)(NH > 0 ? NH : 1) * 8 + 1, 0);
    std::vector<int32_t> esv_cnt(18, 0), esv_alpha((size_t)9 * cap * 2, 0), esv_active((size_t)9 * NA, 0), bp_cnt(N, 0);
    for (int i = 0; i < n; i++)
      for (int r = 0; r < 4; r++)
      {
        ci[4 * i + r] = pwp_init_.coeff_x[i](r);
        ci[32 + 4 * i + r] = pwp_init_.coeff_y[i](r);
        ci[64 + 4 * i + r] = pwp_init_.coeff_z[i](r);
      }
    for (int s = 0; s < NH; s++)
      for (int i = 0; i < 8; i++)
      {
        hptr[s * 8 + i] = (int64_t)hxy.size() / 2;
        if (i < n && i < (int)hulls_[s].size())
          for (int c = 0; c < hulls_[s][i].cols(); c++) hxy.push_back(hulls_[s][i](0, c)), hxy.push_back(hulls_[s][i](1, c));
      }
    hptr[(size_t)NH * 8] = (int64_t)hxy.size() / 2;
    for (int j = 0; j < N && j < (int)hullsNoInflation_.size(); j++)
      for (int i = 0; i < n && i < (int)hullsNoInflation_[j].size(); i++)
        if (hullsNoInflation_[j][i].cols() > 0)
        {
          nih0[((size_t)j * 8 + i) * 2] = hullsNoInflation_[j][i](0, 0);
          nih0[((size_t)j * 8 + i) * 2 + 1] = hullsNoInflation_[j][i](1, 0);
        }
    for (int i = 0; i <= n && i < (int)entStateVec_.size(); i++)
    {
      const eu::ent_state& e = entStateVec_[i];
      esv_cnt[2 * i] = (int)e.alphas.size(), esv_cnt[2 * i + 1] = (int)e.bendPointsIdx.size();
      for (size_t q = 0; q < e.alphas.size(); q++)
        esv_alpha[((size_t)i * cap + q) * 2] = e.alphas[q](0), esv_alpha[((size_t)i * cap + q) * 2 + 1] = e.alphas[q](1);
      for (int q = 0; q < NA && q < (int)e.active_cases.size(); q++) esv_active[(size_t)i * NA + q] = e.active_cases[q];
    }
    for (int j = 0; j < N && j < (int)bendPtsForAgents_.size(); j++)
    {
      bp_cnt[j] = (int)bendPtsForAgents_[j].size();
      for (int q = 0; q < bp_cnt[j]; q++)
        bp_xy[((size_t)j * par_.bp_max + q) * 2] = bendPtsForAgents_[j][q](0),
                                   bp_xy[((size_t)j * par_.bp_max + q) * 2 + 1] = bendPtsForAgents_[j][q](1);
    }
    int32_t agent_id = id_, n_int = n, status = -1, iters[2] = { 0, 0 };
    double co[96], obj = 0.0;
    nb_replan_args a = nb_replan_args();
    a.B = 1, a.space = NB_HOST, a.agent_id = &agent_id, a.n_int = &n_int, a.coeff_init = ci.data();
    a.n_hull_slots = NH, a.hull_ptr = hptr.data(), a.hull_xy = hxy.empty() ? &obj : hxy.data(), a.hull_nvert = (int64_t)hxy.size() / 2;
    a.hull_cnt = nullptr, a.nih0 = nih0.data(), a.esv_cnt = esv_cnt.data(), a.esv_alpha = esv_alpha.data();
    a.esv_active = esv_active.data(), a.bp_cnt = bp_cnt.data(), a.bp_xy = bp_xy.data();
    a.coeff_out = co, a.obj = &obj, a.status = &status, a.iters = iters, a.lines = nullptr, a.line_ok = nullptr;
    const int rc = nb_replan_batch(h_, &a, nullptr);
    if (rc != NB_OK) return rc;
    total_replannings_++;
    pwp_out_ = pwp_init_;
    for (int i = 0; i < n; i++)
    {
      pwp_out_.coeff_x[i] = Eigen::Matrix<double, 4, 1>(co[4 * i], co[4 * i + 1], co[4 * i + 2], co[4 * i + 3]);
      pwp_out_.coeff_y[i] = Eigen::Matrix<double, 4, 1>(co[32 + 4 * i], co[32 + 4 * i + 1], co[32 + 4 * i + 2], co[32 + 4 * i + 3]);
      pwp_out_.coeff_z[i] = Eigen::Matrix<double, 4, 1>(co[64 + 4 * i], co[64 + 4 * i + 1], co[64 + 4 * i + 2], co[64 + 4 * i + 3]);
    }
    last_status_ = status;
    if (status != NB_STATUS_FAILED) objective_value = obj;
    return NB_OK;
  }

public:

  // solver_gurobi_poly.cpp:889-936
  void generatePwpOut(mt::PieceWisePol& pwp_out, std::vector<mt::state>& traj_out, double t_start, double dc)
  {
    pwp_out = pwp_out_;
    const int nt = (int)pwp_out.times.size(), n = nt - 1;
    for (int i = 0; i < nt; i++) pwp_out.times[i] += t_start;
    traj_out.clear();
    double t = 0;
    int i = 0;
    while (i < n)
    {
      const double dt = t - i * par_.T_span;
      mt::state st;
      const Eigen::Matrix<double, 4, 1>* c[3] = { &pwp_out.coeff_x[i], &pwp_out.coeff_y[i], &pwp_out.coeff_z[i] };
      for (int ax = 0; ax < 3; ax++)
      {
        const Eigen::Matrix<double, 4, 1>& q = *c[ax];
        st.pos(ax) = q(0) * dt * dt * dt + q(1) * dt * dt + q(2) * dt + q(3);
        st.vel(ax) = q(0) * 3 * dt * dt + q(1) * 2 * dt + q(2);
        st.accel(ax) = q(0) * 6 * dt + q(1) * 2;
        st.jerk(ax) = q(0) * 6;
      }
      traj_out.push_back(st);
      t += dc;
      if (t > (i + 1) * par_.T_span) i++;
    }
  }

  int lastStatus() const { return last_status_; }  // NB_STATUS_* of the last optimize() (extension)

private:
  void reset()
  {
    if (h_) nb_destroy(h_);
    h_ = nullptr;
  }
  void ensure()
  {
    if (h_) return;
    std::vector<double> pb;
    for (auto& p : pb_) pb.push_back(p(0)), pb.push_back(p(1));
    nb_detail::check(nb_create(&par_, pb.data(), -1, &h_), "nb_create");
    if (par_.num_static > 0) nb_detail::check(nb_set_static(h_, st_ptr_.data(), st_xy_.data(), nullptr), "nb_set_static");
  }
  nb_params par_;
  nb_handle* h_ = nullptr;
  int id_;
  std::vector<Eigen::Vector2d> pb_;
  std::vector<int64_t> st_ptr_{ 0 };
  std::vector<double> st_xy_;
  mt::PieceWisePol pwp_init_, pwp_out_;
  mt::ConvexHullsOfCurves_Std2d hulls_, hullsNoInflation_;
  std::vector<eu::ent_state> entStateVec_;
  std::vector<std::vector<Eigen::Vector2d>> bendPtsForAgents_;
  std::vector<std::vector<Eigen::Vector3d>> vecOfAgentsBetas_;
  std::string last_error_;
  double max_runtime_ = 0.05;
  int total_replannings_ = 0, solutions_found_ = 0, last_status_ = -1;
};

namespace separator
{
// submodules/separator/include/separator.hpp:18-48, 2-D overloads (separator_glpk.cpp:248, :375, :500)
class Separator
{
public:
  Separator()
  {
    nb_params p = nb_params();
    p.num_pol = 8, p.deg_pol = 3, p.num_agents = 1, p.use_linear_constraints = 1, p.T_span = 0.5, p.weight = 1.0;
    p.samples = 3, p.ent_cap = 8, p.bp_max = 2, p.ent_slots = 1, p.ipm_max_iter = 1, p.ipm_tol = 1e-9;
    const double pb[2] = { 0, 0 };
    nb_detail::check(nb_create(&p, pb, -1, &h_), "nb_create");
  }
  ~Separator()
  {
    if (h_) nb_destroy(h_);
  }
  Separator(const Separator&) = delete;
  bool solveModel(Eigen::Vector3d& solutionN, const mt::Polygon_Std& pointsA, const mt::Polygon_Std& pointsB)
  {
    return run(solutionN, pointsA, nullptr, pointsB);
  }
  bool solveModel(Eigen::Vector3d& solutionN, const mt::Polygon_Std& pointsA, const mt::Polygon_Std& pointsAPlus,
                  const mt::Polygon_Std& pointsB)
  {
    return run(solutionN, pointsA, &pointsAPlus, pointsB);
  }
  bool solveModel(const mt::Polygon_Std& pointsA, const mt::Polygon_Std& pointsB)
  {
    Eigen::Vector3d n;
    return run(n, pointsA, nullptr, pointsB);
  }
  long int getNumOfLPsRun() { return num_; }
  double meanSolveTimeMs() { return num_ > 0 ? total_ms_ / (double)num_ : 0.0; }  // separator.hpp:41

private:
  bool run(Eigen::Vector3d& out, const mt::Polygon_Std& A, const mt::Polygon_Std* Ap, const mt::Polygon_Std& B)
  {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<double> a, b;
    for (int c = 0; c < A.cols(); c++) a.push_back(A(0, c)), a.push_back(A(1, c));
    if (Ap)
      for (int c = 0; c < Ap->cols(); c++) a.push_back((*Ap)(0, c)), a.push_back((*Ap)(1, c));
    for (int c = 0; c < B.cols(); c++) b.push_back(B(0, c)), b.push_back(B(1, c));
    const int64_t ap[2] = { 0, (int64_t)a.size() / 2 }, bp[2] = { 0, (int64_t)b.size() / 2 };
    double line[3];
    uint8_t ok = 0;
    const int rc = nb_separate_batch(h_, 1, NB_HOST, ap, a.data(), bp, b.data(), 0, line, &ok, nullptr);
    num_++;
    total_ms_ += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc != NB_OK)
    {  // the reference reports an LP it cannot solve as "not separable" (separator_glpk.cpp:340-346); so does a library error
      std::fprintf(stderr, "separator::Separator (neptune_b200): %s\n", nb_last_error());
      return false;
    }
    out(0) = line[0], out(1) = line[1], out(2) = line[2];
    return ok != 0;
  }
  nb_handle* h_ = nullptr;
  long int num_ = 0;
  double total_ms_ = 0.0;
};
}  // namespace separator

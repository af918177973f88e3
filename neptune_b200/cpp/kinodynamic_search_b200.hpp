// kinodynamic_search_b200.hpp -- source-compatible replacement of the reference's front-end class on top of the
// C-ABI of include/neptune_b200.h (B = 1 per call, host buffers).
//
//   class KinodynamicSearch   <- neptune/include/kinodynamic_search.hpp:115-199 (the members Neptune uses:
//                                 constructor neptune.cpp:88-91, one-time setters :92-97, :659-670, per replan
//                                 setRunTime :1421, setInitZCoeffs :1430, setUp :1450, run :1453,
//                                 getPwpOut_0tstart / getEntStateVector :1509-1510)
//
// Same method names, argument meaning and call order.  run() returns false when no node qualifies (and, with
// use_not_reaching_soln == false, whenever the goal was not reached), exactly like the reference (:1789-1799).
// Two extensions replace non-deterministic state of the reference: setJerkOrder (the reference shuffles
// all_combinations_ with a wall-clock seed, kinodynamic_search.cpp:321, :1462) and setMaxExpansions (open-list pops
// that stand in for the wall-clock budget of setRunTime, :1646).
//
// Like poly_solver_b200.hpp this header defines the include guard of the reference's kinodynamic_search.hpp, so that
// seen first (`g++ -include kinodynamic_search_b200.hpp`) it takes the reference class's place in neptune.cpp.
#pragma once
#ifndef KINODYNAMIC_SEARCH_HPP
#define KINODYNAMIC_SEARCH_HPP  // include guard of the reference's header
#endif
#include <numeric>

#include "poly_solver_b200.hpp"

class KinodynamicSearch
{
public:
  KinodynamicSearch(int num_pol, int deg_pol, int id, double safe_factor, double T_span, int num_sample_per_interval,
                    std::vector<Eigen::Vector2d> pb, bool use_not_reaching_soln, bool enable_entangle_check = true)
    : id_(id), pb_(pb)
  {
    par_ = nb_params();
    par_.num_pol = num_pol, par_.deg_pol = deg_pol, par_.num_agents = (int)pb.size(), par_.num_static = 0;
    par_.samples = num_sample_per_interval, par_.use_linear_constraints = 1, par_.T_span = T_span, par_.weight = 1.0;
    par_.ent_cap = 3 * (int)pb.size() + 16, par_.bp_max = 8, par_.ent_slots = 16, par_.ipm_max_iter = 30, par_.ipm_tol = 1e-9;
    sp_ = nb_search_params();
    sp_.num_samples = 5, sp_.j_max = 5.0, sp_.voxel_size = 0.2, sp_.bias = 1.0, sp_.goal_size = 0.5;
    sp_.enable_entangle_check = enable_entangle_check ? 1 : 0, sp_.use_not_reaching_soln = use_not_reaching_soln ? 1 : 0;
    sp_.max_nodes = 4096, sp_.max_expansions = 600, sp_.ecap = 24;
    (void)safe_factor;  // only read by collidesWithObstacles2d, which run() does not call (:1466-1511)
    comb_.resize(25);
    std::iota(comb_.begin(), comb_.end(), 0);
  }
  ~KinodynamicSearch()
  {
    if (h_) nb_destroy(h_);
  }
  KinodynamicSearch(const KinodynamicSearch&) = delete;
  KinodynamicSearch& operator=(const KinodynamicSearch&) = delete;

  void setTetherLength(double tetherLength) { par_.tether_length = tetherLength, reset(); }
  void setMaxValuesAndSamples(double v_max, double a_max, double j_max, int num_samples)
  {
    par_.v_max = v_max, par_.a_max = a_max, sp_.j_max = j_max, sp_.num_samples = num_samples;
    comb_.resize((size_t)num_samples * num_samples);
    std::iota(comb_.begin(), comb_.end(), 0);
    reset();
  }
  void setXYZMinMaxAndRa(double x_min, double x_max, double y_min, double y_max, double z_min, double z_max, double Ra,
                         double voxel_size)
  {
    par_.lim_min[0] = x_min, par_.lim_min[1] = y_min, par_.lim_min[2] = z_min;
    par_.lim_max[0] = x_max, par_.lim_max[1] = y_max, par_.lim_max[2] = z_max;
    sp_.voxel_size = voxel_size;
    (void)Ra;
    reset();
  }
  void setBias(double bias) { sp_.bias = bias, configured_ = false; }
  void setGoalSize(double goal_size) { sp_.goal_size = goal_size, configured_ = false; }
  void setRunTime(double max_runtime) { max_runtime_ = max_runtime; }  // see setMaxExpansions
  void setMaxExpansions(int pops) { sp_.max_expansions = pops, configured_ = false; }
  void setJerkOrder(const std::vector<int>& order) { comb_.assign(order.begin(), order.end()); }
  void setInitZCoeffs(std::vector<Eigen::Matrix<double, 4, 1>>& coeffs_z) { coeffs_z_ = coeffs_z; }
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
    reset();
  }
  // staticObsRep[m]: 2x2, col(0) and col(1) = the two representative points; staticObsLongestDist[m] = (d0, d1)
  void setStaticObstRep(const std::vector<std::vector<Eigen::Vector2d>>& staticObsRep,
                        const std::vector<Eigen::Vector2d>& staticObsLongestDist)
  {
    strep_.clear(), longest_.clear();
    for (auto& r : staticObsRep)
      for (int c = 0; c < 2; c++) strep_.push_back(r[c](0)), strep_.push_back(r[c](1));
    for (auto& d : staticObsLongestDist) longest_.push_back(d(0)), longest_.push_back(d(1));
    reset();
  }

#ifdef NB_HAVE_EIGEN
  // the reference's own signature (kinodynamic_search.hpp:144-145)
  void setStaticObstRep(std::vector<Eigen::Matrix<double, 2, 2>>& staticObsRep, std::vector<Eigen::Vector2d>& staticObsLongestDist)
  {
    std::vector<std::vector<Eigen::Vector2d>> rep;
    for (auto& m : staticObsRep) rep.push_back({ Eigen::Vector2d(m(0, 0), m(1, 0)), Eigen::Vector2d(m(0, 1), m(1, 1)) });
    setStaticObstRep(rep, staticObsLongestDist);
  }
#endif

  void setUp(mt::state initial_state, Eigen::Vector3d& goal, const mt::ConvexHullsOfCurves_Std2d& hulls,
             const mt::SampledPointsofCurves& SPoC, eu::ent_state& entangle_state,
             std::vector<std::vector<Eigen::Vector2d>>& bendPtsForAgents)
  {
    initial_ = initial_state, goal_ = goal, hulls_ = hulls, SPoC_ = SPoC, es_ = entangle_state, bend_ = bendPtsForAgents;
  }

  // run() of the reference never throws; a library error (no device, a list beyond its storage) is reported as
  // "no solution" (status 0, false) with the text kept in lastError()
  bool run(std::vector<Eigen::Vector3d>& result, int& status)
  {
    last_error_.clear();
    try
    {
      return runImpl(result, status);
    }
    catch (const std::exception& e)
    {
      last_error_ = e.what();
      std::fprintf(stderr, "KinodynamicSearch (neptune_b200): %s -- no solution returned\n", e.what());
      status = 0;
      pwp_out_.clear();
      entStateVec_.clear();
      return false;
    }
  }
  const std::string& lastError() const { return last_error_; }
  void clearProcess() {}  // the node pool lives in the library's workspace and is reset by every search (:1829-1858)

  // KinodynamicSearch::generatePwpOut (kinodynamic_search.cpp:605-655): same sampling loop as the back end's
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

  // KinodynamicSearch::updateSPocAndbendPtsForAgent (:987-993): a late trajectory replaces the samples and the bend
  // points the planner holds for that agent (Neptune::safetyCheckAfterReplan, neptune.cpp:737-742)
  void updateSPocAndbendPtsForAgent(int idx, mt::SampledPointsofIntervals& SampledPtsForOne, std::vector<Eigen::Vector2d>& bendPtsForAgent)
  {
    if (idx >= (int)SPoC_.size()) SPoC_.resize(idx + 1);
    if (idx >= (int)bend_.size()) bend_.resize(idx + 1);
    SPoC_[idx] = SampledPtsForOne;
    bend_[idx] = bendPtsForAgent;
  }

  // KinodynamicSearch::entangleCheckGivenPwp (:897-985), on the samples and bend points the planner holds; the state
  // is advanced in place like the reference's argument.  A library error counts as "entangled" (the replan is dropped).
  bool entangleCheckGivenPwp(mt::PieceWisePol& pwp, eu::ent_state& ent_state_begin)
  {
    try
    {
      ensure();
      const int N = par_.num_agents, M = par_.num_static, NA = N + M, cap = par_.ent_cap, S = par_.samples, np = par_.num_pol;
      const int n = (int)pwp.coeff_x.size();
      if (n < 1 || n > NB_NPOL || (int)ent_state_begin.alphas.size() > cap) throw std::runtime_error("entangleCheckGivenPwp: input beyond storage");
      std::vector<double> co(96, 0.0), samp((size_t)N * np * (S + 1) * 2, 0.0), bp_xy((size_t)N * par_.bp_max * 2, 0.0), beta(cap, 0.0);
      std::vector<int32_t> cnt(2, 0), alpha((size_t)cap * 2, 0), bend(cap, 0), active(NA, 0), bp_cnt(N, 0);
      std::vector<uint8_t> known(N, 0);
      for (int i = 0; i < n; i++)
        for (int r = 0; r < 4; r++)
          co[4 * i + r] = pwp.coeff_x[i](r), co[32 + 4 * i + r] = pwp.coeff_y[i](r), co[64 + 4 * i + r] = pwp.coeff_z[i](r);
      for (int j = 0; j < N && j < (int)SPoC_.size(); j++)
        if (j != id_ - 1 && !SPoC_[j].empty())
        {
          known[j] = 1;
          for (int i = 0; i < np && i < (int)SPoC_[j].size(); i++)
            for (int s2 = 0; s2 <= S && s2 < SPoC_[j][i].cols(); s2++)
              samp[(((size_t)j * np + i) * (S + 1) + s2) * 2] = SPoC_[j][i](0, s2), samp[(((size_t)j * np + i) * (S + 1) + s2) * 2 + 1] = SPoC_[j][i](1, s2);
        }
      for (int j = 0; j < N && j < (int)bend_.size(); j++)
      {
        if ((int)bend_[j].size() > par_.bp_max) throw std::runtime_error("bend-point list longer than bp_max");
        bp_cnt[j] = (int)bend_[j].size();
        for (int q = 0; q < bp_cnt[j]; q++) bp_xy[((size_t)j * par_.bp_max + q) * 2] = bend_[j][q](0), bp_xy[((size_t)j * par_.bp_max + q) * 2 + 1] = bend_[j][q](1);
      }
      eu::ent_state& e = ent_state_begin;
      cnt[0] = (int)e.alphas.size(), cnt[1] = (int)e.bendPointsIdx.size();
      for (size_t q = 0; q < e.alphas.size(); q++) alpha[2 * q] = e.alphas[q](0), alpha[2 * q + 1] = e.alphas[q](1);
      for (size_t q = 0; q < e.betas.size() && q < (size_t)cap; q++) beta[q] = e.betas[q];
      for (size_t q = 0; q < e.bendPointsIdx.size(); q++) bend[q] = e.bendPointsIdx[q];
      for (int q = 0; q < NA && q < (int)e.active_cases.size(); q++) active[q] = e.active_cases[q];
      int32_t agent_id = id_, n_int = n, entangled = 0;
      nb_ent_state st;
      st.cnt = cnt.data(), st.alpha = alpha.data(), st.beta = beta.data(), st.bend = bend.data(), st.active = active.data();
      nb_detail::check(nb_entangle_check_batch(h_, 1, NB_HOST, &agent_id, known.data(), bp_cnt.data(), bp_xy.data(), st, &n_int,
                                               co.data(), samp.data(), 1, &entangled, nullptr),
                       "nb_entangle_check_batch");
      e.alphas.clear(), e.betas.clear(), e.bendPointsIdx.clear();
      for (int q = 0; q < cnt[0]; q++) e.alphas.push_back(Eigen::Vector2i(alpha[2 * q], alpha[2 * q + 1])), e.betas.push_back(beta[q]);
      for (int q = 0; q < cnt[1]; q++) e.bendPointsIdx.push_back(bend[q]);
      e.active_cases.assign(active.begin(), active.end());
      return entangled != 0;
    }
    catch (const std::exception& ex)
    {
      last_error_ = ex.what();
      std::fprintf(stderr, "KinodynamicSearch (neptune_b200): %s -- reported as entangled\n", ex.what());
      return true;
    }
  }

private:
  bool runImpl(std::vector<Eigen::Vector3d>& result, int& status)
  {
    ensure();
    result.clear();
    const int N = par_.num_agents, M = par_.num_static, NA = N + M, cap = par_.ent_cap, S = par_.samples, np = par_.num_pol;
    const int ns2 = sp_.num_samples * sp_.num_samples;
    std::vector<double> init = { initial_.pos(0), initial_.pos(1), initial_.vel(0), initial_.vel(1), initial_.accel(0), initial_.accel(1) };
    std::vector<double> goal = { goal_(0), goal_(1) }, cz(32, 0.0), hxy((size_t)N * 8 * NB_HULL_STRIDE * 2, 0.0);
    std::vector<double> samp((size_t)N * np * (S + 1) * 2, 0.0), bp_xy((size_t)N * par_.bp_max * 2, 0.0), es_beta(cap, 0.0);
    std::vector<int32_t> hcnt((size_t)N * 8, 0), es_cnt(2, 0), es_alpha((size_t)cap * 2, 0), es_bend(cap, 0), es_active(NA, 0), bp_cnt(N, 0);
    std::vector<uint8_t> known(N, 0), comb(ns2, 0);
    for (int i = 0; i < np && i < (int)coeffs_z_.size(); i++)
      for (int r = 0; r < 4; r++) cz[4 * i + r] = coeffs_z_[i](r);
    // SampledPointsForAll_ is indexed by agent id - 1 (empty = unknown); hulls_ holds one entry per known
    // trajectory in arrival order.  The collision tests are an "any" over hulls, so hull k goes to the slot of the
    // k-th known agent.
    std::vector<int> slots;
    for (int j = 0; j < N && j < (int)SPoC_.size(); j++)
      if (j != id_ - 1 && !SPoC_[j].empty())
      {
        known[j] = 1;
        slots.push_back(j);
        for (int i = 0; i < np && i < (int)SPoC_[j].size(); i++)
          for (int s = 0; s <= S && s < SPoC_[j][i].cols(); s++)
            samp[(((size_t)j * np + i) * (S + 1) + s) * 2] = SPoC_[j][i](0, s), samp[(((size_t)j * np + i) * (S + 1) + s) * 2 + 1] = SPoC_[j][i](1, s);
      }
    if (hulls_.size() > slots.size()) throw std::runtime_error("KinodynamicSearch: more hulls than known trajectories");
    for (size_t k = 0; k < hulls_.size(); k++)
      for (int i = 0; i < 8 && i < (int)hulls_[k].size(); i++)
      {
        const int nv = hulls_[k][i].cols();
        if (nv > NB_HULL_STRIDE) throw std::runtime_error("hull with more than NB_HULL_STRIDE vertices");
        hcnt[(size_t)slots[k] * 8 + i] = nv;
        for (int v = 0; v < nv; v++)
          hxy[(((size_t)slots[k] * 8 + i) * NB_HULL_STRIDE + v) * 2] = hulls_[k][i](0, v), hxy[(((size_t)slots[k] * 8 + i) * NB_HULL_STRIDE + v) * 2 + 1] = hulls_[k][i](1, v);
      }
    if ((int)es_.alphas.size() > cap) throw std::runtime_error("ent_state longer than ent_cap");
    es_cnt[0] = (int)es_.alphas.size(), es_cnt[1] = (int)es_.bendPointsIdx.size();
    for (size_t q = 0; q < es_.alphas.size(); q++) es_alpha[2 * q] = es_.alphas[q](0), es_alpha[2 * q + 1] = es_.alphas[q](1);
    for (size_t q = 0; q < es_.betas.size() && q < (size_t)cap; q++) es_beta[q] = es_.betas[q];
    for (size_t q = 0; q < es_.bendPointsIdx.size(); q++) es_bend[q] = es_.bendPointsIdx[q];
    for (int q = 0; q < NA && q < (int)es_.active_cases.size(); q++) es_active[q] = es_.active_cases[q];
    for (int j = 0; j < N && j < (int)bend_.size(); j++)
    {
      if ((int)bend_[j].size() > par_.bp_max) throw std::runtime_error("bend-point list longer than bp_max");
      bp_cnt[j] = (int)bend_[j].size();
      for (int q = 0; q < bp_cnt[j]; q++) bp_xy[((size_t)j * par_.bp_max + q) * 2] = bend_[j][q](0), bp_xy[((size_t)j * par_.bp_max + q) * 2 + 1] = bend_[j][q](1);
    }
    for (int k = 0; k < ns2; k++) comb[k] = (uint8_t)comb_[k];
    int32_t agent_id = id_, st = -1, solved = 0, n_int = 0, stats[4] = { 0, 0, 0, 0 };
    double cost = 0.0;
    std::vector<double> co(96, 0.0), ev_beta((size_t)9 * cap, 0.0);
    std::vector<int32_t> ev_cnt(18, 0), ev_alpha((size_t)9 * cap * 2, 0), ev_bend((size_t)9 * cap, 0), ev_active((size_t)9 * NA, 0);
    nb_search_args a = nb_search_args();
    a.B = 1, a.space = NB_HOST, a.agent_id = &agent_id, a.init = init.data(), a.goal = goal.data(), a.coeffs_z = cz.data();
    a.n_groups = 1, a.group = nullptr, a.hull_xy = hxy.data(), a.hull_cnt = hcnt.data(), a.samp = samp.data(), a.known = known.data();
    a.es.cnt = es_cnt.data(), a.es.alpha = es_alpha.data(), a.es.beta = es_beta.data(), a.es.bend = es_bend.data(), a.es.active = es_active.data();
    a.bp_cnt = bp_cnt.data(), a.bp_xy = bp_xy.data(), a.comb = comb.data(), a.comb_shared = 1;
    a.status = &st, a.solved = &solved, a.n_int = &n_int, a.coeff = co.data();
    a.esv.cnt = ev_cnt.data(), a.esv.alpha = ev_alpha.data(), a.esv.beta = ev_beta.data(), a.esv.bend = ev_bend.data(), a.esv.active = ev_active.data();
    a.stats = stats, a.cost = &cost;
    nb_detail::check(nb_search_batch(h_, &a, nullptr), "nb_search_batch");
    status = st;
    node_used_num_ = stats[0];
    pwp_out_.clear();
    entStateVec_.clear();
    if (!solved) return false;
    for (int i = 0; i <= n_int; i++) pwp_out_.times.push_back(i * par_.T_span);  // recoverPwpOut (:521-553)
    for (int i = 0; i < n_int; i++)
    {
      pwp_out_.coeff_x.push_back(Eigen::Matrix<double, 4, 1>(co[4 * i], co[4 * i + 1], co[4 * i + 2], co[4 * i + 3]));
      pwp_out_.coeff_y.push_back(Eigen::Matrix<double, 4, 1>(co[32 + 4 * i], co[32 + 4 * i + 1], co[32 + 4 * i + 2], co[32 + 4 * i + 3]));
      pwp_out_.coeff_z.push_back(Eigen::Matrix<double, 4, 1>(co[64 + 4 * i], co[64 + 4 * i + 1], co[64 + 4 * i + 2], co[64 + 4 * i + 3]));
    }
    for (int i = 0; i <= n_int; i++)  // recoverEntStateVector (:582-603)
    {
      eu::ent_state e;
      for (int q = 0; q < ev_cnt[2 * i]; q++)
      {
        e.alphas.push_back(Eigen::Vector2i(ev_alpha[((size_t)i * cap + q) * 2], ev_alpha[((size_t)i * cap + q) * 2 + 1]));
        e.betas.push_back(ev_beta[(size_t)i * cap + q]);
      }
      for (int q = 0; q < ev_cnt[2 * i + 1]; q++) e.bendPointsIdx.push_back(ev_bend[(size_t)i * cap + q]);
      for (int q = 0; q < NA; q++) e.active_cases.push_back(ev_active[(size_t)i * NA + q]);
      entStateVec_.push_back(e);
    }
    return true;
  }


public:
  void getPwpOut_0tstart(mt::PieceWisePol& pwp_out) { pwp_out = pwp_out_; }
  void getEntStateVector(std::vector<eu::ent_state>& entStateVec) { entStateVec = entStateVec_; }
  void getRuntime(double& runtime_this_round, double& time_spent_contact_pt, int& node_used_num)
  {
    runtime_this_round = 0.0, time_spent_contact_pt = 0.0, node_used_num = node_used_num_;
  }

private:
  void reset()
  {
    if (h_) nb_destroy(h_);
    h_ = nullptr, configured_ = false;
  }
  void ensure()
  {
    if (!h_)
    {
      std::vector<double> pb;
      for (auto& q : pb_) pb.push_back(q(0)), pb.push_back(q(1));
      nb_detail::check(nb_create(&par_, pb.data(), -1, &h_), "nb_create");
      if (par_.num_static > 0)
      {
        if ((int)strep_.size() != 4 * par_.num_static || (int)longest_.size() != 2 * par_.num_static)
          throw std::runtime_error("KinodynamicSearch: setStaticObstRep must follow setStaticObstVert");
        nb_detail::check(nb_set_static(h_, st_ptr_.data(), st_xy_.data(), strep_.data()), "nb_set_static");
        nb_detail::check(nb_set_static_longest(h_, longest_.data()), "nb_set_static_longest");
      }
    }
    if (!configured_) nb_detail::check(nb_search_configure(h_, &sp_), "nb_search_configure"), configured_ = true;
  }

  nb_params par_;
  nb_search_params sp_;
  nb_handle* h_ = nullptr;
  bool configured_ = false;
  std::string last_error_;
  int id_, node_used_num_ = 0;
  double max_runtime_ = 0.5;
  std::vector<Eigen::Vector2d> pb_;
  std::vector<int> comb_;
  std::vector<Eigen::Matrix<double, 4, 1>> coeffs_z_;
  std::vector<int64_t> st_ptr_ = { 0 };
  std::vector<double> st_xy_, strep_, longest_;
  mt::state initial_;
  Eigen::Vector3d goal_;
  mt::ConvexHullsOfCurves_Std2d hulls_;
  mt::SampledPointsofCurves SPoC_;
  eu::ent_state es_;
  std::vector<std::vector<Eigen::Vector2d>> bend_;
  mt::PieceWisePol pwp_out_;
  std::vector<eu::ent_state> entStateVec_;
};

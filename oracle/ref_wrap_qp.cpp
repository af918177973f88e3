// ref_wrap_qp.cpp -- C entry point around the REFERENCE's own PolySolverGurobi (TEST INFRASTRUCTURE).
//
// oracle/Makefile (target _ref) compiles neptune/src/solver_gurobi_poly.cpp where it lies under /root/reference,
// unmodified, against the Eigen stand-in (oracle/eigen_shim), the Gurobi stand-in (oracle/ref_stubs/gurobi_c++.h: records
// the model, defers the solve to a callback) and the reference's own separator_glpk.cpp (its LP engine likewise a
// callback).  This wrapper drives the object exactly as Neptune does -- constructor and one-time setters as at
// neptune.cpp:102-107 and :663, then setInitTrajectory -> setHulls -> setHullsNoInflation -> setEntStateVector ->
// optimize -> generatePwpOut as at neptune.cpp:1514-1527 -- so the reference's own setInitTrajectory / addObjective /
// addConstraints / addEntangleConstraintForIJCase / optimize build the model and walk the status path;
// tests/test_reference_pin.py compares every recorded model with the oracle's restatement row for row.
#include <cmath>
#include <cstdint>
#include <vector>

#include "solver_gurobi_poly.hpp"

extern "C" int ref_qp_replan(int N, int id, int num_pol, double T_span, double weight, const double* pb, const double* lim,
                             double v_max, double a_max, double j_max, double runtime, double tether, int M,
                             const int64_t* st_ptr, const double* st_xy, int n, const double* coeff_init, int NH,
                             const int64_t* hull_ptr, const double* hull_xy, const double* nih0, const int* esv_cnt,
                             const int* esv_alpha, const int* esv_active, int cap, const int* bp_cnt, const double* bp_xy,
                             int bp_max, int replans, double t_start, double dc, double* coeff_out, double* times_out,
                             double* obj, int* n_states)
{
  std::vector<Eigen::Vector2d> pbv;
  for (int j = 0; j < N; j++) pbv.push_back(Eigen::Vector2d(pb[2 * j], pb[2 * j + 1]));
  PolySolverGurobi solver(num_pol, 3, id, T_span, pbv, weight, 0.5, true);
  solver.setMaxValues(lim[0], lim[1], lim[2], lim[3], lim[4], lim[5], v_max, a_max, j_max);
  solver.setMaxRuntime(runtime);
  solver.setTetherLength(tether);
  std::vector<mt::Polygon_Std> statics;
  for (int m = 0; m < M; m++)
  {
    const int c = (int)(st_ptr[m + 1] - st_ptr[m]);
    mt::Polygon_Std p(2, c);
    for (int q = 0; q < c; q++) p(0, q) = st_xy[2 * (st_ptr[m] + q)], p(1, q) = st_xy[2 * (st_ptr[m] + q) + 1];
    statics.push_back(p);
  }
  solver.setStaticObstVert(statics);

  mt::PieceWisePol pwp;
  for (int i = 0; i <= n; i++) pwp.times.push_back(i * T_span);
  for (int i = 0; i < n; i++)
  {
    const double* c = coeff_init + 4 * i;
    pwp.coeff_x.push_back(Eigen::Matrix<double, 4, 1>(c[0], c[1], c[2], c[3]));
    pwp.coeff_y.push_back(Eigen::Matrix<double, 4, 1>(c[32], c[33], c[34], c[35]));
    pwp.coeff_z.push_back(Eigen::Matrix<double, 4, 1>(c[64], c[65], c[66], c[67]));
  }
  // hulls_: one entry per KNOWN other agent (the reference never holds empty entries), slot order kept
  mt::ConvexHullsOfCurves_Std2d hulls, nih(N);
  for (int s = 0; s < NH; s++)
  {
    if (hull_ptr[s * 8 + 1] == hull_ptr[s * 8]) continue;
    mt::ConvexHullsOfCurve_Std2d per;
    for (int i = 0; i < n; i++)
    {
      const int64_t o = hull_ptr[s * 8 + i];
      const int c = (int)(hull_ptr[s * 8 + i + 1] - o);
      mt::Polygon_Std p(2, c);
      for (int q = 0; q < c; q++) p(0, q) = hull_xy[2 * (o + q)], p(1, q) = hull_xy[2 * (o + q) + 1];
      per.push_back(p);
    }
    hulls.push_back(per);
  }
  for (int j = 0; j < N; j++)
    for (int i = 0; i < n; i++)
    {
      const double x = nih0[((size_t)j * 8 + i) * 2], y = nih0[((size_t)j * 8 + i) * 2 + 1];
      mt::Polygon_Std p(2, x == x ? 1 : 0);
      if (x == x) p(0, 0) = x, p(1, 0) = y;
      nih[j].push_back(p);
    }
  const int NA = N + M;
  std::vector<eu::ent_state> esv(n + 1);
  for (int i = 0; i <= n; i++)
  {
    for (int q = 0; q < esv_cnt[2 * i]; q++)
      esv[i].alphas.push_back(Eigen::Vector2i(esv_alpha[((size_t)i * cap + q) * 2], esv_alpha[((size_t)i * cap + q) * 2 + 1]));
    esv[i].betas.assign(esv[i].alphas.size(), 0.0);
    for (int q = 0; q < NA; q++) esv[i].active_cases.push_back(esv_active[(size_t)i * NA + q]);
  }
  std::vector<std::vector<Eigen::Vector2d>> bend(N);
  for (int j = 0; j < N; j++)
    for (int q = 0; q < bp_cnt[j]; q++)
      bend[j].push_back(Eigen::Vector2d(bp_xy[((size_t)j * bp_max + q) * 2], bp_xy[((size_t)j * bp_max + q) * 2 + 1]));

  bool ok = false;
  double objective = 0.0;
  mt::PieceWisePol out;
  std::vector<mt::state> traj;
  for (int r = 0; r < replans; r++)
  {  // the object is long-lived in the reference (one per agent process): the same replan, `replans` times
    solver.setInitTrajectory(pwp);
    solver.setHulls(hulls);
    solver.setHullsNoInflation(nih);
    solver.setEntStateVector(esv, bend);
    objective = 0.0;
    ok = solver.optimize(objective);
    solver.generatePwpOut(out, traj, t_start, dc);
  }
  for (int q = 0; q < 96; q++) coeff_out[q] = 0.0;
  for (int i = 0; i < (int)out.coeff_x.size() && i < 8; i++)
    for (int c = 0; c < 4; c++)
    {
      coeff_out[4 * i + c] = out.coeff_x[i](c);
      coeff_out[32 + 4 * i + c] = out.coeff_y[i](c);
      coeff_out[64 + 4 * i + c] = out.coeff_z[i](c);
    }
  for (int i = 0; i < (int)out.times.size() && i < 9; i++) times_out[i] = out.times[i];
  *obj = objective;
  *n_states = (int)traj.size();
  return ok ? 1 : 0;
}

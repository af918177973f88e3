// test_shim.cpp -- exercises the C++ drop-in classes (neptune_b200/cpp/poly_solver_b200.hpp) the way
// neptune.cpp does (:102-107, :1514-1527).  Reads one agent's scene from a text file written by the
// Python test, prints the status and the optimised coefficients; the Python side compares them with
// the oracle.  Needs a GPU (it goes through libneptune_b200.so).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../neptune_b200/cpp/poly_solver_b200.hpp"

int main(int argc, char** argv)
{
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "r");
  if (!f) return 2;
  int N, id, n, NH;
  double T, W, lim[6], vmax, amax;
  if (fscanf(f, "%d %d %d %d %lf %lf", &N, &id, &n, &NH, &T, &W) != 6) return 2;
  for (int k = 0; k < 6; k++)
    if (fscanf(f, "%lf", &lim[k]) != 1) return 2;
  if (fscanf(f, "%lf %lf", &vmax, &amax) != 2) return 2;
  std::vector<Eigen::Vector2d> pb(N);
  for (int j = 0; j < N; j++)
    if (fscanf(f, "%lf %lf", &pb[j](0), &pb[j](1)) != 2) return 2;
  mt::PieceWisePol pwp;
  for (int i = 0; i <= n; i++) pwp.times.push_back(i * T);
  for (int ax = 0; ax < 3; ax++)
    for (int i = 0; i < n; i++)
    {
      double c[4];
      if (fscanf(f, "%lf %lf %lf %lf", &c[0], &c[1], &c[2], &c[3]) != 4) return 2;
      Eigen::Matrix<double, 4, 1> v(c[0], c[1], c[2], c[3]);
      (ax == 0 ? pwp.coeff_x : ax == 1 ? pwp.coeff_y : pwp.coeff_z).push_back(v);
    }
  mt::ConvexHullsOfCurves_Std2d hulls(NH), nih(N);
  for (int s = 0; s < NH; s++)
    for (int i = 0; i < n; i++)
    {
      int cnt;
      if (fscanf(f, "%d", &cnt) != 1) return 2;
      mt::Polygon_Std p(2, cnt);
      for (int c = 0; c < cnt; c++)
        if (fscanf(f, "%lf %lf", &p(0, c), &p(1, c)) != 2) return 2;
      hulls[s].push_back(p);
    }
  // optional section "ENT": entStateVec (alphas + active cases per interval), bend points per agent and column 0 of
  // hullsNoInflation -- the inputs of addEntangleConstraintForIJCase (solver_gurobi_poly.cpp:620-642, :715-784)
  std::vector<eu::ent_state> esv(n + 1);
  for (auto& e : esv) e.active_cases.assign(N, 0);
  std::vector<std::vector<Eigen::Vector2d>> bend(N);
  for (int j = 0; j < N; j++) bend[j].push_back(pb[j]);
  char tag[8] = { 0 };
  if (fscanf(f, "%7s", tag) == 1 && tag[0] == 'E')
  {
    for (int i = 0; i <= n; i++)
    {
      int na;
      if (fscanf(f, "%d", &na) != 1) return 2;
      for (int q = 0; q < na; q++)
      {
        int a0, a1;
        if (fscanf(f, "%d %d", &a0, &a1) != 2) return 2;
        esv[i].alphas.push_back(Eigen::Vector2i(a0, a1));
        esv[i].betas.push_back(0.0);
      }
      for (int j = 0; j < N; j++)
        if (fscanf(f, "%d", &esv[i].active_cases[j]) != 1) return 2;
    }
    for (int j = 0; j < N; j++)
    {
      int nb;
      if (fscanf(f, "%d", &nb) != 1) return 2;
      bend[j].assign(nb, Eigen::Vector2d(0, 0));
      for (int q = 0; q < nb; q++)
        if (fscanf(f, "%lf %lf", &bend[j][q](0), &bend[j][q](1)) != 2) return 2;
    }
    for (int j = 0; j < N; j++)
      for (int i = 0; i < n; i++)
      {
        int has;
        double x = 0, y = 0;
        if (fscanf(f, "%d %lf %lf", &has, &x, &y) != 3) return 2;
        mt::Polygon_Std p(2, has ? 1 : 0);
        if (has) p(0, 0) = x, p(1, 0) = y;
        nih[j].push_back(p);
      }
  }
  fclose(f);
  PolySolverGurobi solver(8, 3, id, T, pb, W, 0.5, true);
  solver.setMaxValues(lim[0], lim[1], lim[2], lim[3], lim[4], lim[5], vmax, amax, 5.0);
  solver.setMaxRuntime(0.05);
  solver.setTetherLength(40.0);
  solver.setInitTrajectory(pwp);
  solver.setHulls(hulls);
  solver.setHullsNoInflation(nih);
  solver.setEntStateVector(esv, bend);
  double obj = 0;
  const bool ok = solver.optimize(obj);
  mt::PieceWisePol out;
  std::vector<mt::state> traj;
  solver.generatePwpOut(out, traj, 1.25, 0.05);
  printf("%d %d %.17g %zu\n", ok ? 1 : 0, solver.lastStatus(), obj, traj.size());
  for (int ax = 0; ax < 3; ax++)
    for (int i = 0; i < n; i++)
    {
      const Eigen::Matrix<double, 4, 1>& c = (ax == 0 ? out.coeff_x : ax == 1 ? out.coeff_y : out.coeff_z)[i];
      printf("%.17g %.17g %.17g %.17g\n", c(0), c(1), c(2), c(3));
    }
  // the separator class on the first hull
  separator::Separator sep;
  Eigen::Vector3d line;
  mt::Polygon_Std B(2, 4);
  const double q[8] = { 100, 100, 101, 100, 101, 101, 100, 101 };
  for (int c = 0; c < 4; c++) B(0, c) = q[2 * c], B(1, c) = q[2 * c + 1];
  const bool s_ok = NH > 0 ? sep.solveModel(line, hulls[0][0], B) : false;
  printf("%d %.17g %.17g %.17g\n", s_ok ? 1 : 0, line(0), line(1), line(2));
  return 0;
}

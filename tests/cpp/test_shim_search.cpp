// test_shim_search.cpp -- exercises the C++ drop-in class KinodynamicSearch (neptune_b200/cpp/
// kinodynamic_search_b200.hpp) the way neptune.cpp does (:88-97 once; :1421-1453, :1509-1510 per replan).  Reads
// one agent's search inputs from a text file written by the Python test, prints status / pieces / coefficients /
// entStateVec; the Python side compares them with the oracle.  Needs a GPU (libneptune_b200.so).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../neptune_b200/cpp/kinodynamic_search_b200.hpp"

#define RD(fmt, ...) \
  if (fscanf(f, fmt, __VA_ARGS__) < 1) return 2

int main(int argc, char** argv)
{
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "r");
  if (!f) return 2;
  int N, id, np, S, ns, max_exp, max_nodes;
  double T, lim[4], vmax, amax, jmax, voxel, bias, goal_size, tether;
  RD("%d %d %d %d %d %d %d", &N, &id, &np, &S, &ns, &max_exp, &max_nodes);
  RD("%lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", &T, &lim[0], &lim[1], &lim[2], &lim[3], &vmax, &amax, &jmax, &voxel, &bias, &goal_size, &tether);
  std::vector<Eigen::Vector2d> pb(N);
  for (int j = 0; j < N; j++) RD("%lf %lf", &pb[j](0), &pb[j](1));
  KinodynamicSearch ks(np, 3, id, 0.6, T, S, pb, true, true);
  ks.setTetherLength(tether);
  ks.setMaxValuesAndSamples(vmax, amax, jmax, ns);
  ks.setXYZMinMaxAndRa(lim[0], lim[1], lim[2], lim[3], -1.0, 3.0, 5.0, voxel);
  ks.setBias(bias);
  ks.setGoalSize(goal_size);
  ks.setMaxExpansions(max_exp);
  std::vector<int> order(ns * ns);
  for (auto& o : order) RD("%d", &o);
  ks.setJerkOrder(order);
  mt::state A;
  Eigen::Vector3d goal;
  RD("%lf %lf %lf %lf %lf %lf %lf %lf", &A.pos(0), &A.pos(1), &A.vel(0), &A.vel(1), &A.accel(0), &A.accel(1), &goal(0), &goal(1));
  std::vector<Eigen::Matrix<double, 4, 1>> cz(np);
  for (int i = 0; i < np; i++) RD("%lf %lf %lf %lf", &cz[i](0), &cz[i](1), &cz[i](2), &cz[i](3));
  ks.setInitZCoeffs(cz);
  // known agents: samples [np][S+1] and hulls [np] (vertex count, then vertices)
  mt::SampledPointsofCurves spoc(N);
  mt::ConvexHullsOfCurves_Std2d hulls;
  for (int j = 0; j < N; j++)
  {
    int known;
    RD("%d", &known);
    if (!known) continue;
    mt::ConvexHullsOfCurve_Std2d hj;
    for (int i = 0; i < np; i++)
    {
      mt::PointsofInterval m(2, S + 1);
      for (int s = 0; s <= S; s++) RD("%lf %lf", &m(0, s), &m(1, s));
      spoc[j].push_back(m);
    }
    for (int i = 0; i < np; i++)
    {
      int nv;
      RD("%d", &nv);
      mt::Polygon_Std h(2, nv);
      for (int v = 0; v < nv; v++) RD("%lf %lf", &h(0, v), &h(1, v));
      hj.push_back(h);
    }
    hulls.push_back(hj);
  }
  eu::ent_state es;
  int na, nb;
  RD("%d %d", &na, &nb);
  for (int q = 0; q < na; q++)
  {
    int a0, a1;
    double be;
    RD("%d %d %lf", &a0, &a1, &be);
    es.alphas.push_back(Eigen::Vector2i(a0, a1)), es.betas.push_back(be);
  }
  for (int q = 0; q < nb; q++)
  {
    int bd;
    RD("%d", &bd);
    es.bendPointsIdx.push_back(bd);
  }
  for (int q = 0; q < N; q++)
  {
    int ac;
    RD("%d", &ac);
    es.active_cases.push_back(ac);
  }
  std::vector<std::vector<Eigen::Vector2d>> bend(N);
  for (int j = 0; j < N; j++) bend[j].push_back(pb[j]);
  fclose(f);
  ks.setUp(A, goal, hulls, spoc, es, bend);
  std::vector<Eigen::Vector3d> path;
  int status = -1;
  const bool ok = ks.run(path, status);
  mt::PieceWisePol pwp;
  std::vector<eu::ent_state> esv;
  ks.getPwpOut_0tstart(pwp);
  ks.getEntStateVector(esv);
  printf("%d %d %d\n", ok ? 1 : 0, status, (int)pwp.coeff_x.size());
  for (size_t i = 0; i < pwp.coeff_x.size(); i++)
    for (int ax = 0; ax < 3; ax++)
    {
      const Eigen::Matrix<double, 4, 1>& c = ax == 0 ? pwp.coeff_x[i] : (ax == 1 ? pwp.coeff_y[i] : pwp.coeff_z[i]);
      printf("%.17g %.17g %.17g %.17g\n", c(0), c(1), c(2), c(3));
    }
  for (auto& e : esv)
  {
    printf("%d", (int)e.alphas.size());
    for (auto& a : e.alphas) printf(" %d %d", a(0), a(1));
    printf("\n");
  }
  return 0;
}

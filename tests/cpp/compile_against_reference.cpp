// compile_against_reference.cpp -- the drop-in headers BESIDE the reference's own headers, in one translation unit.
//
// Built by tests/cpp/Makefile (target _build/test_shim_ref, only where /root/reference exists) the way INTEGRATION.md
// tells a maintainer to build neptune.cpp:
//     -I neptune_b200/cpp/dropin            (separator.hpp shadows submodules/separator/include/separator.hpp)
//     -I $REF/neptune/include               (mader_types.hpp, entangle_utils.hpp, utils.hpp, ... the reference's own)
//     -include poly_solver_b200.hpp -include kinodynamic_search_b200.hpp
// and, in this image, oracle/eigen_shim + oracle/ref_stubs for Eigen / ROS messages (test infrastructure; a real build
// uses real Eigen).  The reference's solver-facing headers are then included by name exactly as neptune.hpp includes
// them (neptune.hpp:13-19): their include guards are already defined, so PolySolverGurobi, KinodynamicSearch and
// separator::Separator resolve to the B200 classes while mt:: and eu:: types are the reference's own -- no type is
// defined twice.  The body is tests/cpp/test_shim.cpp: the call sequence of neptune.cpp:102-107 and :1514-1527.
#include "entangle_utils.hpp"
#include "mader_types.hpp"
#include "kinodynamic_search.hpp"
#include "solver_gurobi_poly.hpp"
#include "separator.hpp"
#include "utils.hpp"

#ifndef NB_HAVE_REFERENCE_TYPES
#error "the reference's headers were not found: nb_types.hpp fell back to its stand-ins"
#endif
// the reference's types, not stand-ins: members only the reference's definitions have
static_assert(sizeof(mt::parameters) > 0, "mt::parameters comes from the reference's mader_types.hpp");
static_assert(sizeof(mt::dynTrajCompiled) > 0, "mt::dynTrajCompiled comes from the reference's mader_types.hpp");

// members of the reference classes that Neptune calls and the drop-ins must offer (signatures as in
// solver_gurobi_poly.hpp:30-52, kinodynamic_search.hpp:118-199, separator.hpp:24-42)
template <class S>
void instantiate_solver_api(S& s, mt::PieceWisePol& pwp, mt::ConvexHullsOfCurves_Std2d& hulls, std::vector<mt::Polygon_Std>& st,
                            std::vector<eu::ent_state>& esv, std::vector<std::vector<Eigen::Vector2d>>& bend,
                            std::vector<std::vector<Eigen::Vector3d>>& betas, std::vector<mt::state>& traj)
{
  double obj;
  s.setMaxRuntime(0.05), s.setMaxValues(0, 1, 0, 1, 0, 1, 1, 1, 1), s.setTetherLength(1.0), s.setStaticObstVert(st);
  s.setInitTrajectory(pwp), s.setHulls(hulls), s.setHullsNoInflation(hulls), s.setBetasVector(betas), s.setEntStateVector(esv, bend);
  if (traj.size() == 123456789) s.optimize(obj);  // instantiated and linked, not executed here
  s.generatePwpOut(pwp, traj, 0.0, 0.05);
}
template <class K>
void instantiate_search_api(K& k, mt::state& st, Eigen::Vector3d& goal, mt::ConvexHullsOfCurves_Std2d& hulls, mt::SampledPointsofCurves& spoc,
                            eu::ent_state& es, std::vector<std::vector<Eigen::Vector2d>>& bend, mt::PieceWisePol& pwp,
                            std::vector<mt::state>& traj, std::vector<eu::ent_state>& esv)
{
  std::vector<Eigen::Matrix<double, 4, 1>> cz;
  std::vector<mt::Polygon_Std> sv;
  std::vector<Eigen::Matrix<double, 2, 2>> rep;
  std::vector<Eigen::Vector2d> longest;
  k.setMaxValuesAndSamples(1, 1, 1, 5), k.setXYZMinMaxAndRa(0, 1, 0, 1, 0, 1, 1, 0.2), k.setBias(1.1), k.setGoalSize(0.5), k.setRunTime(0.1);
  k.setTetherLength(1.0), k.setInitZCoeffs(cz), k.setStaticObstVert(sv), k.setStaticObstRep(rep, longest);
  k.setUp(st, goal, hulls, spoc, es, bend);
  std::vector<Eigen::Vector3d> path;
  int status = 0;
  if (traj.size() == 123456789) k.run(path, status), k.entangleCheckGivenPwp(pwp, es);
  k.clearProcess(), k.getPwpOut_0tstart(pwp), k.getEntStateVector(esv), k.generatePwpOut(pwp, traj, 0.0, 0.05);
  mt::SampledPointsofIntervals one;
  std::vector<Eigen::Vector2d> b1;
  k.updateSPocAndbendPtsForAgent(0, one, b1);
}
template void instantiate_solver_api<PolySolverGurobi>(PolySolverGurobi&, mt::PieceWisePol&, mt::ConvexHullsOfCurves_Std2d&,
                                                       std::vector<mt::Polygon_Std>&, std::vector<eu::ent_state>&,
                                                       std::vector<std::vector<Eigen::Vector2d>>&,
                                                       std::vector<std::vector<Eigen::Vector3d>>&, std::vector<mt::state>&);
template void instantiate_search_api<KinodynamicSearch>(KinodynamicSearch&, mt::state&, Eigen::Vector3d&, mt::ConvexHullsOfCurves_Std2d&,
                                                        mt::SampledPointsofCurves&, eu::ent_state&,
                                                        std::vector<std::vector<Eigen::Vector2d>>&, mt::PieceWisePol&,
                                                        std::vector<mt::state>&, std::vector<eu::ent_state>&);
static_assert(sizeof(&separator::Separator::meanSolveTimeMs) > 0 && sizeof(&separator::Separator::getNumOfLPsRun) > 0, "separator API");

#define main shim_main
#include "test_shim.cpp"
#undef main
int main(int argc, char** argv) { return shim_main(argc, argv); }

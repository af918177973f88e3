// Stand-in for DecompROS <decomp_geometry/polyhedron.h> (TEST INFRASTRUCTURE, oracle/_ref build).  The reference's
// solver_gurobi_poly.hpp only names Hyperplane3D as the element type of an unused member (planes_).
#pragma once
#include <Eigen/Dense>
struct Hyperplane3D
{
  Eigen::Vector3d p_, n_;
};

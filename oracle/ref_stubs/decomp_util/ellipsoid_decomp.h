// Stand-in for DecompROS <decomp_util/ellipsoid_decomp.h> (TEST INFRASTRUCTURE): included by
// solver_gurobi_poly.cpp "for Polyhedron definition", nothing of it is used on the path.
#pragma once
#include <decomp_geometry/polyhedron.h>

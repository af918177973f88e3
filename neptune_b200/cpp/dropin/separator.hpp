// separator.hpp -- drop-in for submodules/separator/include/separator.hpp: put this directory BEFORE the reference's
// separator include directory on the include path and every `#include "separator.hpp"` of the reference's tree
// (kinodynamic_search.hpp:15, solver_gurobi_poly.hpp:19) gets separator::Separator of poly_solver_b200.hpp instead of
// the GLPK-backed class.  See INTEGRATION.md.
#pragma once
#include "../poly_solver_b200.hpp"

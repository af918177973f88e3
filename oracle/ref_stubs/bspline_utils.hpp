// Shadows the reference's bspline_utils.hpp (needs unsupported/Eigen/Splines; unused by kinodynamic_search.cpp).
#pragma once

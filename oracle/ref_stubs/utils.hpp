// Shadows the reference's utils.hpp (ROS message helpers; none are used by kinodynamic_search.cpp) for the oracle/_ref build.
#pragma once
#include "mader_types.hpp"

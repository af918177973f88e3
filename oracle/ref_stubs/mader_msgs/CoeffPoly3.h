// Field-only stand-ins for mader_msgs/CoeffPoly3 and PieceWisePolTraj (TEST INFRASTRUCTURE, oracle/_ref build).
#pragma once
#include <vector>
namespace mader_msgs
{
struct CoeffPoly3
{
  double a = 0, b = 0, c = 0, d = 0;
};
struct PieceWisePolTraj
{
  std::vector<double> times;
  std::vector<CoeffPoly3> coeff_x, coeff_y, coeff_z;
};
}  // namespace mader_msgs

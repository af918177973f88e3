// nb_types.hpp -- the reference's value types on the back-end boundary.
//
// Three build situations, chosen by what is on the include path:
//  1. the reference's own headers (neptune/include/mader_types.hpp, entangle_utils.hpp) are reachable: they are
//     included and NOTHING they define is defined here -- this is the drop-in build of neptune.cpp;
//  2. only Eigen (<Eigen/Dense>) is reachable: the reference's typedefs are restated on real Eigen
//     (mader_types.hpp:20-30, :462-548; entangle_utils.hpp:23-29);
//  3. neither (this build image): a minimal Eigen stand-in with the same member access syntax, so that the shim can be
//     compiled and tested.
#pragma once
#include <vector>

#if defined(__has_include)
#if __has_include("mader_types.hpp") && __has_include("entangle_utils.hpp")
#define NB_HAVE_REFERENCE_TYPES 1
#endif
#if __has_include(<Eigen/Dense>)
#define NB_HAVE_EIGEN 1
#endif
#endif

#ifdef NB_HAVE_REFERENCE_TYPES
#include "mader_types.hpp"      // mt::PieceWisePol, mt::state, hull and sample typedefs
#include "entangle_utils.hpp"   // eu::ent_state
#else
#ifdef NB_HAVE_EIGEN
#include <Eigen/Dense>
#else
namespace Eigen
{
const int Dynamic = -1;
struct Vector2d
{
  double v[2];
  Vector2d() : v{ 0, 0 } {}
  Vector2d(double x, double y) : v{ x, y } {}
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
};
struct Vector2i
{
  int v[2];
  Vector2i() : v{ 0, 0 } {}
  Vector2i(int a, int b) : v{ a, b } {}
  int& operator()(int i) { return v[i]; }
  int operator()(int i) const { return v[i]; }
};
struct Vector3d
{
  double v[3];
  Vector3d() : v{ 0, 0, 0 } {}
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};
struct Vector4d
{
  double v[4];
  Vector4d() : v{ 0, 0, 0, 0 } {}
  Vector4d(double a, double b, double c, double d) : v{ a, b, c, d } {}
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};
template <typename T, int R, int C>
struct Matrix;
template <>
struct Matrix<double, 4, 1> : Vector4d
{
  using Vector4d::Vector4d;
};
template <>
struct Matrix<double, 2, Dynamic>
{  // 2 x n, column = point
  std::vector<double> d;
  Matrix() {}
  Matrix(int, int c) : d(2 * c, 0.0) {}
  int cols() const { return (int)d.size() / 2; }
  double& operator()(int r, int c) { return d[2 * c + r]; }
  double operator()(int r, int c) const { return d[2 * c + r]; }
};
}  // namespace Eigen
#endif

namespace mt
{
typedef Eigen::Matrix<double, 2, Eigen::Dynamic> Polygon_Std;          // mader_types.hpp:24
typedef std::vector<Polygon_Std> ConvexHullsOfCurve_Std2d;             // one hull per interval
typedef std::vector<ConvexHullsOfCurve_Std2d> ConvexHullsOfCurves_Std2d;  // one entry per obstacle

struct PieceWisePol  // mader_types.hpp:462-548
{
  std::vector<double> times;
  std::vector<Eigen::Matrix<double, 4, 1>> coeff_x, coeff_y, coeff_z;
  void clear()
  {
    times.clear(), coeff_x.clear(), coeff_y.clear(), coeff_z.clear();
  }
};

struct state  // mader_types.hpp:35-124 (fields written by generatePwpOut)
{
  Eigen::Vector3d pos, vel, accel, jerk;
  void setPos(const Eigen::Vector3d& p) { pos = p; }
  void setVel(const Eigen::Vector3d& p) { vel = p; }
  void setAccel(const Eigen::Vector3d& p) { accel = p; }
  void setJerk(const Eigen::Vector3d& p) { jerk = p; }
};
}  // namespace mt

namespace eu
{
struct ent_state  // entangle_utils.hpp:23-29
{
  std::vector<Eigen::Vector2i> alphas;
  std::vector<double> betas;
  std::vector<int> bendPointsIdx;
  std::vector<int> active_cases;
};
}  // namespace eu

namespace mt
{
typedef Eigen::Matrix<double, 2, Eigen::Dynamic> PointsofInterval;       // mader_types.hpp:22
typedef std::vector<PointsofInterval> SampledPointsofIntervals;           // :26, one 2 x (S+1) matrix per interval
typedef std::vector<SampledPointsofIntervals> SampledPointsofCurves;      // :30, indexed by agent id - 1, empty = unknown
}  // namespace mt
#endif  // NB_HAVE_REFERENCE_TYPES

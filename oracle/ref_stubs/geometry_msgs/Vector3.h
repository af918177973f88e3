// Field-only stand-ins for the geometry_msgs messages utils.hpp names (TEST INFRASTRUCTURE, oracle/_ref build).
#pragma once
namespace geometry_msgs
{
struct Vector3
{
  double x = 0, y = 0, z = 0;
};
struct Point
{
  double x = 0, y = 0, z = 0;
};
struct Quaternion
{
  double x = 0, y = 0, z = 0, w = 0;
};
struct Pose
{
  Point position;
  Quaternion orientation;
};
}  // namespace geometry_msgs

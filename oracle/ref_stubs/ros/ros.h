// Stand-in for <ros/ros.h> (TEST INFRASTRUCTURE): a clock for the reference's timer.hpp and utils.cpp, and the two names
// utils.hpp mentions in a template (NodeHandle, ROS_ERROR).  Nothing else of ROS is touched by the sources compiled into
// oracle/_ref.
#pragma once
#include <chrono>
#include <cstdio>
#include <string>
namespace ros
{
struct Time
{
  double s;
  static Time now()
  {
    return Time{ std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() };
  }
  double toSec() const { return s; }
};
typedef Time WallTime;
struct NodeHandle
{
  template <typename T>
  bool getParam(const std::string&, T&) const
  {
    return false;
  }
  std::string resolveName(const std::string& n, bool = true) const { return n; }
};
}  // namespace ros
#define ROS_ERROR(...) std::fprintf(stderr, __VA_ARGS__)

// Stand-in for <ros/ros.h> (TEST INFRASTRUCTURE): the reference's timer.hpp reads ros::Time::now(); nothing else of ROS is
// touched by the sources compiled into oracle/_ref.
#pragma once
#include <chrono>
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
}  // namespace ros

// Field-only stand-in for the ROS message of the same name (TEST INFRASTRUCTURE, oracle/_ref build).
#pragma once
#include <string>
#include "ros/ros.h"
namespace std_msgs
{
struct ColorRGBA
{
  float r = 0, g = 0, b = 0, a = 0;
};
struct Header
{
  std::string frame_id;
  ros::Time stamp;
};
}  // namespace std_msgs

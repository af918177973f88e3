#pragma once
#include "mader_msgs/CoeffPoly3.h"

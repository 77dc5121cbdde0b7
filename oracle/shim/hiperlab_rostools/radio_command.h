// oracle/shim -- TEST INFRASTRUCTURE: plain struct with the fields of AIFS_ROS/hiperlab_rostools/msg/radio_command.msg
// (roscpp zero-initialises fixed-size arrays and scalars of a default-constructed message)
#pragma once
#include <cstdint>
#include "ros/ros.h"
namespace hiperlab_rostools {
struct radio_command {
  std_msgs::Header header;
  uint8_t raw[32] = {0};
  uint8_t debugflags = 0;
  float debugvals[10] = {0};
  int32_t debugtype = 0;
};
}  // namespace hiperlab_rostools

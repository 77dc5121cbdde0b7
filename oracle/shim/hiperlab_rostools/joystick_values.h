// oracle/shim -- TEST INFRASTRUCTURE: plain struct with the fields of AIFS_ROS/hiperlab_rostools/msg/joystick_values.msg
#pragma once
#include <cstdint>
#include "ros/ros.h"
namespace hiperlab_rostools {
struct joystick_values {
  std_msgs::Header header;
  uint8_t buttonStart = 0, buttonRed = 0, buttonYellow = 0, buttonBlue = 0, buttonGreen = 0;
  float axes[4] = {0};
};
}  // namespace hiperlab_rostools

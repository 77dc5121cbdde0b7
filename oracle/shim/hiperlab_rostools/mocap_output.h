// oracle/shim -- TEST INFRASTRUCTURE: plain struct with the fields of AIFS_ROS/hiperlab_rostools/msg/mocap_output.msg
#pragma once
#include <cstdint>
#include "ros/ros.h"
namespace hiperlab_rostools {
struct mocap_output {
  std_msgs::Header header;
  int64_t vehicleID = 0;
  double posx = 0, posy = 0, posz = 0, attyaw = 0, attpitch = 0, attroll = 0, attq0 = 0, attq1 = 0, attq2 = 0, attq3 = 0;
};
}  // namespace hiperlab_rostools

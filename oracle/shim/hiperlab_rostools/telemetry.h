// oracle/shim -- TEST INFRASTRUCTURE: plain struct with the fields of AIFS_ROS/hiperlab_rostools/msg/telemetry.msg
#pragma once
#include <cstdint>
#include "ros/ros.h"
namespace hiperlab_rostools {
struct telemetry {
  std_msgs::Header header;
  uint8_t vehicleID = 0, type = 0, packetNumber = 0, seqNum = 0;
  double accelerometer[3] = {0}, rateGyro[3] = {0}, position[3] = {0}, attitude[3] = {0}, velocity[3] = {0}, attitudeYPR[3] = {0},
         motorForces[4] = {0}, debugVals[6] = {0}, batteryVoltage = 0;
  uint8_t panicReason = 0, warnings = 0;
};
}  // namespace hiperlab_rostools

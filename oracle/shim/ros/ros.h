// oracle/shim/ros/ros.h -- TEST INFRASTRUCTURE (oracle build only).
// Stand-in for the slice of roscpp that AIFS_ROS/hiperlab_rostools/src/QuadMocapRatesControl/ExampleVehicleStateMachine.{hpp,cpp}
// touches, so that the UNMODIFIED flight-stage state machine compiles into oracle/_ref: a NodeHandle whose subscribe /
// advertise do nothing, a Publisher that keeps the last message of each type where the harness can read it, and ros::Time.
// ROS itself (noetic) is not installed in this image and there is no network.
#pragma once
#include <iostream>  // roscpp pulls it in; the node uses cout
#include <memory>
#include <sstream>
#include <string>

namespace ros {

struct Time {
  double t = 0;
  static Time now() { return Time(); }
};

// the last message published, per message type (one vehicle per harness object at a time: the harness copies it out
// right after Run())
template<typename M>
struct LastPublished {
  static M& get() {
    static thread_local M m;
    return m;
  }
  static bool& fresh() {
    static thread_local bool f = false;
    return f;
  }
};

struct Publisher {
  template<typename M>
  void publish(const M& m) const {
    LastPublished<M>::get() = m;
    LastPublished<M>::fresh() = true;
  }
};
struct Subscriber {};

struct NodeHandle {
  template<typename M, typename T>
  Subscriber subscribe(const std::string&, unsigned, void (T::*)(const M&), T*) { return Subscriber(); }
  template<typename M>
  Subscriber subscribe(const std::string&, unsigned, void (*)(const M&)) { return Subscriber(); }
  template<typename M>
  Publisher advertise(const std::string&, unsigned) { return Publisher(); }
};

}  // namespace ros

namespace std_msgs {
struct Header {
  ros::Time stamp;
};
}  // namespace std_msgs

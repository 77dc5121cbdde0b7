// oracle/shim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE (oracle build only).
// The RAPPIDS planner (Components/DepthImagePlanner/DepthImagePlanner.hpp) only reads
// cv::Mat::{rows, cols, data}; OpenCV C++ headers are absent from this image.
#pragma once
#include <algorithm>
#include <vector>
namespace cv {
struct Mat {
  int rows;
  int cols;
  unsigned char* data;
  Mat() : rows(0), cols(0), data(nullptr) {}
};
}  // namespace cv

// oracle/shim -- see ../opencv.hpp
#pragma once
#include "opencv2/opencv.hpp"

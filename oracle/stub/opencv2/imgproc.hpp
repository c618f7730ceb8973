// stand-in for <opencv2/imgproc.hpp> (TEST INFRASTRUCTURE ONLY) -- see opencv.hpp in this directory
#pragma once
#include "opencv.hpp"

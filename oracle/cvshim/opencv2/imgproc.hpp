#include "imgproc/imgproc.hpp"
